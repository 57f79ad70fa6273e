// Launcher of the L2-resident checkpoint + recompute triangle smoother (pst_tri_l2.cu).
#pragma once
#include <cuda_runtime.h>

// true when the kernel can smooth axis `axis` (0, 1, 2) of an n1 x n2 x n3 volume with radius nb
bool pst_tri_l2_ok(int axis, int n1, int n2, int n3, int nb, const void *src, const void *dst);
// src -> dst (may alias).  0 = launched, -1 = not eligible, -2 / -3 = tensor map set-up refused (nothing launched),
// -4 = attribute, -5 = launch failed
int pst_tri_l2_launch(cudaStream_t stream, int sm_count, int axis, const float *src, float *dst, int n1, int n2, int n3, int nb);
