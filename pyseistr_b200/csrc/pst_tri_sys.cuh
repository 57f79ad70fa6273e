// Systolic triangle smoothing of a strided axis (experimental, see pst_tri_sys.cu / pst_tri_sys_kernels.cuh).
// Results are bit-identical to ps_smooth2 (reference dip_cfuns.c:564-625).
#pragma once
#include <cuda_runtime.h>

// true when the kernel can run this axis (1 or 2 of an n1 x n2 x n3 volume, radius nb)
bool pst_tri_sys_ok(int axis, int n1, int n2, int n3, int nb, const void *src, const void *dst);
// src -> dst (dst may alias src).  Returns 0, -1..-4 (refused / set-up failed, nothing launched), -5 (launch failed).
int pst_tri_sys_launch(cudaStream_t stream, int sm_count, int axis, const float *src, float *dst, int n1, int n2,
                       int n3, int nb, unsigned *d_err);
