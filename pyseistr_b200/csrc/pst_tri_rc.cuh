// Checkpoint + recompute triangle smoothing of one axis (experimental, see pst_tri_rc.cu).
// Results are bit-identical to ps_smooth2 (reference dip_cfuns.c:564-625).
#pragma once
#include <cuda_runtime.h>

// rc_pref: preferred block length (16 or 32; 16 only when 2 nb <= 16, and never on the contiguous axis)
bool pst_tri_rc_ok(int axis, int n1, int n2, int n3, int nb, int rc_pref);
// src -> dst (dst may alias src).  Returns 0, -1 (shape refused, nothing launched), -4 / -5 (CUDA).
int pst_tri_rc_launch(cudaStream_t stream, int axis, const float *src, float *dst, int n1, int n2, int n3, int nb,
                      int rc_pref);
