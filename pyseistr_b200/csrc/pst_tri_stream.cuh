// Streaming (persistent, warp-specialised, TMA-staged) triangle smoothing of one axis.
// See pst_tri_stream.cu.  Results are bit-identical to ps_smooth2 (reference dip_cfuns.c:564-580).
#pragma once
#include <cuda_runtime.h>

// true when the streaming kernel can run this axis (axis 0/1/2 of an n1 x n2 x n3 volume, radius nb)
bool pst_tri_stream_ok(int axis, int n1, int n2, int n3, int nb, const void *src, const void *dst);
// optional epilogue fused into the store side: dst = eps*p + S(src), per-CTA partial sums of dst^2
// (double) at partial[cta * partial_stride]
struct pst_tri_stream_epi { const float *p; float eps; double *partial; int partial_stride; };
// src -> dst (dst may alias src).  Returns 0, or < 0 when the launch could not be set up.
// *grid_out receives the number of CTAs (= partial sums written).
int pst_tri_stream_launch(cudaStream_t stream, int sm_count, int axis, const float *src, float *dst,
                          int n1, int n2, int n3, int nb, unsigned *d_err,
                          const pst_tri_stream_epi *epi = nullptr, int *grid_out = nullptr);
