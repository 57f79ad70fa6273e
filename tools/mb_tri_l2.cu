// Stand-alone check + micro-benchmark of the L2-resident checkpoint + recompute triangle smoother
// (pyseistr_b200/csrc/pst_tri_l2.cu) against a plain CPU restatement of ps_smooth2.
//   nvcc -O3 -std=c++17 -fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo \
//        tools/mb_tri_l2.cu pyseistr_b200/csrc/pst_tri_l2.cu -o tools/mb_tri_l2.bin
//   tools/mb_tri_l2.bin                        correctness: small volumes, all axes, radii, in place
//   tools/mb_tri_l2.bin bench [n1 n2 n3 [nb]]  + timing (default 1000x1024x1024, nb 5); PST_TRI_L2_WARPS / _SLOTS / _HINTS tune it
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../pyseistr_b200/csrc/pst_tri_l2.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)

static void cpu_smooth_line(float *x, long d, int nx, int nb, std::vector<float> &tmp)
{
    const int np = nx + 2 * nb;
    const float wt = (float)(1.0 / ((double)nb * nb));
    const float wm = -wt, w2 = (float)(2. * wt);
    tmp.assign(np, 0.f);
    for (int i = 0; i < nx; i++) tmp[i] += wm * x[i * d];
    for (int i = 0; i < nx; i++) tmp[i + nb] += w2 * x[i * d];
    for (int i = 0; i < nx; i++) tmp[i + 2 * nb] += wm * x[i * d];
    float s = 0.f;
    for (int k = 0; k < np; k++) { s += tmp[k]; tmp[k] = s; }
    s = 0.f;
    for (int k = np - 1; k >= 0; k--) { s += tmp[k]; tmp[k] = s; }
    for (int i = 0; i < nx; i++) {
        float v = tmp[i + nb];
        if (i >= nx - nb) v = v + tmp[nb + nx + (nx - 1 - i)];
        if (i < nb) v = v + tmp[nb - 1 - i];
        x[i * d] = v;
    }
}

static void cpu_smooth(std::vector<float> &v, int axis, int n1, int n2, int n3, int nb)
{
    std::vector<float> tmp;
    if (axis == 0) for (long l = 0; l < (long)n2 * n3; l++) cpu_smooth_line(&v[l * n1], 1, n1, nb, tmp);
    else if (axis == 1) { for (int i3 = 0; i3 < n3; i3++) for (int i1 = 0; i1 < n1; i1++) cpu_smooth_line(&v[i1 + (long)n1 * n2 * i3], n1, n2, nb, tmp); }
    else for (long l = 0; l < (long)n1 * n2; l++) cpu_smooth_line(&v[l], (long)n1 * n2, n3, nb, tmp);
}

static int check(int n1, int n2, int n3, int axis, int nb, bool inplace)
{
    const size_t n = (size_t)n1 * n2 * n3;
    std::vector<float> h(n), ref;
    unsigned s = 12345u + n1 * 7 + n2 * 13 + n3 * 17 + axis + nb * 3;
    for (size_t i = 0; i < n; i++) { s = s * 1664525u + 1013904223u; h[i] = ((int)(s >> 8) % 20001 - 10000) * 1e-4f; }
    ref = h;
    cpu_smooth(ref, axis, n1, n2, n3, nb);
    float *d_in, *d_out;
    CK(cudaMalloc(&d_in, n * 4)); CK(cudaMalloc(&d_out, n * 4));
    if (!pst_tri_l2_ok(axis, n1, n2, n3, nb, d_in, d_out)) { cudaFree(d_in); cudaFree(d_out); printf("  %dx%dx%d axis %d nb %d: not eligible\n", n1, n2, n3, axis, nb); return 0; }
    CK(cudaMemcpy(d_in, h.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xff, n * 4));
    int rc = pst_tri_l2_launch(0, 148, axis, d_in, inplace ? d_in : d_out, n1, n2, n3, nb);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc != 0 || e != cudaSuccess) { printf("  %dx%dx%d axis %d nb %d: launch rc %d, %s\n", n1, n2, n3, axis, nb, rc, cudaGetErrorString(e)); return 1; }
    std::vector<float> out(n);
    CK(cudaMemcpy(out.data(), inplace ? d_in : d_out, n * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0, first = 0;
    for (size_t i = 0; i < n; i++) if (memcmp(&out[i], &ref[i], 4) != 0) { if (!bad) first = i; bad++; }
    printf("  %dx%dx%d axis %d nb %d %s: %s", n1, n2, n3, axis, nb, inplace ? "in-place" : "out-of-place", bad ? "MISMATCH" : "bit-exact");
    if (bad) printf(" (%zu of %zu, first at %zu: got %g want %g)", bad, n, first, out[first], ref[first]);
    printf("\n");
    cudaFree(d_in); cudaFree(d_out);
    return bad ? 1 : 0;
}

int main(int argc, char **argv)
{
    int fails = 0;
    const bool benchonly = argc >= 2 && !strcmp(argv[1], "benchonly");
    const int shapes[][3] = {{64, 40, 36}, {100, 70, 37}, {36, 33, 65}, {1000, 40, 8}, {8, 1024, 5}, {12, 6, 1030}, {256, 12, 12}, {32, 130, 3}, {2000, 9, 70}, {68, 33, 300}};
    const int radii[] = {5, 2, 3, 8, 10, 16};
    for (auto &sh : shapes) {
        if (benchonly) break;
        for (int axis = 0; axis < 3; axis++)
            for (int nb : radii) {
                if (nb > sh[axis]) continue;
                for (int ip = 0; ip < 2; ip++) fails += check(sh[0], sh[1], sh[2], axis, nb, ip != 0);
            }
    }
    printf("correctness: %d failing cases\n", fails);
    if (fails || argc < 2 || (strcmp(argv[1], "bench") && !benchonly)) return fails ? 1 : 0;

    int n1 = 1000, n2 = 1024, n3 = 1024, nb = 5;
    if (argc >= 5) { n1 = atoi(argv[2]); n2 = atoi(argv[3]); n3 = atoi(argv[4]); }
    if (argc >= 6) nb = atoi(argv[5]);
    const size_t n = (size_t)n1 * n2 * n3;
    float *a, *b;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
    CK(cudaMemset(a, 0, n * 4));
    {
        std::vector<float> h(1 << 20);
        unsigned s = 7u;
        for (auto &v : h) { s = s * 1664525u + 1013904223u; v = ((int)(s >> 8) % 2001 - 1000) * 1e-3f; }
        for (size_t o = 0; o < n; o += h.size()) CK(cudaMemcpy(a + o, h.data(), std::min(h.size(), n - o) * 4, cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int axis = 0; axis < 3; axis++)
        for (int inplace = 0; inplace < 2; inplace++) {
            for (int w = 0; w < 2; w++) pst_tri_l2_launch(0, 148, axis, a, inplace ? a : b, n1, n2, n3, nb);
            CK(cudaDeviceSynchronize());
            const int reps = 5;
            cudaEventRecord(e0);
            for (int r = 0; r < reps; r++) {
                if (inplace) pst_tri_l2_launch(0, 148, axis, a, a, n1, n2, n3, nb);
                else pst_tri_l2_launch(0, 148, axis, r & 1 ? b : a, r & 1 ? a : b, n1, n2, n3, nb);
            }
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            ms /= reps;
            printf("bench %dx%dx%d axis %d nb %d %s: %.3f ms/pass  %.1f GB/s algorithmic (8 B/voxel)\n",
                   n1, n2, n3, axis, nb, inplace ? "in-place" : "out-of-place", ms, 8.0 * n / ms * 1e-6);
        }
    return 0;
}
