// Stand-alone check + micro-benchmark of the streaming triangle-smoothing kernel
// (pyseistr_b200/csrc/pst_tri_stream.cu) against a plain CPU restatement of ps_smooth2.
//   nvcc -O3 -std=c++17 -fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo \
//        tools/mb_tri_stream.cu pyseistr_b200/csrc/pst_tri_stream.cu -o tools/mb_tri_stream.bin
//   tools/mb_tri_stream.bin            correctness on small volumes, all axes, several radii
//   tools/mb_tri_stream.bin bench      + timing on 1000x1024x1024
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../pyseistr_b200/csrc/pst_tri_stream.cuh"

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
    float *d_in, *d_out; unsigned *d_err;
    CK(cudaMalloc(&d_in, n * 4)); CK(cudaMalloc(&d_out, n * 4)); CK(cudaMalloc(&d_err, 4));
    CK(cudaMemset(d_err, 0, 4));
    CK(cudaMemcpy(d_in, h.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xff, n * 4));
    if (!pst_tri_stream_ok(axis, n1, n2, n3, nb, d_in, d_out)) { printf("  %dx%dx%d axis %d nb %d: not eligible\n", n1, n2, n3, axis, nb); cudaFree(d_in); cudaFree(d_out); cudaFree(d_err); return 0; }
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    int rc = pst_tri_stream_launch(0, sm, axis, d_in, inplace ? d_in : d_out, n1, n2, n3, nb, d_err);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc != 0 || e != cudaSuccess) { printf("  %dx%dx%d axis %d nb %d: launch rc %d, %s\n", n1, n2, n3, axis, nb, rc, cudaGetErrorString(e)); return 1; }
    std::vector<float> out(n);
    CK(cudaMemcpy(out.data(), inplace ? d_in : d_out, n * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0, first = 0;
    for (size_t i = 0; i < n; i++) if (memcmp(&out[i], &ref[i], 4) != 0) { if (!bad) first = i; bad++; }
    printf("  %dx%dx%d axis %d nb %d %s: %s", n1, n2, n3, axis, nb, inplace ? "in-place" : "out-of-place", bad ? "MISMATCH" : "bit-exact");
    if (bad) printf(" (%zu of %zu, first at %zu: got %g want %g)", bad, n, first, out[first], ref[first]);
    printf("\n");
    // fused epilogue: dst = eps*p + S(src), partial sums of dst^2
    if (!bad && !inplace) {
        std::vector<float> hp(n);
        for (size_t i = 0; i < n; i++) { s = s * 1664525u + 1013904223u; hp[i] = ((int)(s >> 8) % 2001 - 1000) * 1e-3f; }
        float *d_p; double *d_part;
        CK(cudaMalloc(&d_p, n * 4)); CK(cudaMalloc(&d_part, 4096 * 8 * sizeof(double)));
        CK(cudaMemcpy(d_p, hp.data(), n * 4, cudaMemcpyHostToDevice));
        pst_tri_stream_epi e{d_p, 0.75f, d_part, 8};
        int grid = 0;
        rc = pst_tri_stream_launch(0, sm, axis, d_in, d_out, n1, n2, n3, nb, d_err, &e, &grid);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (rc != 0 || e2 != cudaSuccess) { printf("    epilogue launch rc %d, %s\n", rc, cudaGetErrorString(e2)); return 1; }
        CK(cudaMemcpy(out.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
        std::vector<double> part((size_t)grid * 8);
        CK(cudaMemcpy(part.data(), d_part, part.size() * 8, cudaMemcpyDeviceToHost));
        double want = 0., got = 0.;
        size_t bad2 = 0;
        for (size_t i = 0; i < n; i++) {
            float gq = 0.75f * hp[i];
            gq += ref[i];
            want += (double)gq * gq;
            if (memcmp(&out[i], &gq, 4) != 0) bad2++;
        }
        for (int b = 0; b < grid; b++) got += part[(size_t)b * 8];
        const bool oksum = fabs(got - want) <= 1e-9 * fabs(want) + 1e-30;
        if (bad2 || !oksum) { printf("    epilogue MISMATCH: %zu values, sum %.17g vs %.17g\n", bad2, got, want); bad = 1; }
        cudaFree(d_p); cudaFree(d_part);
    }
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_err);
    return bad ? 1 : 0;
}

int main(int argc, char **argv)
{
    int fails = 0;
    const bool benchonly = argc >= 2 && !strcmp(argv[1], "benchonly");
    if (argc >= 7 && !strcmp(argv[1], "case")) return check(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), false);
    const int shapes[][3] = {{64, 40, 36}, {100, 70, 37}, {36, 33, 65}, {1000, 40, 8}, {8, 1024, 5}, {12, 6, 1030}, {256, 12, 12}};
    const int radii[] = {5, 2, 3, 8, 16};
    for (auto &sh : shapes)
        for (int axis = 0; axis < 3 && !benchonly; axis++)
            for (int nb : radii) {
                const int nx = sh[axis];
                if (nb > nx) continue;
                fails += check(sh[0], sh[1], sh[2], axis, nb, (nb & 1) != 0);
            }
    printf("correctness: %d failing cases\n", fails);
    if (fails || argc < 2 || (strcmp(argv[1], "bench") && !benchonly)) return fails ? 1 : 0;

    int n1 = 1000, n2 = 1024, n3 = 1024, nb = 5;
    if (argc >= 5) { n1 = atoi(argv[2]); n2 = atoi(argv[3]); n3 = atoi(argv[4]); }
    const size_t n = (size_t)n1 * n2 * n3;
    float *a, *b; unsigned *d_err;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&d_err, 4));
    CK(cudaMemset(d_err, 0, 4));
    CK(cudaMemset(a, 0, n * 4));
    {   // cheap non-trivial fill
        std::vector<float> h(1 << 20);
        unsigned s = 7u;
        for (auto &v : h) { s = s * 1664525u + 1013904223u; v = ((int)(s >> 8) % 2001 - 1000) * 1e-3f; }
        for (size_t o = 0; o < n; o += h.size()) CK(cudaMemcpy(a + o, h.data(), std::min(h.size(), n - o) * 4, cudaMemcpyHostToDevice));
    }
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int axis = 0; axis < 3; axis++) {
        for (int w = 0; w < 2; w++) pst_tri_stream_launch(0, sm, axis, a, b, n1, n2, n3, nb, d_err);
        CK(cudaDeviceSynchronize());
        const int reps = 5;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) pst_tri_stream_launch(0, sm, axis, r & 1 ? b : a, r & 1 ? a : b, n1, n2, n3, nb, d_err);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        ms /= reps;
        printf("bench %dx%dx%d axis %d nb %d: %.3f ms/pass  %.1f GB/s algorithmic (8 B/voxel)\n", n1, n2, n3, axis, nb, ms, 8.0 * n / ms * 1e-6);
    }
    {   // fused epilogue (dst = eps*p + S(src), sum dst^2): 12 B/voxel
        float *pbuf; double *d_part;
        CK(cudaMalloc(&pbuf, n * 4)); CK(cudaMalloc(&d_part, 4096 * 8 * sizeof(double)));
        CK(cudaMemset(pbuf, 0, n * 4));
        pst_tri_stream_epi e{pbuf, 1.0f, d_part, 8};
        for (int axis = 0; axis < 3; axis++) {
            pst_tri_stream_launch(0, sm, axis, a, b, n1, n2, n3, nb, d_err, &e);
            CK(cudaDeviceSynchronize());
            const int reps = 5;
            cudaEventRecord(e0);
            for (int r = 0; r < reps; r++) pst_tri_stream_launch(0, sm, axis, a, b, n1, n2, n3, nb, d_err, &e);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            ms /= reps;
            printf("bench+gp %dx%dx%d axis %d nb %d: %.3f ms/pass  %.1f GB/s algorithmic (12 B/voxel)\n", n1, n2, n3, axis, nb, ms, 12.0 * n / ms * 1e-6);
        }
    }
    return 0;
}
