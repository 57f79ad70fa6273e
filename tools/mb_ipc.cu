// single-warp issue-rate micro-benchmark (sm_100a): how many independent instructions per cycle can
// ONE warp issue?  nvcc -arch=sm_100a -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ilp(float *out, long long *cyc, int n)
{
    float a = out[threadIdx.x + 32];
    float s0 = out[threadIdx.x], s1 = s0 + 1, s2 = s0 + 2, s3 = s0 + 3, s4 = s0 + 4, s5 = s0 + 5, s6 = s0 + 6, s7 = s0 + 7;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        s0 += a; s1 += a; s2 += a; s3 += a; s4 += a; s5 += a; s6 += a; s7 += a;
        s0 += a; s1 += a; s2 += a; s3 += a; s4 += a; s5 += a; s6 += a; s7 += a;
    }
    long long t1 = clock64();
    out[threadIdx.x] = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// chain of dependent FADD with K independent FMULs interleaved per chain step
template <int K>
__global__ void k_chain_plus(float *out, long long *cyc, int n)
{
    float a = out[threadIdx.x + 32], s = out[threadIdx.x];
    float m[8];
    for (int q = 0; q < 8; q++) m[q] = a + q;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            s += m[u & 7];
#pragma unroll
            for (int q = 0; q < K; q++) m[(u + q + 1) & 7] *= 1.0001f;
        }
    }
    long long t1 = clock64();
    float r = s;
    for (int q = 0; q < 8; q++) r += m[q];
    out[threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 4096);
    for (int rep = 0; rep < 2; rep++) {
        k_ilp<<<1, 32>>>(out, cyc, 1000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("ILP8 independent FADD, 1 warp: %.2f cyc/instr\n", h / 16000.0);
        k_ilp<<<1, 128>>>(out, cyc, 1000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("ILP8 independent FADD, 4 warps (1/SMSP): %.2f cyc/instr/warp\n", h / 16000.0);
        k_chain_plus<0><<<1, 32>>>(out, cyc, 1000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("chain + 0 FMUL per step: %.2f cyc/step\n", h / 8000.0);
        k_chain_plus<1><<<1, 32>>>(out, cyc, 1000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("chain + 1 FMUL per step: %.2f cyc/step\n", h / 8000.0);
        k_chain_plus<2><<<1, 32>>>(out, cyc, 1000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("chain + 2 FMUL per step: %.2f cyc/step\n", h / 8000.0);
        k_chain_plus<3><<<1, 32>>>(out, cyc, 1000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("chain + 3 FMUL per step: %.2f cyc/step\n", h / 8000.0);
        k_chain_plus<5><<<1, 32>>>(out, cyc, 1000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("chain + 5 FMUL per step: %.2f cyc/step\n", h / 8000.0);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
