// Micro-benchmark: cost per element of the serial running-sum step for one warp (smem round trip
// + dependent FADD), in the layouts used by tri_tile_kernel.  nvcc -arch=sm_100a -fmad=false.
#include <cstdio>
#include <cuda_runtime.h>

#define NP 720

__global__ void k_chain_regs(float *out, long long *cyc, int n)
{
    float s = out[threadIdx.x], a = out[threadIdx.x + 32];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { s += a; s += a; s += a; s += a; s += a; s += a; s += a; s += a; }
    long long t1 = clock64();
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// strided layout [k][16], scalar LDS/STS, `lanes` active lanes, plain loop (compiler-scheduled)
__global__ void k_strided(float *out, long long *cyc, int lanes)
{
    __shared__ float tile[(NP + 16) * 16];
    for (int i = threadIdx.x; i < (NP + 16) * 16; i += blockDim.x) tile[i] = (float)(i % 7) * 0.25f;
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x < lanes) {
        float *col = tile + 8 * 16 + (threadIdx.x & 15);
        float s = 0.f;
        for (int k = 0; k + 8 <= NP; k += 8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = col[(k + q) * 16];
#pragma unroll
            for (int q = 0; q < 8; q++) { s += v[q]; col[(k + q) * 16] = s; }
        }
        out[threadIdx.x] = s;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// same, but each thread owns TWO lines (ILP 2): columns c and c+8
__global__ void k_strided_ilp2(float *out, long long *cyc)
{
    __shared__ float tile[(NP + 16) * 16];
    for (int i = threadIdx.x; i < (NP + 16) * 16; i += blockDim.x) tile[i] = (float)(i % 7) * 0.25f;
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x < 8) {
        float *col = tile + 8 * 16 + threadIdx.x;
        float s0 = 0.f, s1 = 0.f;
        for (int k = 0; k + 8 <= NP; k += 8) {
            float v[8], u[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { v[q] = col[(k + q) * 16]; u[q] = col[(k + q) * 16 + 8]; }
#pragma unroll
            for (int q = 0; q < 8; q++) { s0 += v[q]; col[(k + q) * 16] = s0; s1 += u[q]; col[(k + q) * 16 + 8] = s1; }
        }
        out[threadIdx.x] = s0 + s1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// k-minor groups of 4: layout [k/4][16][4], LDS.128/STS.128
__global__ void k_strided_v4(float *out, long long *cyc, int lanes)
{
    __shared__ __align__(16) float tile[(NP + 16) * 16];
    for (int i = threadIdx.x; i < (NP + 16) * 16; i += blockDim.x) tile[i] = (float)(i % 7) * 0.25f;
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x < lanes) {
        float4 *col = reinterpret_cast<float4 *>(tile) + 32 + (threadIdx.x & 15);
        float s = 0.f;
        float4 cur = col[0];
        for (int g = 0; g < NP / 4; g++) {
            float4 nxt = col[(g + 1) * 16];
            s += cur.x; cur.x = s; s += cur.y; cur.y = s; s += cur.z; cur.z = s; s += cur.w; cur.w = s;
            col[g * 16] = cur;
            cur = nxt;
        }
        out[threadIdx.x] = s;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
    cudaMemset(out, 0, 4096);
    for (int rep = 0; rep < 2; rep++) {
        k_chain_regs<<<1, 32>>>(out, cyc, 1000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("chain regs: %.2f cyc/FADD\n", h / 8000.0);
        for (int lanes : {16, 32}) {
            k_strided<<<1, 32>>>(out, cyc, lanes); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("strided scalar lanes=%d: %.2f cyc/elem\n", lanes, h / (double)NP);
        }
        k_strided<<<1, 128>>>(out, cyc, 16); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("strided scalar lanes=16 (128 thr block): %.2f cyc/elem\n", h / (double)NP);
        k_strided_ilp2<<<1, 32>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("strided scalar ILP2 (8 lanes x 2 lines): %.2f cyc/elem-step\n", h / (double)NP);
        k_strided_v4<<<1, 32>>>(out, cyc, 16); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("strided float4 groups lanes=16: %.2f cyc/elem\n", h / (double)NP);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
