// Micro-benchmarks that ground the kernel design: FP64 FMA latency / throughput, L1-hit load latency.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_latency(double* out, long long* cyc, double a, double b) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x = fma(x, b, a);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void dfma_tput(double* out, double a, double b, int iters) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], b, a);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ldg_latency(const int* idx, int* out, long long* cyc) {
    int j = 0;
    for (int i = 0; i < 64; ++i) j = idx[j];  // warm L1
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) j = __ldg(idx + j);
    long long t1 = clock64();
    out[0] = j;
    *cyc = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 24);
    cudaMallocManaged(&cyc, 8);
    for (int warps : {1, 2, 4, 8}) {
        dfma_latency<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999);
        cudaDeviceSynchronize();
        printf("dependent DFMA chain, %d warps/SM: %.2f cycles per DFMA\n", warps, (double) *cyc / (1024.0 * 16));
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        dfma_tput<8><<<148, threads>>>(out, 1.0000001, 0.9999999, iters);
        cudaEventRecord(e0);
        dfma_tput<8><<<148, threads>>>(out, 1.0000001, 0.9999999, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * 148.0 * threads;
        printf("DFMA throughput, 148 CTAs x %4d threads, ILP 8: %.2f TFLOP/s\n", threads, fl / ms * 1e-9);
    }
    int* idx;
    cudaMallocManaged(&idx, 4096 * 4);
    for (int i = 0; i < 4096; ++i) idx[i] = (i + 32) % 4096;
    int* o2;
    cudaMalloc(&o2, 4);
    ldg_latency<<<1, 1>>>(idx, o2, cyc);
    cudaDeviceSynchronize();
    printf("dependent LDG (L1 hit) latency: %.1f cycles\n", (double) *cyc / 1024.0);
    return 0;
}
