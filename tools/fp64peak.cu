// Measured FP64 FMA peak of the device (denominator of the Cholesky update kernel's roofline).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64peak tools/fp64peak.cu && tools/fp64peak
#include <cuda_runtime.h>
#include <cstdio>

__global__ void k_dfma(double *out, double a, double b, int iters) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int threads = 512, blocks = p.multiProcessorCount * 4, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_dfma<<<blocks, threads>>>(out, 0.999999, 1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 16 * iters * (double)threads * blocks / ms / 1e9;
        if (tf > best) best = tf;
    }
    printf("{\"device\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f}\n", p.name, p.multiProcessorCount, best);
    return 0;
}
