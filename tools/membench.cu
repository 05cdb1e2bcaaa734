// Read-only / copy streaming micro-benchmark: what HBM rate can a plain kernel reach on this GPU?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int U>
__global__ void k_read(const double2 *__restrict__ a, size_t n, double *out) {
    double s = 0.0;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = a[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; u++) s += v[u].x + v[u].y;
    }
    for (; i < n; i += stride) s += a[i].x + a[i].y;
    if (s == 1.2345e-300) *out = s;
}
__global__ void k_copy(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = a[i];
}
int main() {
    size_t bytes = (size_t)8 << 30, n = bytes / 16;
    double2 *a, *b; double *o;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 8);
    cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int blocks_per_sm : {4, 8, 16}) {
        for (int rep = 0; rep < 2; rep++) {
            float ms;
            cudaEventRecord(e0); for (int i = 0; i < 5; i++) k_read<8><<<sms * blocks_per_sm, 256>>>(a, n, o); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("read  U=8 %2d CTA/SM: %.1f GB/s\n", blocks_per_sm, 5.0 * bytes / ms / 1e6);
            cudaEventRecord(e0); for (int i = 0; i < 5; i++) k_read<4><<<sms * blocks_per_sm, 256>>>(a, n, o); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("read  U=4 %2d CTA/SM: %.1f GB/s\n", blocks_per_sm, 5.0 * bytes / ms / 1e6);
            cudaEventRecord(e0); for (int i = 0; i < 5; i++) k_copy<<<sms * blocks_per_sm, 256>>>(a, b, n); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("copy      %2d CTA/SM: %.1f GB/s (read+write)\n", blocks_per_sm, 2.0 * 5.0 * bytes / ms / 1e6);
        }
    }
    float ms;
    cudaEventRecord(e0); for (int i = 0; i < 5; i++) cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D: %.1f GB/s (read+write)\n", 2.0 * 5.0 * bytes / ms / 1e6);
    return 0;
}
