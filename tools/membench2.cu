// Which access pattern costs HBM read bandwidth?  (diagnostic for the SpMV roofline)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench2 tools/membench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_read8(const double *__restrict__ a, size_t n, double *out) {       // 8 B per lane
    double s = 0.0;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 7 * stride < n; i += 8 * stride) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = a[i + u * stride];
#pragma unroll
        for (int u = 0; u < 8; u++) s += v[u];
    }
    if (s == 1.2345e-300) *out = s;
}
// persistent CTAs walking 33 KB chunks round-robin (the SpMV tile schedule), plain loads, optional sparse writes
__global__ void k_chunks(const double *__restrict__ a, size_t n, size_t chunk, double *out, double *y, int do_write) {
    double s = 0.0;
    size_t nch = n / chunk;
    for (size_t c = blockIdx.x; c < nch; c += gridDim.x) {
        const double *p = a + c * chunk;
        for (size_t i = threadIdx.x; i < chunk; i += blockDim.x) s += p[i];
        if (do_write && threadIdx.x < 48) y[c * 48 + threadIdx.x] = s;
    }
    if (s == 1.2345e-300) *out = s;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// same schedule, data moved by cp.async.bulk into a 6-stage ring, consumers only wait
__global__ void k_bulk(const double *__restrict__ a, size_t n, size_t chunk, double *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[6];
    const int S = 6;
    size_t nch = n / chunk, my = blockIdx.x < nch ? (nch - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    uint32_t bytes = (uint32_t)(chunk * 8);
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    auto issue = [&](size_t i) {
        int st = i % S;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[st])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + (size_t)st * bytes)), "l"(a + (blockIdx.x + i * gridDim.x) * chunk), "r"(bytes),
                     "r"(smem_u32(&full[st])) : "memory");
    };
    if (threadIdx.x == 0) for (size_t i = 0; i < S - 1 && i < my; i++) issue(i);
    double s = 0;
    for (size_t i = 0; i < my; i++) {
        if (threadIdx.x == 0 && i + S - 1 < my) issue(i + S - 1);
        uint32_t ok, par = (i / S) & 1;
        do { asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&full[i % S])), "r"(par) : "memory"); } while (!ok);
        s += reinterpret_cast<double *>(smem + (i % S) * (size_t)bytes)[threadIdx.x];
        __syncthreads();
    }
    if (s == 1.2345e-300) *out = s;
}
template <typename F> static float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaEventRecord(e0); for (int i = 0; i < 3; i++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 3;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *o, *y; cudaMalloc(&o, 8); cudaMalloc(&y, (size_t)1 << 30);
    for (int mode = 0; mode < 2; mode++) {
        size_t bytes = (size_t)20 << 30, n = bytes / 8;
        double *a;
        if (mode == 0) cudaMalloc(&a, bytes); else cudaMallocAsync(&a, bytes, 0);
        cudaMemset(a, 0, bytes); cudaDeviceSynchronize();
        const char *nm = mode ? "mallocAsync" : "malloc";
        size_t chunk = 4176;   // doubles: 33408 B, multiple of 16
        printf("[%s 20GB] read 8B/lane grid-stride : %.0f GB/s\n", nm, bytes / timeit([&] { k_read8<<<sms * 8, 256>>>(a, n, o); }) / 1e6);
        printf("[%s 20GB] 33KB chunks round-robin   : %.0f GB/s\n", nm, bytes / timeit([&] { k_chunks<<<sms * 4, 512>>>(a, n, chunk, o, y, 0); }) / 1e6);
        printf("[%s 20GB] 33KB chunks + 1%% writes   : %.0f GB/s\n", nm, bytes / timeit([&] { k_chunks<<<sms * 4, 512>>>(a, n, chunk, o, y, 1); }) / 1e6);
        cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 33408);
        printf("[%s 20GB] bulk-copy ring 6x33KB, 1/SM: %.0f GB/s\n", nm, bytes / timeit([&] { k_bulk<<<sms, 256, 6 * 33408>>>(a, n, chunk, o); }) / 1e6);
        cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 16704);
        printf("[%s 20GB] bulk-copy ring 6x16KB, 2/SM: %.0f GB/s\n", nm, bytes / timeit([&] { k_bulk<<<sms * 2, 256, 6 * 16704>>>(a, n, chunk / 2, o); }) / 1e6);
        if (mode == 0) cudaFree(a); else cudaFreeAsync(a, 0);
        cudaDeviceSynchronize();
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
