// Microbenchmark 2: return-less reductions into an L2-resident delta block (design of the delta+fold ingest).
//   mode 0: red.global.add.u32 on 32-bit counters      mode 1: red.global.add.noftz.f16x2 on 16-bit (half) lanes
//   mode 2: same as 1 but 8 independent reds per thread from a coalesced index array (what k_scatter does)
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t v)
{
    v ^= v >> 33; v *= 0xff51afd7ed558ccdull; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ull; v ^= v >> 33;
    return v;
}
__device__ __forceinline__ void red_h2(void* addr, uint32_t v)
{
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(256) k(uint8_t* d, uint64_t bytes, uint64_t n, uint64_t seed)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = mix(i + seed);
        if (MODE == 0) {
            uint32_t bin = (uint32_t)(((r >> 32) * (bytes / 4)) >> 32);
            atomicAdd(reinterpret_cast<unsigned*>(d) + bin, 1u);
        } else {
            uint32_t bin = (uint32_t)(((r >> 32) * (bytes / 2)) >> 32);
            red_h2(d + (uint64_t)(bin >> 1) * 4, (bin & 1) ? 0x3C000000u : 0x00003C00u);
        }
    }
}
__global__ void __launch_bounds__(256) k_idx(uint8_t* d, const uint32_t* __restrict__ idx, uint64_t n)
{
    uint64_t i = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 8;
    if (i + 8 > n) return;
    uint4 a = __ldcs(reinterpret_cast<const uint4*>(idx + i));
    uint4 b = __ldcs(reinterpret_cast<const uint4*>(idx + i + 4));
    uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; j++) red_h2(d + (uint64_t)(v[j] >> 1) * 4, (v[j] & 1) ? 0x3C000000u : 0x00003C00u);
}
__global__ void k_fill(uint32_t* idx, uint64_t n, uint64_t bins, uint64_t seed)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        idx[i] = (uint32_t)(((mix(i + seed) >> 32) * bins) >> 32);
}
__global__ void k_check(const __half* d, uint64_t bins, unsigned long long* sum, unsigned* mx)
{
    unsigned long long s = 0; unsigned m = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < bins; i += (uint64_t)gridDim.x * blockDim.x) {
        unsigned v = (unsigned)__half2float(d[i]); s += v; m = v > m ? v : m;
    }
    atomicAdd(sum, s); atomicMax(mx, m);
}
int main()
{
    const uint64_t n = 1ull << 27;
    uint8_t* d; uint32_t* idx; unsigned long long* sum; unsigned* mx;
    cudaMalloc(&d, 1ull << 30); cudaMalloc(&idx, n * 4); cudaMalloc(&sum, 8); cudaMalloc(&mx, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%-8s %14s %14s %14s  (G updates/s)\n", "MB", "red.u32", "red.f16x2", "red.f16x2 idx8");
    for (uint64_t mb : {16, 50, 100, 128, 200, 400}) {
        uint64_t bytes = mb * 1000000ull;
        printf("%-8llu", (unsigned long long)mb);
        for (int mode = 0; mode < 3; mode++) {
            cudaMemset(d, 0, bytes);
            if (mode == 2) k_fill<<<148 * 8, 256>>>(idx, n, bytes / 2, 77);
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148 * 8, 256>>>(d, bytes, n, rep * n);
                if (mode == 1) k<1><<<148 * 8, 256>>>(d, bytes, n, rep * n);
                if (mode == 2) k_idx<<<(unsigned)(n / 8 / 256), 256>>>(d, idx, n);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf(" %14.2f", n / (best * 1e-3) / 1e9);
        }
        printf("\n");
    }
    // exactness of half lanes: 4 reps x n adds into `bins` lanes must sum exactly when no lane passes 2048
    {
        uint64_t bytes = 100000000ull, bins = bytes / 2;
        cudaMemset(d, 0, bytes); cudaMemset(sum, 0, 8); cudaMemset(mx, 0, 4);
        k_fill<<<148 * 8, 256>>>(idx, n, bins, 5);
        for (int r = 0; r < 3; r++) k_idx<<<(unsigned)(n / 8 / 256), 256>>>(d, idx, n);
        k_check<<<148 * 8, 256>>>((const __half*)d, bins, sum, mx);
        unsigned long long s; unsigned m; cudaMemcpy(&s, sum, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&m, mx, 4, cudaMemcpyDeviceToHost);
        printf("half-lane sum %llu expected %llu max lane %u\n", s, 3ull * n, m);
        // saturation: hammer one lane 5000 times
        cudaMemset(d, 0, 64);
        k_fill<<<1, 256>>>(idx, 8192, 1, 1);
        k_idx<<<4, 256>>>(d, idx, 8192);
        __half h[2]; cudaMemcpy(h, d, 4, cudaMemcpyDeviceToHost);
        printf("lane hammered 8192 times reads %g (saturates at 2048), neighbour %g\n", __half2float(h[0]), __half2float(h[1]));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
