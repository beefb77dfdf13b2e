// Microbenchmark: random counter-update throughput vs table size on one B200 — the second ceiling SURVEY.md §8d
// asks for (L2-resident atomics vs HBM-resident).  Variants: ld.cg + atomicCAS on the containing word (what
// k_ingest does), atomicAdd with return, red (no return), plain load only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_ceiling atomic_ceiling.cu && ./atomic_ceiling
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t v)
{
    v ^= v >> 33; v *= 0xff51afd7ed558ccdull; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ull; v ^= v >> 33;
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_upd(uint8_t* table, uint64_t size, uint64_t n, uint64_t seed, unsigned long long* sink)
{
    unsigned acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t bin = mix(i + seed) % size;
        uint32_t* word = reinterpret_cast<uint32_t*>(table + (bin & ~3ull));
        uint32_t sh = (bin & 3) * 8;
        if (MODE == 0) {  // ld.cg + CAS, saturating
            uint32_t cur = __ldcg(word);
            while (true) {
                uint32_t b = (cur >> sh) & 255u;
                if (b == 255u) break;
                uint32_t seen = atomicCAS(word, cur, cur + (1u << sh));
                if (seen == cur) { acc += b == 0; break; }
                cur = seen;
            }
        } else if (MODE == 1) {  // atomicAdd with return
            uint32_t old = atomicAdd(word, 1u << sh);
            acc += ((old >> sh) & 255u) == 0;
        } else if (MODE == 2) {  // red
            atomicAdd(word, 1u << sh);
        } else {  // load only
            acc += (__ldcg(word) >> sh) & 1u;
        }
    }
    if (acc == 0xffffffffu) *sink = acc;
}

int main()
{
    const uint64_t n = 1ull << 27;  // 134M updates per launch
    uint8_t* t;
    unsigned long long* sink;
    const uint64_t maxb = 4ull << 30;
    cudaMalloc(&t, maxb);
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"ld.cg+CAS", "atomicAdd(ret)", "red.add", "ld.cg only"};
    printf("%-10s", "MB");
    for (int m = 0; m < 4; m++) printf("%18s", names[m]);
    printf("   (G updates/s)\n");
    for (uint64_t mb : {8, 16, 32, 48, 64, 80, 100, 128, 200, 400, 1000, 4000}) {
        uint64_t size = mb * 1000000ull - 11;
        printf("%-10llu", (unsigned long long)mb);
        for (int mode = 0; mode < 4; mode++) {
            cudaMemset(t, 0, size + 16);
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                int g = 148 * 8;
                if (mode == 0) k_upd<0><<<g, 256>>>(t, size, n, rep * n, sink);
                if (mode == 1) k_upd<1><<<g, 256>>>(t, size, n, rep * n, sink);
                if (mode == 2) k_upd<2><<<g, 256>>>(t, size, n, rep * n, sink);
                if (mode == 3) k_upd<3><<<g, 256>>>(t, size, n, rep * n, sink);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("%18.2f", n / (best * 1e-3) / 1e9);
        }
        printf("\n");
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
