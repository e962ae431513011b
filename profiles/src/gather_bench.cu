// gather_bench.cu -- measurement utility (not product): the B200's random-access ceiling for the seed-lookup phase.
// Random aligned accesses of 8 / 16 / 32 / 64 / 128 bytes over a table of T GiB, U independent accesses in flight per thread.
// Prints accesses/s and useful GB/s per configuration; the 32-byte row is the denominator ("random 32 B-sector gather peak")
// bench.py's seed_phase reports against.   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

__device__ __forceinline__ uint64_t fmix64(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }

template <int BYTES, int U>
__global__ void __launch_bounds__(256) gather(const uint8_t *tbl, uint64_t nUnits, uint64_t nAccess, uint64_t salt, unsigned long long *sink)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nT = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = t * U; i < nAccess; i += nT * U) {
        uint4 v[U][(BYTES + 15) / 16];
        #pragma unroll
        for (int u = 0; u < U; u++) {
            uint64_t h = fmix64((i + u) ^ salt);
            uint64_t unit = __umul64hi(h, nUnits);
            const uint8_t *p = tbl + unit * BYTES;
            if (BYTES == 8) { uint2 w = __ldg((const uint2 *)p); v[u][0] = make_uint4(w.x, w.y, 0, 0); }
            else {
                #pragma unroll
                for (int k = 0; k < BYTES / 16; k++) v[u][k] = __ldg((const uint4 *)p + k);
            }
        }
        #pragma unroll
        for (int u = 0; u < U; u++)
            #pragma unroll
            for (int k = 0; k < (BYTES + 15) / 16; k++) acc += v[u][k].x ^ v[u][k].y ^ v[u][k].z ^ v[u][k].w;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ULL);
}

template <int BYTES, int U>
static void run(const uint8_t *tbl, uint64_t tableBytes, uint64_t nAccess, unsigned long long *sink, int blocksPerSM)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * blocksPerSM;
    gather<BYTES, U><<<grid, 256>>>(tbl, tableBytes / BYTES, nAccess / 8, 1, sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        gather<BYTES, U><<<grid, 256>>>(tbl, tableBytes / BYTES, nAccess, 77 + r, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("{\"bytes\": %d, \"in_flight_per_thread\": %d, \"ctas_per_sm\": %d, \"table_gib\": %.1f, \"ms\": %.3f, \"g_access_per_s\": %.3f, \"useful_gbs\": %.1f, \"sector_gbs\": %.1f}\n",
           BYTES, U, blocksPerSM, tableBytes / 1073741824.0, best, nAccess / (best * 1e6), nAccess * (double)BYTES / (best * 1e6), nAccess * (double)(BYTES < 32 ? 32 : BYTES) / (best * 1e6));
    fflush(stdout);
}

int main(int argc, char **argv)
{
    if (argc > 3) {          // L2 fetch granularity hint (bytes): 32, 64 or 128
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[3]));
        size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("{\"l2_fetch_granularity_requested\": %d, \"set_rc\": %d, \"now\": %zu}\n", atoi(argv[3]), (int)e, got);
    }
    double gib = argc > 1 ? atof(argv[1]) : 32.0;
    uint64_t tableBytes = (uint64_t)(gib * 1073741824.0) / 4096 * 4096;
    uint64_t nAccess = argc > 2 ? strtoull(argv[2], 0, 10) : (1ULL << 26);
    uint8_t *tbl; unsigned long long *sink;
    if (cudaMalloc(&tbl, tableBytes) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    cudaMemset(tbl, 1, tableBytes);
    cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
    for (int bps = 4; bps <= 8; bps += 4) {
        run<8, 1>(tbl, tableBytes, nAccess, sink, bps);  run<8, 4>(tbl, tableBytes, nAccess, sink, bps);
        run<16, 1>(tbl, tableBytes, nAccess, sink, bps); run<16, 4>(tbl, tableBytes, nAccess, sink, bps);
        run<32, 1>(tbl, tableBytes, nAccess, sink, bps); run<32, 2>(tbl, tableBytes, nAccess, sink, bps); run<32, 4>(tbl, tableBytes, nAccess, sink, bps);
        run<64, 1>(tbl, tableBytes, nAccess, sink, bps); run<64, 2>(tbl, tableBytes, nAccess, sink, bps);
        run<128, 1>(tbl, tableBytes, nAccess, sink, bps); run<128, 2>(tbl, tableBytes, nAccess, sink, bps);
    }
    return 0;
}
