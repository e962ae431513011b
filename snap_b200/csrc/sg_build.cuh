// sg_build.cuh -- building the lookup structure on the device (replaces, in *results*,
// GenomeIndex::BuildIndexToDirectory, reference SNAPLib/GenomeIndex.cpp:527-1110 and its worker :1447-1530).
//
// Contract kept from the reference: every position whose seedLen bases are all ACGT is indexed under its forward
// seed (small-table layout: key = low keyBytes of the seed, table = the remaining high bases); a seed that occurs
// once stores its location in the entry, a repeated seed stores countOfBases + offset into the overflow table,
// where the list is `count` followed by the locations in DESCENDING order (:879-889).  Slot placement inside a
// table is not part of the contract (the reference's own depends on thread interleaving, :1523-1530).
//
// Method: emit (seed, location) for locations in descending order, stable LSD radix sort on the seed (so equal
// seeds keep descending locations), then one pass sizes the tables and one pass inserts run heads with atomicCAS
// along the very probe sequence lookups use.
#pragma once
#include <cub/device/device_radix_sort.cuh>


__global__ void sg_build_emit_kernel(const uint8_t *bases /* location 0 */, long long nPositions, uint32_t seedLen,
                                     unsigned long long *keys, uint32_t *locs)
{
    const unsigned long long SG_BUILD_INVALID_KEY = 1ULL << (2 * seedLen);
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; j < nPositions; j += stride) {
        long long loc = nPositions - 1 - j;      // descending locations
        unsigned long long b = 0;
        bool ok = true;
        for (uint32_t i = 0; i < seedLen; i++) {
            uint32_t v = sg_base_value(bases[loc + i]);
            ok = ok && v < 4;
            b = (b << 2) | (v & 3);
        }
        keys[j] = ok ? b : (SG_BUILD_INVALID_KEY | b);
        locs[j] = (uint32_t)loc;
    }
}

// pass 1: per run head, count unique seeds per table and overflow words.  The keys are sorted, so a thread sees the same table
// for long stretches: counts are kept in registers, flushed to a per-block shared histogram when the table changes, and each
// block adds its totals to the global counters once (one atomic per seed on three hot words cost 2 s at 3 Gbp).
__global__ void sg_build_count_kernel(const unsigned long long *keys, long long n, uint32_t keyBits, uint32_t seedLen, uint32_t nTables, int hist,
                                      unsigned long long *tableUsed, unsigned long long *overflowWords, unsigned long long *nValid)
{
    // shared per-block histogram when the tables fit (seed length <= 22); for more tables the per-thread run counts go straight to
    // the global histogram (still one atomic per run of equal tables, not per seed)
    extern __shared__ unsigned long long sUsed[];       // [nTables or 0] + nValid + overflowWords
    const bool sharedHist = hist == 1;
    const uint32_t nShared = sharedHist ? nTables : 0;
    for (uint32_t t = threadIdx.x; t < nShared + 2; t += blockDim.x) sUsed[t] = 0;
    __syncthreads();
    const unsigned long long SG_BUILD_INVALID_KEY = 1ULL << (2 * seedLen);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    uint32_t curTable = 0xffffffffu;
    unsigned long long curUsed = 0, valid = 0, over = 0;
    for (; i < n; i += stride) {
        unsigned long long k = keys[i];
        if (k & SG_BUILD_INVALID_KEY) continue;
        if (i > 0 && keys[i - 1] == k) continue;
        long long e = i + 1;
        while (e < n && keys[e] == k) e++;
        unsigned long long count = (unsigned long long)(e - i);
        const uint32_t table = (uint32_t)(k >> keyBits);
        if (table != curTable) {
            if (curUsed) atomicAdd(sharedHist ? &sUsed[curTable] : &tableUsed[curTable], curUsed);
            curTable = table; curUsed = 0;
        }
        curUsed++;
        valid += count;
        if (count > 1) over += count + 1;
    }
    if (curUsed) atomicAdd(sharedHist ? &sUsed[curTable] : &tableUsed[curTable], curUsed);
    if (valid) atomicAdd(&sUsed[nShared], valid);
    if (over) atomicAdd(&sUsed[nShared + 1], over);
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < nShared; t += blockDim.x) if (sUsed[t]) atomicAdd(&tableUsed[t], sUsed[t]);
    if (threadIdx.x == 0) {
        if (sUsed[nShared]) atomicAdd(nValid, sUsed[nShared]);
        if (sUsed[nShared + 1]) atomicAdd(overflowWords, sUsed[nShared + 1]);
    }
}

// pass 2: insert run heads
__global__ void sg_build_insert_kernel(const unsigned long long *keys, const uint32_t *locs, long long n, uint32_t keyBits, uint32_t seedLen,
                                       const uint64_t *tableStart, const uint64_t *tableSize, unsigned long long *tables,
                                       uint32_t *overflow, unsigned long long *overflowCursor, long long nBases, int *failed)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const unsigned long long EMPTY = 0x00000000ffffffffULL;   // key 0, value invalid (HashTable.cpp:66-72)
    const unsigned long long SG_BUILD_INVALID_KEY = 1ULL << (2 * seedLen);
    for (; i < n; i += stride) {
        unsigned long long k = keys[i];
        if (k & SG_BUILD_INVALID_KEY) continue;
        if (i > 0 && keys[i - 1] == k) continue;
        long long e = i + 1;
        while (e < n && keys[e] == k) e++;
        unsigned long long count = (unsigned long long)(e - i);
        uint32_t value;
        if (count == 1) {
            value = locs[i];
        } else {
            unsigned long long off = atomicAdd(overflowCursor, count + 1);
            overflow[off] = (uint32_t)count;
            for (unsigned long long c = 0; c < count; c++) overflow[off + 1 + c] = locs[i + c];   // already descending
            value = (uint32_t)((unsigned long long)nBases + off);
        }
        const uint32_t table = (uint32_t)(k >> keyBits);
        const unsigned long long low = k & ((1ULL << keyBits) - 1);
        const uint64_t size = tableSize[table], base = tableStart[table];
        uint64_t idx = sg_fmix64(low) % size;
        const unsigned long long entry = (low << 32) | value;
        unsigned long long nProbes = 0;
        bool placed = false;
        for (;;) {
            unsigned long long old = atomicCAS(&tables[base + idx], EMPTY, entry);
            if (old == EMPTY) { placed = true; break; }
            nProbes++;
            if (nProbes > size + 5) break;
            if (nProbes < 5) idx = (idx + nProbes * nProbes) % size; else idx = (idx + 1) % size;
        }
        if (!placed) *failed = 1;
    }
}

__global__ void sg_build_fill_kernel(unsigned long long *p, long long n, unsigned long long v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
