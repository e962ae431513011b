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

// Sector-bucket insertion (sg_bucket.h) with atomics: claims the first empty slot of the home bucket, or of the next bucket with room,
// setting the "continue" flag (bit 63 of slot 0) of every full bucket it passes.  Slot 0 can change under us in two ways -- another
// key claims it, or a passer-by sets its flag -- hence the compare-and-swap retry.
__device__ __forceinline__ bool sg_bucket_insert_atomic(unsigned long long *buckets, unsigned long long nBuckets, uint32_t bits, unsigned long long canonical,
                                                        uint32_t orient, uint32_t value)
{
    const unsigned long long h = sg_bucket_mix(canonical, bits);
    unsigned long long b = sg_bucket_home(h, bits, nBuckets);
    const unsigned long long slot = sg_bucket_slot_make(h, orient, value);
    for (int disp = 0; disp <= SG_BUCKET_MAX_DISP; disp++) {
        unsigned long long *p = buckets + b * SG_BUCKET_SLOTS;
        #pragma unroll 1
        for (int k = 0; k < SG_BUCKET_SLOTS; k++) {
            unsigned long long old = ((volatile unsigned long long *)p)[k];
            while ((uint32_t)old == 0xffffffffu) {
                const unsigned long long desired = slot | (k == 0 ? (old & SG_BUCKET_FLAG) : 0ULL);
                const unsigned long long prev = atomicCAS(&p[k], old, desired);
                if (prev == old) return true;
                old = prev;
            }
        }
        atomicOr(&p[0], SG_BUCKET_FLAG);
        b = (b + 1 == nBuckets) ? 0 : b + 1;
    }
    return false;
}

// Re-lays the reference's tables (as uploaded from an index directory: 8-byte default or 12-byte -large entries) into sector buckets.
__global__ void sg_build_relayout_kernel(const uint8_t *tables, const uint64_t *tableStart, const uint64_t *tableSize, uint32_t nTables, uint32_t entryBytes,
                                         uint32_t large, uint32_t keyBits, uint32_t seedLen, uint32_t invalidValue, unsigned long long *buckets,
                                         unsigned long long nBuckets, int *failed)
{
    const uint32_t nv = large ? 2u : 1u, bits = 2 * seedLen;
    for (uint32_t t = blockIdx.y; t < nTables; t += gridDim.y) {
        const uint64_t size = tableSize[t], base = tableStart[t];
        for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < size; k += (uint64_t)gridDim.x * blockDim.x) {
            const uint8_t *p = tables + (base + k) * entryBytes;
            const uint32_t v0 = *(const uint32_t *)p, v1 = nv == 2 ? *(const uint32_t *)(p + 4) : 0xfffffffeu, key = *(const uint32_t *)(p + 4 * nv);
            if (v0 == invalidValue) continue;
            const unsigned long long seed = ((unsigned long long)t << keyBits) | key;
            bool ok = true;
            if (large) {
                if (v0 != 0xfffffffeu) ok = sg_bucket_insert_atomic(buckets, nBuckets, bits, seed, 0, v0);
                if (ok && v1 != 0xfffffffeu) ok = sg_bucket_insert_atomic(buckets, nBuckets, bits, seed, 1, v1);
            } else {
                const unsigned long long rc = sg_seed_revcomp(seed, seedLen);
                const unsigned long long c = seed < rc ? seed : rc;
                ok = sg_bucket_insert_atomic(buckets, nBuckets, bits, c, seed == c ? 0u : 1u, v0);
            }
            if (!ok) *failed = 1;
        }
    }
}

// counts the used values of the reference's tables (= bucket slots needed)
__global__ void sg_build_count_values_kernel(const uint8_t *tables, unsigned long long totalSlots, uint32_t entryBytes, uint32_t large, uint32_t invalidValue,
                                             unsigned long long *count)
{
    unsigned long long local = 0;
    const uint32_t nv = large ? 2u : 1u;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < totalSlots; k += (unsigned long long)gridDim.x * blockDim.x) {
        const uint8_t *p = tables + k * entryBytes;
        for (uint32_t j = 0; j < nv; j++) {
            const uint32_t v = *(const uint32_t *)(p + 4 * j);
            if (v != invalidValue && v != 0xfffffffeu) local++;
        }
    }
    local = __reduce_add_sync(0xffffffffu, (unsigned)local) ;
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// pass 2: insert run heads.  buckets != NULL: into sector buckets (sg_bucket.h); else into the reference's table layout.
__global__ void sg_build_insert_kernel(const unsigned long long *keys, const uint32_t *locs, long long n, uint32_t keyBits, uint32_t seedLen,
                                       const uint64_t *tableStart, const uint64_t *tableSize, unsigned long long *tables,
                                       uint32_t *overflow, unsigned long long *overflowCursor, long long nBases, int *failed,
                                       unsigned long long *buckets, unsigned long long nBuckets)
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
        if (buckets) {
            const unsigned long long rc = sg_seed_revcomp(k, seedLen);
            const unsigned long long c = k < rc ? k : rc;
            if (!sg_bucket_insert_atomic(buckets, nBuckets, 2 * seedLen, c, k == c ? 0u : 1u, value)) *failed = 1;
            continue;
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
