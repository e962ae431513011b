// sg_warp.cuh -- warp-cooperative leaves (device only).
//
// Execution model of the alignment kernel: one warp per read.  The BaseAligner state machine is inherently
// sequential, so all 32 lanes execute it *uniformly* (identical registers, identical loads, same-value stores:
// a warp instruction costs one issue slot whether 1 or 32 lanes are live, so the redundancy is free) and therefore
// arrive converged at every leaf below, where the lanes split the data-parallel work and exchange it with shuffles.
#pragma once
#include "sg_common.h"
#include "sg_seed.h"

#define SG_FULL 0xffffffffu

// Seed::Seed across lanes: lane i encodes base i (seedLen <= 32).  Returns false if any base is not ACGT.
__device__ __forceinline__ bool sg_warp_seed_pack(const uint8_t *text, uint32_t seedLen, int lane, uint64_t *bases, uint64_t *rc)
{
    uint32_t v = 0;
    bool ok = true;
    if ((uint32_t)lane < seedLen) {
        v = sg_base_value(text[lane]);
        ok = v < 4;
    }
    ok = __all_sync(SG_FULL, ok);
    uint64_t b = 0, r = 0;
    if ((uint32_t)lane < seedLen) {
        b = (uint64_t)(v & 3) << ((seedLen - lane - 1) * 2);
        r = (uint64_t)((v & 3) ^ 3) << (lane * 2);
    }
    uint32_t blo = __reduce_or_sync(SG_FULL, (uint32_t)b), bhi = __reduce_or_sync(SG_FULL, (uint32_t)(b >> 32));
    uint32_t rlo = __reduce_or_sync(SG_FULL, (uint32_t)r), rhi = __reduce_or_sync(SG_FULL, (uint32_t)(r >> 32));
    *bases = ((uint64_t)bhi << 32) | blo;
    *rc = ((uint64_t)rhi << 32) | rlo;
    return ok;
}

// Offset of probe number k from the home slot (SNAPHashTable::GetFirstValueForKey, HashTable.h:87-118):
// +1, +4, +9, +16 for nProbes 1..4 (QUADRATIC_CHAINING_DEPTH 5), then +1 each.
__device__ __forceinline__ uint32_t sg_probe_offset(uint32_t k)
{
    return k == 0 ? 0u : k == 1 ? 1u : k == 2 ? 5u : k == 3 ? 14u : (30u + (k - 4u));
}

// Both strand chains of a default (small) index are probed at once: lanes 0-15 take probes 0..15 of the forward
// seed's chain, lanes 16-31 the same for the reverse complement; one round of independent 8-byte loads usually
// resolves both (mean chain 5.5 entries).  A `-large` index has one chain (canonical seed): all 32 lanes probe it.
// Result semantics are exactly sg_lookup_seed32's.
__device__ __forceinline__ void sg_warp_lookup_seed32(const SgIndexView &ix, uint64_t bases, uint64_t rc, int lane, SgHits *out,
                                                      uint32_t *examined, uint32_t *overflowWords)
{
    if (ix.layout == SG_LAYOUT_BUCKET) {
        // one sector resolves the lookup: every lane loads the same 32 bytes (one request) and resolves it redundantly -- there is
        // nothing to spread over the lanes
        sg_bucket_lookup_seed32(ix, bases, rc, out->nHits, out->hits, examined, overflowWords);
        return;
    }
    const uint32_t keyBits = ix.keyBytes * 8;
    out->nHits[0] = out->nHits[1] = 0;
    out->hits[0] = out->hits[1] = ix.overflow;
    const bool large = ix.large != 0;
    const bool lookedUpComplement = large && (bases > rc);
    // chain c (0/1) = strand for the small index; for -large only chain 0 exists
    const uint32_t chainWidth = large ? 32u : 16u;
    const uint32_t myChain = large ? 0u : (uint32_t)(lane >> 4);
    const uint32_t myK0 = large ? (uint32_t)lane : (uint32_t)(lane & 15);
    uint64_t s = large ? (lookedUpComplement ? rc : bases) : (myChain ? rc : bases);
    const uint64_t low = (ix.keyBytes == 8) ? s : (s & ((1ULL << keyBits) - 1));
    const uint32_t table = (ix.keyBytes == 8) ? 0u : (uint32_t)(s >> keyBits);
    const uint64_t size = ix.tableSize[table];
    const uint64_t base = ix.tableStart[table];
    // hash % size without a 64-bit division: multiply-high by floor((2^64-1)/size), then at most two corrections
    uint64_t home;
    {
        const uint64_t hsh = sg_fmix64(low);
        const uint64_t qq = __umul64hi(hsh, ix.tableMagic[table]);
        home = hsh - qq * size;
        if (home >= size) home -= size;
        if (home >= size) home -= size;
    }

    uint64_t foundSlot[2] = {~0ULL, ~0ULL};
    bool done[2] = {false, large};
    uint32_t round = 0;
    while (!(done[0] && done[1])) {
        uint32_t k = round * chainWidth + myK0;
        uint64_t pos = home + sg_probe_offset(k);       // probe offsets are small: wrap with subtractions, falling back to % for tiny tables
        if (pos >= size) { pos -= size; if (pos >= size) pos %= size; }
        uint64_t slot = base + pos;
        uint32_t v; uint64_t key;
        bool active = !done[myChain] && (uint64_t)k <= size + 5;
        if (active) sg_entry_load(ix, slot, &v, &key); else { v = 0; key = ~0ULL; }
        bool match = active && (key == low);
        bool inval = active && (v == ix.invalidValue);
        bool stop = (k == 0) ? (match && !inval) : (match || inval);
        if (active && (uint64_t)k == size + 5) stop = true;          // nProbes > tableSize + 5: give up
        uint32_t stopMask = __ballot_sync(SG_FULL, stop);
        for (uint32_t c = 0; c < 2; c++) {
            if (done[c]) continue;
            uint32_t m = large ? stopMask : ((stopMask >> (16 * c)) & 0xffffu);
            if (m != 0) {
                uint32_t first = __ffs(m) - 1;
                uint32_t srcLane = large ? first : (16 * c + first);
                uint32_t kk = round * chainWidth + first;
                bool hit = __shfl_sync(SG_FULL, (int)(match && !inval), srcLane) != 0;
                uint32_t slo = __shfl_sync(SG_FULL, (uint32_t)slot, srcLane), shi = __shfl_sync(SG_FULL, (uint32_t)(slot >> 32), srcLane);
                foundSlot[c] = hit ? (((uint64_t)shi << 32) | slo) : ~0ULL;
                *examined += kk + 1;
                done[c] = true;
            } else if ((uint64_t)(round + 1) * chainWidth > size + 5) {
                done[c] = true;
            }
        }
        round++;
    }

    if (large) {
        if (foundSlot[0] == ~0ULL) return;
        const uint32_t *entry = (const uint32_t *)(ix.tables + foundSlot[0] * ix.entryBytes);
        sg_fill_hits(ix, lookedUpComplement ? entry + 1 : entry, &out->nHits[0], &out->hits[0], overflowWords);
        if (bases == rc) {
            out->nHits[1] = out->nHits[0];
            out->hits[1] = out->hits[0];
        } else {
            sg_fill_hits(ix, lookedUpComplement ? entry : entry + 1, &out->nHits[1], &out->hits[1], overflowWords);
        }
    } else {
        for (int c = 0; c < 2; c++) {
            if (foundSlot[c] != ~0ULL) {
                const uint32_t *entry = (const uint32_t *)(ix.tables + foundSlot[c] * ix.entryBytes);
                sg_fill_hits(ix, entry, &out->nHits[c], &out->hits[c], overflowWords);
            }
        }
    }
}

// ---- split form of the first probe round, for the batched seed-lookup kernel: issue the 32 entry loads of several seeds
//      back to back (memory-level parallelism), then resolve them one after the other ----
struct SgProbeRound {
    uint64_t slot, key;
    uint32_t v;
    bool active;
};

__device__ __forceinline__ void sg_warp_probe_issue(const SgIndexView &ix, uint64_t bases, uint64_t rc, int lane, SgProbeRound &r)
{
    const uint32_t keyBits = ix.keyBytes * 8;
    const bool large = ix.large != 0;
    const bool lookedUpComplement = large && (bases > rc);
    const uint32_t myChain = large ? 0u : (uint32_t)(lane >> 4);
    const uint32_t k = large ? (uint32_t)lane : (uint32_t)(lane & 15);
    const uint64_t s = large ? (lookedUpComplement ? rc : bases) : (myChain ? rc : bases);
    const uint64_t low = (ix.keyBytes == 8) ? s : (s & ((1ULL << keyBits) - 1));
    const uint32_t table = (ix.keyBytes == 8) ? 0u : (uint32_t)(s >> keyBits);
    const uint64_t size = ix.tableSize[table];
    const uint64_t base = ix.tableStart[table];
    const uint64_t hsh = sg_fmix64(low);
    const uint64_t qq = __umul64hi(hsh, ix.tableMagic[table]);
    uint64_t home = hsh - qq * size;
    if (home >= size) home -= size;
    if (home >= size) home -= size;
    uint64_t pos = home + sg_probe_offset(k);
    if (pos >= size) { pos -= size; if (pos >= size) pos %= size; }
    r.slot = base + pos;
    r.active = (uint64_t)k <= size + 5;
    if (r.active) sg_entry_load(ix, r.slot, &r.v, &r.key); else { r.v = 0; r.key = ~0ULL; }
}

// Returns false when a chain did not stop within the first round (caller falls back to sg_warp_lookup_seed32).
__device__ __forceinline__ bool sg_warp_probe_finish(const SgIndexView &ix, const SgProbeRound &r, uint64_t bases, uint64_t rc, int lane, SgHits *out,
                                                     uint32_t *examined, uint32_t *overflowWords)
{
    const bool large = ix.large != 0;
    const bool lookedUpComplement = large && (bases > rc);
    const uint32_t k = large ? (uint32_t)lane : (uint32_t)(lane & 15);
    const uint32_t keyBits = ix.keyBytes * 8;
    const uint64_t s = large ? (lookedUpComplement ? rc : bases) : ((lane >> 4) ? rc : bases);
    const uint64_t low = (ix.keyBytes == 8) ? s : (s & ((1ULL << keyBits) - 1));
    const uint64_t size = ix.tableSize[(ix.keyBytes == 8) ? 0u : (uint32_t)(s >> keyBits)];
    out->nHits[0] = out->nHits[1] = 0;
    out->hits[0] = out->hits[1] = ix.overflow;
    const bool match = r.active && (r.key == low);
    const bool inval = r.active && (r.v == ix.invalidValue);
    bool stop = (k == 0) ? (match && !inval) : (match || inval);
    if (r.active && (uint64_t)k == size + 5) stop = true;
    const uint32_t stopMask = __ballot_sync(SG_FULL, stop);
    uint64_t foundSlot[2] = {~0ULL, ~0ULL};
    const int nChains = large ? 1 : 2;
    for (int c = 0; c < nChains; c++) {
        const uint32_t m = large ? stopMask : ((stopMask >> (16 * c)) & 0xffffu);
        if (m == 0) return false;
        const uint32_t first = __ffs(m) - 1;
        const uint32_t srcLane = large ? first : (16 * c + first);
        const bool hit = __shfl_sync(SG_FULL, (int)(match && !inval), srcLane) != 0;
        const uint32_t slo = __shfl_sync(SG_FULL, (uint32_t)r.slot, srcLane), shi = __shfl_sync(SG_FULL, (uint32_t)(r.slot >> 32), srcLane);
        foundSlot[c] = hit ? (((uint64_t)shi << 32) | slo) : ~0ULL;
        *examined += first + 1;
    }
    if (large) {
        if (foundSlot[0] == ~0ULL) return true;
        const uint32_t *entry = (const uint32_t *)(ix.tables + foundSlot[0] * ix.entryBytes);
        sg_fill_hits(ix, lookedUpComplement ? entry + 1 : entry, &out->nHits[0], &out->hits[0], overflowWords);
        if (bases == rc) { out->nHits[1] = out->nHits[0]; out->hits[1] = out->hits[0]; }
        else sg_fill_hits(ix, lookedUpComplement ? entry : entry + 1, &out->nHits[1], &out->hits[1], overflowWords);
    } else {
        for (int c = 0; c < 2; c++) {
            if (foundSlot[c] != ~0ULL) {
                const uint32_t *entry = (const uint32_t *)(ix.tables + foundSlot[c] * ix.entryBytes);
                sg_fill_hits(ix, entry, &out->nHits[c], &out->hits[c], overflowWords);
            }
        }
    }
    return true;
}


// ---- hit-list staging: one bulk asynchronous copy (TMA: cp.async.bulk, completion on an mbarrier) of up to SgScratch::hitStageWords
//      words of an overflow list into the warp's shared-memory buffer.  `src` need only be 4-byte aligned: the copy starts at the
//      16-byte boundary below it and *lead (0..3) says how many words precede the first wanted one.  All 32 lanes call this converged;
//      lane 0 issues, every lane waits on the barrier (which also makes the data visible to it). ----
__device__ __forceinline__ void sg_warp_hits_barrier_init(SgScratch &sc, int lane)
{
    if (sc.hitBar == (unsigned long long *)0) return;
    if (lane == 0) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(sc.hitBar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
}

__device__ __forceinline__ uint32_t sg_warp_stage_hits(SgScratch &sc, const uint32_t *src, uint32_t nWords, int lane, uint32_t *lead)
{
    const uintptr_t a = (uintptr_t)src;
    const uint32_t ld = (uint32_t)((a & 15u) >> 2);
    uint32_t words = nWords + ld;
    if (words > sc.hitStageWords) words = sc.hitStageWords;
    const uint32_t bytes = ((words * 4u) + 15u) & ~15u;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(sc.hitBar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sc.hitStage);
    const uint32_t phase = sc.hitPhase;
    __syncwarp();                    // nobody is still reading the previous chunk (or the phase)
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the buffer was last touched through the generic proxy
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"((unsigned long long)(a & ~(uintptr_t)15)), "r"(bytes), "r"(bar) : "memory");
    }
    const uint32_t parity = phase & 1u;
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
    sc.hitPhase = phase ^ 1u;        // (shared by the warp's lanes: every lane stores the same value, all of them read it before the first __syncwarp)
    __syncwarp();
    *lead = ld;
    return words - ld;               // wanted words now at hitStage[ld ...]
}
