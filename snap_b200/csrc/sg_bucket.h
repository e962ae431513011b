// sg_bucket.h -- the B200-native index layout: one 32-byte DRAM sector resolves a lookup.  Host+device.
//
// The reference's tables (HashTable.h:87-198) store one (value, key) entry per slot and chain by quadratic-then-linear probing:
// a `lookupSeed32` on the default index walks two chains (seed and reverse complement, GenomeIndex.cpp:2138-2155), 10.6 entries
// on average at 3 Gbp, each in its own sector -- >= 10 random DRAM sectors per lookup.  Random sectors are what HBM is slowest at
// (measured: 35 G sector accesses/s on a B200, whatever the access width from 8 to 32 bytes; profiles/r02_random_gather_peak.jsonl),
// so the layout here is organised around the sector, not the entry:
//
//   * keyed by the CANONICAL seed c = min(seed, reverse complement), like the reference's own `-large` layout (GenomeIndex.cpp:2131,
//     Seed.h:99): both strands of a lookup come from one place;
//   * bucket = 4 slots x 8 bytes = 32 bytes = one sector, 32-byte aligned.  A slot is
//         [63] "continue" flag (slot 0 of a bucket only)   [62] orientation (0: locations of c, 1: locations of rc(c))
//         [61:32] tag = low 30 bits of h                   [31:0] value
//     with h = a bijective mix of c on 2*seedLen bits and the home bucket = floor(h * nBuckets / 2^(2*seedLen)).  The value is the
//     reference's: a location (< countOfBases: singleton, fillInLookedUpResults32 GenomeIndex.cpp:2159-2202) or countOfBases +
//     offset into the overflow table (count, then the locations descending, :879-889).  An empty slot has value 0xffffffff;
//   * a key that finds its home bucket full goes to the next bucket with room and sets the "continue" flag of every bucket it
//     passed; a lookup scans its home bucket and goes on only while the bucket just scanned carries the flag (and it has not yet
//     seen both orientations).  At load 0.6 that is ~1.1 sectors per lookup, present or absent;
//   * the high 2*seedLen - 30 bits of h are not stored: they follow from the bucket, because a key is never more than
//     SG_BUCKET_MAX_DISP buckets from home and nBuckets is kept >= 2^(2*seedLen - 30) * 2 * (SG_BUCKET_MAX_DISP + 1), so two keys with
//     equal tags whose homes are that close cannot exist (their h differ by a multiple of 2^30).
//
// What is kept from the reference is the RESULT of a lookup: the same hit set, in the same (descending) order, with `hits` pointing
// at 32-bit locations in HBM ("parity is defined on results, so the GPU may re-layout the table").
#pragma once
#include "sg_common.h"

#define SG_BUCKET_SLOTS 4
#define SG_BUCKET_MAX_DISP 127
#define SG_BUCKET_EMPTY 0x3fffffffffffffffULL          // tag all ones, orientation 0, no flag, value invalid
#define SG_BUCKET_FLAG (1ULL << 63)
#define SG_BUCKET_ORIENT (1ULL << 62)

// Bijective mix of a `bits`-wide value (32 <= bits <= 48): odd multiplications and right xor-shifts are permutations of [0, 2^bits).
SG_HD uint64_t sg_bucket_mix(uint64_t c, uint32_t bits)
{
    const uint64_t mask = (bits >= 64) ? ~0ULL : ((1ULL << bits) - 1);
    const uint32_t s = bits / 2;
    uint64_t x = c & mask;
    x = (x * 0xff51afd7ed558ccdULL) & mask;
    x ^= x >> s;
    x = (x * 0xc4ceb9fe1a85ec53ULL) & mask;
    x ^= x >> s;
    x = (x * 0x9e3779b97f4a7c15ULL) & mask;
    x ^= x >> s;
    return x;
}

SG_HD uint64_t sg_mulhi64(uint64_t a, uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// floor(h * nBuckets / 2^bits)
SG_HD uint64_t sg_bucket_home(uint64_t h, uint32_t bits, uint64_t nBuckets)
{
    return sg_mulhi64(h << (64 - bits), nBuckets);
}

// Smallest bucket count for which tags identify keys (see the header comment).
SG_HD uint64_t sg_bucket_min_count(uint32_t bits)
{
    const uint32_t hidden = bits > 30 ? bits - 30 : 0;
    return (1ULL << hidden) * 2ULL * (SG_BUCKET_MAX_DISP + 1);
}

SG_HD uint64_t sg_bucket_slot_make(uint64_t h, uint32_t orient, uint32_t value)
{
    return ((uint64_t)orient << 62) | ((h & 0x3fffffffULL) << 32) | value;
}

// The four slots of bucket b, one sector.
struct SgBucket { uint64_t s[SG_BUCKET_SLOTS]; };

SG_HD SgBucket sg_bucket_load(const uint64_t *buckets, uint64_t b)
{
    SgBucket r;
#if defined(__CUDA_ARCH__)
    const ulonglong2 *p = (const ulonglong2 *)(buckets + b * SG_BUCKET_SLOTS);
    const ulonglong2 lo = __ldg(p), hi = __ldg(p + 1);
    r.s[0] = lo.x; r.s[1] = lo.y; r.s[2] = hi.x; r.s[3] = hi.y;
#else
    for (int k = 0; k < SG_BUCKET_SLOTS; k++) r.s[k] = buckets[b * SG_BUCKET_SLOTS + k];
#endif
    return r;
}

// Host-side (sequential) insertion; the device builder does the same with atomics (sg_build.cuh).  Returns false if the key would
// land more than SG_BUCKET_MAX_DISP buckets from home (table too full: rebuild with more buckets).
SG_HD bool sg_bucket_insert_seq(uint64_t *buckets, uint64_t nBuckets, uint32_t bits, uint64_t canonical, uint32_t orient, uint32_t value)
{
    const uint64_t h = sg_bucket_mix(canonical, bits);
    uint64_t b = sg_bucket_home(h, bits, nBuckets);
    const uint64_t slot = sg_bucket_slot_make(h, orient, value);
    for (int disp = 0; disp <= SG_BUCKET_MAX_DISP; disp++) {
        uint64_t *p = buckets + b * SG_BUCKET_SLOTS;
        for (int k = 0; k < SG_BUCKET_SLOTS; k++) {
            if ((uint32_t)p[k] == 0xffffffffu) {
                p[k] = slot | (k == 0 ? (p[0] & SG_BUCKET_FLAG) : 0ULL);
                return true;
            }
        }
        p[0] |= SG_BUCKET_FLAG;
        b = (b + 1 == nBuckets) ? 0 : b + 1;
    }
    return false;
}

// GenomeIndex::lookupSeed32 on the bucket layout: the semantics of sg_lookup_seed32 (sg_seed.h), i.e. of the reference's
// lookupSeed32 + fillInLookedUpResults32 (GenomeIndex.cpp:2095-2202).  `bases` / `rc` from sg_seed_pack.  *examined counts slots
// examined (4 per bucket), *overflowWords overflow-table count words read.
struct SgHits;
SG_HD void sg_fill_hits(const SgIndexView &ix, const uint32_t *subEntry, uint32_t *nHits, const uint32_t **hits, uint32_t *overflowWords);

SG_HD void sg_bucket_lookup_seed32(const SgIndexView &ix, uint64_t bases, uint64_t rc, uint32_t *nHits /*[2]*/, const uint32_t **hits /*[2]*/,
                                   uint32_t *examined, uint32_t *overflowWords)
{
    nHits[0] = nHits[1] = 0;
    hits[0] = hits[1] = ix.overflow;
    const uint32_t bits = 2 * ix.seedLen;
    const bool lookedUpComplement = bases > rc;               // Seed::isBiggerThanItsReverseComplement (Seed.h:99)
    const uint64_t c = lookedUpComplement ? rc : bases;
    const uint64_t h = sg_bucket_mix(c, bits);
    uint64_t b = sg_bucket_home(h, bits, ix.nBuckets);
    const uint64_t want = (h & 0x3fffffffULL) << 32;
    const uint32_t *found[2] = {(const uint32_t *)0, (const uint32_t *)0};
    #pragma unroll 1
    for (int disp = 0; disp <= SG_BUCKET_MAX_DISP; disp++) {
        const SgBucket B = sg_bucket_load(ix.buckets, b);
        *examined += SG_BUCKET_SLOTS;
        #pragma unroll
        for (int k = 0; k < SG_BUCKET_SLOTS; k++) {
            const uint64_t s = B.s[k];
            if ((s & 0x3fffffff00000000ULL) == want && (uint32_t)s != 0xffffffffu) {
                found[(s >> 62) & 1] = (const uint32_t *)(ix.buckets + b * SG_BUCKET_SLOTS + k);      // little-endian: the value is the slot's first word
            }
        }
        if (!(B.s[0] & SG_BUCKET_FLAG)) break;
        if (found[0] && (found[1] || bases == rc)) break;
        b = (b + 1 == ix.nBuckets) ? 0 : b + 1;
    }
    // found[o]: orientation o of the canonical seed.  Direction 0 of the lookup is the seed as given, direction 1 its reverse complement.
    const uint32_t *e0 = lookedUpComplement ? found[1] : found[0];
    const uint32_t *e1 = lookedUpComplement ? found[0] : found[1];
    if (bases == rc) e1 = e0;                                  // isOwnReverseComplement: both directions share the entry (GenomeIndex.cpp:2131)
    if (e0) sg_fill_hits(ix, e0, &nHits[0], &hits[0], overflowWords);
    if (e1) {
        if (e1 == e0) { nHits[1] = nHits[0]; hits[1] = hits[0]; }
        else sg_fill_hits(ix, e1, &nHits[1], &hits[1], overflowWords);
    }
}
