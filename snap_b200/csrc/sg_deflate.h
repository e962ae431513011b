// sg_deflate.h -- a deflate compressor for BGZF members, one thread block per member.
//
// What the reference does here: GzipCompressWorker (reference SNAPLib/GzipDataWriter.cpp:153-276) hands each 64 KB chunk of the output to zlib's
// deflate() and wraps it as a BGZF member.  zlib is a third-party dependency of the reference (not in /root/reference); what is pinned is the
// FORMAT (RFC 1951 / RFC 1952 / SAM spec 4.1): any inflater must give back exactly the payload, with the right CRC-32 and ISIZE.  The compressed
// bytes themselves are not comparable between compressors (nor between zlib versions or levels), so parity is defined on the inflated stream.
//
// The design is for a thread block, not a port of zlib's serial loop:
//   1. match finding: every position hashes its next 4 bytes; positions are taken in strides of SG_DEFLATE_STRIDE, a stride first LOOKS UP the
//      most recent earlier position with its hash (inserted by earlier strides), then all of it INSERTS -- two barriers per stride, no
//      per-position dependency.  A distance-1 candidate (runs of one byte: quality strings) is tried as well.  Match lengths by byte compare in
//      shared memory.
//   2. greedy parse without a serial walk: next[i] = i + (match at i ? its length : 1); the positions reachable from 0 are found by pointer
//      jumping (mark, then square the jump table), 16 rounds for 65280 positions.
//   3. histograms of the literal/length and distance symbols of the marked positions (shared-memory atomics); code lengths by Moffat &
//      Katajainen's in-place minimum-redundancy algorithm on a rank-sorted frequency list, limited to 15 bits (Kraft sum restored by the usual
//      demote-one / promote-two step), canonical codes.
//   4. the bit offset of every token by a block-wide prefix sum of token bit lengths; tokens OR-ed into a shared-memory image of the block
//      (shared-memory atomics), then copied out.
// One dynamic-Huffman block per member (BFINAL = 1); its code-length alphabet uses a flat 4-bit code for lengths 0..15 and no repeat symbols
// (a ~150-byte header per 64 KB member, no third Huffman construction).  A member that does not shrink is written as a stored block.
//
// The same source builds for the host (the test-only build of these headers): the "threads" are then one thread running every loop to its end, barriers are no-ops.
#ifndef SG_DEFLATE_H
#define SG_DEFLATE_H

#include "sg_common.h"

#define SG_DEFLATE_MAX_PAYLOAD 0xff00u      // BGZF: at most 65280 payload bytes per member, so that a stored member still fits 64 KB
#define SG_DEFLATE_STRIDE 1024u
#define SG_DEFLATE_HASH_BITS 13
#define SG_DEFLATE_MIN_MATCH 4u
#define SG_DEFLATE_MAX_MATCH 258u
#define SG_DEFLATE_MAX_DIST 32768u
#define SG_DEFLATE_NLIT 286
#define SG_DEFLATE_NDIST 30
#define SG_DEFLATE_MEMBER_PITCH (SG_DEFLATE_MAX_PAYLOAD + 31u)    // the largest member: stored payload + 18 header + 5 block header + 8 trailer

// (a test build may bring its own threads: it defines SGD_CUSTOM_THREADS and the six macros before including this file)
#if defined(SGD_CUSTOM_THREADS)
#elif defined(__CUDA_ARCH__)
#define SGD_TID ((uint32_t)threadIdx.x)
#define SGD_NT ((uint32_t)blockDim.x)
#define SGD_WARP ((uint32_t)threadIdx.x >> 5)
#define SGD_SYNC() __syncthreads()
#define SGD_ATOMIC_OR(p, v) atomicOr((p), (v))
#define SGD_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define SGD_TID 0u
#define SGD_NT 1u
#define SGD_WARP 0u
#define SGD_SYNC() do { } while (0)
#define SGD_ATOMIC_OR(p, v) (*(p) |= (v))
#define SGD_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

// Working storage of one block.  "shared": small and hot, in shared memory on the device; "arena": per-block scratch in HBM (L2-resident).
struct SgDeflateShared {
    uint8_t  buf[SG_DEFLATE_MAX_PAYLOAD + 8];          // the payload (+ padding so that 4-byte reads at the end stay inside)
    uint32_t out[SG_DEFLATE_MAX_PAYLOAD / 4 + 8];      // image of the deflate block being assembled; the hash table lives here during match finding
    uint32_t sel[SG_DEFLATE_MAX_PAYLOAD / 32 + 2];     // bit i: position i starts a token of the greedy parse
    uint32_t litFreq[SG_DEFLATE_NLIT + 2], distFreq[SG_DEFLATE_NDIST + 2];
    uint16_t litCode[SG_DEFLATE_NLIT + 2], distCode[SG_DEFLATE_NDIST + 2];     // canonical codes, bit-reversed (deflate packs them MSB first)
    uint8_t  litLen[SG_DEFLATE_NLIT + 2], distLen[SG_DEFLATE_NDIST + 2];
    uint16_t sortSym[SG_DEFLATE_NLIT + 2];             // symbols by ascending (frequency, symbol); scratch of the code construction
    uint32_t sortFreq[SG_DEFLATE_NLIT + 2];
    uint32_t partial[1024 + 1];                        // per-thread sums of the prefix sums (block size <= 1024), CRC partials
    uint32_t crcTable[256];
    uint32_t byteFreq[256];
    uint32_t warpHist[32][SG_DEFLATE_NLIT + 2];        // one histogram per warp (bytes, then literal/length symbols): a payload has few distinct values and
                                                       // same-address shared-memory atomics serialise; summed into byteFreq / litFreq afterwards
    uint32_t crcGroup[32], crcGroupBytes[32];
    uint8_t  litCost[256];                             // estimated cost of a literal of each byte value, in 1/8 bit (from the byte histogram)
    uint32_t x2n[32];
    uint32_t headerBits, totalBits, nUsedLit, nUsedDist, crc;
};
struct SgDeflateArena {
    uint16_t *mlen;      // [MAX_PAYLOAD]   match length at i (0 = none)
    uint16_t *mdist;     // [MAX_PAYLOAD]   its distance
    uint16_t *jumpA;     // [MAX_PAYLOAD+1] pointer-jumping tables (entry n = n)
    uint16_t *jumpB;
};
#define SG_DEFLATE_ARENA_BYTES ((size_t)(4 * (SG_DEFLATE_MAX_PAYLOAD + 8)) * 2)

SG_HD uint32_t sgd_load32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
SG_HD uint32_t sgd_hash(uint32_t v) { return (v * 2654435761u) >> (32 - SG_DEFLATE_HASH_BITS); }
SG_HD uint32_t sgd_log2(uint32_t v) { uint32_t r = 0; while (v >>= 1) r++; return r; }
// 8 * log2(v), v >= 1, to 1/8 bit: the integer part from the leading bit, the fraction read off the next three bits
SG_HD uint32_t sgd_log2_x8(uint32_t v) { const uint32_t e = sgd_log2(v); return 8u * e + (e >= 3 ? ((v >> (e - 3)) & 7u) : ((v << (3 - e)) & 7u)); }
SG_HD uint32_t sgd_reverse(uint32_t code, uint32_t len) { uint32_t r = 0; for (uint32_t k = 0; k < len; k++) { r = (r << 1) | (code & 1u); code >>= 1; } return r; }

// RFC 1951 3.2.5: length 3..258 -> (symbol 257..285, extra bits, extra value); distance 1..32768 -> (symbol 0..29, extra bits, extra value)
SG_HD void sgd_length_symbol(uint32_t length, uint32_t *sym, uint32_t *ebits, uint32_t *eval)
{
    if (length == 258u) { *sym = 285; *ebits = 0; *eval = 0; return; }
    const uint32_t l = length - 3u;
    if (l < 8u) { *sym = 257u + l; *ebits = 0; *eval = 0; return; }
    const uint32_t eb = sgd_log2(l) - 2u;
    *sym = 257u + 4u * eb + 4u + ((l >> eb) & 3u); *ebits = eb; *eval = l & ((1u << eb) - 1u);
}
SG_HD void sgd_dist_symbol(uint32_t dist, uint32_t *sym, uint32_t *ebits, uint32_t *eval)
{
    const uint32_t d = dist - 1u;
    if (d < 4u) { *sym = d; *ebits = 0; *eval = 0; return; }
    const uint32_t eb = sgd_log2(d) - 1u;
    *sym = 2u * eb + 2u + ((d >> eb) & 1u); *ebits = eb; *eval = d & ((1u << eb) - 1u);
}

// value (nbits <= 32) at bit position `at` of the little-endian bit stream held in `words`
SG_HD void sgd_put_bits(uint32_t *words, uint32_t at, uint32_t value, uint32_t nbits)
{
    if (nbits == 0) return;
    const uint32_t w = at >> 5, s = at & 31u;
    SGD_ATOMIC_OR(&words[w], value << s);
    if (s + nbits > 32u) SGD_ATOMIC_OR(&words[w + 1], value >> (32u - s));
}

// Code lengths (<= maxBits) of the n symbols with freq > 0 among freq[0..nSym), by ONE thread.  sortSym / sortFreq: the used symbols in ascending
// (frequency, symbol) order, nUsed of them (>= 2: the caller adds dummies).  Moffat & Katajainen, "In-place calculation of minimum-redundancy
// codes" (1995), then the length limit.
SG_HD void sgd_code_lengths(const uint16_t *sortSym, uint32_t *A /* = sortFreq, destroyed */, int n, int maxBits, uint8_t *len)
{
    if (n == 1) { len[sortSym[0]] = 1; return; }
    // phase 1: internal nodes
    A[0] += A[1];
    int root = 0, leaf = 2, next;
    for (next = 1; next < n - 1; next++) {
        if (leaf >= n || A[root] < A[leaf]) { A[next] = A[root]; A[root++] = (uint32_t)next; } else A[next] = A[leaf++];
        if (leaf >= n || (root < next && A[root] < A[leaf])) { A[next] += A[root]; A[root++] = (uint32_t)next; } else A[next] += A[leaf++];
    }
    // phase 2: depths of the internal nodes
    A[n - 2] = 0;
    for (next = n - 3; next >= 0; next--) A[next] = A[A[next]] + 1;
    // phase 3: depths of the leaves
    int avbl = 1, used = 0, dpth = 0;
    root = n - 2; next = n - 1;
    while (avbl > 0) {
        while (root >= 0 && (int)A[root] == dpth) { used++; root--; }
        while (avbl > used) { A[next--] = (uint32_t)dpth; avbl--; }
        avbl = 2 * used; dpth++; used = 0;
    }
    // A[i] = code length of the i-th least frequent symbol (non-increasing in i).  Limit to maxBits: count per length, fold the overlong ones
    // into maxBits, then restore the Kraft sum by turning one maxBits code and one shorter code into ... (the classic fix, as in zlib / miniz)
    int count[33];
    for (int b = 0; b <= 32; b++) count[b] = 0;
    for (int i = 0; i < n; i++) count[A[i] > 32u ? 32 : A[i]]++;
    for (int b = maxBits + 1; b <= 32; b++) { count[maxBits] += count[b]; count[b] = 0; }
    uint32_t total = 0;
    for (int b = maxBits; b > 0; b--) total += (uint32_t)count[b] << (maxBits - b);
    while (total != (1u << maxBits)) {
        count[maxBits]--;
        for (int b = maxBits - 1; b > 0; b--) if (count[b]) { count[b]--; count[b + 1] += 2; break; }
        total--;
    }
    // longest codes to the least frequent symbols
    int i = 0;
    for (int b = maxBits; b > 0; b--) for (int c = count[b]; c > 0; c--) len[sortSym[i++]] = (uint8_t)b;
}

// canonical codes (RFC 1951 3.2.2) of lengths len[0..nSym), bit-reversed for the LSB-first bit stream; by one thread
SG_HD void sgd_canonical_codes(const uint8_t *len, int nSym, uint16_t *code)
{
    uint32_t blCount[16], nextCode[16];
    for (int b = 0; b < 16; b++) blCount[b] = 0;
    for (int s = 0; s < nSym; s++) blCount[len[s]]++;
    blCount[0] = 0;
    uint32_t c = 0;
    nextCode[0] = 0;
    for (int b = 1; b < 16; b++) { c = (c + blCount[b - 1]) << 1; nextCode[b] = c; }
    for (int s = 0; s < nSym; s++) code[s] = len[s] ? (uint16_t)sgd_reverse(nextCode[len[s]]++, len[s]) : 0;
}

// rank sort of the used symbols of one alphabet into sortSym / sortFreq (all threads), then lengths and codes (thread `owner`)
SG_HD void sgd_build_alphabet(SgDeflateShared &S, uint32_t *freq, int nSym, uint8_t *len, uint16_t *code, uint32_t *nUsedOut, uint32_t owner)
{
    for (uint32_t s = SGD_TID; s < (uint32_t)nSym; s += SGD_NT) len[s] = 0;
    SGD_SYNC();
    for (uint32_t s = SGD_TID; s < (uint32_t)nSym; s += SGD_NT) {
        const uint32_t f = freq[s];
        if (f == 0) continue;
        uint32_t rank = 0;
        for (int t = 0; t < nSym; t++) { const uint32_t g = freq[t]; rank += (g != 0 && (g < f || (g == f && (uint32_t)t < s))) ? 1u : 0u; }
        S.sortSym[rank] = (uint16_t)s; S.sortFreq[rank] = f;
    }
    SGD_SYNC();
    if (SGD_TID == owner % SGD_NT) {
        uint32_t n = 0;
        for (int t = 0; t < nSym; t++) n += freq[t] != 0;
        sgd_code_lengths(S.sortSym, S.sortFreq, (int)n, 15, len);
        sgd_canonical_codes(len, nSym, code);
        int last = nSym;
        while (last > 0 && len[last - 1] == 0) last--;
        *nUsedOut = (uint32_t)last;
    }
    SGD_SYNC();
}

SG_HD uint32_t sgd_crc_multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1u)) == 0) break; }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xedb88320u : b >> 1;
    }
    return p;
}
SG_HD uint32_t sgd_crc_x8n(const uint32_t *x2n, uint32_t nBytes)
{
    uint32_t p = 1u << 31, k = 3;
    while (nBytes) { if (nBytes & 1u) p = sgd_crc_multmodp(x2n[k & 31u], p); nBytes >>= 1; k++; }
    return p;
}

// token at position i of the parse: its symbols and the number of bits it takes with the current code lengths
struct SgDeflateToken { uint32_t litSym, lebits, leval, distSym, debits, deval; bool match; };
SG_HD void sgd_token(const SgDeflateShared &S, const SgDeflateArena &G, uint32_t i, SgDeflateToken *t)
{
    const uint32_t ml = G.mlen[i];
    if (ml >= SG_DEFLATE_MIN_MATCH) {
        t->match = true;
        sgd_length_symbol(ml, &t->litSym, &t->lebits, &t->leval);
        sgd_dist_symbol(G.mdist[i], &t->distSym, &t->debits, &t->deval);
    } else {
        t->match = false; t->litSym = S.buf[i]; t->lebits = t->leval = 0; t->distSym = t->debits = t->deval = 0;
    }
}

// One BGZF member from n <= SG_DEFLATE_MAX_PAYLOAD bytes at `src`, written at `member` (room for SG_DEFLATE_MEMBER_PITCH bytes); returns its size
// (the same value in every thread).  Called by all threads of the block (one host thread in the test build).
SG_HD uint32_t sg_deflate_member(SgDeflateShared &S, const SgDeflateArena &G, const uint8_t *src, uint32_t n, uint8_t *member)
{
    const uint32_t tid = SGD_TID, nt = SGD_NT;
    uint16_t *hashTab = (uint16_t *)S.out;                 // [1 << HASH_BITS], 0xffff = empty
    // ---- load, tables ----
    for (uint32_t i = tid; i < n; i += nt) S.buf[i] = src[i];
    for (uint32_t i = n + tid; i < n + 8u; i += nt) S.buf[i] = 0;
    for (uint32_t i = tid; i < (1u << SG_DEFLATE_HASH_BITS); i += nt) hashTab[i] = 0xffffu;
    for (uint32_t i = tid; i < SG_DEFLATE_MAX_PAYLOAD / 32 + 2; i += nt) S.sel[i] = 0;
    for (uint32_t i = tid; i < SG_DEFLATE_NLIT + 2; i += nt) S.litFreq[i] = 0;
    for (uint32_t i = tid; i < SG_DEFLATE_NDIST + 2; i += nt) S.distFreq[i] = 0;
    for (uint32_t i = tid; i < 256u; i += nt) S.byteFreq[i] = 0;
    for (uint32_t i = tid; i < 256u; i += nt) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xedb88320u ^ (c >> 1)) : (c >> 1); S.crcTable[i] = c; }
    if (tid == 0) { uint32_t v = 1u << 30; S.x2n[0] = v; for (int k = 1; k < 32; k++) { v = sgd_crc_multmodp(v, v); S.x2n[k] = v; } }
    SGD_SYNC();
    // ---- CRC-32 of the payload: 1024 slices, joined over GF(2) in two levels (crc(A || B) = crc(A) x^(8|B|) + crc(B)): 32 threads join 32 slices
    //      each, one joins the 32 groups ----
    {
        const uint32_t slice = (n + 1023u) / 1024u;
        for (uint32_t t = tid; t < 1024u; t += nt) {
            const uint32_t lo = t * slice < n ? t * slice : n, hi = lo + slice < n ? lo + slice : n;
            uint32_t crc = 0xffffffffu;
            for (uint32_t k = lo; k < hi; k++) crc = S.crcTable[(crc ^ S.buf[k]) & 0xffu] ^ (crc >> 8);
            S.partial[t] = crc ^ 0xffffffffu;
        }
        SGD_SYNC();
        for (uint32_t g = tid; g < 32u; g += nt) {
            const uint32_t xs = sgd_crc_x8n(S.x2n, slice);
            uint32_t crc = 0, bytes = 0;
            for (uint32_t t = 32u * g; t < 32u * g + 32u; t++) {
                const uint32_t lo = t * slice < n ? t * slice : n, hi = lo + slice < n ? lo + slice : n;
                if (hi == lo) break;
                crc = bytes ? (sgd_crc_multmodp(hi - lo == slice ? xs : sgd_crc_x8n(S.x2n, hi - lo), crc) ^ S.partial[t]) : S.partial[t];
                bytes += hi - lo;
            }
            S.crcGroup[g] = crc; S.crcGroupBytes[g] = bytes;
        }
        SGD_SYNC();
        if (tid == 0) {
            uint32_t crc = 0, bytes = 0;
            for (uint32_t g = 0; g < 32u; g++) {
                const uint32_t b = S.crcGroupBytes[g];
                if (b == 0) break;
                crc = bytes ? (sgd_crc_multmodp(sgd_crc_x8n(S.x2n, b), crc) ^ S.crcGroup[g]) : S.crcGroup[g];
                bytes += b;
            }
            S.crc = crc;
        }
        SGD_SYNC();
    }
    // ---- what a literal costs, roughly: -log2 of the byte's share of the payload.  A match is only worth taking when the literals it replaces
    //      would cost more than its own length / distance symbols (short far matches in low-entropy data -- 4-bit packed bases -- would not) ----
    for (uint32_t i = tid; i < 32u * (SG_DEFLATE_NLIT + 2); i += nt) (&S.warpHist[0][0])[i] = 0;
    SGD_SYNC();
    for (uint32_t i = tid; i < n; i += nt) SGD_ATOMIC_ADD(&S.warpHist[SGD_WARP][S.buf[i]], 1u);
    SGD_SYNC();
    for (uint32_t b = tid; b < 256u; b += nt) {
        uint32_t f = 0;
        for (uint32_t w = 0; w < 32u; w++) f += S.warpHist[w][b];
        S.byteFreq[b] = f;
        uint32_t c = f ? sgd_log2_x8(n) - sgd_log2_x8(f) : 120u;
        S.litCost[b] = (uint8_t)(c < 8u ? 8u : (c > 120u ? 120u : c));
    }
    SGD_SYNC();
    // ---- 1. matches ----
    for (uint32_t base = 0; base < n; base += SG_DEFLATE_STRIDE) {
        const uint32_t end = base + SG_DEFLATE_STRIDE < n ? base + SG_DEFLATE_STRIDE : n;
        for (uint32_t i = base + tid; i < end; i += nt) {
            uint32_t bestLen = 0, bestDist = 0;
            if (i + SG_DEFLATE_MIN_MATCH <= n) {
                const uint32_t limit = n - i < SG_DEFLATE_MAX_MATCH ? n - i : SG_DEFLATE_MAX_MATCH;
                const uint32_t v = sgd_load32(S.buf + i);
                const uint32_t cand = hashTab[sgd_hash(v)];
                uint32_t bestGain = 0;                                     // literal cost saved minus the match's own estimated cost, 1/8 bit
                if (cand != 0xffffu && i - cand <= SG_DEFLATE_MAX_DIST && sgd_load32(S.buf + cand) == v) {
                    uint32_t l = 4, cost = (uint32_t)S.litCost[S.buf[i]] + S.litCost[S.buf[i + 1]] + S.litCost[S.buf[i + 2]] + S.litCost[S.buf[i + 3]];
                    while (l < limit && S.buf[cand + l] == S.buf[i + l]) { cost += S.litCost[S.buf[i + l]]; l++; }
                    const uint32_t d = i - cand;
                    const uint32_t own = 8u * (12u + (l < 11u ? 0u : sgd_log2(l - 3u) - 2u) + (d < 5u ? 0u : sgd_log2(d - 1u) - 1u));
                    if (cost > own) { bestLen = l; bestDist = d; bestGain = cost - own; }
                }
                if (i > 0 && sgd_load32(S.buf + i - 1) == v) {             // a run of one byte
                    uint32_t l = 4;
                    while (l < limit && S.buf[i - 1 + l] == S.buf[i + l]) l++;
                    const uint32_t cost = l * S.litCost[S.buf[i]];
                    const uint32_t own = 8u * (12u + (l < 11u ? 0u : sgd_log2(l - 3u) - 2u));
                    if (cost > own && cost - own > bestGain) { bestLen = l; bestDist = 1; }
                }
            }
            G.mlen[i] = (uint16_t)bestLen; G.mdist[i] = (uint16_t)bestDist;
            G.jumpA[i] = (uint16_t)(i + (bestLen >= SG_DEFLATE_MIN_MATCH ? bestLen : 1u));
        }
        SGD_SYNC();
        for (uint32_t i = base + tid; i < end; i += nt) if (i + SG_DEFLATE_MIN_MATCH <= n) hashTab[sgd_hash(sgd_load32(S.buf + i))] = (uint16_t)i;
        SGD_SYNC();
    }
    if (tid == 0) { G.jumpA[n] = (uint16_t)n; G.jumpB[n] = (uint16_t)n; if (n) S.sel[0] = 1u; }
    SGD_SYNC();
    // ---- 2. the parse: positions reachable from 0 ----
    {
        uint16_t *cur = G.jumpA, *nxt = G.jumpB;
        for (uint32_t span = 1; span < n; span <<= 1) {
            // four positions per step: the two dependent loads of each (cur[i], then cur[cur[i]]) are issued together, so the L2 latency of the
            // jump tables is paid once per four positions
            for (uint32_t i0 = tid; i0 < n; i0 += 4u * nt) {
                uint32_t j[4], k[4];
                #pragma unroll
                for (int u = 0; u < 4; u++) { const uint32_t i = i0 + (uint32_t)u * nt; j[u] = i < n ? cur[i] : n; }
                #pragma unroll
                for (int u = 0; u < 4; u++) k[u] = cur[j[u]];
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)u * nt;
                    if (i >= n) continue;
                    if ((S.sel[i >> 5] >> (i & 31u)) & 1u) { if (j[u] < n) SGD_ATOMIC_OR(&S.sel[j[u] >> 5], 1u << (j[u] & 31u)); }
                    nxt[i] = (uint16_t)k[u];
                }
            }
            SGD_SYNC();
            uint16_t *t = cur; cur = nxt; nxt = t;
        }
    }
    // ---- 3. histograms, codes ----
    for (uint32_t i = tid; i < 32u * (SG_DEFLATE_NLIT + 2); i += nt) (&S.warpHist[0][0])[i] = 0;
    SGD_SYNC();
    for (uint32_t i = tid; i < n; i += nt) {
        if (!((S.sel[i >> 5] >> (i & 31u)) & 1u)) continue;
        SgDeflateToken t; sgd_token(S, G, i, &t);
        SGD_ATOMIC_ADD(&S.warpHist[SGD_WARP][t.litSym], 1u);
        if (t.match) SGD_ATOMIC_ADD(&S.distFreq[t.distSym], 1u);
    }
    SGD_SYNC();
    for (uint32_t b = tid; b < (uint32_t)SG_DEFLATE_NLIT; b += nt) {
        uint32_t f = 0;
        for (uint32_t w = 0; w < 32u; w++) f += S.warpHist[w][b];
        S.litFreq[b] = f;
    }
    SGD_SYNC();
    if (tid == 0) {
        S.litFreq[256] = 1;                                  // end of block
        // at least two symbols per alphabet, so that both codes are complete (zlib's inflate rejects an incomplete literal/length code)
        uint32_t used = 0;
        for (int s = 0; s < SG_DEFLATE_NLIT; s++) used += S.litFreq[s] != 0;
        if (used < 2) S.litFreq[S.litFreq[0] ? 1 : 0] = 1;
        used = 0;
        for (int s = 0; s < SG_DEFLATE_NDIST; s++) used += S.distFreq[s] != 0;
        for (int s = 0; used < 2 && s < 2; s++) if (!S.distFreq[s]) { S.distFreq[s] = 1; used++; }
    }
    SGD_SYNC();
    sgd_build_alphabet(S, S.litFreq, SG_DEFLATE_NLIT, S.litLen, S.litCode, &S.nUsedLit, 0);
    sgd_build_alphabet(S, S.distFreq, SG_DEFLATE_NDIST, S.distLen, S.distCode, &S.nUsedDist, 0);
    // ---- 4. the block image ----
    for (uint32_t i = tid; i < SG_DEFLATE_MAX_PAYLOAD / 4 + 8; i += nt) S.out[i] = 0;      // (the hash table is dead)
    SGD_SYNC();
    const uint32_t nLit = S.nUsedLit < 257u ? 257u : S.nUsedLit, nDist = S.nUsedDist < 1u ? 1u : S.nUsedDist;
    if (tid == 0) {
        uint32_t at = 0;
        sgd_put_bits(S.out, at, 1u, 1); at += 1;            // BFINAL
        sgd_put_bits(S.out, at, 2u, 2); at += 2;            // BTYPE = 10: dynamic Huffman
        sgd_put_bits(S.out, at, nLit - 257u, 5); at += 5;
        sgd_put_bits(S.out, at, nDist - 1u, 5); at += 5;
        sgd_put_bits(S.out, at, 15u, 4); at += 4;           // HCLEN: all 19 code length code lengths follow
        // in the order 16, 17, 18, 0, 8, 7, ...: the repeat symbols unused (0), a flat 4-bit code for the lengths 0..15
        for (int k = 0; k < 19; k++) { sgd_put_bits(S.out, at, k < 3 ? 0u : 4u, 3); at += 3; }
        S.headerBits = at + 4u * (nLit + nDist);
    }
    SGD_SYNC();
    {
        const uint32_t at0 = 17u + 57u;
        for (uint32_t s = tid; s < nLit + nDist; s += nt) {
            const uint32_t l = s < nLit ? S.litLen[s] : S.distLen[s - nLit];
            sgd_put_bits(S.out, at0 + 4u * s, sgd_reverse(l, 4), 4);
        }
    }
    // bit length of every token, prefix sums: each thread owns a contiguous range of positions
    const uint32_t per = (n + nt - 1u) / nt;
    const uint32_t lo = tid * per < n ? tid * per : n, hi = lo + per < n ? lo + per : n;
    {
        uint32_t bits = 0;
        for (uint32_t i = lo; i < hi; i++) {
            if (!((S.sel[i >> 5] >> (i & 31u)) & 1u)) continue;
            SgDeflateToken t; sgd_token(S, G, i, &t);
            bits += S.litLen[t.litSym] + t.lebits;
            if (t.match) bits += S.distLen[t.distSym] + t.debits;
        }
        S.partial[tid] = bits;
    }
    SGD_SYNC();
    if (tid == 0) {
        uint32_t run = S.headerBits;
        for (uint32_t t = 0; t < nt; t++) { const uint32_t b = S.partial[t]; S.partial[t] = run; run += b; }
        S.totalBits = run + S.litLen[256];
    }
    SGD_SYNC();
    const uint32_t totalBits = S.totalBits;
    const uint32_t deflateBytes = (totalBits + 7u) / 8u;
    const bool stored = deflateBytes >= n + 5u || deflateBytes + 26u > 65536u;
    if (!stored) {
        uint32_t at = S.partial[tid];
        for (uint32_t i = lo; i < hi; i++) {
            if (!((S.sel[i >> 5] >> (i & 31u)) & 1u)) continue;
            SgDeflateToken t; sgd_token(S, G, i, &t);
            uint32_t l = S.litLen[t.litSym];
            sgd_put_bits(S.out, at, S.litCode[t.litSym], l); at += l;
            if (t.match) {
                sgd_put_bits(S.out, at, t.leval, t.lebits); at += t.lebits;
                l = S.distLen[t.distSym];
                sgd_put_bits(S.out, at, S.distCode[t.distSym], l); at += l;
                sgd_put_bits(S.out, at, t.deval, t.debits); at += t.debits;
            }
        }
        if (tid == 0) sgd_put_bits(S.out, totalBits - S.litLen[256], S.litCode[256], S.litLen[256]);
    }
    SGD_SYNC();
    // ---- the member: gzip header with the BGZF extra field, the deflate data, CRC-32, ISIZE ----
    const uint32_t dataBytes = stored ? n + 5u : deflateBytes;
    const uint32_t total = 18u + dataBytes + 8u;
    if (stored) {
        for (uint32_t i = tid; i < n; i += nt) member[23 + i] = S.buf[i];
        if (tid == 0) { member[18] = 1; member[19] = (uint8_t)(n & 0xffu); member[20] = (uint8_t)(n >> 8); member[21] = (uint8_t)(~n & 0xffu); member[22] = (uint8_t)((~n >> 8) & 0xffu); }
    } else {
        for (uint32_t i = tid; i < deflateBytes; i += nt) member[18 + i] = (uint8_t)(S.out[i >> 2] >> (8u * (i & 3u)));
    }
    if (tid == 0) {
        const uint8_t hdr[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, (uint8_t)((total - 1u) & 0xffu), (uint8_t)((total - 1u) >> 8)};
        for (int k = 0; k < 18; k++) member[k] = hdr[k];
        uint8_t *t = member + 18 + dataBytes;
        const uint32_t crc = S.crc;
        t[0] = (uint8_t)crc; t[1] = (uint8_t)(crc >> 8); t[2] = (uint8_t)(crc >> 16); t[3] = (uint8_t)(crc >> 24);
        t[4] = (uint8_t)n; t[5] = (uint8_t)(n >> 8); t[6] = 0; t[7] = 0;
    }
    SGD_SYNC();
    return total;
}

#endif
