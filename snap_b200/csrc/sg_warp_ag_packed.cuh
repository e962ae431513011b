// sg_warp_ag_packed.cuh -- the unbanded affine-gap DP for patterns of up to 24 striped vectors (192 columns), two cells
// per lane as s16x2 on the DPX min/max unit (VIMNMX.S16x2 / VIADDMNMX.S16x2[.RELU] with their per-half predicates).
//
// Same cells, same order of lazy-F visits and same commit rule as sg_warp_ag_compute's generic path (and therefore as
// AffineGapVectorized::computeScore, AffineGapVectorized.h:821-1339), re-scheduled:
//
//   lane = qq*4 + a     qq = vector within a block of 8 consecutive striped vectors, a = pair of SSE lanes (2a, 2a+1)
//
// so one 32-bit word per lane covers a 128-byte piece of the H / E rows and 64 bytes of the traceback row.  A whole row is
// at most 3 blocks, which stay in registers (H packed, 6 traceback bits per cell) from the main pass through every lazy-F
// pass and are written once.  The reference works in saturating int16; here every quantity stays within +-16384 (scores
// <= scoreInit + patternLen <= 2*MAX_READ_LENGTH, padding columns use -16384 instead of INT16_MIN: any value below every
// real score behaves identically), so plain packed adds are exact.
#pragma once

#define SG_AGP_MAX_VEC 24
#define SG_AGP_BLOCKS 3
#define SG_AGP_PAD (-16384)

struct SgAgBests { int gScore, gText, lScore, lText, lPat; };

__device__ __forceinline__ unsigned sg_pk(int lo, int hi) { return ((unsigned)hi << 16) | ((unsigned)lo & 0xffffu); }
__device__ __forceinline__ unsigned sg_pk2(int v) { return sg_pk(v, v); }
// 0xffff in every half that is non-zero (halves are >= 0): sign of (0 - x) replicated over the half
__device__ __forceinline__ unsigned sg_nzmask2(unsigned x)
{
    unsigned neg = __vsub2(0u, x), m;
#if defined(__CUDA_ARCH__)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(neg), "r"(0u), "r"(0xbb99u));
#else
    m = ((neg & 0x8000u) ? 0xffffu : 0u) | ((neg & 0x80000000u) ? 0xffff0000u : 0u);       // (host build: the SIMT emulator of the CPU test suite)
#endif
    return m;
}
__device__ __forceinline__ unsigned sg_relu2(unsigned x) { return __vimax_s16x2_relu(x, x); }

// UNROLLED: the three block loops unrolled with the blocks in fixed registers (fewer instructions, ~3x the loop code); used where
// this function is nearly all a kernel runs (stage 2 of the paired launch, which is issue-bound).  Otherwise loop-compact.
template <bool UNROLLED>
__device__ __noinline__ void sg_warp_ag_rows_packed(const SgScratch &S, const SgAgParams &P, int dir, const uint8_t *text, int textLen,
                                                    const uint8_t *pattern, int patternLen, int scoreInit, SgAgLayout &lay, uint8_t *bt,
                                                    int lane, SgAgBests *res)
{
    constexpr int kBlockUnroll = UNROLLED ? SG_AGP_BLOCKS : 1;
    const int numVec = lay.numVec;
    const int strideW = numVec * 4;                  // 32-bit words per row (8 int16 per vector)
    const int nBlocks = (numVec + 7) >> 3;
    const int qq = lane >> 2, a = lane & 3;
    const int open = P.gapOpenPenalty, ext = P.gapExtendPenalty;
    const unsigned nOpen2 = sg_pk2(-open), nExt2 = sg_pk2(-ext), nExt2x2 = sg_pk2(-2 * ext), nExt4x2 = sg_pk2(-4 * ext);
    const unsigned nQqExt2 = sg_pk2(-qq * ext), nRowExt2 = sg_pk2(-numVec * ext);
    unsigned *Hp = (unsigned *)S.agH, *Hm = (unsigned *)S.agHm1, *Ew = (unsigned *)S.agE;
    unsigned *prof = (unsigned *)S.agProf;           // [5][strideW] packed substitution scores of the pattern columns

    for (int w = lane; w < strideW; w += 32) {
        const int vec = w >> 2, aa = w & 3;
        const int c0 = (2 * aa) * numVec + vec, c1 = (2 * aa + 1) * numVec + vec;
        const uint32_t p0 = (c0 < patternLen) ? sg_base_value(pattern[c0]) : 5u;
        const uint32_t p1 = (c1 < patternLen) ? sg_base_value(pattern[c1]) : 5u;
        for (uint32_t t = 0; t < 5; t++) {
            prof[t * strideW + w] = sg_pk(p0 == 5u ? SG_AGP_PAD : sg_ag_sub(P, t, p0), p1 == 5u ? SG_AGP_PAD : sg_ag_sub(P, t, p1));
        }
    }
    __syncwarp();

    int bestG = -1, bestGT = -1, bestL = -1, bestLT = -1, bestLP = -1;
    const int globalIdx = lay.cellIndex(patternLen - 1);

    #pragma unroll 1
    for (int i = 0; i < textLen; i++) {
        lay.nRows = i + 1;
        const uint32_t tb = sg_base_value(text[i * dir]);
        const unsigned *profRow = prof + tb * strideW;
        uint8_t *btRow = bt + (size_t)i * (numVec * SG_VEC);
        int hInit = scoreInit;
        if (i > 0) { hInit = scoreInit - open - (i - 1) * ext; if (hInit < 0) hInit = 0; }

        // The row's (up to) three blocks live in three register pairs.  To keep the hot loop bodies single-copy (instruction
        // cache: the kernel is fetch-bound) the block loops are NOT unrolled: each iteration works on (h0, a0) and then rotates
        // (h0,h1,h2) <- (h1,h2,h0); every loop makes exactly SG_AGP_BLOCKS rotations, so the registers end up in place again.
        unsigned h0 = 0, h1r = 0, h2 = 0, a0 = 0, a1 = 0, a2r = 0;
        unsigned hreg[SG_AGP_BLOCKS], areg[SG_AGP_BLOCKS];          // UNROLLED form: block b in (hreg[b], areg[b])
        #define SG_HB (*(UNROLLED ? &hreg[b] : &h0))
        #define SG_AB (*(UNROLLED ? &areg[b] : &a0))
        unsigned fcarry = 0;                         // F entering the next vector of SSE lanes (2a, 2a+1); identical for all qq

        // ---------------- main pass: 8 vectors per step ----------------
        #pragma unroll kBlockUnroll
        for (int b = 0; b < SG_AGP_BLOCKS; b++) {
            unsigned hnew = 0, anew = 0;
            if (b < nBlocks) {
                const int k = 8 * b + qq;
                const bool valid = k < numVec;
                const int w = k * 4 + a;
                unsigned hd = 0, e = 0, pv = 0;
                if (valid) {
                    hd = Hp[(k == 0) ? (numVec - 1) * 4 + a : w - 4];
                    e = Ew[w];
                    pv = profRow[w];
                }
                if (b == 0) {
                    // vector 0 takes the previous row's LAST vector moved up one SSE lane, hInit entering lane 0 (:1027-1031)
                    const unsigned below = __shfl_up_sync(0xffffffffu, hd, 1);
                    if (qq == 0) hd = (hd << 16) | (a == 0 ? ((unsigned)hInit & 0xffffu) : (below >> 16));
                }
                const unsigned m = __vadd2(hd, pv) & sg_nzmask2(hd);              // (hdiag > 0) ? hdiag + profile : 0
                bool mgeHi, mgeLo, tgeHi, tgeLo, hgeHi, hgeLo, t2Hi, t2Lo;
                const unsigned h1 = __vibmax_s16x2(m, e, &mgeHi, &mgeLo);          // bit 1: e > m
                const unsigned temp = __viaddmax_s16x2_relu(m, nOpen2, 0u);        // max(m - open, 0)
                const unsigned e2 = __vadd2(e, nExt2);
                const unsigned ne = __vibmax_s16x2(temp, e2, &tgeHi, &tgeLo);      // bit 4: e - ext > temp
                if (valid) Ew[w] = ne;
                // F entering vector qq of this block: max-plus prefix over the earlier vectors (f' = max(f - ext, temp))
                unsigned g = temp, y;
                y = __shfl_up_sync(0xffffffffu, g, 4);  if (qq >= 1) g = __viaddmax_s16x2(y, nExt2, g);
                y = __shfl_up_sync(0xffffffffu, g, 8);  if (qq >= 2) g = __viaddmax_s16x2(y, nExt2x2, g);
                y = __shfl_up_sync(0xffffffffu, g, 16); if (qq >= 4) g = __viaddmax_s16x2(y, nExt4x2, g);
                const unsigned gi = __shfl_up_sync(0xffffffffu, g, 4);             // g of vector qq-1
                unsigned fin = fcarry;
                if (qq > 0) fin = __vimax_s16x2_relu(gi, __vadd2(fcarry, nQqExt2));
                const unsigned h = __vibmax_s16x2(h1, fin, &hgeHi, &hgeLo);        // bit 2: f > h
                (void)__vibmax_s16x2(temp, __vadd2(fin, nExt2), &t2Hi, &t2Lo);     // bit 32: f - ext > temp
                hnew = valid ? h : 0u;
                anew = ((mgeLo ? 0u : 1u) | (hgeLo ? 0u : 2u) | (tgeLo ? 0u : 4u) | (t2Lo ? 0u : 32u)) |
                       (((mgeHi ? 0u : 1u) | (hgeHi ? 0u : 2u) | (tgeHi ? 0u : 4u) | (t2Hi ? 0u : 32u)) << 8);
                // F leaving the block
                int nv = numVec - 8 * b; if (nv > 8) nv = 8;
                const unsigned gl = __shfl_sync(0xffffffffu, g, (nv - 1) * 4 + a);
                fcarry = __vimax_s16x2_relu(gl, __vadd2(fcarry, sg_pk2(-nv * ext)));
            }
            if (UNROLLED) { hreg[b] = hnew; areg[b] = anew; }
            else {
                h0 = h1r; h1r = h2; h2 = hnew;       // rotate: after the loop (h0,h1r,h2) = blocks (0,1,2)
                a0 = a1; a1 = a2r; a2r = anew;
            }
        }

        // ---------------- lazy F (:1080-1112): whole blocks evaluated at once, committed up to the first vector at which
        //                  no SSE lane can still change H ----------------
        unsigned fl = fcarry;
        bool converged = false;
        #pragma unroll 1
        for (int kk = 0; kk < SG_VEC && !converged; kk++) {
            const unsigned below = __shfl_up_sync(0xffffffffu, fl, 1);
            fl = (fl << 16) | (a == 0 ? 0u : (below >> 16));                       // f = f << one SSE lane
            #pragma unroll kBlockUnroll
            for (int b = 0; b < SG_AGP_BLOCKS; b++) {
                if (b < nBlocks && !converged) {
                    const int k = 8 * b + qq;
                    const bool valid = k < numVec;
                    bool hgeHi, hgeLo, tgeHi, tgeLo;
                    const unsigned fv = __viaddmax_s16x2_relu(fl, sg_pk2(-k * ext), 0u);
                    const unsigned newh = __vibmax_s16x2(SG_HB, fv, &hgeHi, &hgeLo);               // f > h
                    const unsigned temp = __viaddmax_s16x2_relu(newh, nOpen2, 0u);
                    const unsigned fn = __viaddmax_s16x2_relu(fv, nExt2, 0u);
                    (void)__vibmax_s16x2(temp, fn, &tgeHi, &tgeLo);                                // f - ext > h - open
                    const unsigned liveMask = __ballot_sync(0xffffffffu, valid && (!tgeHi || !tgeLo));
                    int nv = numVec - 8 * b; if (nv > 8) nv = 8;
                    // lowest all-zero nibble among the valid vectors = first vector at which no SSE lane is live
                    const unsigned z = (liveMask - 0x11111111u) & ~liveMask & 0x88888888u & (nv >= 8 ? 0xffffffffu : ((1u << (4 * nv)) - 1u));
                    const int firstConv = z ? ((__ffs(z) - 1) >> 2) : 8;
                    if (valid && qq <= firstConv) {
                        SG_HB = newh;
                        SG_AB |= ((hgeLo ? 0u : 2u) | (tgeLo ? 0u : 32u)) | (((hgeHi ? 0u : 2u) | (tgeHi ? 0u : 32u)) << 8);
                    }
                    if (firstConv < 8) converged = true;
                }
                { const unsigned t = h0; h0 = h1r; h1r = h2; h2 = t; }
                { const unsigned t = a0; a0 = a1; a1 = a2r; a2r = t; }
            }
            if (!converged) fl = __viaddmax_s16x2_relu(fl, nRowExt2, 0u);
        }

        // ---------------- write the row once; per-lane row maximum and the largest column holding it ----------------
        unsigned rmax = 0; int kLo = -1, kHi = -1;
        #pragma unroll kBlockUnroll
        for (int b = 0; b < SG_AGP_BLOCKS; b++) {
            const int k = 8 * b + qq;
            if (b < nBlocks && k < numVec) {
                const int w = k * 4 + a;
                Hm[w] = SG_HB;
                *(uint16_t *)(btRow + 2 * w) = (uint16_t)SG_AB;
                bool geHi, geLo;
                rmax = __vibmax_s16x2(SG_HB, rmax, &geHi, &geLo);
                if (geLo) kLo = k;
                if (geHi) kHi = k;
            }
            { const unsigned t = h0; h0 = h1r; h1r = h2; h2 = t; }
            { const unsigned t = a0; a0 = a1; a1 = a2r; a2r = t; }
        }
        __syncwarp();
        const int mLo = (int)(short)(rmax & 0xffffu), mHi = (int)(short)(rmax >> 16);
        const int maxScoreRow = __reduce_max_sync(0xffffffffu, mLo > mHi ? mLo : mHi);

        {
            const int globalAlignmentScore = ((const int16_t *)Hm)[globalIdx];
            if (globalAlignmentScore >= bestG) { bestG = globalAlignmentScore; bestGT = i; }
        }
        if (maxScoreRow == 0) break;
        if (maxScoreRow > bestL) {
            int cand = -1;
            if (mLo == maxScoreRow && kLo >= 0) cand = (2 * a) * numVec + kLo;
            if (mHi == maxScoreRow && kHi >= 0) { const int c = (2 * a + 1) * numVec + kHi; if (c > cand) cand = c; }
            bestL = maxScoreRow; bestLT = i; bestLP = __reduce_max_sync(0xffffffffu, cand);
        }
        if (sg_ag_can_stop_after_row(P, i, patternLen, scoreInit, bestL, bestLT, bestLP, bestG)) break;      // row pruning, see sg_ag.h
        unsigned *tmp = Hm; Hm = Hp; Hp = tmp;
    }
    __syncwarp();
    #undef SG_HB
    #undef SG_AB
    res->gScore = bestG; res->gText = bestGT; res->lScore = bestL; res->lText = bestLT; res->lPat = bestLP;
}
