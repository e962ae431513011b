// sg_ag_cigar.h -- affine-gap alignment with CIGAR output: scalar literal restatement of AffineGapVectorizedWithCigar
// (reference SNAPLib/AffineGapVectorized.cpp): computeGlobalScore (:159-518), computeGlobalScoreBanded (:520-943),
// computeFinalCigarString (:945-1041) and the dispatch computeGlobalScoreNormalized (:1043-1128) -- what SAMFormat::computeCigar
// (SAM.cpp:2470-2592) runs on every read that WAS rescored with affine gap.  Third piece of the output stage.
// STATUS: verified on the host against the compiled reference (tests/test_output_stage.py); no device entry point yet, nothing in
// include/snapgpu.h refers to this file.
//
// Like the scoring kernels (sg_ag.h) this keeps the reference's striped coordinates -- cell (vector j, SSE lane l) holds pattern
// column l * numVec + j -- and walks them in the same order, because the lazy-F loop's joint convergence test makes the
// traceback bits depend on that order.  Unlike them it is a GLOBAL alignment: rows and columns start at -(open + n * extend), the
// padding is INT16_MIN and the 16-bit arithmetic really saturates there, so every add / subtract below saturates like
// _mm_adds_epi16 / _mm_subs_epi16.
#pragma once
#include "sg_ag.h"
#include "sg_lv_cigar.h"

struct SgAgCigarScratch {
    int16_t *H, *Hm1, *E;            // [numVecMax * 8]
    int16_t *prof;                   // [5][numVecMax * 8]
    uint8_t *bt;                     // [rowsMax][numVecMax * 8] backtraceAction (values use bits 0,1,2,5)
    uint8_t *resAction;              // LocalCigarResult (:1430-1435), [resMax]
    int     *resCount;
    int      numVecMax, rowsMax, resMax;
};

struct SgAgCigarOut {
    int score;                       // computeFinalCigarString's edit count, -1 (no alignment) or -2 (operation buffer full)
    int nOps, netDel, tailIns;
};

// ---- The 8 x int16 vector of the reference (__m128i), in two bodies with one interface.
//   * host build: eight values, every operation a loop over them (what the CPU-side tests diff against the compiled reference);
//   * device build: ONE value per thread -- the eight threads of an aligned "octet" of a warp (lanes 8q .. 8q+7) are the eight SSE lanes of
//     one read's vectors (thread l holds element l), lane shifts are shuffles inside the octet and the lazy-F loop's joint "is any lane
//     still live" test is a vote over the octet.  Four reads per warp; everything that is not a vector operation (the traceback, the
//     heuristics, the text of the record) is executed identically by the eight threads of the octet, like the alignment kernels'
//     warp-uniform state machine.  The DP functions below are written once, against this interface.
#if defined(__CUDA_ARCH__)
struct SgV8 {
    int v;
    __device__ __forceinline__ static int lane() { return (int)(threadIdx.x & 7u); }
    __device__ __forceinline__ static unsigned mask() { return 0xffu << (threadIdx.x & 24u); }
    __device__ __forceinline__ static SgV8 splat(int x) { SgV8 r; r.v = x; return r; }
    template <class F> __device__ __forceinline__ static SgV8 gen(F f) { SgV8 r; r.v = f(lane()); return r; }                 // element l = f(l)
    template <class F> __device__ __forceinline__ void each(F f) const { f(lane(), v); }                                      // f(l, element l)
    __device__ __forceinline__ static SgV8 load16(const int16_t *p) { SgV8 r; r.v = p[lane()]; return r; }
    __device__ __forceinline__ static SgV8 load8(const uint8_t *p) { SgV8 r; r.v = p[lane()]; return r; }
    __device__ __forceinline__ void store16(int16_t *p) const { p[lane()] = (int16_t)v; }
    __device__ __forceinline__ void store8(uint8_t *p) const { p[lane()] = (uint8_t)v; }
    __device__ __forceinline__ SgV8 shiftUp(int fill) const { SgV8 r; const int up = __shfl_sync(mask(), v, (int)((threadIdx.x & 24u) | ((threadIdx.x + 7u) & 7u))); r.v = lane() == 0 ? fill : up; return r; }
    __device__ __forceinline__ int elem(int l) const { return __shfl_sync(mask(), v, (int)((threadIdx.x & 24u) | (unsigned)l)); }
    __device__ __forceinline__ bool any() const { return (__ballot_sync(mask(), v != 0) & mask()) != 0u; }
    __device__ __forceinline__ static void sync() { __syncwarp(mask()); }
    __device__ __forceinline__ static bool first() { return lane() == 0; }
};
#define SG_V8_OP(expr) { SgV8 r; { const int a = x.v, b = y.v; (void)a; (void)b; r.v = (expr); } return r; }
#else
struct SgV8 {
    int v[SG_VEC];
    static SgV8 splat(int x) { SgV8 r; for (int l = 0; l < SG_VEC; l++) r.v[l] = x; return r; }
    template <class F> static SgV8 gen(F f) { SgV8 r; for (int l = 0; l < SG_VEC; l++) r.v[l] = f(l); return r; }
    template <class F> void each(F f) const { for (int l = 0; l < SG_VEC; l++) f(l, v[l]); }
    static SgV8 load16(const int16_t *p) { SgV8 r; for (int l = 0; l < SG_VEC; l++) r.v[l] = p[l]; return r; }
    static SgV8 load8(const uint8_t *p) { SgV8 r; for (int l = 0; l < SG_VEC; l++) r.v[l] = p[l]; return r; }
    void store16(int16_t *p) const { for (int l = 0; l < SG_VEC; l++) p[l] = (int16_t)v[l]; }
    void store8(uint8_t *p) const { for (int l = 0; l < SG_VEC; l++) p[l] = (uint8_t)v[l]; }
    SgV8 shiftUp(int fill) const { SgV8 r; for (int l = SG_VEC - 1; l >= 1; l--) r.v[l] = v[l - 1]; r.v[0] = fill; return r; }
    int elem(int l) const { return v[l]; }
    bool any() const { for (int l = 0; l < SG_VEC; l++) if (v[l] != 0) return true; return false; }
    static void sync() {}
    static bool first() { return true; }
};
#define SG_V8_OP(expr) { SgV8 r; for (int l = 0; l < SG_VEC; l++) { const int a = x.v[l], b = y.v[l]; (void)a; (void)b; r.v[l] = (expr); } return r; }
#endif
SG_HD SgV8 sg_v8_adds(const SgV8 &x, const SgV8 &y) SG_V8_OP(sg_sat16(a + b))         // _mm_adds_epi16
SG_HD SgV8 sg_v8_subs(const SgV8 &x, const SgV8 &y) SG_V8_OP(sg_sat16(a - b))         // _mm_subs_epi16
SG_HD SgV8 sg_v8_max(const SgV8 &x, const SgV8 &y) SG_V8_OP(a > b ? a : b)
SG_HD SgV8 sg_v8_gt(const SgV8 &x, const SgV8 &y) SG_V8_OP(a > b ? 1 : 0)
SG_HD SgV8 sg_v8_or(const SgV8 &x, const SgV8 &y) SG_V8_OP(a | b)
SG_HD SgV8 sg_v8_bit(const SgV8 &x, const SgV8 &y) SG_V8_OP(a ? b : 0)                // x ? y : 0 (action bits under a comparison mask)

SG_HD int sg_agc_adds(int a, int b) { return sg_sat16(a + b); }

// computeFinalCigarString with format == BAM_CIGAR_OPS
SG_HD int sg_ag_cigar_final(const SgAgCigarScratch &S, const uint8_t *text, const uint8_t *pattern, int n_res, int min_i, uint32_t *ops, int maxOps,
                            bool useM, int *nOpsOut, int *netDel)
{
    int nEdits = 0, rowIdx = 0, colIdx = 0, nOps = 0;
    for (int i = n_res - 1; i >= min_i; --i) {
        const int cnt = S.resCount[i];
        if (S.resAction[i] == 1) {                                // D
            rowIdx += cnt; *netDel += cnt; nEdits += cnt;
            if (!sg_lvc_write(ops, maxOps, &nOps, cnt, SG_CIGAR_D)) return -2;
        } else if (S.resAction[i] == 2) {                         // I
            colIdx += cnt; nEdits += cnt;
            if (!sg_lvc_write(ops, maxOps, &nOps, cnt, SG_CIGAR_I)) return -2;
        } else if (S.resAction[i] == 0) {                         // M
            if (useM) {
                for (int j = 0; j < cnt; ++j) if (text[rowIdx + j] != pattern[colIdx + j]) nEdits++;
                if (!sg_lvc_write(ops, maxOps, &nOps, cnt, SG_CIGAR_M)) return -2;
            } else {
                int currentRunSize = 1;
                bool currentRunIsX = text[rowIdx] != pattern[colIdx];
                nEdits = currentRunIsX ? nEdits + 1 : nEdits;
                for (int j = 1; j < cnt; j++) {
                    if ((text[rowIdx + j] != pattern[colIdx + j]) == currentRunIsX) {
                        currentRunSize++;
                        if (text[rowIdx + j] != pattern[colIdx + j]) nEdits++;
                    } else {
                        if (!sg_lvc_write(ops, maxOps, &nOps, currentRunSize, currentRunIsX ? SG_CIGAR_X : SG_CIGAR_EQ)) return -2;
                        currentRunSize = 1;
                        currentRunIsX = !currentRunIsX;
                        nEdits = currentRunIsX ? nEdits + 1 : nEdits;
                    }
                }
                if (!sg_lvc_write(ops, maxOps, &nOps, currentRunSize, currentRunIsX ? SG_CIGAR_X : SG_CIGAR_EQ)) return -2;
            }
            rowIdx += cnt; colIdx += cnt;
        }
    }
    *nOpsOut = nOps;
    return nEdits;
}

// Traceback from (textUsed, patternLen - 1) through the action bits, the two "flip" heuristics and computeFinalCigarString: the common
// tail of computeGlobalScore (:371-512) and computeGlobalScoreBanded (:797-938).  The unbanded layout is the banded one with a
// single segment of numVec * 8 columns.
SG_HDN void sg_agc_finish(const SgAgCigarScratch &S, const uint8_t *text, const uint8_t *pattern, const uint8_t *quality, int patternLen, int textUsed,
                          int numVec, int segLen, int numSeg, uint32_t *ops, int maxOps, bool useM, SgAgCigarOut *out)
{
    // ---- traceback (:374-442) ----
    int n_res = 0;
    int rowIdx = textUsed, colIdx = patternLen - 1;
    int action = 0 /* M */, prevAction = 3 /* X */, actionCount = 1;
    while (rowIdx >= 0 && colIdx >= 0) {
        const int matrixIdx = action << 1;
        const int vecIdx = (colIdx / segLen) * numVec + ((colIdx % segLen) % numVec), elemIdx = (colIdx % segLen) / numVec;
        action = (S.bt[((size_t)rowIdx * numVec * numSeg + vecIdx) * SG_VEC + elemIdx] >> matrixIdx) & 3;
        if (action == 0) { rowIdx--; colIdx--; }
        else if (action == 1) { rowIdx--; }
        else { colIdx--; action = 2; }
        if (prevAction == action) {
            actionCount++;
        } else if (prevAction != 3) {
            if (n_res >= S.resMax) { out->score = -2; return; }
            S.resAction[n_res] = (uint8_t)prevAction; S.resCount[n_res] = actionCount; n_res++;
            actionCount = 1;
        }
        prevAction = action;
    }
    if (n_res + 3 > S.resMax) { out->score = -2; return; }
    if (prevAction == action) { S.resAction[n_res] = (uint8_t)prevAction; S.resCount[n_res] = actionCount; n_res++; }
    if (rowIdx >= 0) { S.resAction[n_res] = 1; S.resCount[n_res] = rowIdx + 1; n_res++; }
    if (colIdx >= 0) { S.resAction[n_res] = 2; S.resCount[n_res] = colIdx + 1; out->tailIns = colIdx + 1; n_res++; }

    // tail insertions, which the caller soft-clips (:444-452)
    int min_i = 0;
    if (S.resAction[0] == 2) { min_i = 1; out->tailIns = S.resCount[0]; }

    // The two heuristics below read-modify-write the shared result arrays: on the device ONE thread of the octet does them (the eight
    // run the rest of this function identically, storing the same values, which is harmless; an increment is not).
    SgV8::sync();
    if (SgV8::first()) {
        // "flip order of insertions followed by substitutions" (:454-476)
        rowIdx = 0; colIdx = 0;
        for (int i = n_res - 1; i >= min_i; --i) {
            if (S.resAction[i] == 0) { rowIdx += S.resCount[i]; colIdx += S.resCount[i]; }
            else if (S.resAction[i] == 1) { rowIdx += S.resCount[i]; }
            else {
                if (i > 0 && rowIdx < textUsed && colIdx < patternLen - 1) {
                    if ((pattern[colIdx + 1] == pattern[colIdx]) && (pattern[colIdx + 1] != text[rowIdx]) && (quality[colIdx] < 65)) {
                        if ((i + 1 <= n_res - 1) && S.resAction[i + 1] == 0 && S.resCount[i - 1] > 1) { S.resCount[i + 1] += 1; rowIdx++; colIdx++; }
                        if (S.resAction[i - 1] == 0 && S.resCount[i - 1] > 1) S.resCount[i - 1] -= 1;
                    }
                }
                colIdx += S.resCount[i];
            }
        }
        // "flip order of insertions and substitution with match in between" (:478-502)
        rowIdx = 0; colIdx = 0;
        for (int i = n_res - 1; i >= min_i; --i) {
            if (S.resAction[i] == 0) { rowIdx += S.resCount[i]; colIdx += S.resCount[i]; }
            else if (S.resAction[i] == 1) { rowIdx += S.resCount[i]; }
            else {
                if (i > 0 && rowIdx + 1 < textUsed && colIdx + S.resCount[i] < patternLen - 1) {
                    if ((pattern[colIdx + S.resCount[i]] == pattern[colIdx]) && (pattern[colIdx + S.resCount[i] + 1] != text[rowIdx + 1]) && (quality[colIdx] < 65)) {
                        if ((i + 1 <= n_res - 1) && S.resAction[i + 1] == 0 && S.resCount[i - 1] > 2) { S.resCount[i + 1] += 2; rowIdx += 2; colIdx += 2; }
                        if (S.resAction[i - 1] == 0 && S.resCount[i - 1] > 2) S.resCount[i - 1] -= 2;
                    }
                }
                colIdx += S.resCount[i];
            }
        }
    }
    SgV8::sync();
    out->score = sg_ag_cigar_final(S, text, pattern, n_res, min_i, ops, maxOps, useM, &out->nOps, &out->netDel);
}

// One DP vector step, shared by the main passes of the unbanded and the banded recurrence (:262-311 / :681-731): h = H of the diagonal
// neighbours, f = the running horizontal gap.  Returns the vector's action bits; updates E, f, Hm1.
SG_HD void sg_agc_vector_step(const int16_t *prow, int16_t *Eptr, int16_t *Hm1, const int16_t *Hptr, uint8_t *btRow, int idx8, SgV8 &h, SgV8 &f, const SgV8 &vOpen,
                              const SgV8 &vExt)
{
    const SgV8 m = sg_v8_adds(h, SgV8::load16(prow + idx8));
    SgV8 e = SgV8::load16(Eptr + idx8);
    SgV8 act = sg_v8_gt(e, m);                                                  // bit 0
    SgV8 hh = sg_v8_max(m, e);
    act = sg_v8_or(act, sg_v8_bit(sg_v8_gt(f, hh), SgV8::splat(2)));
    hh = sg_v8_max(hh, f);
    hh.store16(Hm1 + idx8);
    e = sg_v8_subs(e, vExt);
    const SgV8 temp = sg_v8_subs(m, vOpen);
    act = sg_v8_or(act, sg_v8_bit(sg_v8_gt(e, temp), SgV8::splat(4)));
    e = sg_v8_max(e, temp);
    e.store16(Eptr + idx8);
    SgV8 ff = sg_v8_subs(f, vExt);
    act = sg_v8_or(act, sg_v8_bit(sg_v8_gt(ff, temp), SgV8::splat(32)));
    f = sg_v8_max(ff, temp);
    act.store8(btRow + idx8);
    h = SgV8::load16(Hptr + idx8);
}

// One lazy-F vector step (:325-347 / :747-770): returns whether any lane is still live.
SG_HD bool sg_agc_lazy_step(int16_t *Hm1, uint8_t *btRow, int idx8, SgV8 &f, const SgV8 &vOpen, const SgV8 &vExt)
{
    SgV8 hh = SgV8::load16(Hm1 + idx8);
    SgV8 act = SgV8::load8(btRow + idx8);
    act = sg_v8_or(act, sg_v8_bit(sg_v8_gt(f, hh), SgV8::splat(2)));
    hh = sg_v8_max(hh, f);
    hh.store16(Hm1 + idx8);
    const SgV8 temp = sg_v8_subs(hh, vOpen);
    f = sg_v8_subs(f, vExt);
    const SgV8 live = sg_v8_gt(f, temp);
    act = sg_v8_or(act, sg_v8_bit(live, SgV8::splat(32)));
    act.store8(btRow + idx8);
    return live.any();
}

// AffineGapVectorizedWithCigar::computeGlobalScore with format == BAM_CIGAR_OPS.  P: sg_ag_params() of the scoring scheme
// (gapOpenPenalty = open + extend, subPenalty negative), as in the reference's init (:64-92).
SG_HDN void sg_ag_cigar_global(const SgAgParams &P, const SgAgCigarScratch &S, const uint8_t *text, int textLen, const uint8_t *pattern,
                               const uint8_t *quality, int patternLen, uint32_t *ops, int maxOps, bool useM, SgAgCigarOut *out)
{
    out->score = -1; out->nOps = 0; out->netDel = 0; out->tailIns = 0;
    if (text == (const uint8_t *)0) return;
    const int open = P.gapOpenPenalty, ext = P.gapExtendPenalty;
    const int numVec = (patternLen + SG_VEC - 1) / SG_VEC;
    const int stride = numVec * SG_VEC;
    if (numVec > S.numVecMax || textLen > S.rowsMax) { out->score = -2; return; }
    const SgV8 vOpen = SgV8::splat(open), vExt = SgV8::splat(ext);
    // query profile (:186-203) and first row (:222-239), striped: index j * 8 + l  <->  column l * numVec + j
    for (uint32_t t = 0; t < 5; t++) {
        for (int j = 0; j < numVec; j++) {
            SgV8::gen([&](int l) { const int k = l * numVec + j; return (k < patternLen) ? (int)sg_ag_sub(P, t, sg_base_value(pattern[k])) : -32768; })
                .store16(S.prof + t * stride + j * SG_VEC);
        }
    }
    for (int j = 0; j < numVec; j++) {
        SgV8::gen([&](int l) { const int k = l * numVec + j; return (k < patternLen) ? -(open + k * ext) : -32768; }).store16(S.H + j * SG_VEC);
        SgV8::splat(-32768).store16(S.E + j * SG_VEC);
    }
    SgV8::sync();
    int score = -32768, textUsed = -1;
    int16_t *Hptr = S.H, *Hm1 = S.Hm1;
    for (int i = 0; i < textLen; i++) {
        const int16_t *prow = S.prof + sg_base_value(text[i]) * stride;
        uint8_t *btRow = S.bt + (size_t)i * stride;
        SgV8 f = SgV8::splat(-32768);
        const int hInit = (i > 0) ? -(open + (i - 1) * ext) : 0;
        SgV8 h = SgV8::load16(Hptr + (numVec - 1) * SG_VEC).shiftUp((int)(int16_t)hInit);      // h << one lane, lane 0 <- hInit
        for (int j = 0; j < numVec; ++j) sg_agc_vector_step(prow, S.E, Hm1, Hptr, btRow, j * SG_VEC, h, f, vOpen, vExt);
        // lazy F (:317-349)
        bool converged = false;
        for (int k = 0; k < SG_VEC - 1 && !converged; k++) {
            f = f.shiftUp(-32768);
            for (int j = 0; j < numVec && !converged; j++) {
                if (!sg_agc_lazy_step(Hm1, btRow, j * SG_VEC, f, vOpen, vExt)) converged = true;
            }
        }
        SgV8::sync();
        const int g = Hm1[((patternLen - 1) % numVec) * SG_VEC + (patternLen - 1) / numVec];
        if (g >= score) { score = g; textUsed = i; }
        int16_t *tmp = Hm1; Hm1 = Hptr; Hptr = tmp;
    }
    if (!(score > -32768)) { out->score = -1; return; }
    sg_agc_finish(S, text, pattern, quality, patternLen, textUsed, numVec, numVec * SG_VEC, 1, ops, maxOps, useM, out);
}

// AffineGapVectorizedWithCigar::computeGlobalScoreBanded (:520-943) with format == BAM_CIGAR_OPS.  Despite the name it is the banded
// LOCAL-style recurrence of the scoring kernels shifted up by scoreInit (the caller passes MAX_READ_LENGTH) so that nothing goes
// negative: first row max(0, scoreInit - gap), E and the other H row zeroed, F carried between segments through X.  Vectors outside
// the band are never touched, so H / E keep what earlier rows left there, and the action array is never cleared: the traceback can
// step onto cells this call did not write and then reads what an EARLIER call left (same layout arithmetic) -- results are a
// function of the call history in those cases, exactly like the scoring kernels' traceback.
SG_HDN void sg_ag_cigar_banded(const SgAgParams &P, const SgAgCigarScratch &S, const uint8_t *text, int textLen, const uint8_t *pattern,
                               const uint8_t *quality, int patternLen, int w, int scoreInit, uint32_t *ops, int maxOps, bool useM, SgAgCigarOut *out)
{
    out->score = -1; out->nOps = 0; out->netDel = 0; out->tailIns = 0;
    if (w > SG_MAX_K - 1) w = SG_MAX_K - 1;
    if (text == (const uint8_t *)0) return;
    const int open = P.gapOpenPenalty, ext = P.gapExtendPenalty;
    const int bandWidth = (2 * w + 1) < patternLen ? (2 * w + 1) : patternLen;
    const int numVec = (bandWidth + SG_VEC - 1) / SG_VEC;
    const int segLen = numVec * SG_VEC;
    const int numSeg = (patternLen + segLen - 1) / segLen;
    const int stride = numVec * numSeg * SG_VEC;
    if (numVec * numSeg > S.numVecMax || textLen > S.rowsMax) { out->score = -2; return; }
    const SgV8 vOpen = SgV8::splat(open), vExt = SgV8::splat(ext);
    for (uint32_t t = 0; t < 5; t++) {
        for (int sgi = 0; sgi < numSeg; sgi++) for (int j = 0; j < numVec; j++) {
            SgV8::gen([&](int l) { const int idx = sgi * segLen + l * numVec + j; return (idx < patternLen) ? (int)sg_ag_sub(P, t, sg_base_value(pattern[idx])) : -32768; })
                .store16(S.prof + t * stride + (sgi * numVec + j) * SG_VEC);
        }
    }
    {   // first row (:611-627): scoreFirstRow[] is declared outside the loops and only assigned for columns inside the pattern, so a
        // padding lane repeats the value its lane had in the previous vector
        SgV8 scoreFirstRow = SgV8::splat(0);
        for (int sgi = 0; sgi < numSeg; sgi++) for (int j = 0; j < numVec; j++) {
            const SgV8 prev = scoreFirstRow;
            int lanePrev[SG_VEC];
            prev.each([&](int l, int v) { lanePrev[l] = v; });
            scoreFirstRow = SgV8::gen([&](int l) {
                const int idx = sgi * segLen + l * numVec + j;
                if (idx < patternLen) { const int v = scoreInit - (open + idx * ext); return (int)(int16_t)(v > 0 ? v : 0); }
                return lanePrev[l];
            });
            scoreFirstRow.store16(S.H + (sgi * numVec + j) * SG_VEC);
            SgV8::splat(0).store16(S.Hm1 + (sgi * numVec + j) * SG_VEC);
            SgV8::splat(0).store16(S.E + (sgi * numVec + j) * SG_VEC);
        }
    }
    SgV8::sync();
    int score = scoreInit, textUsed = -1;
    int16_t *Hptr = S.H, *Hm1 = S.Hm1;
    for (int i = 0; i < textLen; i++) {
        const int16_t *prow = S.prof + sg_base_value(text[i]) * stride;
        uint8_t *btRow = S.bt + (size_t)i * stride;
        SgV8 f = SgV8::splat(0);
        int X0 = 0;
        const int bandBeg = (i - w) > 0 ? (i - w) : 0;
        const int bandEnd = (i + w) < (patternLen - 1) ? (i + w) : (patternLen - 1);
        const int segBeg = bandBeg / segLen, segEnd = bandEnd / segLen;
        for (int j = segBeg; j <= segEnd; j++) {
            int hInit;
            if (j == 0) {
                hInit = scoreInit;
                if (i > 0) hInit = scoreInit - (open + (i - 1) * ext);
                hInit = (int16_t)hInit;
            } else {
                hInit = (bandBeg > j * segLen) ? 0 : (int)Hptr[(j * numVec - 1) * SG_VEC + (SG_VEC - 1)];
            }
            SgV8 h = SgV8::load16(Hptr + (j * numVec + numVec - 1) * SG_VEC).shiftUp(hInit);
            for (int k = 0; (k < numVec) && (j * segLen + k) <= bandEnd; k++) sg_agc_vector_step(prow, S.E, Hm1, Hptr, btRow, (j * numVec + k) * SG_VEC, h, f, vOpen, vExt);
            bool converged = false;
            for (int k = 0; k < SG_VEC - 1 && !converged; k++) {
                { const int f7 = f.elem(SG_VEC - 1); if (f7 > X0) X0 = f7; }          // X = max(X, f >> 7 lanes)
                f = f.shiftUp(0);
                for (int v = 0; (v < numVec) && (j * segLen + v) <= bandEnd && !converged; v++) {
                    if (!sg_agc_lazy_step(Hm1, btRow, (j * numVec + v) * SG_VEC, f, vOpen, vExt)) converged = true;
                }
            }
            { const int x0 = X0; f = SgV8::gen([&](int l) { return l == 0 ? x0 : 0; }); }      // f = X: (X0, 0, ..., 0)
        }
        SgV8::sync();
        if (bandEnd == patternLen - 1) {
            const int vecIdx = (bandEnd / segLen) * numVec + ((bandEnd % segLen) % numVec), elemIdx = (bandEnd % segLen) / numVec;
            const int g = Hm1[vecIdx * SG_VEC + elemIdx];
            if (g > score) { score = g; textUsed = i; }
        }
        int16_t *tmp = Hm1; Hm1 = Hptr; Hptr = tmp;
    }
    if (!(score > 0)) { out->score = -1; return; }
    sg_agc_finish(S, text, pattern, quality, patternLen, textUsed, numVec, segLen, numSeg, ops, maxOps, useM, out);
}

// AffineGapVectorizedWithCigar::computeGlobalScoreNormalized (:1043-1128) for format == BAM_CIGAR_OPS: the banded form for patterns of
// at least 3 * (2k + 1) columns, falling back to the unbanded one when it finds nothing usable; then the front-clipping verdict
// (a leading deletion sends the caller back with a new start: returns 0; a leading insertion is reported as a negative adjustment).
SG_HD int sg_ag_cigar_normalized(const SgAgParams &P, const SgAgCigarScratch &S, const uint8_t *text, int textLen, const uint8_t *pattern,
                                 const uint8_t *quality, int patternLen, int k, uint32_t *ops, int maxOps, bool useM, SgAgCigarOut *out, int *addFrontClipping)
{
    if (patternLen >= (3 * (2 * k + 1))) {
        sg_ag_cigar_banded(P, S, text, textLen, pattern, quality, patternLen, k, 1000 /* MAX_READ_LENGTH */, ops, maxOps, useM, out);
        if (out->score < 0 || out->score > k || out->tailIns >= patternLen) {
            sg_ag_cigar_global(P, S, text, textLen, pattern, quality, patternLen, ops, maxOps, useM, out);
        }
    } else {
        sg_ag_cigar_global(P, S, text, textLen, pattern, quality, patternLen, ops, maxOps, useM, out);
    }
    if (out->score < 0) return out->score;
    if (addFrontClipping) {
        const uint32_t first = ops[0] & 0xfu;       // (with no operations at all the reference reads whatever is in its buffer; see the test)
        if (out->nOps > 0 && first == SG_CIGAR_D) {
            *addFrontClipping = (int)(ops[0] >> 4);
            if (*addFrontClipping != 0) return 0;
        } else if (out->nOps > 0 && first == SG_CIGAR_I) {
            *addFrontClipping = -(int)(ops[0] >> 4);
        } else {
            *addFrontClipping = 0;
        }
    }
    return out->score;
}
