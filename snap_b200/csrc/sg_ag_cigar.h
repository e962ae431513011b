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
    out->score = sg_ag_cigar_final(S, text, pattern, n_res, min_i, ops, maxOps, useM, &out->nOps, &out->netDel);
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
    // query profile (:186-203) and first row (:222-239), striped: index j * 8 + l  <->  column l * numVec + j
    for (uint32_t t = 0; t < 5; t++) {
        for (int j = 0; j < numVec; j++) {
            for (int l = 0; l < SG_VEC; l++) {
                const int k = l * numVec + j;
                S.prof[t * stride + j * SG_VEC + l] = (k < patternLen) ? (int16_t)sg_ag_sub(P, t, sg_base_value(pattern[k])) : (int16_t)-32768;
            }
        }
    }
    for (int j = 0; j < numVec; j++) {
        for (int l = 0; l < SG_VEC; l++) {
            const int k = l * numVec + j;
            S.H[j * SG_VEC + l] = (k < patternLen) ? (int16_t)(-(open + k * ext)) : (int16_t)-32768;
            S.E[j * SG_VEC + l] = (int16_t)-32768;
        }
    }
    int score = -32768, textUsed = -1;
    int16_t *Hptr = S.H, *Hm1 = S.Hm1;
    for (int i = 0; i < textLen; i++) {
        const int16_t *prow = S.prof + sg_base_value(text[i]) * stride;
        uint8_t *btRow = S.bt + (size_t)i * stride;
        int f[SG_VEC], h[SG_VEC];
        for (int l = 0; l < SG_VEC; l++) f[l] = -32768;
        const int hInit = (i > 0) ? -(open + (i - 1) * ext) : 0;
        for (int l = SG_VEC - 1; l >= 1; l--) h[l] = Hptr[(numVec - 1) * SG_VEC + l - 1];      // h << one lane, lane 0 <- hInit
        h[0] = (int16_t)hInit;
        for (int j = 0; j < numVec; ++j) {
            for (int l = 0; l < SG_VEC; l++) {
                const int idx = j * SG_VEC + l;
                const int m = sg_agc_adds(h[l], prow[idx]);
                int e = S.E[idx];
                int act = (e > m) ? 1 : 0;
                int hh = m > e ? m : e;
                if (f[l] > hh) act |= 2;
                if (f[l] > hh) hh = f[l];
                Hm1[idx] = (int16_t)hh;
                e = sg_sat16(e - ext);
                const int temp = sg_sat16(m - open);
                if (e > temp) act |= 4;
                if (temp > e) e = temp;
                S.E[idx] = (int16_t)e;
                int ff = sg_sat16(f[l] - ext);
                if (ff > temp) act |= 32;
                if (temp > ff) ff = temp;
                f[l] = ff;
                btRow[idx] = (uint8_t)act;
                h[l] = Hptr[idx];
            }
        }
        // lazy F (:317-349)
        bool converged = false;
        for (int k = 0; k < SG_VEC - 1 && !converged; k++) {
            for (int l = SG_VEC - 1; l >= 1; l--) f[l] = f[l - 1];
            f[0] = -32768;
            for (int j = 0; j < numVec && !converged; j++) {
                bool any = false;
                for (int l = 0; l < SG_VEC; l++) {
                    const int idx = j * SG_VEC + l;
                    int hh = Hm1[idx];
                    int act = btRow[idx];
                    if (f[l] > hh) { act |= 2; hh = f[l]; }
                    Hm1[idx] = (int16_t)hh;
                    const int temp = sg_sat16(hh - open);
                    f[l] = sg_sat16(f[l] - ext);
                    if (f[l] > temp) { act |= 32; any = true; }
                    btRow[idx] = (uint8_t)act;
                }
                if (!any) converged = true;
            }
        }
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
    for (uint32_t t = 0; t < 5; t++) {
        for (int sgi = 0; sgi < numSeg; sgi++) for (int j = 0; j < numVec; j++) for (int l = 0; l < SG_VEC; l++) {
            const int idx = sgi * segLen + l * numVec + j;
            S.prof[t * stride + (sgi * numVec + j) * SG_VEC + l] = (idx < patternLen) ? (int16_t)sg_ag_sub(P, t, sg_base_value(pattern[idx])) : (int16_t)-32768;
        }
    }
    {   // first row (:611-627): scoreFirstRow[] is declared outside the loops and only assigned for columns inside the pattern, so a
        // padding lane repeats the value its lane had in the previous vector
        int16_t scoreFirstRow[SG_VEC] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int sgi = 0; sgi < numSeg; sgi++) for (int j = 0; j < numVec; j++) {
            for (int l = 0; l < SG_VEC; l++) {
                const int idx = sgi * segLen + l * numVec + j;
                if (idx < patternLen) { int v = scoreInit - (open + idx * ext); scoreFirstRow[l] = (int16_t)(v > 0 ? v : 0); }
            }
            for (int l = 0; l < SG_VEC; l++) {
                S.H[(sgi * numVec + j) * SG_VEC + l] = scoreFirstRow[l];
                S.Hm1[(sgi * numVec + j) * SG_VEC + l] = 0;
                S.E[(sgi * numVec + j) * SG_VEC + l] = 0;
            }
        }
    }
    int score = scoreInit, textUsed = -1;
    int16_t *Hptr = S.H, *Hm1 = S.Hm1;
    for (int i = 0; i < textLen; i++) {
        const int16_t *prow = S.prof + sg_base_value(text[i]) * stride;
        uint8_t *btRow = S.bt + (size_t)i * stride;
        int f[SG_VEC], h[SG_VEC];
        for (int l = 0; l < SG_VEC; l++) f[l] = 0;
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
            for (int l = SG_VEC - 1; l >= 1; l--) h[l] = Hptr[(j * numVec + numVec - 1) * SG_VEC + l - 1];
            h[0] = hInit;
            for (int k = 0; (k < numVec) && (j * segLen + k) <= bandEnd; k++) {
                for (int l = 0; l < SG_VEC; l++) {
                    const int idx = (j * numVec + k) * SG_VEC + l;
                    const int m = sg_agc_adds(h[l], prow[idx]);
                    int e = S.E[idx];
                    int act = (e > m) ? 1 : 0;
                    int hh = m > e ? m : e;
                    if (f[l] > hh) { act |= 2; hh = f[l]; }
                    Hm1[idx] = (int16_t)hh;
                    e = sg_sat16(e - ext);
                    const int temp = sg_sat16(m - open);
                    if (e > temp) act |= 4;
                    if (temp > e) e = temp;
                    S.E[idx] = (int16_t)e;
                    int ff = sg_sat16(f[l] - ext);
                    if (ff > temp) act |= 32;
                    if (temp > ff) ff = temp;
                    f[l] = ff;
                    btRow[idx] = (uint8_t)act;
                    h[l] = Hptr[idx];
                }
            }
            bool converged = false;
            for (int k = 0; k < SG_VEC - 1 && !converged; k++) {
                if (f[SG_VEC - 1] > X0) X0 = f[SG_VEC - 1];          // X = max(X, f >> 7 lanes)
                for (int l = SG_VEC - 1; l >= 1; l--) f[l] = f[l - 1];
                f[0] = 0;
                for (int v = 0; (v < numVec) && (j * segLen + v) <= bandEnd && !converged; v++) {
                    bool any = false;
                    for (int l = 0; l < SG_VEC; l++) {
                        const int idx = (j * numVec + v) * SG_VEC + l;
                        int hh = Hm1[idx];
                        int act = btRow[idx];
                        if (f[l] > hh) { act |= 2; hh = f[l]; }
                        Hm1[idx] = (int16_t)hh;
                        const int temp = sg_sat16(hh - open);
                        f[l] = sg_sat16(f[l] - ext);
                        if (f[l] > temp) { act |= 32; any = true; }
                        btRow[idx] = (uint8_t)act;
                    }
                    if (!any) converged = true;
                }
            }
            f[0] = X0;                                               // f = X: (X0, 0, ..., 0)
            for (int l = 1; l < SG_VEC; l++) f[l] = 0;
        }
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
