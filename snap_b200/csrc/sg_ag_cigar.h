// sg_ag_cigar.h -- affine-gap global alignment with CIGAR output: scalar literal restatement of
// AffineGapVectorizedWithCigar::computeGlobalScore (reference SNAPLib/AffineGapVectorized.cpp:159-518) and
// computeFinalCigarString (:945-1041), the unbanded form SAMFormat::computeCigar (SAM.cpp:2470-2592) falls back to (and uses
// outright for short patterns) on every read that WAS rescored with affine gap.  Third piece of the output stage (SURVEY 8f row
// N1).  STATUS: verified on the host against the compiled reference (tests/test_lv_cigar.py); the banded form
// (computeGlobalScoreBanded, :520-943) and the dispatch around both (computeGlobalScoreNormalized, :1043-1128) are NOT restated
// yet; no device entry point, nothing in include/snapgpu.h refers to this file.
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

    // ---- traceback (:374-442) ----
    int n_res = 0;
    int rowIdx = textUsed, colIdx = patternLen - 1;
    int action = 0 /* M */, prevAction = 3 /* X */, actionCount = 1;
    while (rowIdx >= 0 && colIdx >= 0) {
        const int matrixIdx = action << 1;
        const int stripedColIdx = (colIdx % numVec) * SG_VEC + (colIdx / numVec);
        action = (S.bt[(size_t)rowIdx * stride + stripedColIdx] >> matrixIdx) & 3;
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
