// sg_lv_cigar.h -- Landau-Vishkin with CIGAR output: literal restatement of LandauVishkinWithCigar::computeEditDistance
// (reference SNAPLib/LandauVishkin.cpp:141-505) and computeEditDistanceNormalized (:507-650), the routine SAMFormat::computeCigar
// (SAM.cpp:2354-2468) runs for every aligned read that was NOT rescored with affine gap.  First piece of the output stage
// (SURVEY 8f row N1).  STATUS: verified on the host against the compiled reference (tests/test_output_stage.py, incl. the 30 known
// answers of the reference's tests/LandauVishkinTest.cpp:34-129); no device entry point yet, nothing in include/snapgpu.h refers to it.
//
// Output is BAM cigar operations, (count << 4) | code with codes "MIDNSHP=X" (Bam.cpp:268) -- the form
// computeEditDistanceNormalized itself works on; the SAM text form is a decode of it (BAMAlignment::decodeCigar).
//
// What makes this different from the scoring LV (sg_lv.h): the order in which the diagonals of one error level are visited
// (0, -1, 1, -2, 2, ...: few-indel answers first), PrevDelta's preference order with ties broken on the TOTAL number of indels
// carried along each path, the early answer when a zero-indel path reaches the end, the "straight mismatches" shortcut that
// discards the DP path when the same edit distance is reachable without indels, and the backtrace that merges repeated actions.
#pragma once
#include "sg_common.h"

#define SG_CIGAR_M 0u
#define SG_CIGAR_I 1u
#define SG_CIGAR_D 2u
#define SG_CIGAR_S 4u
#define SG_CIGAR_EQ 7u
#define SG_CIGAR_X 8u
#define SG_LVC_BUFFER_FULL (-2)      // computeEditDistance's "ran out of space in cigarBuf"

struct SgLvCigarScratch {            // LandauVishkinWithCigar's members, sized for error levels 0..kmax
    int     *L;                      // [(kmax+1) * (2*kmax+3)]
    int     *totalIndels;            // same shape
    uint8_t *A;                      // same shape: 'D' / 'X' / 'I'
    uint8_t *btAction;               // [kmax+1]
    int     *btMatched, *btD;        // [kmax+1]
    int      kmax;
    SG_HD int at(int e, int d) const { return e * (2 * kmax + 3) + d + kmax + 1; }
    // (a diagonal outside [-e, e] of row e is never written by the reference, in any call: it keeps the constructor's -2)
    SG_HD int getL(int e, int d) const { return (d < -e || d > e) ? -2 : L[at(e, d)]; }
};

SG_HD size_t sg_lv_cigar_scratch_ints(int kmax) { return (size_t)(kmax + 1) * (size_t)(2 * kmax + 3); }

struct SgLvCigarOut {
    int score;                       // edit distance, SG_SCORE_ABOVE_LIMIT, or SG_LVC_BUFFER_FULL
    int nOps;                        // BAM operations written
    int textUsed;                    // o_textUsed
    int netIndel;                    // o_netIndel (deletions minus insertions)
};

SG_HD bool sg_lvc_write(uint32_t *ops, int maxOps, int *nOps, int count, uint32_t code)   // writeCigar, BAM_CIGAR_OPS case (:128-135)
{
    if (count <= 0) return true;
    if (*nOps >= maxOps || count >= (1 << 28)) return false;
    ops[(*nOps)++] = ((uint32_t)count << 4) | code;
    return true;
}

// first index in [from, end) at which pattern and text (shifted by d) differ, or end
SG_HD int sg_lvc_extend(const uint8_t *pattern, const uint8_t *text, int d, int from, int end)
{
    int i = from;
    while (i < end && pattern[i] == text[d + i]) i++;
    return i;
}

// LandauVishkinWithCigar::computeEditDistance with format == BAM_CIGAR_OPS.  k <= S.kmax.
SG_HDN void sg_lv_cigar_compute(const SgLvCigarScratch &S, const uint8_t *text, int textLen, const uint8_t *pattern, int patternLen, int k,
                                uint32_t *ops, int maxOps, bool useM, SgLvCigarOut *out)
{
    out->nOps = 0; out->textUsed = 0; out->netIndel = 0; out->score = SG_SCORE_ABOVE_LIMIT;
    if (text == (const uint8_t *)0) return;                        // :165-167
    int nOps = 0;
    const int end = patternLen < textLen ? patternLen : textLen;
    const int L00 = sg_lvc_extend(pattern, text, 0, 0, end);      // :169-184 (8 bytes at a time there; clamped to end)
    S.L[S.at(0, 0)] = L00;
    S.totalIndels[S.at(0, 0)] = 0;
    if (L00 == end) {                                             // :185-210: exact match
        bool ok;
        if (useM) ok = sg_lvc_write(ops, maxOps, &nOps, patternLen, SG_CIGAR_M);
        else {
            ok = sg_lvc_write(ops, maxOps, &nOps, end, SG_CIGAR_EQ);
            if (ok && patternLen > end) ok = sg_lvc_write(ops, maxOps, &nOps, patternLen - end, SG_CIGAR_X);
        }
        out->nOps = nOps;
        if (!ok) { out->score = SG_LVC_BUFFER_FULL; return; }
        out->textUsed = end; out->score = 0;
        return;
    }

    int e;
    int lastBestIndels = SG_MAX_K + 1, lastBestD = SG_MAX_K + 1, lastBestBest = 0;
    bool gotAnswer = false;
    for (e = 1; e <= k && !gotAnswer; e++) {
        // d = 0, -1, 1, -2, 2, ..., -e, e (:221)
        for (int d = 0; d != -(e + 1); d = (d >= 0 ? -(d + 1) : -d)) {
            int bestdelta = 0, bestbest = -1, bestBestIndels = SG_MAX_K + 1;
            const int dy = (d >= 0) + (d > 0);
            for (int dx = 0; dx < 3; dx++) {
                // PrevDelta (:66-69): {0,+1,-1} for d <= 0, {0,-1,+1} for d > 0
                const int delta = dx == 0 ? 0 : (dy == 2 ? (dx == 1 ? -1 : 1) : (dx == 1 ? 1 : -1));
                int best = S.getL(e - 1, d + delta) + (delta >= 0);
                if (best < 0) continue;
                const int bestIndels = S.totalIndels[S.at(e - 1, d + delta)] + (delta != 0);
                if (pattern[best] == text[d + best]) {            // (the reference reads these even at index patternLen; so do we: both
                    const int endd = patternLen < textLen - d ? patternLen : textLen - d;      //  buffers are padded by their callers)
                    int b = sg_lvc_extend(pattern, text, d, best, endd);
                    // the reference's loop compares 8 bytes at a time starting at `best` and clamps to endd; when best >= endd
                    // already, it still takes min(first difference, endd) = endd
                    best = (best >= endd) ? endd : b;
                }
                if (best > bestbest || (best == bestbest && bestIndels < bestBestIndels)) {
                    bestbest = best; bestdelta = delta; bestBestIndels = bestIndels;
                }
            }
            S.A[S.at(e, d)] = (uint8_t)("DXI"[bestdelta + 1]);
            S.L[S.at(e, d)] = bestbest;
            S.totalIndels[S.at(e, d)] = bestBestIndels;
            if (bestbest == patternLen) {
                if (bestBestIndels == 0) { lastBestIndels = 0; lastBestD = d; lastBestBest = bestbest; gotAnswer = true; break; }
                if ((lastBestIndels < 0 ? -lastBestIndels : lastBestIndels) > bestBestIndels) {
                    lastBestIndels = bestBestIndels; lastBestD = d; lastBestBest = bestbest;
                }
            }
        }
        if (gotAnswer) break;
        if (lastBestD != SG_MAX_K + 1) { gotAnswer = true; break; }
    }
    if (!gotAnswer) { out->score = SG_SCORE_ABOVE_LIMIT; out->nOps = 0; return; }      // :286-288

    // ---- got_answer (:290-): can e errors be had with no indels at all? ----
    int straightMismatches = 0;
    for (int i = 0; i < end; i++) if (pattern[i] != text[i]) straightMismatches++;
    straightMismatches += patternLen - end;
    bool ok = true;
    if (straightMismatches == e) {
        if (useM) {
            ok = sg_lvc_write(ops, maxOps, &nOps, patternLen, SG_CIGAR_M);
        } else {
            int streakStart = 0;
            bool matching = (pattern[0] == text[0]);
            for (int i = 0; i < end && ok; i++) {
                const bool newMatching = (pattern[i] == text[i]);
                if (newMatching != matching) {
                    ok = sg_lvc_write(ops, maxOps, &nOps, i - streakStart, matching ? SG_CIGAR_EQ : SG_CIGAR_X);
                    matching = newMatching;
                    streakStart = i;
                }
            }
            if (ok && patternLen > streakStart) {
                if (!matching) {
                    ok = sg_lvc_write(ops, maxOps, &nOps, patternLen - streakStart, SG_CIGAR_X);
                } else {
                    ok = sg_lvc_write(ops, maxOps, &nOps, end - streakStart, SG_CIGAR_EQ);
                    if (ok && patternLen > end) ok = sg_lvc_write(ops, maxOps, &nOps, patternLen - end, SG_CIGAR_X);
                }
            }
        }
        out->nOps = nOps;
        if (!ok) { out->score = SG_LVC_BUFFER_FULL; return; }
        out->textUsed = end; out->score = e;
        return;
    }

    // ---- trace back (:392-413) ----
    int curD = lastBestD;
    for (int curE = e; curE >= 1; curE--) {
        const uint8_t a = S.A[S.at(curE, curD)];
        S.btAction[curE] = a;
        if (a == 'I') {
            S.btD[curE] = curD + 1;
            S.btMatched[curE] = S.L[S.at(curE, curD)] - S.getL(curE - 1, curD + 1) - 1;
        } else if (a == 'D') {
            S.btD[curE] = curD - 1;
            S.btMatched[curE] = S.L[S.at(curE, curD)] - S.getL(curE - 1, curD - 1);
        } else {
            S.btD[curE] = curD;
            S.btMatched[curE] = S.L[S.at(curE, curD)] - S.getL(curE - 1, curD) - 1;
        }
        curD = S.btD[curE];
    }
    int accumulatedMs = 0;
    if (useM) accumulatedMs = L00;
    else if (L00 > 0) ok = sg_lvc_write(ops, maxOps, &nOps, L00, SG_CIGAR_EQ);
    int curE = 1;
    while (curE <= e && ok) {
        const uint8_t action = S.btAction[curE];
        int actionCount = 1;
        while (curE + 1 <= e && S.btMatched[curE] == 0 && S.btAction[curE + 1] == action) { actionCount++; curE++; }
        if (action == 'I') out->netIndel -= actionCount;
        else if (action == 'D') out->netIndel += actionCount;
        const uint32_t code = action == 'I' ? SG_CIGAR_I : action == 'D' ? SG_CIGAR_D : SG_CIGAR_X;
        if (useM) {
            if (action == 'X') {                                   // (:448: `action == '=' || action == 'X'`; '=' never occurs)
                accumulatedMs += actionCount;
            } else {
                if (accumulatedMs != 0) { ok = sg_lvc_write(ops, maxOps, &nOps, accumulatedMs, SG_CIGAR_M); accumulatedMs = 0; }
                if (ok) ok = sg_lvc_write(ops, maxOps, &nOps, actionCount, code);
            }
        } else {
            ok = sg_lvc_write(ops, maxOps, &nOps, actionCount, code);
        }
        if (ok && S.btMatched[curE] > 0) {
            if (useM) accumulatedMs += S.btMatched[curE];
            else ok = sg_lvc_write(ops, maxOps, &nOps, S.btMatched[curE], SG_CIGAR_EQ);
        }
        curE++;
    }
    if (ok && useM && accumulatedMs != 0) ok = sg_lvc_write(ops, maxOps, &nOps, accumulatedMs, SG_CIGAR_M);
    out->nOps = nOps;
    if (!ok) { out->score = SG_LVC_BUFFER_FULL; return; }
    out->textUsed = textLen < lastBestBest + lastBestD ? textLen : lastBestBest + lastBestD;
    out->score = e;
}

// LandauVishkinWithCigar::computeEditDistanceNormalized (:507-650) for format == BAM_CIGAR_OPS: the same operations, plus the
// front-clipping verdict of SAMFormat's retry protocol: a leading deletion asks the caller to move the alignment start and run
// again (returns 0 with *addFrontClipping = its length), a leading insertion is reported as a negative adjustment.
SG_HD int sg_lv_cigar_normalized(const SgLvCigarScratch &S, const uint8_t *text, int textLen, const uint8_t *pattern, int patternLen, int k,
                                 uint32_t *ops, int maxOps, bool useM, SgLvCigarOut *out, int *addFrontClipping)
{
    sg_lv_cigar_compute(S, text, textLen, pattern, patternLen, k, ops, maxOps, useM, out);
    if (out->score < 0) return out->score;
    if (addFrontClipping) {
        const uint32_t first = ops[0] & 0xfu;
        if (first == SG_CIGAR_D) {
            *addFrontClipping = (int)(ops[0] >> 4);
            if (*addFrontClipping != 0) return 0;
        } else if (first == SG_CIGAR_I) {
            *addFrontClipping = -(int)(ops[0] >> 4);
        } else {
            *addFrontClipping = 0;
        }
    }
    return out->score;
}
