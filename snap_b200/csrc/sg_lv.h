// sg_lv.h -- Landau-Vishkin bounded-k edit distance with match-probability backtrace.  Scalar form, host+device.
// Restates LandauVishkin<TEXT_DIRECTION>::computeEditDistance (reference SNAPLib/LandauVishkin.h:100-351)
// including its d visiting order (0,1,-1,2,-2,...), strict '>' tie-breaks, 'X'-finish preference and the
// order of the floating-point multiplies.
#pragma once
#include "sg_common.h"

struct SgLvResult {
    int score;            // edit distance or SG_SCORE_ABOVE_LIMIT
    int netIndel;         // o_netIndel
    int totalIndels;
    int textSpan;
    double matchProbability;
};

// L(e,d) storage: row e holds d in [-e, e]; cells outside were never written in the reference and hold -2
// (LandauVishkin.h:55), which we return analytically.
struct SgLvState {
    int16_t *L; uint8_t *A; int stride; int kmax;
    SG_HD int getL(int e, int d) const { return (d < -e || d > e) ? -2 : (int)L[e * stride + d + kmax]; }
    SG_HD void setL(int e, int d, int v) { L[e * stride + d + kmax] = (int16_t)v; }
    SG_HD uint8_t getA(int e, int d) const { return A[e * stride + d + kmax]; }
    SG_HD void setA(int e, int d, uint8_t v) { A[e * stride + d + kmax] = v; }
};

// countPerfectMatch (LandauVishkin.h:377-407): number of equal leading characters of pattern[pi..] and the text
// read in direction dir starting at text index ti, clamped to availBytes (which may be <= 0: the reference then
// returns availBytes itself because the first characters were already known equal).
// lane >= 0 (device, all 32 lanes converged with identical arguments): 32 characters are compared per step and the first
// mismatch found with a ballot.  lane < 0: scalar.
SG_HD int sg_lv_cpm(const uint8_t *pattern, int pi, const uint8_t *text, int ti, int dir, int availBytes, int lane)
{
    if (availBytes <= 0) return availBytes;
#if defined(__CUDA_ARCH__)
    if (lane >= 0) {
        #pragma unroll 1
        for (int n = 0; n < availBytes; n += 32) {
            int i = n + lane;
            bool mism = (i < availBytes) ? (pattern[pi + i] != text[(ti + i) * dir]) : true;
            unsigned m = __ballot_sync(0xffffffffu, mism);
            if (m != 0) {
                int r = n + (__ffs(m) - 1);
                return r < availBytes ? r : availBytes;
            }
        }
        return availBytes;
    }
#endif
    int n = 0;
    #pragma unroll 1
    while (n < availBytes && pattern[pi + n] == text[(ti + n) * dir]) n++;
    return n;
}

// text: as the reference's callers pass it (for dir == -1: one past the first text character).
SG_HDN void sg_lv_compute(const SgTables &T, const SgScratch &S, int dir, const uint8_t *text, int textLen,
                          const uint8_t *pattern, const uint8_t *quality, int patternLen, int k, SgLvResult *out, int lane = -1)
{
    out->netIndel = 0; out->totalIndels = 0; out->textSpan = 0; out->matchProbability = 0.0;
    if (k < 0) { out->score = SG_SCORE_ABOVE_LIMIT; return; }
    if (k > SG_MAX_K - 1) k = SG_MAX_K - 1;
    if (text == (const uint8_t *)0) { out->matchProbability = 0.0; out->score = SG_SCORE_ABOVE_LIMIT; return; }
    out->matchProbability = 1.0;
    if (dir == -1) text--;

    SgLvState st;
    const bool smallLA = (uint32_t)((k + 1) * (2 * k + 1)) <= S.lvSmallCells;      // the cells of this call fit the shared-memory copy
    st.L = smallLA ? S.lvLs : S.lvL; st.A = smallLA ? S.lvAs : S.lvA; st.kmax = k; st.stride = 2 * k + 1;

    int end = patternLen < textLen ? patternLen : textLen;
    int L00 = end > 0 ? sg_lv_cpm(pattern, 0, text, 0, dir, end, lane) : 0;      // countPerfectMatch(p, t, end) with end >= 0
    st.setL(0, 0, L00);

    if (L00 == end) {
        int result = (patternLen > end ? patternLen - end : 0);
        out->matchProbability = T.perfect[patternLen];
        if (result > k) { out->score = SG_SCORE_ABOVE_LIMIT; return; }
        out->textSpan += patternLen;
        out->score = result;
        return;
    }

    int lastBestD = SG_MAX_K + 1;
    int e;
    bool gotAnswer = false;
    for (e = 1; e <= k; e++) {
        // d = 0, 1, -1, 2, -2, ..., e, -e
        int d = 0;
        for (int i = 0; d != e + 1; i++, d = (d > 0 ? -d : -d + 1)) {
            int endd = patternLen < textLen - d ? patternLen : textLen - d;
            // the three ways into L(e,d): 'X' up from (e-1,d) +1, 'D' from (e-1,d-1), 'I' from (e-1,d+1) +1; each is extended
            // along the diagonal by countPerfectMatch if its first characters agree (the reference compares *p == *t even
            // at index patternLen, where the extension is min(.., 0) = 0).  One loop body instead of three copies.
            int best = -3;
            uint8_t a = 'X';
            #pragma unroll 1
            for (int way = 0; way < 3; way++) {
                int start = (way == 0) ? st.getL(e - 1, d) + 1 : (way == 1) ? st.getL(e - 1, d - 1) : st.getL(e - 1, d + 1) + 1;
                if (start >= 0 && endd != start) {
                    if (pattern[start] == text[(d + start) * dir]) {
                        start += (endd - start > 0) ? sg_lv_cpm(pattern, start, text, d + start, dir, endd - start, lane) : (endd - start);
                    }
                }
                if (way == 0) { best = start; }
                else if (start > best) { best = start; a = (way == 1) ? 'D' : 'I'; }
            }
            st.setA(e, d, a);

            if (best == patternLen) {
                if (a == 'X') {
                    lastBestD = d;
                    gotAnswer = true;      // goto got_answer: L(e,d) is NOT written on this path
                    break;
                } else {
                    int ad = d < 0 ? -d : d, al = lastBestD < 0 ? -lastBestD : lastBestD;
                    if (ad < al) lastBestD = d;
                }
            }
            st.setL(e, d, best);
        }
        if (gotAnswer) break;
        if (SG_MAX_K + 1 != lastBestD) break;
    }

    if (SG_MAX_K + 1 == lastBestD) { out->score = SG_SCORE_ABOVE_LIMIT; return; }
    // note: when the e-loop ran to completion without an answer e == k+1 and we returned above

    // Backtrace (LandauVishkin.h:286-304).  On the 'X' finish path L(e, lastBestD) was not written: the reference
    // then reads a stale cell into backtraceMatched[e], which only feeds `offset` after its last use.  We use 0.
    const bool smallBt = (uint32_t)(k + 2) <= S.lvBtSmall;
    int16_t *btMatched = smallBt ? S.lvBtMatchedS : S.lvBtMatched; int16_t *btD = smallBt ? S.lvBtDS : S.lvBtD;
    uint8_t *btAction = smallBt ? S.lvBtActionS : S.lvBtAction;
    int curD = lastBestD;
    for (int curE = e; curE >= 1; curE--) {
        uint8_t a = st.getA(curE, curD);
        btAction[curE] = a;
        int Lcur = (gotAnswer && curE == e) ? 0 : st.getL(curE, curD);
        if (a == 'I') {
            btD[curE] = (int16_t)(curD + 1);
            btMatched[curE] = (int16_t)(Lcur - st.getL(curE - 1, curD + 1) - 1);
        } else if (a == 'D') {
            btD[curE] = (int16_t)(curD - 1);
            btMatched[curE] = (int16_t)(Lcur - st.getL(curE - 1, curD - 1));
        } else {
            btD[curE] = (int16_t)curD;
            btMatched[curE] = (int16_t)(Lcur - st.getL(curE - 1, curD) - 1);
        }
        curD = btD[curE];
    }

    double mp = 1.0;
    int curE = 1;
    int offset = L00;
    int netIndel = 0, totalIndels = 0, textSpan = 0;
    while (curE <= e) {
        uint8_t action = btAction[curE];
        int actionCount = 1;
        while (curE + 1 <= e && btMatched[curE] == 0 && btAction[curE + 1] == action) {
            actionCount++;
            curE++;
        }
        if (action == 'I') {
            mp *= T.indel[actionCount];
            offset += actionCount;
            netIndel += actionCount;
            totalIndels += actionCount;
        } else if (action == 'D') {
            mp *= T.indel[actionCount];
            offset -= actionCount;
            netIndel -= actionCount;
            totalIndels += actionCount;
            textSpan += actionCount;
        } else {
            for (int i = 0; i < actionCount; i++) {
                int qi = offset > 0 ? offset : 0;
                if (qi > patternLen - 1) qi = patternLen - 1;
                mp *= T.phred[quality[qi]];
                offset++;
            }
        }
        offset += btMatched[curE];
        curE++;
    }
    mp *= T.perfect[patternLen - e];
    textSpan += patternLen;
    out->matchProbability = mp;
    out->netIndel = netIndel; out->totalIndels = totalIndels; out->textSpan = textSpan;
    out->score = e;
}
