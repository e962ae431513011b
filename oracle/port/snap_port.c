/*
 * snap_port.c -- TEST INFRASTRUCTURE ONLY: plain-C restatement of the leaf algorithms of SNAP's seed-and-extend
 * path, written independently of the product code (snap_b200/csrc) and pinned, in tests/test_oracle.py, against
 * (1) the reference's own known-answer vectors (tests/LandauVishkinTest.cpp:11-32) and
 * (2) the compiled reference in oracle/_ref (bit-exact on random inputs).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product never does.
 * The whole-read state machine and the affine-gap kernels are checked against oracle/_ref directly (the reference
 * compiles here), so they are not restated a second time in this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PORT_MAX_K 127                   /* LandauVishkin.h:11 */
#define PORT_MAX_READ 1000               /* Read.h:49 */

/* ---- Tables.cpp:52-58 ---- */
static int port_base_value(unsigned char c)
{
    switch (c) { case 'A': return 0; case 'G': return 1; case 'C': return 2; case 'T': return 3; default: return 4; }
}

/* ---- Seed::Seed, Seed.h:40-53; returns 0 if the text is not a seed (Seed.cpp:28-42) ---- */
int port_seed_pack(const char *text, unsigned seedLen, uint64_t *bases, uint64_t *reverseComplement)
{
    uint64_t b = 0, rc = 0;
    unsigned i;
    for (i = 0; i < seedLen; i++) {
        int v = port_base_value((unsigned char)text[i]);
        if (v > 3) return 0;
        b |= (uint64_t)v << ((seedLen - i - 1) * 2);
        rc |= (uint64_t)(v ^ 0x3) << (i * 2);
    }
    *bases = b;
    *reverseComplement = rc;
    return 1;
}

/* ---- SNAPHashTable::hash, HashTable.h:72-85 ---- */
uint64_t port_hash(uint64_t key)
{
    key ^= key >> 33;
    key *= 0xff51afd7ed558ccdULL;
    key ^= key >> 33;
    key *= 0xc4ceb9fe1a85ec53ULL;
    key ^= key >> 33;
    return key;
}

/*
 * ---- SNAPHashTable::GetFirstValueForKey, HashTable.h:87-118, for 4-byte values + 4-byte keys ----
 * table: tableSize entries of {uint32 value, uint32 key}.  Returns the slot index or -1; *examined = entries looked at.
 */
int64_t port_probe(const uint32_t *table, uint64_t tableSize, uint32_t key, uint32_t invalidValue, unsigned *examined)
{
    uint64_t idx = port_hash(key) % tableSize;
    unsigned nProbes = 0;
    *examined = 1;
    if (table[2 * idx + 1] == key && table[2 * idx] != invalidValue) return (int64_t)idx;
    do {
        nProbes++;
        if (nProbes > tableSize + 5) return -1;
        if (nProbes < 5) idx = (idx + (uint64_t)nProbes * nProbes) % tableSize; else idx = (idx + 1) % tableSize;
        (*examined)++;
    } while (table[2 * idx + 1] != key && table[2 * idx] != invalidValue);
    if (table[2 * idx] == invalidValue) return -1;
    return (int64_t)idx;
}

/* ---- SeedSequencer::SeedSequencer, SeedSequencer.cpp:36-103 (the work list is a FIFO) ---- */
void port_seed_sequencer(unsigned seedSize, unsigned *offsets)
{
    unsigned lo[64], hi[64], head = 0, tail = 0, nFilled = 1, i;
    for (i = 0; i < seedSize; i++) offsets[i] = 0;
    if (seedSize <= 1) return;
    lo[tail] = 1; hi[tail] = seedSize - 1; tail++;
    while (head != tail) {
        unsigned l = lo[head % 64], h = hi[head % 64], sel;
        head++;
        sel = (l + h) / 2;
        offsets[sel] = nFilled++;
        if (h > sel) { lo[tail % 64] = sel + 1; hi[tail % 64] = h; tail++; }
        if (l < sel) { lo[tail % 64] = l; hi[tail % 64] = sel - 1; tail++; }
    }
}

/* ---- computeMAPQ, mapq.h:32-68 ---- */
int port_mapq(double probabilityOfAllCandidates, double probabilityOfBestCandidate, int popularSeedsSkipped)
{
    double correctnessProbability;
    int baseMAPQ, pen;
    if (probabilityOfAllCandidates < probabilityOfBestCandidate) probabilityOfAllCandidates = probabilityOfBestCandidate;
    correctnessProbability = probabilityOfBestCandidate / probabilityOfAllCandidates;
    if (correctnessProbability >= 1) {
        baseMAPQ = 70;
    } else {
        baseMAPQ = (int)(-10 * log10(1 - correctnessProbability));
        if (baseMAPQ > 70) baseMAPQ = 70;
    }
    pen = popularSeedsSkipped - 10;
    if (pen < 0) pen = 0;
    baseMAPQ -= pen / 2;
    return baseMAPQ < 0 ? 0 : baseMAPQ;
}

/* ---- initializeLVProbabilitiesToPhredPlus33, LandauVishkin.cpp:715-763 ---- */
static double g_phred[256], g_indel[PORT_MAX_READ + PORT_MAX_K + 2], g_perfect[PORT_MAX_READ + 1];
static int g_tables_ready = 0;

static void port_init_tables(void)
{
    int i;
    const double SNP_PROB = 0.001, GAP_OPEN_PROB = 0.001, GAP_EXTEND_PROB = 0.5;
    if (g_tables_ready) return;
    g_indel[0] = 1.0;
    g_indel[1] = GAP_OPEN_PROB;
    for (i = 2; i < PORT_MAX_READ + PORT_MAX_K + 2; i++) g_indel[i] = g_indel[i - 1] * GAP_EXTEND_PROB;
    for (i = 0; i < 33; i++) g_phred[i] = SNP_PROB;
    for (i = 33; i <= 93 + 33; i++) g_phred[i] = 1.0 - (1.0 - pow(10.0, -1.0 * (i - 33.0) / 10.0)) * (1.0 - SNP_PROB);
    for (i = 93 + 33 + 1; i < 256; i++) g_phred[i] = SNP_PROB;
    g_perfect[0] = 1.0;
    for (i = 1; i <= PORT_MAX_READ; i++) g_perfect[i] = g_perfect[i - 1] * (1 - SNP_PROB);
    g_tables_ready = 1;
}

void port_tables(double *phred, double *indel, int nIndel, double *perfect, int nPerfect)
{
    int i;
    port_init_tables();
    memcpy(phred, g_phred, sizeof(g_phred));
    for (i = 0; i < nIndel; i++) indel[i] = g_indel[i];
    for (i = 0; i < nPerfect; i++) perfect[i] = g_perfect[i];
}

/*
 * ---- LandauVishkin<TEXT_DIRECTION>::computeEditDistance, LandauVishkin.h:100-351 ----
 * Full (MAX_K+1) x (2*MAX_K+1) L and A arrays like the reference, initialised to -2 once and never cleared (:55).
 * text is as the reference's callers pass it (for dir == -1 one past the first text character).
 */
static int  L_space[2][(PORT_MAX_K + 1) * (2 * PORT_MAX_K + 1)];
static char A_space[2][(PORT_MAX_K + 1) * (2 * PORT_MAX_K + 1)];
static int  lv_ready = 0;
#define LL(e, d) Lz[(e) * (2 * PORT_MAX_K + 1) + (d)]
#define AA(e, d) Az[(e) * (2 * PORT_MAX_K + 1) + (d)]

static int port_cpm(const char *p, const char *t, int dir, int avail)
{
    /* countPerfectMatch, :377-407: 8 bytes at a time in the reference; the result is min(run length, avail) */
    int n = 0;
    if (avail <= 0) return avail;
    while (n < avail && p[n] == t[n * dir]) n++;
    return n;
}

int port_lv(int dir, const char *text, int textLen, const char *pattern, const char *quality, int patternLen, int k,
            double *matchProbability, int *o_netIndel, int *o_totalIndels, int *o_textSpan)
{
    int *Lz; char *Az;
    int which = dir == 1 ? 0 : 1, e, d, i, end, lastBestD = PORT_MAX_K + 1, gotX = 0;
    char btAction[PORT_MAX_K + 1]; int btMatched[PORT_MAX_K + 1], btD[PORT_MAX_K + 1];
    port_init_tables();
    if (!lv_ready) {
        int z, w;
        for (w = 0; w < 2; w++) for (z = 0; z < (PORT_MAX_K + 1) * (2 * PORT_MAX_K + 1); z++) L_space[w][z] = -2;
        lv_ready = 1;
    }
    Lz = L_space[which] + PORT_MAX_K; Az = A_space[which] + PORT_MAX_K;
    *o_netIndel = 0; *o_totalIndels = 0; *o_textSpan = 0; *matchProbability = 0.0;
    if (k < 0) return -1;
    if (k > PORT_MAX_K - 1) k = PORT_MAX_K - 1;
    if (text == NULL) return -1;
    *matchProbability = 1.0;
    if (dir == -1) text--;
    end = patternLen < textLen ? patternLen : textLen;
    LL(0, 0) = end > 0 ? port_cpm(pattern, text, dir, end) : 0;
    if (LL(0, 0) == end) {
        int result = patternLen > end ? patternLen - end : 0;
        *matchProbability = g_perfect[patternLen];
        if (result > k) return -1;
        *o_textSpan += patternLen;
        return result;
    }
    for (e = 1; e <= k; e++) {
        d = 0;
        for (i = 0; d != e + 1; i++, d = (d > 0 ? -d : -d + 1)) {
            int endd = patternLen < textLen - d ? patternLen : textLen - d;
            int best = LL(e - 1, d) + 1, left, right;
            const char *p, *t;
            AA(e, d) = 'X';
            p = pattern + best; t = text + (d + best) * dir;
            if (best >= 0 && (endd - best != 0) && *p == *t) best += port_cpm(p, t, dir, endd - best);
            left = LL(e - 1, d - 1);
            p = pattern + left; t = text + (d + left) * dir;
            if (left >= 0 && (endd - left != 0) && *p == *t) left += port_cpm(p, t, dir, endd - left);
            if (left > best) { best = left; AA(e, d) = 'D'; }
            right = LL(e - 1, d + 1) + 1;
            p = pattern + right; t = text + (d + right) * dir;
            if (right >= 0 && (endd - right != 0) && *p == *t) right += port_cpm(p, t, dir, endd - right);
            if (right > best) { best = right; AA(e, d) = 'I'; }
            if (best == patternLen) {
                if (AA(e, d) == 'X') { lastBestD = d; gotX = 1; break; }
                if (abs(d) < abs(lastBestD)) lastBestD = d;
            }
            LL(e, d) = best;
        }
        if (gotX || lastBestD != PORT_MAX_K + 1) break;
    }
    if (lastBestD == PORT_MAX_K + 1) return -1;
    {
        int curD = lastBestD, curE, offset = LL(0, 0);
        for (curE = e; curE >= 1; curE--) {
            btAction[curE] = AA(curE, curD);
            if (btAction[curE] == 'I') { btD[curE] = curD + 1; btMatched[curE] = LL(curE, curD) - LL(curE - 1, curD + 1) - 1; }
            else if (btAction[curE] == 'D') { btD[curE] = curD - 1; btMatched[curE] = LL(curE, curD) - LL(curE - 1, curD - 1); }
            else { btD[curE] = curD; btMatched[curE] = LL(curE, curD) - LL(curE - 1, curD) - 1; }
            curD = btD[curE];
        }
        curE = 1;
        while (curE <= e) {
            char action = btAction[curE];
            int actionCount = 1;
            while (curE + 1 <= e && btMatched[curE] == 0 && btAction[curE + 1] == action) { actionCount++; curE++; }
            if (action == 'I') { *matchProbability *= g_indel[actionCount]; offset += actionCount; *o_netIndel += actionCount; *o_totalIndels += actionCount; }
            else if (action == 'D') { *matchProbability *= g_indel[actionCount]; offset -= actionCount; *o_netIndel -= actionCount; *o_totalIndels += actionCount; *o_textSpan += actionCount; }
            else {
                int z;
                for (z = 0; z < actionCount; z++) {
                    int qi = offset > 0 ? offset : 0;
                    if (qi > patternLen - 1) qi = patternLen - 1;
                    *matchProbability *= g_phred[(unsigned char)quality[qi]];
                    offset++;
                }
            }
            offset += btMatched[curE];
            curE++;
        }
        *matchProbability *= g_perfect[patternLen - e];
        *o_textSpan += patternLen;
    }
    return e;
}

typedef struct { uint64_t textOff, patOff; int32_t textLen, patternLen, k, dir; } port_lv_job;
typedef struct { int32_t score, netIndel, totalIndels, textSpan; double matchProbability; } port_lv_out;

void port_lv_batch(const char *textBuf, const char *patBuf, const char *qualBuf, const port_lv_job *jobs, int64_t nJobs, port_lv_out *out)
{
    int64_t j;
    for (j = 0; j < nJobs; j++) {
        out[j].score = port_lv(jobs[j].dir, textBuf + jobs[j].textOff, jobs[j].textLen, patBuf + jobs[j].patOff, qualBuf + jobs[j].patOff,
                               jobs[j].patternLen, jobs[j].k, &out[j].matchProbability, &out[j].netIndel, &out[j].totalIndels, &out[j].textSpan);
    }
}
