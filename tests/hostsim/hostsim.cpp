// hostsim.cpp -- TEST-ONLY host build of the engine's scalar algorithm headers (snap_b200/csrc/sg_*.h).
//
// The same headers are what the CUDA kernels compile; building them with g++ lets the CPU-side test suite
// (`-m "not gpu"`) diff the restated state machine, LV and affine-gap code against the compiled reference
// without a GPU.  This library is built into tests/_build/ and is never loaded by the snap_b200 package:
// the product path has no CPU fallback.
#define SG_AG_POISON_CHECK 1
#include <string>
#include <vector>
#include <string.h>
#include <stdlib.h>
#define SG_WITH_PAIRED 1
#include "../../snap_b200/csrc/sg_align.h"
#include "../../snap_b200/csrc/sg_paired.h"
#include "../../snap_b200/csrc/sg_host.h"
#include "../../snap_b200/csrc/sg_lv_cigar.h"
#include "../../snap_b200/csrc/sg_cigar.h"
#include "../../snap_b200/csrc/sg_ag_cigar.h"
#include "../../snap_b200/csrc/sg_sam.h"
#include "../../snap_b200/csrc/sg_bam.h"
#include "../../snap_b200/csrc/sg_bampost.h"
#include "../../snap_b200/csrc/sg_deflate.h"
#include "../../snap_b200/csrc/sg_samheader.h"

struct HsIndex {
    SgHostIndex host;
    SgIndexView view;
};

struct HsAligner {
    HsIndex *index;
    SgParams params;
    SgTables tables;
    std::vector<uint8_t> scratch;
    SgAligner A;
    bool twoPass = false;
    int64_t deferredReads = 0;
};

static std::string g_err;

extern "C" {

const char *hs_last_error(void) { return g_err.c_str(); }

void *hs_index_open(const char *dir)
{
    HsIndex *ix = new HsIndex;
    if (!sg_load_index_directory(dir, ix->host, g_err)) { delete ix; return NULL; }
    ix->view = ix->host.view();
    return ix;
}

void hs_index_close(void *v) { delete (HsIndex *)v; }

// switches the index to the sector-bucket layout (sg_bucket.h), built on the host from the loaded reference-format tables
int hs_index_relayout(void *v, double load)
{
    HsIndex *ix = (HsIndex *)v;
    if (!sg_host_relayout(ix->host, load > 0 ? load : sg_bucket_default_load(), g_err)) return 1;
    ix->view = ix->host.view();
    return 0;
}

int hs_index_info(void *v, snapgpu_index_info *info)
{
    HsIndex *ix = (HsIndex *)v;
    memset(info, 0, sizeof(*info));
    info->countOfBases = ix->host.nBases; info->seedLen = ix->host.seedLen; info->hashTableKeySize = ix->host.keyBytes;
    info->nHashTables = ix->host.nTables; info->locationSize = 4; info->largeHashTable = ix->host.large;
    info->chromosomePadding = ix->host.chromosomePadding; info->nContigs = (uint32_t)ix->host.contigStart.size();
    info->overflowTableSize = ix->host.overflowSize; info->hashTableSlots = ix->host.layout == SG_LAYOUT_BUCKET ? ix->host.nBuckets * SG_BUCKET_SLOTS : ix->host.totalSlots;
    info->reserved = ix->host.layout;
    return 0;
}

int hs_lookup_seeds(void *v, const char *seeds, int64_t nSeeds, uint32_t maxHitsPerSeed, int64_t *nHits, uint32_t *hits, uint32_t *probes)
{
    HsIndex *ix = (HsIndex *)v;
    const uint32_t sl = ix->host.seedLen;
    for (int64_t i = 0; i < nSeeds; i++) {
        uint64_t b, rc;
        nHits[2 * i] = nHits[2 * i + 1] = 0;
        uint32_t examined = 0, ow = 0;
        if (sg_seed_pack((const uint8_t *)seeds + i * sl, sl, &b, &rc)) {
            SgHits h;
            sg_lookup_seed32(ix->view, b, rc, &h, &examined, &ow);
            for (int d = 0; d < 2; d++) {
                nHits[2 * i + d] = h.nHits[d];
                for (uint32_t k = 0; k < h.nHits[d] && k < maxHitsPerSeed; k++) hits[(2 * i + d) * (int64_t)maxHitsPerSeed + k] = h.hits[d][k];
            }
        }
        if (probes) probes[i] = examined;
    }
    return 0;
}

void hs_tables(unsigned seedLen, double *phred, double *indel, int nIndel, double *perfect, int nPerfect, double *mapqThr, uint32_t *wrap)
{
    SgTables T;
    sg_init_tables(T, seedLen);
    memcpy(phred, T.phred, sizeof(T.phred));
    for (int i = 0; i < nIndel; i++) indel[i] = T.indel[i];
    for (int i = 0; i < nPerfect; i++) perfect[i] = T.perfect[i];
    memcpy(mapqThr, T.mapqThreshold, sizeof(T.mapqThreshold));
    memcpy(wrap, T.wrapSeed, sizeof(T.wrapSeed));
}

int hs_mapq(double pAll, double pBest, int popularSeedsSkipped)
{
    static SgTables T; static bool init = false;
    if (!init) { sg_init_tables(T, 20); init = true; }
    return sg_compute_mapq(T, pAll, pBest, popularSeedsSkipped);
}

static void make_scratch(const SgParams &p, std::vector<uint8_t> &mem, SgScratch *s)
{
    mem.assign(sg_scratch_bytes(p) + 256, 0);
    uint8_t *base = (uint8_t *)(((uintptr_t)mem.data() + 255) & ~(uintptr_t)255);
    sg_scratch_carve(p, base, s);
}

void hs_lv_batch(const char *textBuf, const char *patBuf, const char *qualBuf, const snapgpu_lv_job *jobs, int64_t nJobs, snapgpu_lv_out *out)
{
    SgTables T; sg_init_tables(T, 20);
    SgParams p; memset(&p, 0, sizeof(p));
    p.poolSize = 1; p.tableSlots = 2; p.numWeightLists = 2; p.maxReadLen = 1000;
    std::vector<uint8_t> mem; SgScratch s; make_scratch(p, mem, &s);
    for (int64_t j = 0; j < nJobs; j++) {
        SgLvResult r;
        sg_lv_compute(T, s, jobs[j].dir, (const uint8_t *)textBuf + jobs[j].textOff, jobs[j].textLen, (const uint8_t *)patBuf + jobs[j].patOff,
                      (const uint8_t *)qualBuf + jobs[j].patOff, jobs[j].patternLen, jobs[j].k, &r);
        out[j].score = r.score; out[j].netIndel = r.netIndel; out[j].totalIndels = r.totalIndels; out[j].textSpan = r.textSpan;
        out[j].matchProbability = r.matchProbability;
    }
}

// poisoned[j] (optional): 1 if the traceback of job j read a cell this call never wrote -- on such inputs the
// reference reads state left behind by *earlier* calls (its backtraceAction array is never cleared), i.e. its result
// is history-dependent and no restatement can match it.
void hs_ag_batch(const snapgpu_ag_params *ap, const char *textBuf, const char *patBuf, const char *qualBuf, const snapgpu_ag_job *jobs,
                 int64_t nJobs, snapgpu_ag_out *out, int *poisoned)
{
    SgTables T; sg_init_tables(T, 20);
    SgParams p; memset(&p, 0, sizeof(p));
    p.poolSize = 1; p.tableSlots = 2; p.numWeightLists = 2; p.maxReadLen = 1000;
    std::vector<uint8_t> mem; SgScratch s; make_scratch(p, mem, &s);
    SgAgParams P = sg_ag_params(ap->matchReward, ap->subPenalty, ap->gapOpenPenalty, ap->gapExtendPenalty, ap->fivePrimeEndBonus, ap->threePrimeEndBonus);
    for (int64_t j = 0; j < nJobs; j++) {
        SgAgResult r;
        r.agScore = -1; r.textOffset = 0; r.patternOffset = 0; r.nEdits = 0; r.matchProbability = 0.0;
        sg_ag_compute(T, s, P, jobs[j].dir, jobs[j].banded != 0, (const uint8_t *)textBuf + jobs[j].textOff, jobs[j].textLen,
                      (const uint8_t *)patBuf + jobs[j].patOff, (const uint8_t *)qualBuf + jobs[j].patOff, jobs[j].patternLen, jobs[j].w,
                      jobs[j].scoreInit, jobs[j].isRC != 0, jobs[j].useClippingOptimizations != 0, &r);
        out[j].agScore = r.agScore; out[j].textOffset = r.textOffset; out[j].patternOffset = r.patternOffset; out[j].nEdits = r.nEdits;
        out[j].matchProbability = r.matchProbability;
        if (poisoned) poisoned[j] = r.poisoned;
    }
}

void *hs_aligner_create(void *vix, const snapgpu_params *params, uint32_t maxReadLen)
{
    HsIndex *ix = (HsIndex *)vix;
    HsAligner *a = new HsAligner;
    a->index = ix;
    if (!sg_derive_params(*params, ix->host.seedLen, maxReadLen, a->params, g_err)) { delete a; return NULL; }
    sg_init_tables(a->tables, ix->host.seedLen);
    memset(&a->A, 0, sizeof(a->A));
    make_scratch(a->params, a->scratch, &a->A.sc);
    a->A.ix = &ix->view; a->A.pr = &a->params; a->A.tb = &a->tables;
    a->A.ag = sg_ag_params(params->matchReward, params->subPenalty, params->gapOpenPenalty, params->gapExtendPenalty,
                           params->fivePrimeEndBonus, params->threePrimeEndBonus);
    a->A.invalidLocation = SNAPGPU_INVALID_LOCATION_32;
    a->A.nUsedElements = 0;
    a->A.maxK = a->params.maxK;
    return a;
}

void hs_aligner_destroy(void *v) { delete (HsAligner *)v; }
void hs_aligner_set_two_pass(void *v, int on) { ((HsAligner *)v)->twoPass = on != 0; }
int64_t hs_aligner_deferred(void *v) { return ((HsAligner *)v)->deferredReads; }

// Same contract as snapgpu_align_single: pre-filter (SingleAligner.cpp:213), AlignRead, stats.
int hs_align_single(void *v, int64_t n, const char *bases, const char *quals, const uint64_t *offsets, const uint32_t *lens,
                    snapgpu_single_result *results, snapgpu_counters *ctr)
{
    HsAligner *a = (HsAligner *)v;
    memset(&a->A.work, 0, sizeof(a->A.work));
    for (int64_t i = 0; i < n; i++) {
        const uint8_t *rd = (const uint8_t *)bases + offsets[i];
        const uint8_t *rq = (const uint8_t *)quals + offsets[i];
        snapgpu_single_result *r = &results[i];
        memset(r, 0, sizeof(*r));
        if (lens[i] > a->params.maxReadLen) { g_err = "read longer than maxReadLen"; return 1; }
        uint32_t countOfNs = 0;
        for (uint32_t k = 0; k < lens[i]; k++) countOfNs += (rd[k] == 'N');
        if (ctr) ctr->totalReads++;
        if (lens[i] < a->params.minReadLength || countOfNs > a->params.maxK) {
            r->status = SNAPGPU_NOT_FOUND; r->location = a->A.invalidLocation; r->mapq = 0; r->direction = SNAPGPU_FORWARD;
            if (ctr) ctr->uselessReads++;
            continue;
        }
        if (a->twoPass) {
            // the two-pass launch of sg_align_kernel: the instantiation without affine gap first, and the read again from scratch if it bailed out
            const SgWork before = a->A.work;
            sg_align_read_t<false, true>(a->A, rd, rq, lens[i], r);
            if (a->A.deferred) {
                a->A.work = before; a->deferredReads++;
                memset(r, 0, sizeof(*r));
                sg_align_read(a->A, rd, rq, lens[i], r);
            }
        } else {
            sg_align_read(a->A, rd, rq, lens[i], r);
        }
        if (ctr) {
            if (r->status == SNAPGPU_SINGLE_HIT) ctr->singleHits++;
            else if (r->status == SNAPGPU_MULTIPLE_HITS) ctr->multiHits++;
            else ctr->notFound++;
            if (r->status != SNAPGPU_NOT_FOUND && r->mapq >= 0 && r->mapq <= 70) ctr->mapqHistogram[r->mapq]++;
        }
    }
    if (ctr) {
        ctr->nHashTableLookups += a->A.work.lookups; ctr->nHashEntriesProbed += a->A.work.entriesProbed;
        ctr->nOverflowWordsRead += a->A.work.overflowWords; ctr->lvCalls += a->A.work.lvCalls; ctr->affineGapCalls += a->A.work.agCalls;
        ctr->nHitsIgnoredBecauseOfTooHighPopularity += a->A.work.popularIgnored;
    }
    return 0;
}


// Same contract as snapgpu_align_single_secondary (-om): the raw buffer grows like SingleAligner.cpp:250-263's.
int hs_align_single_secondary(void *v, int64_t n, const char *bases, const char *quals, const uint64_t *offsets, const uint32_t *lens,
                              snapgpu_single_result *results, int maxSecondaryAlignments, int maxSecondaryAlignmentsPerContig, int maxEditDistanceForSecondaryResults,
                              int64_t capacityPerRead, snapgpu_single_result *secondary, int32_t *nSecondary, int rawCap, snapgpu_counters *ctr)
{
    HsAligner *a = (HsAligner *)v;
    memset(&a->A.work, 0, sizeof(a->A.work));
    std::vector<snapgpu_single_result> raw((size_t)(rawCap > 0 ? rawCap : 1));
    for (int64_t i = 0; i < n; i++) {
        const uint8_t *rd = (const uint8_t *)bases + offsets[i];
        const uint8_t *rq = (const uint8_t *)quals + offsets[i];
        snapgpu_single_result *r = &results[i];
        memset(r, 0, sizeof(*r));
        nSecondary[i] = 0;
        if (lens[i] > a->params.maxReadLen) { g_err = "read longer than maxReadLen"; return 1; }
        uint32_t countOfNs = 0;
        for (uint32_t k = 0; k < lens[i]; k++) countOfNs += (rd[k] == 'N');
        if (ctr) ctr->totalReads++;
        if (lens[i] < a->params.minReadLength || countOfNs > a->params.maxK) {
            r->status = SNAPGPU_NOT_FOUND; r->location = a->A.invalidLocation; r->mapq = 0; r->direction = SNAPGPU_FORWARD;
            if (ctr) ctr->uselessReads++;
            continue;
        }
        for (;;) {
            a->A.secResults = raw.data(); a->A.maxSec = (int)raw.size(); a->A.nSec = 0; a->A.secOverflow = 0;
            a->A.secMaxEditDist = maxEditDistanceForSecondaryResults; a->A.secMaxResults = maxSecondaryAlignments; a->A.secMaxPerContig = maxSecondaryAlignmentsPerContig;
            memset(r, 0, sizeof(*r));
            const SgWork before = a->A.work;
            sg_align_read_t<false, false, true>(a->A, rd, rq, lens[i], r);
            if (!a->A.secOverflow) break;
            a->A.work = before;
            raw.resize(raw.size() * 2);
        }
        if ((int64_t)a->A.nSec <= capacityPerRead) {
            for (int k = 0; k < a->A.nSec; k++) secondary[i * capacityPerRead + k] = raw[(size_t)k];
            nSecondary[i] = a->A.nSec;
        } else {
            nSecondary[i] = -a->A.nSec;
        }
        if (ctr) {
            if (r->status == SNAPGPU_SINGLE_HIT) ctr->singleHits++;
            else if (r->status == SNAPGPU_MULTIPLE_HITS) ctr->multiHits++;
            else ctr->notFound++;
            if (r->status != SNAPGPU_NOT_FOUND && r->mapq >= 0 && r->mapq <= 70) ctr->mapqHistogram[r->mapq]++;
        }
    }
    a->A.secResults = nullptr;
    if (ctr) {
        ctr->nHashTableLookups += a->A.work.lookups; ctr->nHashEntriesProbed += a->A.work.entriesProbed;
        ctr->nOverflowWordsRead += a->A.work.overflowWords; ctr->lvCalls += a->A.work.lvCalls; ctr->affineGapCalls += a->A.work.agCalls;
        ctr->nHitsIgnoredBecauseOfTooHighPopularity += a->A.work.popularIgnored;
    }
    return 0;
}


// snapgpu_bgzf_deflate_device on the host: the block-cooperative compressor of sg_deflate.h run by one thread per member.  Returns the BGZF stream's
// size, or -1 when `cap` is too small.  memberSizes (optional): one entry per member.
int64_t hs_bgzf_deflate(const uint8_t *in, int64_t n, uint8_t *out, int64_t cap, uint32_t *memberSizes)
{
    static SgDeflateShared S;
    std::vector<uint16_t> arena(SG_DEFLATE_ARENA_BYTES / 2);
    SgDeflateArena G;
    G.mlen = arena.data(); G.mdist = G.mlen + (SG_DEFLATE_MAX_PAYLOAD + 8); G.jumpA = G.mdist + (SG_DEFLATE_MAX_PAYLOAD + 8); G.jumpB = G.jumpA + (SG_DEFLATE_MAX_PAYLOAD + 8);
    std::vector<uint8_t> member(SG_DEFLATE_MEMBER_PITCH + 64);
    int64_t used = 0, m = 0;
    for (int64_t off = 0; off < n; off += SG_DEFLATE_MAX_PAYLOAD, m++) {
        const uint32_t len = (uint32_t)((n - off) < (int64_t)SG_DEFLATE_MAX_PAYLOAD ? (n - off) : (int64_t)SG_DEFLATE_MAX_PAYLOAD);
        const uint32_t sz = sg_deflate_member(S, G, in + off, len, member.data());
        if (used + sz > cap) return -1;
        memcpy(out + used, member.data(), sz);
        if (memberSizes) memberSizes[m] = sz;
        used += sz;
    }
    return used;
}


// snapgpu_sam_header on the host: the SAM header text (bam = 0) or the BAM header block (bam = 1) of a file over this index.  Returns its size, -1 if cap is
// too small, -2 if the index has no usable contig table.
int64_t hs_sam_header(void *vix, int bam, int sorted, const char *commandLine, const char *version, const char *rgLine, uint8_t *out, int64_t cap)
{
    HsIndex *ix = (HsIndex *)vix;
    std::vector<SgHeaderContig> contigs;
    if (!sg_header_contigs(ix->host.contigName, ix->host.contigStart, ix->host.contigIsAlt, ix->host.contigOriginal, ix->view.nBases, ix->host.chromosomePadding, &contigs)) return -2;
    std::vector<uint8_t> o;
    if (bam) o = sg_bam_header(contigs, sorted != 0, commandLine, version, rgLine);
    else { const std::string t = sg_sam_header_text(contigs, sorted != 0, commandLine, version, rgLine); o.assign(t.begin(), t.end()); }
    if ((int64_t)o.size() > cap) return -1;
    memcpy(out, o.data(), o.size());
    return (int64_t)o.size();
}


struct HsPaired {
    HsIndex *index;
    SgParams pr, prSingle;
    SgPairedParams pp;
    SgTables tables;
    std::vector<uint8_t> scratch, pscratch;
    SgAligner S;
    SgPairedAligner P;
    HsPaired *big = nullptr;         // full-size fallback when this one runs with reduced pool caps (mirrors the GPU retry pass)
    int64_t retried = 0;
    bool staged = false;             // mirror of the CUDA path's staged launch: stage 1 of every pair first, then stage 2 of every pair
};

// poolCap / candCap: 0 = the reference's sizes; otherwise this aligner's pools are capped and a pair that needs more is
// re-aligned from scratch by a second, full-size aligner -- exactly what the CUDA path's retry launch does.
void *hs_paired_create(void *vix, const snapgpu_params *params, const snapgpu_paired_params *pparams, uint32_t maxReadLen, uint32_t poolCap, uint32_t candCap)
{
    HsIndex *ix = (HsIndex *)vix;
    HsPaired *a = new HsPaired;
    a->index = ix;
    if (!sg_derive_paired_params(*params, *pparams, ix->host.seedLen, maxReadLen, a->pr, a->prSingle, a->pp, g_err)) { delete a; return NULL; }
    if (poolCap || candCap) {
        a->big = (HsPaired *)hs_paired_create(vix, params, pparams, maxReadLen, 0, 0);
        if (poolCap && poolCap < a->pp.poolSize) a->pp.poolCap = poolCap & ~1u;
        if (candCap && candCap < SG_MAX_AG_CANDIDATES) a->pp.agCandCap = candCap;
    }
    sg_init_tables(a->tables, ix->host.seedLen);
    memset(&a->S, 0, sizeof(a->S));
    make_scratch(a->prSingle, a->scratch, &a->S.sc);
    a->S.ix = &ix->view; a->S.pr = &a->prSingle; a->S.tb = &a->tables;
    a->S.ag = sg_ag_params(params->matchReward, params->subPenalty, params->gapOpenPenalty, params->gapExtendPenalty,
                           params->fivePrimeEndBonus, params->threePrimeEndBonus);
    a->S.invalidLocation = SNAPGPU_INVALID_LOCATION_32;
    a->S.maxK = a->prSingle.maxK;
    memset(&a->P, 0, sizeof(a->P));
    a->pscratch.assign(sg_paired_scratch_bytes(a->pr, a->pp) + 256, 0);
    sg_paired_scratch_carve(a->pr, a->pp, (uint8_t *)(((uintptr_t)a->pscratch.data() + 255) & ~(uintptr_t)255), &a->P.ps);
    a->P.single = &a->S; a->P.ix = &ix->view; a->P.pr = &a->pr; a->P.pp = &a->pp; a->P.tb = &a->tables;
    a->P.ag = a->S.ag; a->P.invalidLocation = SNAPGPU_INVALID_LOCATION_32;
    return a;
}

void hs_paired_destroy(void *v) { HsPaired *a = (HsPaired *)v; if (a->big) hs_paired_destroy(a->big); delete a; }
int64_t hs_paired_retried(void *v) { return ((HsPaired *)v)->retried; }
void hs_paired_set_staged(void *v, int on) { ((HsPaired *)v)->staged = on != 0; }

// Same contract as oracle ref_paired_align: pair i = reads 2i, 2i+1; the pre-filter of PairedAligner.cpp:669-707.
int hs_align_paired(void *v, int64_t nPairs, const char *bases, const char *quals, const uint64_t *offsets, const uint32_t *lens,
                    snapgpu_paired_result *results, int64_t *nLV, int64_t *nAG)
{
    HsPaired *a = (HsPaired *)v;
    a->P.lvCalls = a->P.agCalls = 0;
    memset(&a->S.work, 0, sizeof(a->S.work));
    struct Handoff { int stage = 0; int nLVCand = 0; std::vector<snapgpu_paired_result> cands; };
    std::vector<Handoff> hand(a->staged ? (size_t)nPairs : 0);
    for (int64_t i = 0; i < nPairs; i++) {
        snapgpu_paired_result *r = &results[i];
        memset(r, 0, sizeof(*r));
        const uint8_t *rb[2], *rq[2]; uint32_t ln[2]; bool useful[2];
        for (int w = 0; w < 2; w++) {
            rb[w] = (const uint8_t *)bases + offsets[2 * i + w]; rq[w] = (const uint8_t *)quals + offsets[2 * i + w]; ln[w] = lens[2 * i + w];
            if (ln[w] > a->pr.maxReadLen) { g_err = "read longer than maxReadLen"; return 1; }
            uint32_t countOfNs = 0;
            for (uint32_t k = 0; k < ln[w]; k++) countOfNs += (rb[w][k] == 'N');
            useful[w] = ln[w] >= a->pr.minReadLength && countOfNs <= a->pr.maxK;
        }
        if (!useful[0] && !useful[1]) {
            for (int w = 0; w < 2; w++) { r->status[w] = SNAPGPU_NOT_FOUND; r->location[w] = a->P.invalidLocation; }
            continue;
        }
        a->P.error = 0;
        if (a->staged) {
            Handoff &h = hand[i];
            h.stage = sg_paired_align_stage1(a->P, rb, rq, ln, r, &h.nLVCand);
            if (a->P.error == 0 && h.stage != 0) {
                h.cands.assign(a->P.ps.lvCandidates, a->P.ps.lvCandidates + h.nLVCand);
                continue;
            }
            h.stage = 0;
        } else {
            sg_paired_align(a->P, rb, rq, ln, r);
        }
        if (a->P.error == 4 && a->big) {
            a->retried++;
            memset(r, 0, sizeof(*r));
            a->big->P.error = 0;
            sg_paired_align(a->big->P, rb, rq, ln, r);
            if (a->big->P.error) { g_err = "paired candidate pool / buffer overflow"; return 2; }
            continue;
        }
        if (a->P.error) { g_err = "paired candidate pool / buffer overflow"; return 2; }
    }
    for (int64_t i = 0; a->staged && i < nPairs; i++) {
        Handoff &h = hand[i];
        if (h.stage == 0) continue;
        snapgpu_paired_result *r = &results[i];
        const uint8_t *rb[2], *rq[2]; uint32_t ln[2];
        for (int w = 0; w < 2; w++) { rb[w] = (const uint8_t *)bases + offsets[2 * i + w]; rq[w] = (const uint8_t *)quals + offsets[2 * i + w]; ln[w] = lens[2 * i + w]; }
        // by now the aligner's scratch has been through every other pair of the batch
        sg_paired_restore_reads(a->P, rb, rq, ln);
        for (int k = 0; k < h.nLVCand; k++) a->P.ps.lvCandidates[k] = h.cands[k];
        a->P.error = 0;
        h.stage = sg_paired_align_stage2(a->P, r, h.stage, h.nLVCand);
        if (a->P.error == 0 && h.stage != 0) continue;          // on to stage 3
        h.stage = 0;
        if (a->P.error == 4 && a->big) {
            a->retried++;
            memset(r, 0, sizeof(*r));
            a->big->P.error = 0;
            sg_paired_align(a->big->P, rb, rq, ln, r);
            if (a->big->P.error) { g_err = "paired candidate pool / buffer overflow"; return 2; }
            continue;
        }
        if (a->P.error) { g_err = "paired candidate pool / buffer overflow"; return 2; }
    }
    for (int64_t i = 0; a->staged && i < nPairs; i++) {
        Handoff &h = hand[i];
        if (h.stage == 0) continue;
        snapgpu_paired_result *r = &results[i];
        const uint8_t *rb[2], *rq[2]; uint32_t ln[2];
        for (int w = 0; w < 2; w++) { rb[w] = (const uint8_t *)bases + offsets[2 * i + w]; rq[w] = (const uint8_t *)quals + offsets[2 * i + w]; ln[w] = lens[2 * i + w]; }
        a->P.error = 0;
        sg_paired_align_stage3(a->P, rb, rq, ln, r, h.stage);
        if (a->P.error == 4 && a->big) {
            a->retried++;
            memset(r, 0, sizeof(*r));
            a->big->P.error = 0;
            sg_paired_align(a->big->P, rb, rq, ln, r);
            if (a->big->P.error) { g_err = "paired candidate pool / buffer overflow"; return 2; }
            continue;
        }
        if (a->P.error) { g_err = "paired candidate pool / buffer overflow"; return 2; }
    }
    if (nLV) *nLV = a->P.lvCalls + a->S.work.lvCalls;
    if (nAG) *nAG = a->P.agCalls + a->S.work.agCalls;
    return 0;
}

// LandauVishkinWithCigar restated (sg_lv_cigar.h) over the same test-local job / out records as oracle ref_lv_cigar_batch
struct HsLvCigarJob { unsigned long long textOff, patOff; int textLen, patternLen, k, useM; };
struct HsLvCigarOut { int score, nOps, textUsed, netIndel, normalizedScore, addFrontClipping; unsigned ops[32]; };

void hs_lv_cigar_batch(const char *textBuf, const char *patBuf, const HsLvCigarJob *jobs, int64_t nJobs, HsLvCigarOut *out)
{
    const int kmax = SG_MAX_K - 1;
    std::vector<int> L(sg_lv_cigar_scratch_ints(kmax)), TI(sg_lv_cigar_scratch_ints(kmax)), btM(kmax + 2), btD(kmax + 2);
    std::vector<uint8_t> A(sg_lv_cigar_scratch_ints(kmax)), btA(kmax + 2);
    SgLvCigarScratch S;
    S.L = L.data(); S.totalIndels = TI.data(); S.A = A.data(); S.btAction = btA.data(); S.btMatched = btM.data(); S.btD = btD.data(); S.kmax = kmax;
    for (int64_t j = 0; j < nJobs; j++) {
        const HsLvCigarJob &b = jobs[j];
        HsLvCigarOut &o = out[j];
        memset(&o, 0, sizeof(o));
        SgLvCigarOut r;
        sg_lv_cigar_compute(S, (const uint8_t *)textBuf + b.textOff, b.textLen, (const uint8_t *)patBuf + b.patOff, b.patternLen, b.k, o.ops, 32, b.useM != 0, &r);
        o.score = r.score;
        if (r.score >= 0) { o.nOps = r.nOps; o.textUsed = r.textUsed; o.netIndel = r.netIndel; } else memset(o.ops, 0, sizeof(o.ops));
        unsigned ops2[32]; int clip = 0;
        SgLvCigarOut r2;
        o.normalizedScore = sg_lv_cigar_normalized(S, (const uint8_t *)textBuf + b.textOff, b.textLen, (const uint8_t *)patBuf + b.patOff, b.patternLen, b.k, ops2, 32,
                                                   b.useM != 0, &r2, &clip);
        o.addFrontClipping = o.normalizedScore >= 0 ? clip : 0;
    }
}

// sg_cigar.h (SAMFormat::computeCigarString, LV overload) over the same job records as oracle ref_cigar_lv_batch
struct HsCigarJob { unsigned long long dataOff; long long location; int dataLength, basesClippedBefore, extraBasesClippedBefore, basesClippedAfter,
                    frontHardClipping, backHardClipping, direction, useM; };
struct HsCigarOut { int kind, editDistance, addFrontClipping, refSpan, nOps; unsigned ops[40]; };

void hs_cigar_lv_batch(void *vix, const char *dataBuf, const HsCigarJob *jobs, int64_t nJobs, HsCigarOut *out)
{
    HsIndex *ix = (HsIndex *)vix;
    const int kmax = SG_MAX_K - 1;
    std::vector<int> L(sg_lv_cigar_scratch_ints(kmax)), TI(sg_lv_cigar_scratch_ints(kmax)), btM(kmax + 2), btD(kmax + 2);
    std::vector<uint8_t> A(sg_lv_cigar_scratch_ints(kmax)), btA(kmax + 2);
    SgLvCigarScratch S;
    S.L = L.data(); S.totalIndels = TI.data(); S.A = A.data(); S.btAction = btA.data(); S.btMatched = btM.data(); S.btD = btD.data(); S.kmax = kmax;
    for (int64_t j = 0; j < nJobs; j++) {
        const HsCigarJob &b = jobs[j];
        HsCigarOut &o = out[j];
        memset(&o, 0, sizeof(o));
        SgCigarOut r;
        sg_cigar_lv(ix->view, S, (const uint8_t *)dataBuf + b.dataOff, b.dataLength, (uint32_t)b.basesClippedBefore, b.extraBasesClippedBefore,
                    (uint32_t)b.basesClippedAfter, (uint32_t)b.frontHardClipping, (uint32_t)b.backHardClipping, b.location, b.useM != 0, o.ops, 40, &r);
        o.kind = r.kind; o.editDistance = r.editDistance; o.addFrontClipping = r.addFrontClipping;
        if (r.kind == 2) { o.refSpan = r.refSpan; o.nOps = r.nOps; } else memset(o.ops, 0, sizeof(o.ops));
    }
}

// sg_ag_cigar.h (AffineGapVectorizedWithCigar::computeGlobalScore) over the same job records as oracle ref_ag_cigar_global_batch
struct HsAgCigarJob { unsigned long long textOff, patOff; int textLen, patternLen, w, useM; };
struct HsAgCigarOut { int score, nOps, netDel, tailIns; unsigned ops[64]; };

void hs_ag_cigar_global_batch(const int *params, const char *textBuf, const char *patBuf, const char *qualBuf, const HsAgCigarJob *jobs, int64_t nJobs,
                              HsAgCigarOut *out)
{
    const int numVecMax = (1000 + 7) / 8, rowsMax = 1000 + SG_MAX_K + 8, resMax = 2 * rowsMax;
    std::vector<int16_t> H(numVecMax * 8), Hm1(numVecMax * 8), E(numVecMax * 8), prof(5 * numVecMax * 8);
    std::vector<uint8_t> bt((size_t)rowsMax * numVecMax * 8), ra(resMax);
    std::vector<int> rc(resMax);
    SgAgCigarScratch S;
    S.H = H.data(); S.Hm1 = Hm1.data(); S.E = E.data(); S.prof = prof.data(); S.bt = bt.data(); S.resAction = ra.data(); S.resCount = rc.data();
    S.numVecMax = numVecMax; S.rowsMax = rowsMax; S.resMax = resMax;
    SgAgParams P = sg_ag_params(params[0], params[1], params[2], params[3], 0, 0);
    for (int64_t j = 0; j < nJobs; j++) {
        const HsAgCigarJob &b = jobs[j];
        HsAgCigarOut &o = out[j];
        memset(&o, 0, sizeof(o));
        SgAgCigarOut r;
        sg_ag_cigar_global(P, S, (const uint8_t *)textBuf + b.textOff, b.textLen, (const uint8_t *)patBuf + b.patOff, (const uint8_t *)qualBuf + b.patOff, b.patternLen,
                           o.ops, 64, b.useM != 0, &r);
        o.score = r.score;
        if (r.score >= 0) { o.nOps = r.nOps; o.netDel = r.netDel; o.tailIns = r.tailIns; } else memset(o.ops, 0, sizeof(o.ops));
    }
}

struct HsAgCigarNormOut { int score, nOps, netDel, tailIns, addFrontClipping; unsigned ops[64]; };

// one persistent scratch, like the reference object's members: the banded form's traceback can read what earlier calls left
void hs_ag_cigar_norm_batch(const int *params, const char *textBuf, const char *patBuf, const char *qualBuf, const HsAgCigarJob *jobs, int64_t nJobs,
                            HsAgCigarNormOut *out)
{
    const int numVecMax = (1000 + 7) / 8 + 16, rowsMax = 1000 + SG_MAX_K + 8, resMax = 2 * rowsMax;
    static std::vector<int16_t> H(numVecMax * 8), Hm1(numVecMax * 8), E(numVecMax * 8), prof(5 * numVecMax * 8);
    static std::vector<uint8_t> bt((size_t)rowsMax * numVecMax * 8), ra(resMax);
    static std::vector<int> rc(resMax);
    SgAgCigarScratch S;
    S.H = H.data(); S.Hm1 = Hm1.data(); S.E = E.data(); S.prof = prof.data(); S.bt = bt.data(); S.resAction = ra.data(); S.resCount = rc.data();
    S.numVecMax = numVecMax; S.rowsMax = rowsMax; S.resMax = resMax;
    SgAgParams P = sg_ag_params(params[0], params[1], params[2], params[3], 0, 0);
    for (int64_t j = 0; j < nJobs; j++) {
        const HsAgCigarJob &b = jobs[j];
        HsAgCigarNormOut &o = out[j];
        memset(&o, 0, sizeof(o));
        SgAgCigarOut r; int clip = 0;
        o.score = sg_ag_cigar_normalized(P, S, (const uint8_t *)textBuf + b.textOff, b.textLen, (const uint8_t *)patBuf + b.patOff, (const uint8_t *)qualBuf + b.patOff,
                                         b.patternLen, b.w, o.ops, 64, b.useM != 0, &r, &clip);
        o.netDel = r.netDel; o.tailIns = r.tailIns; o.addFrontClipping = clip;
        if (o.score > 0 || (o.score == 0 && clip == 0)) o.nOps = r.nOps; else memset(o.ops, 0, sizeof(o.ops));
        for (int q = o.nOps; q < 64; q++) o.ops[q] = 0;       // (the banded attempt may have left operations behind the fall-back's)
    }
}

struct HsCigarAgJob { unsigned long long dataOff; long long location; int dataLength, basesClippedBefore, extraBasesClippedBefore, basesClippedAfter,
                      frontHardClipping, backHardClipping, direction, useM, score, pad; };

void hs_cigar_ag_batch(void *vix, const int *params, const char *dataBuf, const char *qualBuf, const HsCigarAgJob *jobs, int64_t nJobs, HsCigarOut *out)
{
    HsIndex *ix = (HsIndex *)vix;
    const int numVecMax = (1000 + 7) / 8 + 16, rowsMax = 1000 + SG_MAX_K + 8, resMax = 2 * rowsMax;
    static std::vector<int16_t> H(numVecMax * 8), Hm1(numVecMax * 8), E(numVecMax * 8), prof(5 * numVecMax * 8);
    static std::vector<uint8_t> bt((size_t)rowsMax * numVecMax * 8), ra(resMax);
    static std::vector<int> rc(resMax);
    SgAgCigarScratch S;
    S.H = H.data(); S.Hm1 = Hm1.data(); S.E = E.data(); S.prof = prof.data(); S.bt = bt.data(); S.resAction = ra.data(); S.resCount = rc.data();
    S.numVecMax = numVecMax; S.rowsMax = rowsMax; S.resMax = resMax;
    SgAgParams P = sg_ag_params(params[0], params[1], params[2], params[3], 0, 0);
    for (int64_t j = 0; j < nJobs; j++) {
        const HsCigarAgJob &b = jobs[j];
        HsCigarOut &o = out[j];
        memset(&o, 0, sizeof(o));
        SgCigarOut r;
        sg_cigar_ag(ix->view, P, S, (const uint8_t *)dataBuf + b.dataOff, (const uint8_t *)qualBuf + b.dataOff, b.dataLength, b.score, (uint32_t)b.basesClippedBefore,
                    b.extraBasesClippedBefore, (uint32_t)b.basesClippedAfter, (uint32_t)b.frontHardClipping, (uint32_t)b.backHardClipping, b.location, b.useM != 0,
                    o.ops, 40, &r);
        o.kind = r.kind; o.editDistance = r.editDistance; o.addFrontClipping = r.addFrontClipping;
        if (r.kind == 2) { o.refSpan = r.refSpan; o.nOps = r.nOps; for (int q = o.nOps; q < 40; q++) o.ops[q] = 0; } else memset(o.ops, 0, sizeof(o.ops));
    }
}

// sg_sam.h: one SAM record per read from the result records, text appended to `out` (returns the bytes written, -1 if `outCap` is too small)
int64_t hs_sam_single_batch(void *vix, const int *agParams, int useM, int useAffineGap, int64_t n, const char *bases, const char *quals, const uint64_t *offsets,
                            const uint32_t *lens, const char *ids, const uint64_t *idOffsets, const uint32_t *idLens, const snapgpu_single_result *results,
                            const snapgpu_paired_result *pairedResults, const uint32_t *frontClipped, const uint32_t *clippedLens, char *out, int64_t outCap)
{
    HsIndex *ix = (HsIndex *)vix;
    const int kmax = SG_MAX_K - 1;
    std::vector<int> L(sg_lv_cigar_scratch_ints(kmax)), TI(sg_lv_cigar_scratch_ints(kmax)), btM(kmax + 2), btD(kmax + 2);
    std::vector<uint8_t> A(sg_lv_cigar_scratch_ints(kmax)), btA(kmax + 2);
    const int numVecMax = (1000 + 7) / 8 + 16, rowsMax = 1000 + SG_MAX_K + 8, resMax = 2 * rowsMax;
    std::vector<int16_t> H(numVecMax * 8), Hm1(numVecMax * 8), E(numVecMax * 8), prof(5 * numVecMax * 8);
    std::vector<uint8_t> bt((size_t)rowsMax * numVecMax * 8), ra(resMax), data(1024), quality(1024);
    std::vector<int> rc(resMax);
    std::vector<const char *> names;
    for (size_t c = 0; c < ix->host.contigName.size(); c++) names.push_back(ix->host.contigName[c].c_str());
    SgSamContext C;
    C.ix = &ix->view; C.contigName = names.data();
    C.ag = sg_ag_params(agParams[0], agParams[1], agParams[2], agParams[3], 0, 0);
    C.readGroupAux = "\tRG:Z:FASTQ\tPL:Z:Illumina\tPU:Z:pu\tLB:Z:lb\tSM:Z:sm";
    C.useM = useM != 0; C.useAffineGap = useAffineGap != 0;
    C.lv.L = L.data(); C.lv.totalIndels = TI.data(); C.lv.A = A.data(); C.lv.btAction = btA.data(); C.lv.btMatched = btM.data(); C.lv.btD = btD.data(); C.lv.kmax = kmax;
    C.agS.H = H.data(); C.agS.Hm1 = Hm1.data(); C.agS.E = E.data(); C.agS.prof = prof.data(); C.agS.bt = bt.data(); C.agS.resAction = ra.data(); C.agS.resCount = rc.data();
    C.agS.numVecMax = numVecMax; C.agS.rowsMax = rowsMax; C.agS.resMax = resMax;
    C.data = data.data(); C.quality = quality.data();
    std::vector<uint8_t> data2(1024), quality2(1024);
    C.data2 = data2.data(); C.quality2 = quality2.data();
    int64_t used = 0;
    if (pairedResults != NULL) {
        const bool scrubP = getenv("HS_SAM_SCRUB") != NULL;
        for (int64_t i = 0; i < n / 2; i++) {
            if (scrubP) memset(bt.data(), 0x2a + (int)(i % 7), bt.size());
            if (lens[2 * i] > 1000 || lens[2 * i + 1] > 1000 || used + 8192 > outCap) return -1;
            SgSamRead R[2];
            for (int w = 0; w < 2; w++) {
                const int64_t k = 2 * i + w;
                R[w].unclippedData = (const uint8_t *)bases + offsets[k]; R[w].unclippedQuality = (const uint8_t *)quals + offsets[k]; R[w].unclippedLength = lens[k];
                R[w].frontClipped = 0; R[w].dataLength = lens[k]; R[w].id = (const uint8_t *)ids + idOffsets[k]; R[w].idLength = idLens[k];
                R[w].additionalFrontClipping = 0; R[w].additionalBackClipping = 0;
            }
            const snapgpu_paired_result &r = pairedResults[i];
            SgSamPairResult pr;
            for (int w = 0; w < 2; w++) {
                pr.status[w] = r.status[w]; pr.location[w] = r.location[w]; pr.direction[w] = r.direction[w]; pr.mapq[w] = r.mapq[w]; pr.score[w] = r.score[w];
                pr.usedAffineGapScoring[w] = r.usedAffineGapScoring[w]; pr.basesClippedBefore[w] = r.basesClippedBefore[w]; pr.basesClippedAfter[w] = r.basesClippedAfter[w];
                pr.clippingForReadAdjustment[w] = r.clippingForReadAdjustment[w];
            }
            pr.alignedAsPair = r.alignedAsPair;
            used += sg_sam_write_pair(C, R[0], R[1], pr, out + used);
        }
        return used;
    }
    const bool scrub = getenv("HS_SAM_SCRUB") != NULL;      // fill the never-cleared action array with junk before every read: does any record depend on call history?
    for (int64_t i = 0; i < n; i++) {
        if (lens[i] > 1000 || used + 4096 > outCap) return -1;
        if (scrub) memset(bt.data(), 0x2a + (int)(i % 7), bt.size());
        SgSamRead R;
        R.unclippedData = (const uint8_t *)bases + offsets[i]; R.unclippedQuality = (const uint8_t *)quals + offsets[i]; R.unclippedLength = lens[i];
        R.frontClipped = frontClipped ? frontClipped[i] : 0; R.dataLength = clippedLens ? clippedLens[i] : lens[i];      // quality clipping (Read::clip)
        R.id = (const uint8_t *)ids + idOffsets[i]; R.idLength = idLens[i];
        R.additionalFrontClipping = 0; R.additionalBackClipping = 0;
        const snapgpu_single_result &r = results[i];
        SgSamResult sr;
        sr.status = r.status; sr.location = r.status == SNAPGPU_NOT_FOUND ? -1 : r.location; sr.direction = r.direction; sr.mapq = r.mapq; sr.score = r.score;
        sr.scorePriorToClipping = r.scorePriorToClipping; sr.usedAffineGapScoring = r.usedAffineGapScoring; sr.basesClippedBefore = r.basesClippedBefore;
        sr.basesClippedAfter = r.basesClippedAfter; sr.clippingForReadAdjustment = r.clippingForReadAdjustment;
        used += sg_sam_write_single(C, R, sr, out + used);
    }
    return used;
}

// sg_bam.h: one BAM record per read (unpaired), appended to `out`
int64_t hs_bam_single_batch(void *vix, const int *agParams, int useM, int useAffineGap, int64_t n, const char *bases, const char *quals, const uint64_t *offsets,
                            const uint32_t *lens, const char *ids, const uint64_t *idOffsets, const uint32_t *idLens, const snapgpu_single_result *results,
                            const snapgpu_paired_result *pairedResults, char *out, int64_t outCap)
{
    HsIndex *ix = (HsIndex *)vix;
    const int kmax = SG_MAX_K - 1;
    std::vector<int> L(sg_lv_cigar_scratch_ints(kmax)), TI(sg_lv_cigar_scratch_ints(kmax)), btM(kmax + 2), btD(kmax + 2);
    std::vector<uint8_t> A(sg_lv_cigar_scratch_ints(kmax)), btA(kmax + 2);
    const int numVecMax = (1000 + 7) / 8 + 16, rowsMax = 1000 + SG_MAX_K + 8, resMax = 2 * rowsMax;
    std::vector<int16_t> H(numVecMax * 8), Hm1(numVecMax * 8), E(numVecMax * 8), prof(5 * numVecMax * 8);
    std::vector<uint8_t> bt((size_t)rowsMax * numVecMax * 8), ra(resMax), data(1024), quality(1024), data2(1024), quality2(1024);
    std::vector<int> rc(resMax);
    std::vector<const char *> names;
    for (size_t c = 0; c < ix->host.contigName.size(); c++) names.push_back(ix->host.contigName[c].c_str());
    SgSamContext C;
    C.ix = &ix->view; C.contigName = names.data();
    C.ag = sg_ag_params(agParams[0], agParams[1], agParams[2], agParams[3], 0, 0);
    C.readGroupAux = "";
    C.useM = useM != 0; C.useAffineGap = useAffineGap != 0;
    C.lv.L = L.data(); C.lv.totalIndels = TI.data(); C.lv.A = A.data(); C.lv.btAction = btA.data(); C.lv.btMatched = btM.data(); C.lv.btD = btD.data(); C.lv.kmax = kmax;
    C.agS.H = H.data(); C.agS.Hm1 = Hm1.data(); C.agS.E = E.data(); C.agS.prof = prof.data(); C.agS.bt = bt.data(); C.agS.resAction = ra.data(); C.agS.resCount = rc.data();
    C.agS.numVecMax = numVecMax; C.agS.rowsMax = rowsMax; C.agS.resMax = resMax;
    C.data = data.data(); C.quality = quality.data(); C.data2 = data2.data(); C.quality2 = quality2.data();
    static const char rg[] = "RGZFASTQ\0PLZIllumina\0PUZpu\0LBZlb\0SMZsm";       // + the terminating NUL of the literal = the last tag's
    SgBamContext B;
    B.readGroupAux = (const uint8_t *)rg; B.readGroupAuxLen = (int)sizeof(rg);
    int64_t used = 0;
    if (pairedResults != NULL) {
        for (int64_t i = 0; i < n / 2; i++) {
            if (lens[2 * i] > 1000 || lens[2 * i + 1] > 1000 || used + 8192 > outCap) return -1;
            SgSamRead R[2];
            for (int w = 0; w < 2; w++) {
                const int64_t k = 2 * i + w;
                R[w].unclippedData = (const uint8_t *)bases + offsets[k]; R[w].unclippedQuality = (const uint8_t *)quals + offsets[k]; R[w].unclippedLength = lens[k];
                R[w].frontClipped = 0; R[w].dataLength = lens[k]; R[w].id = (const uint8_t *)ids + idOffsets[k]; R[w].idLength = idLens[k];
                R[w].additionalFrontClipping = 0; R[w].additionalBackClipping = 0;
            }
            const snapgpu_paired_result &r = pairedResults[i];
            SgSamPairResult pr;
            for (int w = 0; w < 2; w++) {
                pr.status[w] = r.status[w]; pr.location[w] = r.location[w]; pr.direction[w] = r.direction[w]; pr.mapq[w] = r.mapq[w]; pr.score[w] = r.score[w];
                pr.usedAffineGapScoring[w] = r.usedAffineGapScoring[w]; pr.basesClippedBefore[w] = r.basesClippedBefore[w]; pr.basesClippedAfter[w] = r.basesClippedAfter[w];
                pr.clippingForReadAdjustment[w] = r.clippingForReadAdjustment[w];
            }
            pr.alignedAsPair = r.alignedAsPair;
            used += sg_bam_write_pair(C, B, R[0], R[1], pr, out + used);
        }
        return used;
    }
    for (int64_t i = 0; i < n; i++) {
        if (lens[i] > 1000 || used + 4096 > outCap) return -1;
        SgSamRead R;
        R.unclippedData = (const uint8_t *)bases + offsets[i]; R.unclippedQuality = (const uint8_t *)quals + offsets[i]; R.unclippedLength = lens[i];
        R.frontClipped = 0; R.dataLength = lens[i]; R.id = (const uint8_t *)ids + idOffsets[i]; R.idLength = idLens[i];
        R.additionalFrontClipping = 0; R.additionalBackClipping = 0;
        const snapgpu_single_result &r = results[i];
        SgSamResult sr;
        sr.status = r.status; sr.location = r.location; sr.direction = r.direction; sr.mapq = r.mapq; sr.score = r.score;
        sr.scorePriorToClipping = r.scorePriorToClipping; sr.usedAffineGapScoring = r.usedAffineGapScoring; sr.basesClippedBefore = r.basesClippedBefore;
        sr.basesClippedAfter = r.basesClippedAfter; sr.clippingForReadAdjustment = r.clippingForReadAdjustment;
        used += sg_bam_write_single(C, B, R, sr, (uint8_t *)out + used);
    }
    return used;
}

// ---- row N4 after the sort: duplicate marking and the .bai of a coordinate-sorted stream of BAM records (sg_bampost.h), host orchestration ----
static bool split_records(const uint8_t *records, int64_t nBytes, std::vector<unsigned long long> &off)
{
    int64_t p = 0;
    while (p + 4 <= nBytes) { off.push_back((unsigned long long)p); SgBamRec r; r.p = records + p; if (r.size() < 36) return false; p += r.size(); }
    return p == nBytes;
}

// BAMDupMarkFilter over the whole stream as one batch: sets FLAG 0x400 in place, returns the number of records it set it on (-1: malformed stream)
int64_t hs_bam_markdup(uint8_t *records, int64_t nBytes, const int64_t *contigStartByOriginal, int32_t nRef)
{
    std::vector<unsigned long long> off;
    if (!split_records(records, nBytes, off)) return -1;
    const long long n = (long long)off.size();
    std::vector<SgDupFields> f((size_t)n);
    for (long long i = 0; i < n; i++) { SgBamRec r; r.p = records + off[i]; sg_dup_fields(r, contigStartByOriginal, nRef, &f[i]); }
    // the runs actually visited (the device marks this orbit by pointer jumping)
    std::vector<long long> rs, re;
    for (long long s = 0; s < n;) {
        if (f[s].logical == SG_DUP_INVALID_LOCATION) { s++; continue; }
        const long long e = sg_dup_first_beyond(f.data(), n, s, 2 * SG_DUP_RUN_REACH);
        rs.push_back(s); re.push_back(e);
        if (e == n) break;
        s = sg_dup_first_beyond(f.data(), n, s, SG_DUP_RUN_REACH);
    }
    SgDupView V; V.n = n; V.records = records; V.offsets = off.data(); V.f = f.data(); V.nRuns = (long long)rs.size(); V.runStart = rs.data(); V.runEnd = re.data();
    std::vector<int32_t> flagRun((size_t)n, -1);
    std::vector<uint8_t> fragFlag((size_t)n, 0);
    // pair keys
    {
        std::vector<uint32_t> idx;
        for (long long i = 0; i < n; i++) if (f[i].flag & SG_BAM_FLAG_PAIRED) idx.push_back((uint32_t)i);
        auto lo = [&](uint32_t i) { return std::min(f[i].info, f[i].mateInfo); };
        auto hi = [&](uint32_t i) { return std::max(f[i].info, f[i].mateInfo); };
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
            if (f[a].lib != f[b].lib) return f[a].lib < f[b].lib;
            if (lo(a) != lo(b)) return lo(a) < lo(b);
            return hi(a) < hi(b); });
        for (size_t j = 0; j < idx.size();) {
            size_t e = j + 1;
            while (e < idx.size() && f[idx[e]].lib == f[idx[j]].lib && lo(idx[e]) == lo(idx[j]) && hi(idx[e]) == hi(idx[j])) e++;
            sg_dup_walk_pair_key(V, idx.data() + j, (long long)(e - j), lo(idx[j]), hi(idx[j]), flagRun.data());
            j = e;
        }
    }
    // fragment keys
    {
        std::vector<uint32_t> idx((size_t)n);
        for (long long i = 0; i < n; i++) idx[(size_t)i] = (uint32_t)i;
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return f[a].lib != f[b].lib ? f[a].lib < f[b].lib : f[a].info < f[b].info; });
        for (size_t j = 0; j < idx.size();) {
            size_t e = j + 1;
            while (e < idx.size() && f[idx[e]].lib == f[idx[j]].lib && f[idx[e]].info == f[idx[j]].info) e++;
            sg_dup_walk_fragment_key(V, idx.data() + j, (long long)(e - j), flagRun.data(), fragFlag.data());
            j = e;
        }
    }
    int64_t marked = 0;
    for (long long i = 0; i < n; i++) {
        if ((flagRun[i] >= 0 || fragFlag[i]) && !(f[i].flag & SG_BAM_FLAG_DUPLICATE)) {
            const uint32_t fl = f[i].flag | SG_BAM_FLAG_DUPLICATE;
            records[off[i] + 18] = (uint8_t)fl; records[off[i] + 19] = (uint8_t)(fl >> 8);
            marked++;
        }
    }
    return marked;
}

// The .bai of the file header ‖ records wrapped into BGZF members of 0xff00 payload bytes + the end-of-file member.  Returns its size (-1: malformed, -2: bai too small).
static int64_t bam_index_impl(const uint8_t *records, int64_t nBytes, int64_t headerBytes, int32_t nRef, const uint64_t *memberOffsets, uint8_t *bai, int64_t cap);
int64_t hs_bam_index(const uint8_t *records, int64_t nBytes, int64_t headerBytes, int32_t nRef, uint8_t *bai, int64_t cap)
{
    return bam_index_impl(records, nBytes, headerBytes, nRef, nullptr, bai, cap);
}
// the same for a file whose members are compressed: memberOffsets as hs_bgzf_deflate / snapgpu_bgzf_deflate_device report them (one per member + the end)
int64_t hs_bam_index_members(const uint8_t *records, int64_t nBytes, int64_t headerBytes, int32_t nRef, const uint64_t *memberOffsets, uint8_t *bai, int64_t cap)
{
    return bam_index_impl(records, nBytes, headerBytes, nRef, memberOffsets, bai, cap);
}
static int64_t bam_index_impl(const uint8_t *records, int64_t nBytes, int64_t headerBytes, int32_t nRef, const uint64_t *memberOffsets, uint8_t *bai, int64_t cap)
{
    std::vector<unsigned long long> off;
    if (!split_records(records, nBytes, off)) return -1;
    const long long n = (long long)off.size();
    std::vector<SgBaiChunk> chunks;
    std::vector<SgBaiRef> refs((size_t)nRef);
    for (long long i = 0; i < n; i++) {
        SgBamRec r; r.p = records + off[i];
        const uint64_t at = (uint64_t)headerBytes + off[i];
        SgBamRec prev; prev.p = i ? records + off[i - 1] : nullptr;
        if (i == 0 || prev.refID() != r.refID() || prev.bin() != r.bin()) {
            if (!chunks.empty()) chunks.back().end = at;
            SgBaiChunk c; c.ref = r.refID(); c.bin = r.bin(); c.start = at; c.end = 0; chunks.push_back(c);
        }
        if (r.refID() >= 0 && r.refID() < nRef) {
            SgBaiRef &R = refs[r.refID()];
            if (!R.any) { R.any = true; R.firstStart = at; }
            R.lastEnd = at + (uint64_t)r.size();
            if (r.flag() & SG_BAM_FLAG_UNMAPPED) R.unmapped++;
            else {
                R.mapped++;
                const int32_t slot = sg_bai_linear_slot(r);
                if ((size_t)slot >= R.intervals.size()) { R.intervals.resize((size_t)slot, ~0ULL); R.intervals.push_back(at); }
            }
        }
    }
    const uint64_t total = (uint64_t)headerBytes + (uint64_t)nBytes;
    if (!chunks.empty()) chunks.back().end = total;
    std::vector<uint8_t> o = sg_bai_compose(nRef, chunks, refs, total, memberOffsets);
    if ((int64_t)o.size() > cap) return -2;
    memcpy(bai, o.data(), o.size());
    return (int64_t)o.size();
}

} // extern "C"
