/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.  Thin C-ABI shim over the UNMODIFIED reference
 * (amplab/snap 2.0.5) compiled from the sources where they lie under /root/reference by oracle/Makefile
 * into oracle/_ref/libsnapref.so.  Nothing in the product path (snap_b200/, include/) may link, load or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * It contains no algorithm of its own: every function forwards to the reference class named in its comment
 * and copies the answer into the POD structs of include/snapgpu.h so that tests can memcmp them against
 * the CUDA path.  Construction mirrors SingleAligner.cpp:133-180 (BigAllocator arena with 16-byte
 * granularity -- required).
 *
 * Must be compiled as C++98 like the reference (Makefile:2).
 */
#include "stdafx.h"
#include "Compat.h"
// SAMFormat::computeCigarString is a private static member; this harness is its only outside caller (test infrastructure)
#define private public
#include "SAM.h"
#undef private
#include "BigAlloc.h"
#include "Genome.h"
#include "GenomeIndex.h"
#include "Seed.h"
#include "SeedSequencer.h"
#include "LandauVishkin.h"
#include "AffineGapVectorized.h"
#include "BaseAligner.h"
#include "IntersectingPairedEndAligner.h"
#include "ChimericPairedEndAligner.h"
#include "AlignerOptions.h"
#include "Read.h"
#include "mapq.h"
#include "Tables.h"
#include "FASTQ.h"
#include "DataReader.h"
#include "DataWriter.h"
#include "Bam.h"

#include <pthread.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <unistd.h>
#include <sched.h>
#include <sys/syscall.h>

#include "../include/snapgpu.h"

extern GenomeIndex *g_index;                 // AlignerContext.cpp:58
extern _int64 nProbesInGetEntryForKey;       // HashTable.cpp:263 (probes beyond the first slot)

static bool g_inited = false;

extern "C" {

int ref_init(void)
{
    if (!g_inited) {
        InitializeSeedSequencers();                    // SeedSequencer.cpp:29
        initializeLVProbabilitiesToPhredPlus33();      // LandauVishkin.cpp:716
        g_inited = true;
    }
    return 0;
}

void *ref_index_load_ex(const char *dir, int map, int prefetch)
{
    ref_init();
    GenomeIndex *index = GenomeIndex::loadFromDirectory((char *)dir, map != 0, prefetch != 0);
    if (NULL != index) {
        g_index = index;
    }
    return index;
}

void *ref_index_load(const char *dir)
{
    return ref_index_load_ex(dir, /*map*/0, /*prefetch*/0);
}

/*
 * Page placement for the CPU baseline.  The index is tens of GB touched at random by every thread; written (tmpfs) or read in
 * (BigAlloc) by ONE thread it lands on that thread's NUMA node and all other sockets go through the interconnect -- measured
 * 0.67 vs 3.1 M reads/s on two boxes with the same binary.  MPOL_INTERLEAVE on the calling thread (inherited by the threads it
 * creates; obeyed by tmpfs pages it instantiates) spreads the pages round-robin, which is what `numactl --interleave=all
 * snap-aligner ...` does.  Returns the number of memory nodes interleaved over (1 = nothing to do), or -1 if the kernel refused.
 */
int ref_numa_interleave(void)
{
    int nNodes = 0;
    for (int i = 0; i < 64; i++) {
        char path[64];
        snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/meminfo", i);
        FILE *f = fopen(path, "r");
        if (!f) continue;
        fclose(f);
        nNodes = i + 1;
    }
    if (nNodes <= 1) return 1;
    unsigned long mask = 0;
    int n = 0;
    for (int i = 0; i < nNodes; i++) {
        char path[64];
        snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/meminfo", i);
        FILE *f = fopen(path, "r");
        if (!f) continue;
        // nodes without memory (MemTotal 0) must not be in the mask
        char line[256]; unsigned long long kb = 0; int node = 0;
        while (fgets(line, sizeof(line), f)) { if (2 == sscanf(line, "Node %d MemTotal: %llu", &node, &kb)) break; }
        fclose(f);
        if (kb > 0) { mask |= 1UL << i; n++; }
    }
    if (n <= 1) return 1;
    const int MPOL_INTERLEAVE_ = 3;
    long rc = syscall(SYS_set_mempolicy, MPOL_INTERLEAVE_, &mask, (unsigned long)(8 * sizeof(mask) + 1));
    return rc == 0 ? n : -1;
}

int ref_index_info(void *vidx, snapgpu_index_info *info)
{
    GenomeIndex *index = (GenomeIndex *)vidx;
    memset(info, 0, sizeof(*info));
    info->countOfBases = index->getGenome()->getCountOfBases();
    info->seedLen = index->getSeedLength();
    info->locationSize = index->doesGenomeIndexHave64BitLocations() ? 8 : 4;
    info->nContigs = index->getGenome()->getNumContigs();
    return 0;
}

/* Raw base array of the loaded genome (for building windows in leaf tests). */
const char *ref_genome_bases(void *vidx, _int64 loc, _int64 len)
{
    GenomeIndex *index = (GenomeIndex *)vidx;
    return index->getGenome()->getSubstring(loc, len);
}

/* GenomeIndex::lookupSeed32 (GenomeIndex.cpp:2095).  Returns probes beyond first slot via *extraProbes. */
int ref_lookup_seed(void *vidx, const char *bases, _int64 *nHits /*[2]*/, unsigned *hitsOut /*[2*maxOut]*/,
                    unsigned maxOut, _int64 *extraProbes)
{
    GenomeIndex *index = (GenomeIndex *)vidx;
    unsigned seedLen = index->getSeedLength();
    nHits[0] = nHits[1] = 0;
    if (!Seed::DoesTextRepresentASeed(bases, seedLen)) {
        if (extraProbes) *extraProbes = 0;
        return 0;
    }
    Seed seed(bases, seedLen);
    const unsigned *hits[2];
    _int64 before = nProbesInGetEntryForKey;
    index->lookupSeed32(seed, &nHits[0], &hits[0], &nHits[1], &hits[1]);
    if (extraProbes) *extraProbes = nProbesInGetEntryForKey - before;
    for (int d = 0; d < 2; d++) {
        for (_int64 i = 0; i < nHits[d] && i < (_int64)maxOut; i++) {
            hitsOut[d * maxOut + i] = hits[d][i];
        }
    }
    return 0;
}

unsigned ref_wrapped_seed(unsigned seedLen, unsigned wrapCount)   // SeedSequencer.cpp:105
{
    ref_init();
    return GetWrappedNextSeedToTest(seedLen, wrapCount);
}

int ref_mapq(double pAll, double pBest, int score, int popularSeedsSkipped)  // mapq.h:32
{
    return computeMAPQ(pAll, pBest, score, popularSeedsSkipped);
}

/* Copies of the global probability tables (LandauVishkin.cpp:715-763). */
void ref_tables(double *phred /*[256]*/, double *indel /*[n]*/, int nIndel, double *perfect /*[n]*/, int nPerfect)
{
    ref_init();
    for (int i = 0; i < 256; i++) phred[i] = lv_phredToProbability[i];
    for (int i = 0; i < nIndel; i++) indel[i] = lv_indelProbabilities[i];
    for (int i = 0; i < nPerfect; i++) perfect[i] = lv_perfectMatchProbability[i];
}

/*
 * LandauVishkin<dir>::computeEditDistance (LandauVishkin.h:100).  One persistent object per direction,
 * exactly like BaseAligner keeps (its L/A arrays carry state between calls).
 */
static LandauVishkin<1>  *g_lvF = NULL;
static LandauVishkin<-1> *g_lvR = NULL;

int ref_lv(int dir, const char *text, int textLen, const char *pattern, const char *quality, int patternLen, int k,
           double *matchProbability, int *netIndel, int *totalIndels, int *textSpan)
{
    ref_init();
    if (NULL == g_lvF) { g_lvF = new LandauVishkin<1>; g_lvR = new LandauVishkin<-1>; }
    if (dir == 1) {
        return g_lvF->computeEditDistance(text, textLen, pattern, quality, patternLen, k, matchProbability, netIndel, totalIndels, textSpan);
    }
    return g_lvR->computeEditDistance(text, textLen, pattern, quality, patternLen, k, matchProbability, netIndel, totalIndels, textSpan);
}

void ref_lv_batch(const char *textBuf, const char *patBuf, const char *qualBuf, const snapgpu_lv_job *jobs, _int64 nJobs,
                  snapgpu_lv_out *out)
{
    for (_int64 j = 0; j < nJobs; j++) {
        const snapgpu_lv_job *b = &jobs[j];
        snapgpu_lv_out *o = &out[j];
        o->matchProbability = 0; o->netIndel = 0; o->totalIndels = 0; o->textSpan = 0;
        o->score = ref_lv(b->dir, textBuf + b->textOff, b->textLen, patBuf + b->patOff, qualBuf + b->patOff, b->patternLen, b->k,
                          &o->matchProbability, &o->netIndel, &o->totalIndels, &o->textSpan);
    }
}

/*
 * LandauVishkinWithCigar::computeEditDistance / computeEditDistanceNormalized (LandauVishkin.cpp:141-505 / :507-650) with
 * BAM_CIGAR_OPS output: the oracle for snap_b200/csrc/sg_lv_cigar.h (output stage).  POD job / out records are
 * test-local (mirrored in tests/hostsim and oracle/reflib.py), not part of include/snapgpu.h.
 */
struct RefLvCigarJob { unsigned long long textOff, patOff; int textLen, patternLen, k, useM; };
struct RefLvCigarOut { int score, nOps, textUsed, netIndel, normalizedScore, addFrontClipping; unsigned ops[32]; };

void ref_lv_cigar_batch(const char *textBuf, const char *patBuf, const RefLvCigarJob *jobs, _int64 nJobs, RefLvCigarOut *out)
{
    static LandauVishkinWithCigar *lvc = NULL;
    if (lvc == NULL) lvc = new LandauVishkinWithCigar();
    for (_int64 j = 0; j < nJobs; j++) {
        const RefLvCigarJob *b = &jobs[j];
        RefLvCigarOut *o = &out[j];
        memset(o, 0, sizeof(*o));
        int used = 0, textUsed = 0, netIndel = 0;
        o->score = lvc->computeEditDistance(textBuf + b->textOff, b->textLen, patBuf + b->patOff, b->patternLen, b->k, (char *)o->ops, (int)sizeof(o->ops),
                                            b->useM != 0, BAM_CIGAR_OPS, &used, &textUsed, &netIndel);
        o->nOps = o->score >= 0 ? used / 4 : 0; o->textUsed = o->score >= 0 ? textUsed : 0; o->netIndel = o->score >= 0 ? netIndel : 0;
        if (o->score < 0) memset(o->ops, 0, sizeof(o->ops));
        unsigned ops2[32]; int used2 = 0, clip = 0, netIndel2 = 0;
        o->normalizedScore = lvc->computeEditDistanceNormalized(textBuf + b->textOff, b->textLen, patBuf + b->patOff, b->patternLen, b->k, (char *)ops2,
                                                                (int)sizeof(ops2), b->useM != 0, BAM_CIGAR_OPS, &used2, &clip, &netIndel2);
        o->addFrontClipping = o->normalizedScore >= 0 ? clip : 0;
    }
}

// the SAM text form of BAM operations (BAMAlignment::decodeCigar, Bam.cpp:345-375), for the known-answer vectors
int ref_decode_cigar(const unsigned *ops, int nOps, char *buf, int bufLen)
{
    return BAMAlignment::decodeCigar(buf, bufLen, (_uint32 *)ops, nOps) ? 1 : 0;
}

/*
 * SAMFormat::computeCigarString, LandauVishkinWithCigar overload (SAM.cpp:2595-2671, around computeCigar :2354-2468): the oracle
 * for snap_b200/csrc/sg_cigar.h.  kind: 0 = NULL (caller retries with addFrontClipping), 1 = "*", 2 = a CIGAR string.
 */
struct RefCigarJob { unsigned long long dataOff; long long location; int dataLength, basesClippedBefore, extraBasesClippedBefore, basesClippedAfter,
                     frontHardClipping, backHardClipping, direction, useM; };
struct RefCigarOut { int kind, editDistance, addFrontClipping, refSpan; char cigar[240]; };

void ref_cigar_lv_batch(void *vidx, const char *dataBuf, const RefCigarJob *jobs, _int64 nJobs, RefCigarOut *out)
{
    GenomeIndex *index = (GenomeIndex *)vidx;
    static LandauVishkinWithCigar *lvc = NULL;
    if (lvc == NULL) lvc = new LandauVishkinWithCigar();
    const int bufSize = MAX_READ_LENGTH * 2;
    char *cigarBuf = new char[bufSize], *withClipping = new char[bufSize + 32];
    for (_int64 j = 0; j < nJobs; j++) {
        const RefCigarJob *b = &jobs[j];
        RefCigarOut *o = &out[j];
        memset(o, 0, sizeof(*o));
        int editDistance = -1, addFrontClipping = 0, refSpan = 0;
        const char *c = SAMFormat::computeCigarString(index->getGenome(), lvc, cigarBuf, bufSize, withClipping, bufSize + 32, dataBuf + b->dataOff, b->dataLength,
                                                      (unsigned)b->basesClippedBefore, (GenomeDistance)b->extraBasesClippedBefore, (unsigned)b->basesClippedAfter,
                                                      (unsigned)b->frontHardClipping, (unsigned)b->backHardClipping, GenomeLocation(b->location),
                                                      b->direction ? RC : FORWARD, b->useM != 0, &editDistance, &addFrontClipping, &refSpan);
        o->editDistance = editDistance; o->addFrontClipping = addFrontClipping;
        if (c == NULL) { o->kind = 0; }
        else if (c[0] == '*') { o->kind = 1; }
        else { o->kind = 2; o->refSpan = refSpan; strncpy(o->cigar, c, sizeof(o->cigar) - 1); }
    }
    delete[] cigarBuf; delete[] withClipping;
}

/*
 * AffineGapVectorizedWithCigar::computeGlobalScore (AffineGapVectorized.cpp:159-518) with BAM_CIGAR_OPS output: the oracle for
 * snap_b200/csrc/sg_ag_cigar.h.  The object holds __m128i members => 64-byte aligned storage, init() instead of a constructor.
 */
struct RefAgCigarJob { unsigned long long textOff, patOff; int textLen, patternLen, w, useM; };
struct RefAgCigarOut { int score, nOps, netDel, tailIns; unsigned ops[64]; };

void ref_ag_cigar_global_batch(const int *params /* match, sub, open, extend */, const char *textBuf, const char *patBuf, const char *qualBuf,
                               const RefAgCigarJob *jobs, _int64 nJobs, RefAgCigarOut *out)
{
    static AffineGapVectorizedWithCigar *agc = NULL;
    if (agc == NULL) {
        void *m = NULL;
        if (posix_memalign(&m, 64, sizeof(AffineGapVectorizedWithCigar))) abort();
        memset(m, 0, sizeof(AffineGapVectorizedWithCigar));
        agc = (AffineGapVectorizedWithCigar *)m;
    }
    agc->init(params[0], params[1], params[2], params[3]);
    for (_int64 j = 0; j < nJobs; j++) {
        const RefAgCigarJob *b = &jobs[j];
        RefAgCigarOut *o = &out[j];
        memset(o, 0, sizeof(*o));
        int used = 0, netDel = 0, tailIns = 0;
        o->score = agc->computeGlobalScore(textBuf + b->textOff, b->textLen, patBuf + b->patOff, qualBuf + b->patOff, b->patternLen, b->w, (char *)o->ops,
                                           (int)sizeof(o->ops), b->useM != 0, BAM_CIGAR_OPS, &used, &netDel, &tailIns);
        if (o->score >= 0) { o->nOps = used / 4; o->netDel = netDel; o->tailIns = tailIns; } else memset(o->ops, 0, sizeof(o->ops));
    }
}

// computeGlobalScoreNormalized (:1043-1128): banded / unbanded dispatch + front-clipping verdict.  `out->score` = its return value.
struct RefAgCigarNormOut { int score, nOps, netDel, tailIns, addFrontClipping; unsigned ops[64]; };

void ref_ag_cigar_norm_batch(const int *params, const char *textBuf, const char *patBuf, const char *qualBuf, const RefAgCigarJob *jobs, _int64 nJobs,
                             RefAgCigarNormOut *out)
{
    static AffineGapVectorizedWithCigar *agc = NULL;
    if (agc == NULL) {
        void *m = NULL;
        if (posix_memalign(&m, 64, sizeof(AffineGapVectorizedWithCigar))) abort();
        memset(m, 0, sizeof(AffineGapVectorizedWithCigar));
        agc = (AffineGapVectorizedWithCigar *)m;
    }
    agc->init(params[0], params[1], params[2], params[3]);
    for (_int64 j = 0; j < nJobs; j++) {
        const RefAgCigarJob *b = &jobs[j];
        RefAgCigarNormOut *o = &out[j];
        memset(o, 0, sizeof(*o));
        int used = 0, netDel = 0, tailIns = 0, clip = 0;
        o->score = agc->computeGlobalScoreNormalized(textBuf + b->textOff, b->textLen, patBuf + b->patOff, qualBuf + b->patOff, b->patternLen, b->w, (char *)o->ops,
                                                     (int)sizeof(o->ops), b->useM != 0, BAM_CIGAR_OPS, &used, &clip, &netDel, &tailIns);
        o->netDel = netDel; o->tailIns = tailIns; o->addFrontClipping = clip;
        if (o->score > 0 || (o->score == 0 && clip == 0)) o->nOps = used / 4; else memset(o->ops, 0, sizeof(o->ops));
    }
}

// SAMFormat::computeCigarString, AffineGapVectorizedWithCigar overload (SAM.cpp:2677-2766 around computeCigar :2470-2592)
struct RefCigarAgJob { unsigned long long dataOff; long long location; int dataLength, basesClippedBefore, extraBasesClippedBefore, basesClippedAfter,
                       frontHardClipping, backHardClipping, direction, useM, score, pad; };

void ref_cigar_ag_batch(void *vidx, const int *params, const char *dataBuf, const char *qualBuf, const RefCigarAgJob *jobs, _int64 nJobs, RefCigarOut *out)
{
    GenomeIndex *index = (GenomeIndex *)vidx;
    static AffineGapVectorizedWithCigar *agc = NULL;
    if (agc == NULL) {
        void *m = NULL;
        if (posix_memalign(&m, 64, sizeof(AffineGapVectorizedWithCigar))) abort();
        memset(m, 0, sizeof(AffineGapVectorizedWithCigar));
        agc = (AffineGapVectorizedWithCigar *)m;
    }
    agc->init(params[0], params[1], params[2], params[3]);
    const int bufSize = MAX_READ_LENGTH * 2;
    char *cigarBuf = new char[bufSize], *withClipping = new char[bufSize + 32];
    for (_int64 j = 0; j < nJobs; j++) {
        const RefCigarAgJob *b = &jobs[j];
        RefCigarOut *o = &out[j];
        memset(o, 0, sizeof(*o));
        int editDistance = -1, addFrontClipping = 0, refSpan = 0;
        const char *c = SAMFormat::computeCigarString(index->getGenome(), agc, cigarBuf, bufSize, withClipping, bufSize + 32, dataBuf + b->dataOff, qualBuf + b->dataOff,
                                                      b->dataLength, b->score, (unsigned)b->basesClippedBefore, (GenomeDistance)b->extraBasesClippedBefore,
                                                      (unsigned)b->basesClippedAfter, (unsigned)b->frontHardClipping, (unsigned)b->backHardClipping,
                                                      GenomeLocation(b->location), b->direction ? RC : FORWARD, b->useM != 0, &editDistance, &addFrontClipping, &refSpan);
        o->editDistance = editDistance; o->addFrontClipping = addFrontClipping;
        if (c == NULL) { o->kind = 0; }
        else if (c[0] == '*') { o->kind = 1; }
        else { o->kind = 2; o->refSpan = refSpan; strncpy(o->cigar, c, sizeof(o->cigar) - 1); }
    }
    delete[] cigarBuf; delete[] withClipping;
}

/*
 * SimpleReadWriter::writeReads (ReadWriter.cpp:170-360) into memory: the oracle for sg_sam_write_single's retry loop (the loop that
 * applies the CIGAR routines' front-clipping verdicts), which real alignments hardly ever enter.  The DataWriter below hands out one
 * big buffer; the ReadWriter comes from the reference's own factory with the SAM format object.
 */
class RefMemWriter : public DataWriter {
public:
    RefMemWriter(char *b, size_t c) : DataWriter(NULL), buf(b), cap(c), used(0) {}
    virtual bool getBuffer(char **o_buffer, size_t *o_size) { *o_buffer = buf + used; *o_size = cap - used; return true; }
    virtual void advance(_int64 bytes, GenomeLocation location = 0) { used += (size_t)bytes; }
    virtual bool getBatch(int relative, char **o_buffer, size_t *o_size = NULL, size_t *o_used = NULL, size_t *o_offset = NULL, size_t *o_logicalUsed = 0,
                          size_t *o_logicalOffset = NULL) { return false; }
    virtual bool nextBatch(bool lastBatch = false) { return true; }
    virtual void close() {}
    char *buf; size_t cap, used;
};
class RefMemSupplier : public DataWriterSupplier {
public:
    RefMemSupplier(RefMemWriter *i_w) : w(i_w) {}
    virtual DataWriter *getWriter() { return w; }
    virtual void close() {}
    RefMemWriter *w;
};

_int64 ref_write_reads_batch(void *vidx, const int *agParams, int useM, int useAffineGap, _int64 n, const char *bases, const char *quals, const _uint64 *offsets,
                             const unsigned *lens, const char *ids, const _uint64 *idOffsets, const unsigned *idLens, const snapgpu_single_result *results,
                             char *out, _int64 outCap)
{
    GenomeIndex *index = (GenomeIndex *)vidx;
    RefMemWriter *mw = new RefMemWriter(out, (size_t)outCap);
    RefMemSupplier *ms = new RefMemSupplier(mw);
    ReadWriterSupplier *rws = ReadWriterSupplier::create(FileFormat::SAM[useM ? 1 : 0], ms, index->getGenome(), false, false, (char *)"", false,
                                                        agParams[0], agParams[1], agParams[2], agParams[3], false);
    ReadWriter *rw = rws->getWriter();
    ReaderContext ctx;
    memset(&ctx, 0, sizeof(ctx));
    ctx.genome = index->getGenome();
    ctx.defaultReadGroup = "FASTQ";
    static const char rg[] = "\tRG:Z:FASTQ\tPL:Z:Illumina\tPU:Z:pu\tLB:Z:lb\tSM:Z:sm";
    ctx.defaultReadGroupAux = rg; ctx.defaultReadGroupAuxLen = (int)strlen(rg);
    ctx.headerMatchesIndex = false;
    for (_int64 i = 0; i < n; i++) {
        Read read;
        read.init(ids + idOffsets[i], idLens[i], bases + offsets[i], quals + offsets[i], lens[i], NULL, 0);
        read.setReadGroup(ctx.defaultReadGroup);
        const snapgpu_single_result &r = results[i];
        SingleAlignmentResult sr;
        memset(&sr, 0, sizeof(sr));
        sr.status = (AlignmentResult)r.status; sr.location = r.status == 0 ? InvalidGenomeLocation : GenomeLocation(r.location); sr.origLocation = sr.location;
        sr.direction = r.direction ? RC : FORWARD; sr.score = r.score; sr.scorePriorToClipping = r.scorePriorToClipping; sr.mapq = r.mapq;
        sr.clippingForReadAdjustment = r.clippingForReadAdjustment; sr.usedAffineGapScoring = r.usedAffineGapScoring != 0;
        sr.basesClippedBefore = r.basesClippedBefore; sr.basesClippedAfter = r.basesClippedAfter; sr.agScore = r.agScore; sr.supplementary = false;
        if (!rw->writeReads(ctx, &read, &sr, 1, true, useAffineGap != 0)) return -1;
    }
    return (_int64)mw->used;
}

// SimpleReadWriter::writePairs (ReadWriter.cpp:362-560) into memory, same set-up as ref_write_reads_batch: the oracle for sg_sam_write_pair
_int64 ref_write_pairs_batch(void *vidx, const int *agParams, int useM, int useAffineGap, _int64 nReads, const char *bases, const char *quals, const _uint64 *offsets,
                             const unsigned *lens, const char *ids, const _uint64 *idOffsets, const unsigned *idLens, const snapgpu_paired_result *results,
                             char *out, _int64 outCap)
{
    GenomeIndex *index = (GenomeIndex *)vidx;
    RefMemWriter *mw = new RefMemWriter(out, (size_t)outCap);
    RefMemSupplier *ms = new RefMemSupplier(mw);
    ReadWriterSupplier *rws = ReadWriterSupplier::create(FileFormat::SAM[useM ? 1 : 0], ms, index->getGenome(), false, false, (char *)"", false,
                                                        agParams[0], agParams[1], agParams[2], agParams[3], false);
    ReadWriter *rw = rws->getWriter();
    ReaderContext ctx;
    memset(&ctx, 0, sizeof(ctx));
    ctx.genome = index->getGenome();
    ctx.defaultReadGroup = "FASTQ";
    static const char rg[] = "\tRG:Z:FASTQ\tPL:Z:Illumina\tPU:Z:pu\tLB:Z:lb\tSM:Z:sm";
    ctx.defaultReadGroupAux = rg; ctx.defaultReadGroupAuxLen = (int)strlen(rg);
    ctx.paired = true;
    for (_int64 i = 0; i < nReads / 2; i++) {
        Read r0, r1;
        Read *reads[2] = {&r0, &r1};
        for (int w = 0; w < 2; w++) {
            const _int64 k = 2 * i + w;
            reads[w]->init(ids + idOffsets[k], idLens[k], bases + offsets[k], quals + offsets[k], lens[k], NULL, 0);
            reads[w]->setReadGroup(ctx.defaultReadGroup);
        }
        const snapgpu_paired_result &r = results[i];
        PairedAlignmentResult pr;
        memset(&pr, 0, sizeof(pr));
        for (int w = 0; w < 2; w++) {
            pr.status[w] = (AlignmentResult)r.status[w]; pr.location[w] = r.status[w] == 0 ? InvalidGenomeLocation : GenomeLocation(r.location[w]);
            pr.origLocation[w] = pr.location[w]; pr.direction[w] = r.direction[w] ? RC : FORWARD; pr.score[w] = r.score[w];
            pr.scorePriorToClipping[w] = r.scorePriorToClipping[w]; pr.mapq[w] = r.mapq[w]; pr.clippingForReadAdjustment[w] = r.clippingForReadAdjustment[w];
            pr.usedAffineGapScoring[w] = r.usedAffineGapScoring[w] != 0; pr.basesClippedBefore[w] = r.basesClippedBefore[w]; pr.basesClippedAfter[w] = r.basesClippedAfter[w];
            pr.agScore[w] = r.agScore[w];
        }
        pr.alignedAsPair = r.alignedAsPair != 0;
        if (!rw->writePairs(ctx, reads, &pr, 1, NULL, NULL, true, useAffineGap != 0)) return -1;
    }
    return (_int64)mw->used;
}

/*
 * AffineGapVectorized<dir>::computeScore / computeScoreBanded (AffineGapVectorized.h:821 / 256).
 * The objects hold __m128i members => allocate 16-byte aligned.
 */
struct AGPair {
    AffineGapVectorized<1>  *f;
    AffineGapVectorized<-1> *r;
    snapgpu_ag_params p;
};
static AGPair g_ag = {NULL, NULL, {0, 0, 0, 0, 0, 0}};

static void ensure_ag(const snapgpu_ag_params *p)
{
    if (g_ag.f != NULL && 0 == memcmp(&g_ag.p, p, sizeof(*p))) return;
    if (g_ag.f == NULL) {
        void *m1 = NULL, *m2 = NULL;
        if (posix_memalign(&m1, 64, sizeof(AffineGapVectorized<1>)) || posix_memalign(&m2, 64, sizeof(AffineGapVectorized<-1>))) {
            fprintf(stderr, "ref_harness: posix_memalign failed\n");
            abort();
        }
        memset(m1, 0, sizeof(AffineGapVectorized<1>));
        memset(m2, 0, sizeof(AffineGapVectorized<-1>));
        g_ag.f = (AffineGapVectorized<1> *)m1;
        g_ag.r = (AffineGapVectorized<-1> *)m2;
    }
    g_ag.f->init(p->matchReward, p->subPenalty, p->gapOpenPenalty, p->gapExtendPenalty, p->fivePrimeEndBonus, p->threePrimeEndBonus);
    g_ag.r->init(p->matchReward, p->subPenalty, p->gapOpenPenalty, p->gapExtendPenalty, p->fivePrimeEndBonus, p->threePrimeEndBonus);
    g_ag.p = *p;
}

void ref_ag_batch(const snapgpu_ag_params *p, const char *textBuf, const char *patBuf, const char *qualBuf,
                  const snapgpu_ag_job *jobs, _int64 nJobs, snapgpu_ag_out *out)
{
    ref_init();
    ensure_ag(p);
    for (_int64 j = 0; j < nJobs; j++) {
        const snapgpu_ag_job *b = &jobs[j];
        snapgpu_ag_out *o = &out[j];
        o->textOffset = 0; o->patternOffset = 0; o->nEdits = 0; o->matchProbability = 0;
        const char *text = textBuf + b->textOff;
        const char *pat = patBuf + b->patOff;
        const char *qual = qualBuf + b->patOff;
        if (b->dir == 1) {
            o->agScore = b->banded
                ? g_ag.f->computeScoreBanded(text, b->textLen, pat, qual, b->patternLen, b->w, b->scoreInit, b->isRC != 0, &o->textOffset, &o->patternOffset, &o->nEdits, &o->matchProbability, b->useClippingOptimizations != 0)
                : g_ag.f->computeScore(text, b->textLen, pat, qual, b->patternLen, b->w, b->scoreInit, b->isRC != 0, &o->textOffset, &o->patternOffset, &o->nEdits, &o->matchProbability, b->useClippingOptimizations != 0);
        } else {
            o->agScore = b->banded
                ? g_ag.r->computeScoreBanded(text, b->textLen, pat, qual, b->patternLen, b->w, b->scoreInit, b->isRC != 0, &o->textOffset, &o->patternOffset, &o->nEdits, &o->matchProbability, b->useClippingOptimizations != 0)
                : g_ag.r->computeScore(text, b->textLen, pat, qual, b->patternLen, b->w, b->scoreInit, b->isRC != 0, &o->textOffset, &o->patternOffset, &o->nEdits, &o->matchProbability, b->useClippingOptimizations != 0);
        }
    }
}

/*
 * Single-end aligner, constructed exactly like SingleAligner.cpp:145-173.
 */
struct RefSingle {
    GenomeIndex *index;
    BigAllocator *allocator;
    BaseAligner *aligner;
    snapgpu_params params;
};

void *ref_single_create(void *vidx, const snapgpu_params *p)
{
    ref_init();
    GenomeIndex *index = (GenomeIndex *)vidx;
    RefSingle *rs = new RefSingle;
    rs->index = index;
    rs->params = *p;
    int maxReadSize = MAX_READ_LENGTH;
    rs->allocator = new BigAllocator(BaseAligner::getBigAllocatorReservation(index, true, p->maxHits, maxReadSize, index->getSeedLength(),
                                        p->numSeedsFromCommandLine, p->seedCoverage, -1, p->extraSearchDepth) + 4096, 16);
    DisabledOptimizations dis;
    dis.noUkkonen = p->noUkkonen != 0;
    dis.noOrderedEvaluation = p->noOrderedEvaluation != 0;
    dis.noTruncation = p->noTruncation != 0;
    dis.noEditDistance = p->noEditDistance != 0;
    dis.noBandedAffineGap = p->noBandedAffineGap != 0;
    rs->aligner = new (rs->allocator) BaseAligner(index, p->maxHits, p->maxDist, maxReadSize, p->numSeedsFromCommandLine, p->seedCoverage,
        p->minWeightToCheck, p->extraSearchDepth, dis, p->useAffineGap != 0, p->ignoreAlignmentAdjustmentsForOm != 0,
        p->altAwareness != 0, /*emitALT*/false, p->maxScoreGapToPreferNonAltAlignment, /*maxSecondaryAlignmentsPerContig*/-1,
        NULL, NULL, p->matchReward, p->subPenalty, p->gapOpenPenalty, p->gapExtendPenalty, p->fivePrimeEndBonus, p->threePrimeEndBonus,
        NULL, rs->allocator);
    rs->aligner->setExplorePopularSeeds(p->explorePopularSeeds != 0);
    rs->aligner->setStopOnFirstHit(p->stopOnFirstHit != 0);
    return rs;
}

void ref_single_destroy(void *v)
{
    RefSingle *rs = (RefSingle *)v;
    rs->aligner->~BaseAligner();
    delete rs->allocator;
    delete rs;
}

static void copy_result(const SingleAlignmentResult *r, snapgpu_single_result *o)
{
    memset(o, 0, sizeof(*o));
    o->status = (int)r->status;
    o->direction = (int)r->direction;
    o->location = GenomeLocationAsInt64(r->location);
    o->origLocation = GenomeLocationAsInt64(r->origLocation);
    o->score = r->score;
    o->scorePriorToClipping = r->scorePriorToClipping;
    o->mapq = r->mapq;
    o->clippingForReadAdjustment = r->clippingForReadAdjustment;
    o->usedAffineGapScoring = r->usedAffineGapScoring ? 1 : 0;
    o->basesClippedBefore = r->basesClippedBefore;
    o->basesClippedAfter = r->basesClippedAfter;
    o->agScore = r->agScore;
    o->supplementary = r->supplementary ? 1 : 0;
    o->seedOffset = r->seedOffset;
    o->matchProbability = r->matchProbability;
    o->probabilityAllCandidates = r->probabilityAllCandidates;
    o->popularSeedsSkipped = r->popularSeedsSkipped;
}

/*
 * The per-thread loop body of SingleAligner.cpp:197-338 without the writer: pre-filter (:213), AlignRead (:250),
 * updateStats (:354-374).  Reads are given as the already-clipped view (Read::getData()/getQuality()).
 */
static void align_range(RefSingle *rs, _int64 begin, _int64 end, const char *bases, const char *quals,
                        const _uint64 *offsets, const unsigned *lens, snapgpu_single_result *results, snapgpu_counters *ctr)
{
    const snapgpu_params *p = &rs->params;
    _int64 lv0 = rs->aligner->getLocationsScoredWithLandauVishkin();
    _int64 ag0 = rs->aligner->getLocationsScoredWithAffineGap();
    _int64 lk0 = rs->aligner->getNHashTableLookups();
    _int64 pop0 = rs->aligner->getNHitsIgnoredBecauseOfTooHighPopularity();
    for (_int64 i = begin; i < end; i++) {
        Read read;
        read.init("r", 1, bases + offsets[i], quals + offsets[i], lens[i], NULL, 0);
        SingleAlignmentResult res, alt;
        memset(&res, 0, sizeof(res));
        memset(&alt, 0, sizeof(alt));
        ctr->totalReads++;
        if (read.getDataLength() < p->minReadLength || read.countOfNs() > (int)p->maxDist) {
            res.status = NotFound;
            res.location = InvalidGenomeLocation;
            res.mapq = 0;
            res.direction = FORWARD;
            ctr->uselessReads++;
            copy_result(&res, &results[i]);
            continue;
        }
        _int64 nSecondary = 0;
        rs->aligner->AlignRead(&read, &res, &alt, p->maxSecondaryAlignmentAdditionalEditDistance, 0, &nSecondary, 0, NULL, 0, NULL, NULL);
        copy_result(&res, &results[i]);
        if (res.status == SingleHit) ctr->singleHits++;
        else if (res.status == MultipleHits) ctr->multiHits++;
        else ctr->notFound++;
        if (res.status != NotFound && res.mapq >= 0 && res.mapq <= 70) ctr->mapqHistogram[res.mapq]++;
    }
    ctr->lvCalls += rs->aligner->getLocationsScoredWithLandauVishkin() - lv0;
    ctr->affineGapCalls += rs->aligner->getLocationsScoredWithAffineGap() - ag0;
    ctr->nHashTableLookups += rs->aligner->getNHashTableLookups() - lk0;
    ctr->nHitsIgnoredBecauseOfTooHighPopularity += rs->aligner->getNHitsIgnoredBecauseOfTooHighPopularity() - pop0;
}

int ref_single_align(void *v, _int64 n, const char *bases, const char *quals, const _uint64 *offsets, const unsigned *lens,
                     snapgpu_single_result *results, snapgpu_counters *counters)
{
    RefSingle *rs = (RefSingle *)v;
    snapgpu_counters local;
    memset(&local, 0, sizeof(local));
    _int64 probes0 = nProbesInGetEntryForKey;
    align_range(rs, 0, n, bases, quals, offsets, lens, results, &local);
    // entries examined = first slot of each probe chain (2 per lookup on a small index) + extra probes
    local.nHashEntriesProbed = (nProbesInGetEntryForKey - probes0) + 2 * local.nHashTableLookups;
    if (counters) {
        _int64 *dst = (_int64 *)counters;
        const _int64 *src = (const _int64 *)&local;
        for (size_t i = 0; i < sizeof(local) / sizeof(_int64); i++) dst[i] += src[i];
    }
    return 0;
}

/*
 * The same loop with secondary alignments on (`-om`): AlignRead with a secondaryResults buffer that doubles when it overflows, exactly
 * SingleAligner.cpp:137-142 and :250-263.  The aligner must have been created with maxSecondaryAlignmentsPerContig (see
 * ref_single_create_om).  secondary[i * capacityPerRead ...] receives read i's records in buffer order; nSecondary[i] < 0 = -count when
 * they do not fit.
 */
void *ref_single_create_om(void *vidx, const snapgpu_params *p, int maxSecondaryAlignmentsPerContig)
{
    ref_init();
    GenomeIndex *index = (GenomeIndex *)vidx;
    RefSingle *rs = new RefSingle;
    rs->index = index;
    rs->params = *p;
    int maxReadSize = MAX_READ_LENGTH;
    rs->allocator = new BigAllocator(BaseAligner::getBigAllocatorReservation(index, true, p->maxHits, maxReadSize, index->getSeedLength(),
                                        p->numSeedsFromCommandLine, p->seedCoverage, maxSecondaryAlignmentsPerContig, p->extraSearchDepth) + 4096, 16);
    DisabledOptimizations dis;
    dis.noUkkonen = p->noUkkonen != 0;
    dis.noOrderedEvaluation = p->noOrderedEvaluation != 0;
    dis.noTruncation = p->noTruncation != 0;
    dis.noEditDistance = p->noEditDistance != 0;
    dis.noBandedAffineGap = p->noBandedAffineGap != 0;
    rs->aligner = new (rs->allocator) BaseAligner(index, p->maxHits, p->maxDist, maxReadSize, p->numSeedsFromCommandLine, p->seedCoverage,
        p->minWeightToCheck, p->extraSearchDepth, dis, p->useAffineGap != 0, p->ignoreAlignmentAdjustmentsForOm != 0,
        p->altAwareness != 0, /*emitALT*/false, p->maxScoreGapToPreferNonAltAlignment, maxSecondaryAlignmentsPerContig,
        NULL, NULL, p->matchReward, p->subPenalty, p->gapOpenPenalty, p->gapExtendPenalty, p->fivePrimeEndBonus, p->threePrimeEndBonus,
        NULL, rs->allocator);
    rs->aligner->setExplorePopularSeeds(p->explorePopularSeeds != 0);
    rs->aligner->setStopOnFirstHit(p->stopOnFirstHit != 0);
    return rs;
}

int ref_single_align_om(void *v, _int64 n, const char *bases, const char *quals, const _uint64 *offsets, const unsigned *lens,
                        snapgpu_single_result *results, int maxSecondaryAlignments, _int64 capacityPerRead, snapgpu_single_result *secondary,
                        int *nSecondaryOut, snapgpu_counters *counters)
{
    RefSingle *rs = (RefSingle *)v;
    const snapgpu_params *p = &rs->params;
    g_index = rs->index;            // SingleAlignmentResult::compareByContigAndScore reads it (AlignmentResult.cpp:31)
    snapgpu_counters local;
    memset(&local, 0, sizeof(local));
    _int64 bufferCount = 32;
    SingleAlignmentResult *buf = (SingleAlignmentResult *)BigAlloc(sizeof(SingleAlignmentResult) * bufferCount);
    for (_int64 i = 0; i < n; i++) {
        Read read;
        read.init("r", 1, bases + offsets[i], quals + offsets[i], lens[i], NULL, 0);
        SingleAlignmentResult alt;
        memset(&alt, 0, sizeof(alt));
        memset(buf, 0, sizeof(SingleAlignmentResult) * bufferCount);
        nSecondaryOut[i] = 0;
        local.totalReads++;
        if (read.getDataLength() < p->minReadLength || read.countOfNs() > (int)p->maxDist) {
            buf[0].status = NotFound;
            buf[0].location = InvalidGenomeLocation;
            buf[0].mapq = 0;
            buf[0].direction = FORWARD;
            local.uselessReads++;
            copy_result(&buf[0], &results[i]);
            continue;
        }
        _int64 nSecondary = 0;
        while (!rs->aligner->AlignRead(&read, buf, &alt, p->maxSecondaryAlignmentAdditionalEditDistance, bufferCount - 1, &nSecondary, maxSecondaryAlignments,
                                       buf + 1, 0, NULL, NULL)) {
            BigDealloc(buf);
            bufferCount *= 2;
            buf = (SingleAlignmentResult *)BigAlloc(sizeof(SingleAlignmentResult) * bufferCount);
            memset(buf, 0, sizeof(SingleAlignmentResult) * bufferCount);
        }
        copy_result(&buf[0], &results[i]);
        if (nSecondary <= capacityPerRead) {
            for (_int64 k = 0; k < nSecondary; k++) copy_result(&buf[1 + k], &secondary[i * capacityPerRead + k]);
            nSecondaryOut[i] = (int)nSecondary;
        } else {
            nSecondaryOut[i] = -(int)nSecondary;
        }
        if (buf[0].status == SingleHit) local.singleHits++;
        else if (buf[0].status == MultipleHits) local.multiHits++;
        else local.notFound++;
        if (buf[0].status != NotFound && buf[0].mapq >= 0 && buf[0].mapq <= 70) local.mapqHistogram[buf[0].mapq]++;
    }
    BigDealloc(buf);
    if (counters) {
        _int64 *dst = (_int64 *)counters;
        const _int64 *src = (const _int64 *)&local;
        for (size_t i = 0; i < sizeof(local) / sizeof(_int64); i++) dst[i] += src[i];
    }
    return 0;
}

/*
 * Multi-threaded variant: nThreads pthreads, each with its own BaseAligner over a contiguous range, the
 * way ParallelTask.h:40-120 runs SingleAlignerContext::runIterationThread.  Used for the CPU baseline.
 * Returns wall seconds spent aligning (aligner construction excluded, like AlignerContext.cpp:420).
 */
/* Thread pinning for the multi-threaded baselines (stock SNAP's -b): thread t -> logical CPU t, so that a 32-thread run sits on the
 * physical cores of one socket and a 64-thread run on the physical cores of both (the usual Linux numbering: SMT siblings last). */
static int g_pin_threads = 0;
void ref_set_thread_pinning(int on) { g_pin_threads = on; }

static void pin_self(int t)
{
    if (!g_pin_threads) return;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET((int)(t % n), &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
}

struct MTArg {
    RefSingle *rs; _int64 begin, end; int threadNo;
    const char *bases; const char *quals; const _uint64 *offsets; const unsigned *lens;
    snapgpu_single_result *results; snapgpu_counters ctr;
    pthread_barrier_t *start; int reps;
};

static void *mt_main(void *v)
{
    MTArg *a = (MTArg *)v;
    pin_self(a->threadNo);
    pthread_barrier_wait(a->start);              // the clock starts when every thread exists and is waiting here
    for (int r = 0; r < a->reps; r++) {
        if (r > 0) memset(&a->ctr, 0, sizeof(a->ctr));
        align_range(a->rs, a->begin, a->end, a->bases, a->quals, a->offsets, a->lens, a->results, &a->ctr);
    }
    return NULL;
}

double ref_single_align_mt_reps(void *vidx, const snapgpu_params *p, int nThreads, int reps, _int64 n, const char *bases, const char *quals,
                                const _uint64 *offsets, const unsigned *lens, snapgpu_single_result *results, snapgpu_counters *counters)
{
    if (nThreads < 1) nThreads = 1;
    if (reps < 1) reps = 1;
    MTArg *args = new MTArg[nThreads];
    pthread_t *threads = new pthread_t[nThreads];
    pthread_barrier_t start;
    pthread_barrier_init(&start, NULL, nThreads + 1);
    for (int t = 0; t < nThreads; t++) {
        args[t].start = &start; args[t].reps = reps; args[t].threadNo = t;
        args[t].rs = (RefSingle *)ref_single_create(vidx, p);
        args[t].begin = n * t / nThreads;
        args[t].end = n * (t + 1) / nThreads;
        args[t].bases = bases; args[t].quals = quals; args[t].offsets = offsets; args[t].lens = lens;
        args[t].results = results;
        memset(&args[t].ctr, 0, sizeof(args[t].ctr));
    }
    struct timespec t0, t1;
    for (int t = 0; t < nThreads; t++) pthread_create(&threads[t], NULL, mt_main, &args[t]);
    pthread_barrier_wait(&start);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nThreads; t++) pthread_join(threads[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    pthread_barrier_destroy(&start);
    for (int t = 0; t < nThreads; t++) {
        if (counters) {
            _int64 *dst = (_int64 *)counters;
            const _int64 *src = (const _int64 *)&args[t].ctr;
            for (size_t i = 0; i < sizeof(snapgpu_counters) / sizeof(_int64); i++) dst[i] += src[i];
        }
        ref_single_destroy(args[t].rs);
    }
    delete[] args;
    delete[] threads;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

double ref_single_align_mt(void *vidx, const snapgpu_params *p, int nThreads, _int64 n, const char *bases, const char *quals,
                           const _uint64 *offsets, const unsigned *lens, snapgpu_single_result *results, snapgpu_counters *counters)
{
    return ref_single_align_mt_reps(vidx, p, nThreads, 1, n, bases, quals, offsets, lens, results, counters);
}

/*
 * Paired-end stack, constructed exactly like PairedAligner.cpp:547-638: one arena holding the
 * IntersectingPairedEndAligner, the ChimericPairedEndAligner (with its single-end BaseAligner) and the result /
 * candidate buffers; align() called with the arguments of PairedAligner.cpp:727-730 (-om unset).
 */
struct RefPaired {
    GenomeIndex *index;
    BigAllocator *allocator;
    IntersectingPairedEndAligner *intersecting;
    ChimericPairedEndAligner *aligner;
    PairedAlignmentResult *results;
    PairedAlignmentResult *pairedCandidates;
    SingleAlignmentResult *singleCandidates;
    SingleAlignmentResult *singleSecondary;
    _int64 maxPairedCandidates, maxSingleCandidates;
    snapgpu_params params;
    snapgpu_paired_params pp;
};

void *ref_paired_create(void *vidx, const snapgpu_params *p, const snapgpu_paired_params *pp)
{
    ref_init();
    GenomeIndex *index = (GenomeIndex *)vidx;
    RefPaired *rp = new RefPaired;
    rp->index = index; rp->params = *p; rp->pp = *pp;
    int maxReadSize = MAX_READ_LENGTH;
    size_t memoryPoolSize = IntersectingPairedEndAligner::getBigAllocatorReservation(index, pp->intersectingAlignerMaxHits, maxReadSize, index->getSeedLength(),
                                p->numSeedsFromCommandLine, p->seedCoverage, MAX_K, p->extraSearchDepth, pp->maxCandidatePoolSize, -1);
    memoryPoolSize += ChimericPairedEndAligner::getBigAllocatorReservation(index, maxReadSize, p->maxHits, index->getSeedLength(), pp->maxSeedsSingleEnd, p->seedCoverage,
                                MAX_K, p->extraSearchDepth, pp->maxCandidatePoolSize, -1);
    rp->maxPairedCandidates = p->useAffineGap ? 4096 : 0;
    rp->maxSingleCandidates = p->useAffineGap ? 4096 : 0;
    memoryPoolSize += (1 + rp->maxPairedCandidates) * sizeof(PairedAlignmentResult) + (rp->maxSingleCandidates) * sizeof(SingleAlignmentResult) + 65536;
    rp->allocator = new BigAllocator(memoryPoolSize, 16);
    DisabledOptimizations dis;
    dis.noUkkonen = p->noUkkonen != 0; dis.noOrderedEvaluation = p->noOrderedEvaluation != 0; dis.noTruncation = p->noTruncation != 0;
    dis.noEditDistance = p->noEditDistance != 0; dis.noBandedAffineGap = p->noBandedAffineGap != 0;
    rp->intersecting = new (rp->allocator) IntersectingPairedEndAligner(index, maxReadSize, p->maxHits, p->maxDist, pp->maxDistForIndels, p->numSeedsFromCommandLine,
        p->seedCoverage, pp->minSpacing, pp->maxSpacing, pp->intersectingAlignerMaxHits, p->extraSearchDepth, pp->maxCandidatePoolSize, -1, rp->allocator, dis,
        p->useAffineGap != 0, p->ignoreAlignmentAdjustmentsForOm != 0, p->altAwareness != 0, p->maxScoreGapToPreferNonAltAlignment,
        p->matchReward, p->subPenalty, p->gapOpenPenalty, p->gapExtendPenalty, pp->useSoftClipping != 0);
    rp->aligner = new (rp->allocator) ChimericPairedEndAligner(index, maxReadSize, p->maxHits, p->maxDist, pp->maxSeedsSingleEnd, p->seedCoverage, p->minWeightToCheck,
        pp->forceSpacing != 0, p->extraSearchDepth, dis, p->useAffineGap != 0, p->ignoreAlignmentAdjustmentsForOm != 0, p->altAwareness != 0, /*emitALT*/false,
        rp->intersecting, p->minReadLength, -1, p->maxScoreGapToPreferNonAltAlignment, pp->flattenMAPQAtOrBelow, pp->useSoftClipping != 0,
        p->matchReward, p->subPenalty, p->gapOpenPenalty, p->gapExtendPenalty, p->fivePrimeEndBonus, p->threePrimeEndBonus,
        pp->minScoreRealignment, pp->minScoreGapRealignmentALT, pp->minAGScoreImprovement, pp->enableHammingScoringBaseAligner != 0, rp->allocator);
    rp->results = (PairedAlignmentResult *)rp->allocator->allocate(sizeof(PairedAlignmentResult));
    rp->singleSecondary = (SingleAlignmentResult *)rp->allocator->allocate(16);
    rp->pairedCandidates = NULL; rp->singleCandidates = NULL;
    if (p->useAffineGap) {
        rp->pairedCandidates = (PairedAlignmentResult *)rp->allocator->allocate(rp->maxPairedCandidates * sizeof(PairedAlignmentResult));
        rp->singleCandidates = (SingleAlignmentResult *)rp->allocator->allocate(rp->maxSingleCandidates * sizeof(SingleAlignmentResult));
    }
    return rp;
}

void ref_paired_destroy(void *v)
{
    RefPaired *rp = (RefPaired *)v;
    delete rp->allocator;
    delete rp;
}

static void copy_paired(const PairedAlignmentResult *r, snapgpu_paired_result *o)
{
    memset(o, 0, sizeof(*o));
    for (int i = 0; i < 2; i++) {
        o->status[i] = (int)r->status[i]; o->direction[i] = (int)r->direction[i];
        o->location[i] = GenomeLocationAsInt64(r->location[i]); o->origLocation[i] = GenomeLocationAsInt64(r->origLocation[i]);
        o->score[i] = r->score[i]; o->scorePriorToClipping[i] = r->scorePriorToClipping[i]; o->mapq[i] = r->mapq[i];
        o->clippingForReadAdjustment[i] = r->clippingForReadAdjustment[i]; o->usedAffineGapScoring[i] = r->usedAffineGapScoring[i] ? 1 : 0;
        o->basesClippedBefore[i] = r->basesClippedBefore[i]; o->basesClippedAfter[i] = r->basesClippedAfter[i]; o->agScore[i] = r->agScore[i];
        o->supplementary[i] = r->supplementary[i] ? 1 : 0; o->seedOffset[i] = r->seedOffset[i]; o->lvIndels[i] = r->lvIndels[i];
        o->usedGaplessClipping[i] = r->usedGaplessClipping[i] ? 1 : 0; o->refSpan[i] = r->refSpan[i]; o->liftover[i] = r->liftover[i] ? 1 : 0;
        o->popularSeedsSkipped[i] = r->popularSeedsSkipped[i]; o->matchProbability[i] = r->matchProbability[i];
    }
    o->alignedAsPair = r->alignedAsPair ? 1 : 0; o->agForcedSingleAlignerCall = r->agForcedSingleAlignerCall ? 1 : 0;
    o->probabilityAllPairs = r->probabilityAllPairs;
}

/*
 * The per-thread loop body of PairedAligner.cpp:676-790 without the writer.  Pair i = reads 2i and 2i+1 of the batch.
 * nLV / nAG (optional): locations scored, summed over the call.
 */
int ref_paired_align(void *v, _int64 nPairs, const char *bases, const char *quals, const _uint64 *offsets, const unsigned *lens,
                     snapgpu_paired_result *results, _int64 *nLV, _int64 *nAG)
{
    RefPaired *rp = (RefPaired *)v;
    const snapgpu_params *p = &rp->params;
    _int64 lv0 = rp->intersecting->getLocationsScoredWithLandauVishkin(), ag0 = rp->intersecting->getLocationsScoredWithAffineGap();
    for (_int64 i = 0; i < nPairs; i++) {
        Read r0, r1;
        r0.init("r/1", 3, bases + offsets[2 * i], quals + offsets[2 * i], lens[2 * i], NULL, 0);
        r1.init("r/2", 3, bases + offsets[2 * i + 1], quals + offsets[2 * i + 1], lens[2 * i + 1], NULL, 0);
        bool useful0 = r0.getDataLength() >= p->minReadLength && (int)r0.countOfNs() <= (int)p->maxDist;
        bool useful1 = r1.getDataLength() >= p->minReadLength && (int)r1.countOfNs() <= (int)p->maxDist;
        PairedAlignmentResult *res = rp->results;
        memset(res, 0, sizeof(*res));
        if (!useful0 && !useful1) {
            res->status[0] = res->status[1] = NotFound;
            res->location[0] = res->location[1] = InvalidGenomeLocation;
            copy_paired(res, &results[i]);
            continue;
        }
        PairedAlignmentResult firstALT;
        memset(&firstALT, 0, sizeof(firstALT));
        _int64 nSecondary = 0, nPairedCand = 0, nSingleSecondary[2] = {0, 0}, nSingleCand[2] = {0, 0};
        bool ok = rp->aligner->align(&r0, &r1, res, &firstALT, p->maxSecondaryAlignmentAdditionalEditDistance, 0, &nSecondary, res + 1,
            0, 0, &nSingleSecondary[0], &nSingleSecondary[1], rp->singleSecondary,
            rp->maxPairedCandidates, &nPairedCand, rp->pairedCandidates, rp->maxSingleCandidates, &nSingleCand[0], &nSingleCand[1], rp->singleCandidates, p->maxDist);
        if (!ok) {
            fprintf(stderr, "ref_paired_align: candidate buffer overflow on pair %lld (not handled by the harness)\n", (long long)i);
            return 1;
        }
        copy_paired(res, &results[i]);
    }
    if (nLV) *nLV += rp->intersecting->getLocationsScoredWithLandauVishkin() - lv0;
    if (nAG) *nAG += rp->intersecting->getLocationsScoredWithAffineGap() - ag0;
    return 0;
}


/* Multi-threaded paired variant for the CPU baseline: one ChimericPairedEndAligner stack per thread over a contiguous
 * range of pairs (ParallelTask.h:40-120 / PairedAligner.cpp:520-800).  Returns wall seconds spent aligning. */
struct MTPairedArg {
    void *rp; _int64 begin, end;
    const char *bases; const char *quals; const _uint64 *offsets; const unsigned *lens;
    snapgpu_paired_result *results; _int64 nLV, nAG; int rc;
    pthread_barrier_t *start; int reps; int threadNo;
};

static void *mt_paired_main(void *v)
{
    MTPairedArg *a = (MTPairedArg *)v;
    pin_self(a->threadNo);
    pthread_barrier_wait(a->start);
    for (int r = 0; r < a->reps && !a->rc; r++) {
        a->nLV = a->nAG = 0;
        a->rc = ref_paired_align(a->rp, a->end - a->begin, a->bases, a->quals, a->offsets + 2 * a->begin, a->lens + 2 * a->begin,
                                 a->results + a->begin, &a->nLV, &a->nAG);
    }
    return NULL;
}

double ref_paired_align_mt_reps(void *vidx, const snapgpu_params *p, const snapgpu_paired_params *pp, int nThreads, int reps, _int64 nPairs, const char *bases,
                                const char *quals, const _uint64 *offsets, const unsigned *lens, snapgpu_paired_result *results, _int64 *nLV, _int64 *nAG)
{
    if (nThreads < 1) nThreads = 1;
    if (reps < 1) reps = 1;
    MTPairedArg *args = new MTPairedArg[nThreads];
    pthread_t *threads = new pthread_t[nThreads];
    pthread_barrier_t start;
    pthread_barrier_init(&start, NULL, nThreads + 1);
    for (int t = 0; t < nThreads; t++) {
        args[t].start = &start; args[t].reps = reps; args[t].threadNo = t;
        args[t].rp = ref_paired_create(vidx, p, pp);
        args[t].begin = nPairs * t / nThreads;
        args[t].end = nPairs * (t + 1) / nThreads;
        args[t].bases = bases; args[t].quals = quals; args[t].offsets = offsets; args[t].lens = lens;
        args[t].results = results; args[t].nLV = args[t].nAG = 0; args[t].rc = 0;
    }
    struct timespec t0, t1;
    for (int t = 0; t < nThreads; t++) pthread_create(&threads[t], NULL, mt_paired_main, &args[t]);
    pthread_barrier_wait(&start);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nThreads; t++) pthread_join(threads[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    pthread_barrier_destroy(&start);
    double secs = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    for (int t = 0; t < nThreads; t++) {
        if (nLV) *nLV += args[t].nLV;
        if (nAG) *nAG += args[t].nAG;
        if (args[t].rc) secs = -1.0;
        ref_paired_destroy(args[t].rp);
    }
    delete[] args;
    delete[] threads;
    return secs;
}

double ref_paired_align_mt(void *vidx, const snapgpu_params *p, const snapgpu_paired_params *pp, int nThreads, _int64 nPairs, const char *bases,
                           const char *quals, const _uint64 *offsets, const unsigned *lens, snapgpu_paired_result *results, _int64 *nLV, _int64 *nAG)
{
    return ref_paired_align_mt_reps(vidx, p, pp, nThreads, 1, nPairs, bases, quals, offsets, lens, results, nLV, nAG);
}


/*
 * FASTQ ingest: FASTQReader::getReadFromBuffer (FASTQ.cpp:229-300) + Read::init upper-casing (Read.h:465-492)
 * + Read::clip (Read.h:567-619) over a whole in-memory buffer.  The parser only asks its DataReader for error context and the
 * batch id, so a stub that answers "end of file" is enough.  `buf` must be readable one byte past nBytes.
 * Outputs per read: clipped bases / qualities copied back to back, their offset and length, the id's offset and length in
 * `buf`, the number of bases clipped from the front.  Returns the bytes consumed.
 */
class RefStubDataReader : public DataReader {
public:
    virtual bool init(const char *) { return true; }
    virtual char *readHeader(_int64 *) { return NULL; }
    virtual void reinit(_int64, _int64) {}
    virtual bool getData(char **, _int64 *, _int64 *) { return false; }
    virtual void advance(_int64) {}
    virtual void nextBatch() {}
    virtual bool isEOF() { return true; }
    virtual DataBatch getBatch() { return DataBatch(); }
    virtual void holdBatch(DataBatch) {}
    virtual bool releaseBatch(DataBatch) { return true; }
    virtual _int64 getFileOffset() { return 0; }
    virtual void getExtra(char **, _int64 *) {}
    virtual const char *getFilename() { return "buffer"; }
};

_int64 ref_fastq_parse(char *buf, _int64 nBytes, int clippingType, _int64 maxReads, _int64 *nReads, char *basesOut, char *qualsOut,
                       _uint64 *outOff, unsigned *outLen, _uint64 *idOff, unsigned *idLen, unsigned *frontClipped)
{
    ref_init();
    RefStubDataReader stub;
    ReaderContext context;
    memset(&context, 0, sizeof(context));
    context.clipping = ReadClippingType((ClippingType)clippingType);
    context.preserveFASTQComments = false;
    _int64 pos = 0, n = 0;
    _uint64 out = 0;
    while (pos < nBytes && n < maxReads) {
        // stop at an incomplete trailing record instead of letting the parser exit the process: count the newlines that are left
        int nl = 0;
        for (_int64 k = pos; k < nBytes && nl < 4; k++) nl += (buf[k] == '\n');
        if (nl < 4) break;
        Read read;
        _int64 consumed = FASTQReader::getReadFromBuffer(buf + pos, nBytes - pos, &read, "buffer", &stub, context);
        if (consumed == 0) break;
        memcpy(basesOut + out, read.getData(), read.getDataLength());
        memcpy(qualsOut + out, read.getQuality(), read.getDataLength());
        outOff[n] = out; outLen[n] = read.getDataLength();
        idOff[n] = (_uint64)(read.getId() - buf); idLen[n] = read.getIdLength();
        frontClipped[n] = read.getFrontClippedLength();
        out += read.getDataLength();
        pos += consumed;
        n++;
    }
    *nReads = n;
    return pos;
}

} // extern "C"
