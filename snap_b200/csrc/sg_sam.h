// sg_sam.h -- one SAM record of an unpaired read: SimpleReadWriter::writeReads' per-result loop (reference SNAPLib/ReadWriter.cpp:
// 170-330) around SAMFormat::writeRead (SAM.cpp:1897-2112 for Landau-Vishkin results, :2113-2352 for affine-gap ones) and
// SAMFormat::createSAMLine (:1423-1573).  Fourth piece of the output stage; same status as sg_cigar.h: verified on
// the host (tests/test_output_stage.py: against the SAM file the reference binary writes for the same reads), no device entry point
// yet, nothing in include/snapgpu.h refers to it.  Primary alignments only (no secondary results), default tags (PG, NM, the default
// read group line, QS for pairs); sg_sam_write_pair adds SimpleReadWriter::writePairs (ReadWriter.cpp:362-560) around
// SAMFormat::writePairs (SAM.cpp:1574-1896) and fillMateInfo (:1308-1422).
//
// The loop: format the record; if the CIGAR routine answers with a front-clipping verdict (a leading deletion / insertion, or a
// read that starts before its contig) move the alignment or clip the read and format again; give the read up (unmapped) when that
// would cross a contig boundary or does not settle.
#pragma once
#include "sg_cigar.h"

// CIGAR operations one record can carry: an alignment with e edits has at most 2e + 1 operations (e <= MAX_K - 1 = 126), plus two
// soft / hard clips on either side.  (The reference formats into a 2000-character buffer, which is never the limit.)
#define SG_SAM_MAX_OPS 264

struct SgSamRead {                   // the Read object's view of one read (Read.h:412-560)
    const uint8_t *unclippedData, *unclippedQuality;
    uint32_t unclippedLength;
    uint32_t frontClipped;           // getFrontClippedLength(): quality clipping + additional front clipping
    uint32_t dataLength;             // getDataLength(): the clipped view's length
    const uint8_t *id; uint32_t idLength;
    int additionalFrontClipping, additionalBackClipping;
    SG_HD void setAdditionalFrontClipping(int c) { frontClipped += (uint32_t)(c - additionalFrontClipping); dataLength -= (uint32_t)(c - additionalFrontClipping); additionalFrontClipping = c; }
    SG_HD void setAdditionalBackClipping(int c) { dataLength -= (uint32_t)(c - additionalBackClipping); additionalBackClipping = c; }
};

// What SimpleReadWriter hands DataWriter::advance for each record it wrote (ReadWriter.cpp:331, :605-615): its bytes and the genome location it
// is filed under by a sorting writer (SortedDataFilter::onAdvance, SortedDataWriter.cpp:905-939) -- the record's own final location, or its
// mate's when it is unaligned itself; SG_SORT_UNALIGNED when neither is.  Row N4 of SURVEY 8(f) sorts on these.
#define SG_SORT_UNALIGNED (-1LL)
struct SgSortInfo { int64_t location[2]; uint32_t bytes[2]; int nRecords; };

struct SgSamContext {
    const SgIndexView *ix;
    const char *const *contigName;   // [nContigs], NUL-terminated
    SgAgParams ag;
    const char *readGroupAux;        // ReaderContext::defaultReadGroupAux, e.g. "\tRG:Z:FASTQ\tPL:Z:Illumina\tPU:Z:pu\tLB:Z:lb\tSM:Z:sm"
    bool useM, useAffineGap;
    SgLvCigarScratch lv;
    SgAgCigarScratch agS;
    uint8_t *data, *quality;         // [maxReadLen] scratch for the oriented read
    uint8_t *data2, *quality2;       // the same for the second read of a pair
    SgSortInfo *sort = (SgSortInfo *)0;   // NULL, or where the writers leave the sort keys of the records of this call
};

struct SgSamResult {                 // the SingleAlignmentResult fields the writer reads (ReadWriter.cpp:223-310)
    int status; int64_t location; int direction, mapq, score, scorePriorToClipping, usedAffineGapScoring, basesClippedBefore, basesClippedAfter,
        clippingForReadAdjustment;
};

SG_HD char *sg_put_str(char *p, const char *s) { while (*s) *p++ = *s++; return p; }
SG_HD char *sg_put_i64(char *p, long long v)
{
    char tmp[24]; int n = 0;
    unsigned long long u = v < 0 ? (unsigned long long)(-(v + 1)) + 1ULL : (unsigned long long)v;
    if (v < 0) *p++ = '-';
    do { tmp[n++] = (char)('0' + (int)(u % 10)); u /= 10; } while (u);
    while (n) *p++ = tmp[--n];
    return p;
}

// Genome::getContigForRead (Genome.cpp:734-758): the contig of `location`, or the next one when the read starts before it
SG_HD int sg_contig_for_read(const SgIndexView &ix, int64_t location, uint32_t readLength, int64_t *extraBasesClippedBefore)
{
    int c = sg_contig_at(ix, location);
    const int64_t end = (c < 0) ? 0 : ((c == (int)ix.nContigs - 1) ? ix.nBases : ix.contigStart[c + 1]);
    if (c < 0 || location + readLength > end) {
        // getNextContigAfterLocation (Genome.cpp:603-634)
        int next = (ix.nContigs > 0 && location < ix.contigStart[0]) ? 0 : c + 1;
        if (next >= (int)ix.nContigs) next = (int)ix.nContigs - 1;
        *extraBasesClippedBefore = ix.contigStart[next] - location;
        return next;
    }
    *extraBasesClippedBefore = 0;
    return c;
}

// SAMFormat::writeRead (either overload) for a primary, unpaired record.  Returns the record's length, or 0 with *addFrontClipping != 0.
SG_HDN int sg_sam_format(const SgSamContext &C, const SgSamRead &R, int status, int mapQuality, int64_t genomeLocation, int direction, bool affineGap, int score,
                         int bpClippedBefore, int bpClippedAfter, char *out, int *addFrontClipping)
{
    *addFrontClipping = 0;
    const SgIndexView &ix = *C.ix;
    int flags = 0;
    // ---- createSAMLine (:1470-1572) ----
    if (status == SNAPGPU_NOT_FOUND) genomeLocation = -1;
    const bool mapped = genomeLocation != -1;
    if (!mapped) direction = SNAPGPU_FORWARD;
    uint32_t clippedLength = R.dataLength;
    const uint32_t fullLength = R.unclippedLength;
    uint32_t basesClippedBefore, basesClippedAfter;
    const uint8_t *clippedData, *clippedQuality;
    if (direction == SNAPGPU_RC) {
        for (uint32_t i = 0; i < fullLength; i++) {
            C.data[fullLength - 1 - i] = sg_complement(R.unclippedData[i]);
            C.quality[fullLength - 1 - i] = R.unclippedQuality[i];
        }
        clippedData = &C.data[fullLength - clippedLength - R.frontClipped];
        clippedQuality = &C.quality[fullLength - clippedLength - R.frontClipped];
        basesClippedBefore = fullLength - clippedLength - R.frontClipped;
        basesClippedAfter = R.frontClipped;
    } else {
        for (uint32_t i = 0; i < fullLength; i++) { C.data[i] = R.unclippedData[i]; C.quality[i] = R.unclippedQuality[i]; }
        clippedData = C.data + R.frontClipped;
        clippedQuality = C.quality + R.frontClipped;
        basesClippedBefore = R.frontClipped;
        basesClippedAfter = fullLength - clippedLength - basesClippedBefore;
    }
    basesClippedBefore += (uint32_t)bpClippedBefore; basesClippedAfter += (uint32_t)bpClippedAfter;
    clippedData += bpClippedBefore; clippedQuality += bpClippedBefore;
    clippedLength -= (uint32_t)(bpClippedBefore + bpClippedAfter);
    int64_t extraBasesClippedBefore = 0, positionInContig = 0;
    const char *contigName = "*";
    if (mapped) {
        if (direction == SNAPGPU_RC) flags |= 0x10;
        const int c = sg_contig_for_read(ix, genomeLocation, R.dataLength, &extraBasesClippedBefore);
        genomeLocation += extraBasesClippedBefore;
        contigName = C.contigName[c];
        positionInContig = genomeLocation - ix.contigStart[c] + 1;
        mapQuality = mapQuality < 0 ? 0 : (mapQuality > 70 ? 70 : mapQuality);
    } else {
        flags |= 0x4;
        mapQuality = 0;
    }
    // ---- writeRead: the CIGAR (:1976-1983 / :2194-2204) ----
    uint32_t ops[SG_SAM_MAX_OPS];
    SgCigarOut co;
    co.kind = 1; co.editDistance = -1; co.nOps = 0;
    int editDistance = -1;
    if (affineGap && extraBasesClippedBefore != 0) { *addFrontClipping = (int)extraBasesClippedBefore; return 0; }
    if (mapped) {
        // (createSAMLine already moved genomeLocation; computeCigarString adds extraBasesClippedBefore again to the location it is given,
        //  which is the caller's original one: pass that)
        const int64_t locForCigar = genomeLocation - extraBasesClippedBefore;
        if (affineGap) sg_cigar_ag(ix, C.ag, C.agS, clippedData, clippedQuality, clippedLength, score, basesClippedBefore, extraBasesClippedBefore, basesClippedAfter, 0, 0,
                                   locForCigar, C.useM, ops, SG_SAM_MAX_OPS, &co);
        else sg_cigar_lv(ix, C.lv, clippedData, clippedLength, basesClippedBefore, extraBasesClippedBefore, basesClippedAfter, 0, 0, locForCigar, C.useM, ops, SG_SAM_MAX_OPS, &co);
        editDistance = co.editDistance;
        if (co.addFrontClipping != 0) { *addFrontClipping = co.addFrontClipping; return 0; }
    }
    // ---- the text (:2074-2096) ----
    char *p = out;
    uint32_t qnameLen = R.idLength;
    for (uint32_t i = 0; i < qnameLen; i++) if (R.id[i] == ' ') { qnameLen = i; break; }
    for (uint32_t i = 0; i < qnameLen; i++) *p++ = (char)R.id[i];
    *p++ = '\t'; p = sg_put_i64(p, flags);
    *p++ = '\t'; p = sg_put_str(p, contigName);
    *p++ = '\t'; p = sg_put_i64(p, positionInContig);
    *p++ = '\t'; p = sg_put_i64(p, mapQuality);
    *p++ = '\t';
    if (mapped && co.kind == 2) {
        for (int i = 0; i < co.nOps; i++) { p = sg_put_i64(p, (long long)(ops[i] >> 4)); *p++ = "MIDNSHP=X"[ops[i] & 15]; }
    } else {
        *p++ = '*';
    }
    p = sg_put_str(p, "\t*\t0\t0\t");
    for (uint32_t i = 0; i < fullLength; i++) *p++ = (char)C.data[i];
    *p++ = '\t';
    for (uint32_t i = 0; i < fullLength; i++) *p++ = (char)C.quality[i];
    p = sg_put_str(p, "\tPG:Z:SNAP\tNM:i:"); p = sg_put_i64(p, editDistance);
    p = sg_put_str(p, C.readGroupAux);
    *p++ = '\n';
    return (int)(p - out);
}

// the per-result loop of SimpleReadWriter::writeReads (ReadWriter.cpp:223-310) for one primary result.  R and res are working copies.
SG_HDN int sg_sam_write_single(const SgSamContext &C, SgSamRead R, SgSamResult res, char *out)
{
    const SgIndexView &ix = *C.ix;
    int addFrontClipping = 0;
    R.setAdditionalFrontClipping(res.clippingForReadAdjustment);
    int cumulativeAddFrontClipping = 0;
    int64_t finalLocation = res.location;
    unsigned nAdjustments = 0;
    const bool affineGap = C.useAffineGap && (res.usedAffineGapScoring || res.score > 0);
    for (;;) {
        const int n = affineGap ? sg_sam_format(C, R, res.status, res.mapq, finalLocation, res.direction, true, res.score, res.basesClippedBefore, res.basesClippedAfter, out, &addFrontClipping)
                                : sg_sam_format(C, R, res.status, res.mapq, finalLocation, res.direction, false, 0, 0, 0, out, &addFrontClipping);
        if (n > 0) {
            if (C.sort) { C.sort->nRecords = 1; C.sort->bytes[0] = (uint32_t)n; C.sort->location[0] = finalLocation < 0 ? SG_SORT_UNALIGNED : finalLocation; }
            return n;
        }
        nAdjustments++;
        if (addFrontClipping == 0) return 0;                      // (cannot happen here: out of buffer space in the reference)
        const int origC = res.status == SNAPGPU_NOT_FOUND ? -1 : sg_contig_at(ix, res.location);
        const int newC = res.status == SNAPGPU_NOT_FOUND ? -1 : sg_contig_at(ix, res.location + addFrontClipping);
        const int64_t endOf = (origC < 0) ? 0 : (((origC == (int)ix.nContigs - 1) ? ix.nBases : ix.contigStart[origC + 1]) - (int64_t)ix.chromosomePadding);
        if (newC < 0 || newC != origC || finalLocation + addFrontClipping > endOf || nAdjustments > R.dataLength) {
            res.status = SNAPGPU_NOT_FOUND; res.location = -1; res.score = -1; res.direction = SNAPGPU_FORWARD;
            finalLocation = -1;
        } else if (affineGap) {
            if (addFrontClipping < 0) {                           // leading insertion: soft-clip it
                cumulativeAddFrontClipping += addFrontClipping;
                if (res.direction == SNAPGPU_FORWARD) R.setAdditionalFrontClipping(-cumulativeAddFrontClipping);
                else R.setAdditionalBackClipping(-cumulativeAddFrontClipping);
            } else {                                              // leading deletion: move the start
                finalLocation = res.location + addFrontClipping;
            }
        } else {
            if (addFrontClipping > 0) {
                cumulativeAddFrontClipping += addFrontClipping;
                R.setAdditionalFrontClipping(cumulativeAddFrontClipping);
            }
            finalLocation += addFrontClipping;
        }
    }
}

#define SG_SAM_INVALID_LOCATION 0xffffffffLL   /* InvalidGenomeLocation of a 4-byte-location index: sorts after every real location */

struct SgSamLine {                   // what createSAMLine hands back (:1424-1450)
    int flags, contig, mapQuality;
    int64_t positionInContig, extraBasesClippedBefore;
    uint32_t fullLength, clippedLength, basesClippedBefore, basesClippedAfter;
    const uint8_t *clippedData, *clippedQuality;
};

// SAMFormat::createSAMLine (:1470-1572); data / quality: where the oriented read goes.  genomeLocation: SG_SAM_INVALID_LOCATION if unmapped.
SG_HD void sg_sam_create_line(const SgIndexView &ix, const SgSamRead &R, int status, int64_t genomeLocation, int direction, int mapQuality, int bpClippedBefore,
                              int bpClippedAfter, uint8_t *data, uint8_t *quality, SgSamLine *o)
{
    o->flags = 0; o->contig = -1; o->positionInContig = 0; o->extraBasesClippedBefore = 0;
    if (status == SNAPGPU_NOT_FOUND) genomeLocation = SG_SAM_INVALID_LOCATION;
    const bool mapped = genomeLocation != SG_SAM_INVALID_LOCATION;
    if (!mapped) direction = SNAPGPU_FORWARD;
    o->clippedLength = R.dataLength;
    o->fullLength = R.unclippedLength;
    const uint32_t fullLength = o->fullLength;
    if (direction == SNAPGPU_RC) {
        for (uint32_t i = 0; i < fullLength; i++) { data[fullLength - 1 - i] = sg_complement(R.unclippedData[i]); quality[fullLength - 1 - i] = R.unclippedQuality[i]; }
        o->clippedData = &data[fullLength - o->clippedLength - R.frontClipped];
        o->clippedQuality = &quality[fullLength - o->clippedLength - R.frontClipped];
        o->basesClippedBefore = fullLength - o->clippedLength - R.frontClipped;
        o->basesClippedAfter = R.frontClipped;
    } else {
        for (uint32_t i = 0; i < fullLength; i++) { data[i] = R.unclippedData[i]; quality[i] = R.unclippedQuality[i]; }
        o->clippedData = data + R.frontClipped;
        o->clippedQuality = quality + R.frontClipped;
        o->basesClippedBefore = R.frontClipped;
        o->basesClippedAfter = fullLength - o->clippedLength - o->basesClippedBefore;
    }
    o->basesClippedBefore += (uint32_t)bpClippedBefore; o->basesClippedAfter += (uint32_t)bpClippedAfter;
    o->clippedData += bpClippedBefore; o->clippedQuality += bpClippedBefore;
    o->clippedLength -= (uint32_t)(bpClippedBefore + bpClippedAfter);
    if (mapped) {
        if (direction == SNAPGPU_RC) o->flags |= 0x10;
        o->contig = sg_contig_for_read(ix, genomeLocation, R.dataLength, &o->extraBasesClippedBefore);
        genomeLocation += o->extraBasesClippedBefore;
        o->positionInContig = genomeLocation - ix.contigStart[o->contig] + 1;
        o->mapQuality = mapQuality < 0 ? 0 : (mapQuality > 70 ? 70 : mapQuality);
    } else {
        o->flags |= 0x4;
        o->mapQuality = 0;
    }
}

struct SgSamPairResult {             // the PairedAlignmentResult fields the writer reads
    int status[2]; int64_t location[2]; int direction[2], mapq[2], score[2], usedAffineGapScoring[2], basesClippedBefore[2], basesClippedAfter[2],
        clippingForReadAdjustment[2];
    int alignedAsPair;
};

// SimpleReadWriter::writePairs for one (primary) pair result: both records, in genome order.  R[] and res are working copies.
SG_HDN int sg_sam_write_pair(const SgSamContext &C, SgSamRead R0, SgSamRead R1, SgSamPairResult res, char *out)
{
    const SgIndexView &ix = *C.ix;
    SgSamRead R[2] = {R0, R1};
    uint8_t *dataBuf[2] = {C.data, C.data2}, *qualBuf[2] = {C.quality, C.quality2};
    // QNAME: a trailing /1 /2 pair is cut (ReadWriter.cpp:409-421)
    uint32_t idLen[2] = {R[0].idLength, R[1].idLength};
    if (idLen[0] == idLen[1] && idLen[0] > 2 && R[0].id[idLen[0] - 2] == '/' && R[1].id[idLen[0] - 2] == '/') {
        const uint8_t a = R[0].id[idLen[0] - 1], b = R[1].id[idLen[1] - 1];
        if ((a == '1' || a == '2') && (b == '1' || b == '2') && a != b) { idLen[0] -= 2; idLen[1] -= 2; }
    }
    R[0].setAdditionalFrontClipping(res.clippingForReadAdjustment[0]);
    R[1].setAdditionalFrontClipping(res.clippingForReadAdjustment[1]);
    int64_t locations[2];
    for (int w = 0; w < 2; w++) locations[w] = res.status[w] != SNAPGPU_NOT_FOUND ? res.location[w] : SG_SAM_INVALID_LOCATION;
    int cumulative[2] = {0, 0};
    bool secondReadLocationChanged, writeOrderChanged;
    int n = 0;
    do {
        secondReadLocationChanged = false; writeOrderChanged = false;
        int writeOrder[2];
        if (locations[0] <= locations[1]) { writeOrder[0] = 0; writeOrder[1] = 1; } else { writeOrder[0] = 1; writeOrder[1] = 0; }
        // ---- SAMFormat::writePairs (:1628-1716): line fields and CIGAR of each read, in write order ----
        SgSamLine line[2];
        uint32_t ops[2][SG_SAM_MAX_OPS];
        SgCigarOut co[2];
        int editDistance[2] = {-1, -1}, refSpan[2] = {0, 0};
        for (int fs = 0; fs < 2; fs++) {
            const int w = writeOrder[fs];
            int addFrontClipping;
            do {
                addFrontClipping = 0;
                sg_sam_create_line(ix, R[w], res.status[w], locations[w], res.direction[w], res.mapq[w], res.basesClippedBefore[w], res.basesClippedAfter[w],
                                   dataBuf[w], qualBuf[w], &line[w]);
                co[w].kind = 1; co[w].nOps = 0;
                if (locations[w] != SG_SAM_INVALID_LOCATION) {
                    const bool ag = C.useAffineGap && (res.usedAffineGapScoring[w] || res.score[w] > 0);
                    if (ag) sg_cigar_ag(ix, C.ag, C.agS, line[w].clippedData, line[w].clippedQuality, line[w].clippedLength, res.score[w], line[w].basesClippedBefore,
                                        line[w].extraBasesClippedBefore, line[w].basesClippedAfter, 0, 0, locations[w], C.useM, ops[w], SG_SAM_MAX_OPS, &co[w]);
                    else sg_cigar_lv(ix, C.lv, line[w].clippedData, line[w].clippedLength, line[w].basesClippedBefore, line[w].extraBasesClippedBefore,
                                     line[w].basesClippedAfter, 0, 0, locations[w], C.useM, ops[w], SG_SAM_MAX_OPS, &co[w]);
                    editDistance[w] = co[w].editDistance; refSpan[w] = co[w].kind == 2 ? co[w].refSpan : 0;
                    addFrontClipping = co[w].addFrontClipping;
                    if (addFrontClipping != 0) {
                        secondReadLocationChanged = fs == 1;
                        const int origC = sg_contig_at(ix, locations[w]), newC = sg_contig_at(ix, locations[w] + addFrontClipping);
                        const int64_t endOf = (origC < 0) ? 0 : (((origC == (int)ix.nContigs - 1) ? ix.nBases : ix.contigStart[origC + 1]) - (int64_t)ix.chromosomePadding);
                        if (newC != origC || newC < 0 || locations[w] + addFrontClipping > endOf) {
                            res.status[w] = SNAPGPU_NOT_FOUND; res.location[w] = SG_SAM_INVALID_LOCATION; locations[w] = SG_SAM_INVALID_LOCATION;
                            co[w].kind = 1; editDistance[w] = -1; res.direction[w] = SNAPGPU_FORWARD;
                        } else if (ag) {
                            if (addFrontClipping < 0) {
                                cumulative[fs] += addFrontClipping;
                                if (res.direction[w] == SNAPGPU_FORWARD) R[w].setAdditionalFrontClipping(-cumulative[fs]);
                                else R[w].setAdditionalBackClipping(-cumulative[fs]);
                            } else {
                                locations[w] += addFrontClipping;
                            }
                        } else {
                            if (addFrontClipping > 0) { cumulative[fs] += addFrontClipping; R[w].setAdditionalFrontClipping(cumulative[fs]); }
                            locations[w] += addFrontClipping;
                        }
                    }
                }
            } while (addFrontClipping != 0);
        }
        // ---- fillMateInfo (:1308-1422) + the text (:1733-1893) ----
        n = 0;
        char *p = out;
        for (int fs = 0; fs < 2; fs++) {
            const int w = writeOrder[fs], m = 1 - w;
            const char *const recStart = p;
            if (C.sort) {        // ReadWriter.cpp:601-606: filed under its own location, or its mate's when it has none
                const int64_t l = locations[w] != SG_SAM_INVALID_LOCATION ? locations[w] : locations[m];
                C.sort->nRecords = 2; C.sort->location[fs] = l == SG_SAM_INVALID_LOCATION ? SG_SORT_UNALIGNED : l;
            }
            const bool firstInPair = w == 0;
            int flags = line[w].flags | 0x1 | (firstInPair ? 0x40 : 0x80);
            int contig = line[w].contig; int64_t pos = line[w].positionInContig;
            int mateContig = -1; int64_t matePos = 0; bool mateIsEq = false;
            long long templateLength = 0;
            int64_t mateLocation = locations[m], genomeLocation = locations[w];
            int64_t mateExtra = 0, extra = 0;
            if (mateLocation != SG_SAM_INVALID_LOCATION) {
                mateContig = sg_contig_for_read(ix, mateLocation, R[m].dataLength, &mateExtra);
                mateLocation += mateExtra;
                matePos = mateLocation - ix.contigStart[mateContig] + 1;
                if (res.direction[m] == SNAPGPU_RC) flags |= 0x20;
                if (genomeLocation == SG_SAM_INVALID_LOCATION) { contig = mateContig; mateIsEq = true; pos = matePos; }
            } else {
                flags |= 0x8;
                mateIsEq = true; mateContig = contig; matePos = pos;
            }
            if (genomeLocation != SG_SAM_INVALID_LOCATION && mateLocation != SG_SAM_INVALID_LOCATION) {
                if (res.alignedAsPair) flags |= 0x2;
                sg_contig_for_read(ix, genomeLocation, R[w].dataLength, &extra);
                genomeLocation += extra;
                const int64_t myStart = genomeLocation - line[w].basesClippedBefore - extra, myEnd = genomeLocation + refSpan[w];
                const int64_t mateStart = mateLocation - line[m].basesClippedBefore - mateExtra, mateEnd = mateLocation + refSpan[m];
                const bool fwd = res.direction[w] == SNAPGPU_FORWARD, mfwd = res.direction[m] == SNAPGPU_FORWARD;
                if (myStart < mateStart) {
                    if (fwd) templateLength = !mfwd ? mateEnd - myStart : mateStart - myStart;
                    else templateLength = mfwd ? mateStart - myEnd : mateEnd - myEnd;
                } else {
                    if (!fwd) templateLength = mfwd ? -(myEnd - mateStart) : -(myEnd - mateEnd);
                    else templateLength = mfwd ? -(myStart - mateStart) : -(myStart - mateEnd);
                }
            }
            // (the reference compares the two name POINTERS: equal for the same contig, and "*" is never equal to a contig's name)
            if (!mateIsEq && contig >= 0 && contig == mateContig) mateIsEq = true;
            // text
            uint32_t qnameLen = idLen[w];
            for (uint32_t i = 0; i < qnameLen; i++) if (R[w].id[i] == ' ') { qnameLen = i; break; }
            for (uint32_t i = 0; i < qnameLen; i++) *p++ = (char)R[w].id[i];
            *p++ = '\t'; p = sg_put_i64(p, flags);
            *p++ = '\t'; p = sg_put_str(p, contig >= 0 ? C.contigName[contig] : "*");
            *p++ = '\t'; p = sg_put_i64(p, pos);
            *p++ = '\t'; p = sg_put_i64(p, line[w].mapQuality);                           // (createSAMLine clamps / zeroes result->mapq in place)
            *p++ = '\t';
            if (locations[w] != SG_SAM_INVALID_LOCATION && co[w].kind == 2) {
                for (int i = 0; i < co[w].nOps; i++) { p = sg_put_i64(p, (long long)(ops[w][i] >> 4)); *p++ = "MIDNSHP=X"[ops[w][i] & 15]; }
            } else *p++ = '*';
            *p++ = '\t'; p = sg_put_str(p, mateIsEq ? "=" : (mateContig >= 0 ? C.contigName[mateContig] : "*"));
            *p++ = '\t'; p = sg_put_i64(p, matePos);
            *p++ = '\t'; p = sg_put_i64(p, (long long)(int)templateLength);
            *p++ = '\t';
            for (uint32_t i = 0; i < line[w].fullLength; i++) *p++ = (char)dataBuf[w][i];
            *p++ = '\t';
            for (uint32_t i = 0; i < line[w].fullLength; i++) *p++ = (char)qualBuf[w][i];
            p = sg_put_str(p, "\tPG:Z:SNAP\tNM:i:"); p = sg_put_i64(p, editDistance[w]);
            p = sg_put_str(p, C.readGroupAux);
            int mqs = 0;
            for (uint32_t i = 0; i < line[m].fullLength; i++) { const int q = (int)qualBuf[m][i] - '!'; mqs += (q >= 15) ? (q != 255) * q : 0; }
            p = sg_put_str(p, "\tQS:i:"); p = sg_put_i64(p, mqs);
            *p++ = '\n';
            if (C.sort) C.sort->bytes[fs] = (uint32_t)(p - recStart);
        }
        n = (int)(p - out);
        int newOrder0 = (locations[0] <= locations[1]) ? 0 : 1;
        if (writeOrder[0] != newOrder0) writeOrderChanged = true;
    } while (secondReadLocationChanged || writeOrderChanged);
    return n;
}
