// sg_sam.h -- one SAM record of an unpaired read: SimpleReadWriter::writeReads' per-result loop (reference SNAPLib/ReadWriter.cpp:
// 170-330) around SAMFormat::writeRead (SAM.cpp:1897-2112 for Landau-Vishkin results, :2113-2352 for affine-gap ones) and
// SAMFormat::createSAMLine (:1423-1573).  Fourth piece of the output stage; same status as sg_cigar.h: verified on
// the host (tests/test_lv_cigar.py: against the SAM file the reference binary writes for the same reads), no device entry point
// yet, nothing in include/snapgpu.h refers to it.  Primary alignments of single-end runs only (no mate fields, no secondary
// results, default tags: PG, NM, the default read group line).
//
// The loop: format the record; if the CIGAR routine answers with a front-clipping verdict (a leading deletion / insertion, or a
// read that starts before its contig) move the alignment or clip the read and format again; give the read up (unmapped) when that
// would cross a contig boundary or does not settle.
#pragma once
#include "sg_cigar.h"

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

struct SgSamContext {
    const SgIndexView *ix;
    const char *const *contigName;   // [nContigs], NUL-terminated
    SgAgParams ag;
    const char *readGroupAux;        // ReaderContext::defaultReadGroupAux, e.g. "\tRG:Z:FASTQ\tPL:Z:Illumina\tPU:Z:pu\tLB:Z:lb\tSM:Z:sm"
    bool useM, useAffineGap;
    SgLvCigarScratch lv;
    SgAgCigarScratch agS;
    uint8_t *data, *quality;         // [maxReadLen] scratch for the oriented read
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
    uint32_t ops[48];
    SgCigarOut co;
    co.kind = 1; co.editDistance = -1; co.nOps = 0;
    int editDistance = -1;
    if (affineGap && extraBasesClippedBefore != 0) { *addFrontClipping = (int)extraBasesClippedBefore; return 0; }
    if (mapped) {
        // (createSAMLine already moved genomeLocation; computeCigarString adds extraBasesClippedBefore again to the location it is given,
        //  which is the caller's original one: pass that)
        const int64_t locForCigar = genomeLocation - extraBasesClippedBefore;
        if (affineGap) sg_cigar_ag(ix, C.ag, C.agS, clippedData, clippedQuality, clippedLength, score, basesClippedBefore, extraBasesClippedBefore, basesClippedAfter, 0, 0,
                                   locForCigar, C.useM, ops, 48, &co);
        else sg_cigar_lv(ix, C.lv, clippedData, clippedLength, basesClippedBefore, extraBasesClippedBefore, basesClippedAfter, 0, 0, locForCigar, C.useM, ops, 48, &co);
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
        if (n > 0) return n;
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
