// sg_bam.h -- one BAM record of an unpaired read: BAMFormat::writeRead (reference SNAPLib/Bam.cpp:1312-1509 for Landau-Vishkin results,
// :1812-2031 for affine-gap ones), computeCigarOps (:2032-2210) and buildAUX (:1510-1810, default tags) inside the same per-result
// loop of SimpleReadWriter::writeReads as the SAM form (sg_sam.h).  Output stage, HOST-VERIFIED ONLY
// (tests/test_output_stage.py: against the BAM file the reference binary writes, BGZF blocks inflated); not compiled into the CUDA
// library yet and nothing in include/snapgpu.h refers to it.
//
// The CIGAR is the one sg_cigar.h already produces (BAM operations are its native output: computeCigarOps lays out the same H / S /
// operations / S / H words and getRefSpanFromCigar's arithmetic is the same table); what is new here is the fixed part of the record,
// the bin, the 4-bit sequence, the binary tags.
#pragma once
#include "sg_sam.h"

struct SgBamContext {
    const uint8_t *readGroupAux;     // ReaderContext::defaultReadGroupAux in BAM form ("RGZFASTQ\0PLZIllumina\0..."), readGroupAuxLen bytes
    int readGroupAuxLen;
};

SG_HD int sg_bam_reg2bin(int beg, int end)     // BAMAlignment::reg2bin (Bam.cpp:523-534)
{
    --end;
    if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
    if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
    if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
    if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
    if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
    return 0;
}

SG_HD uint8_t sg_bam_seq_code(uint8_t c)       // BAMAlignment::SeqToCode (Bam.cpp:504-509): "=ACMGRSVTWYHKDBN", anything else 15
{
    const char *codes = "=ACMGRSVTWYHKDBN";
    for (int i = 1; i < 16; i++) if ((uint8_t)codes[i] == c) return (uint8_t)i;
    return 15;
}

SG_HD void sg_put_le32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
SG_HD void sg_put_le16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

// BAMFormat::writeRead (either overload) for a primary, unpaired record.  Returns the record's size (block_size + 4), or 0 with
// *addFrontClipping != 0.
SG_HDN int sg_bam_format(const SgSamContext &C, const SgBamContext &B, const SgSamRead &R, int status, int mapQuality, int64_t genomeLocation, int direction,
                         bool affineGap, int score, int bpClippedBefore, int bpClippedAfter, uint8_t *out, int *addFrontClipping)
{
    *addFrontClipping = 0;
    const SgIndexView &ix = *C.ix;
    SgSamLine line;
    const int64_t loc = status == SNAPGPU_NOT_FOUND ? SG_SAM_INVALID_LOCATION : genomeLocation;
    sg_sam_create_line(ix, R, status, loc, direction, mapQuality, bpClippedBefore, bpClippedAfter, C.data, C.quality, &line);
    const bool mapped = loc != SG_SAM_INVALID_LOCATION;
    uint32_t ops[48];
    SgCigarOut co;
    co.kind = 1; co.editDistance = -1; co.nOps = 0;
    int editDistance = -1;
    if (affineGap && line.extraBasesClippedBefore != 0) { *addFrontClipping = (int)line.extraBasesClippedBefore; return 0; }       // (:1892-1895)
    if (mapped) {
        if (affineGap) sg_cigar_ag(ix, C.ag, C.agS, line.clippedData, line.clippedQuality, line.clippedLength, score, line.basesClippedBefore, line.extraBasesClippedBefore,
                                   line.basesClippedAfter, 0, 0, loc, C.useM, ops, 48, &co);
        else sg_cigar_lv(ix, C.lv, line.clippedData, line.clippedLength, line.basesClippedBefore, line.extraBasesClippedBefore, line.basesClippedAfter, 0, 0, loc, C.useM,
                         ops, 48, &co);
        editDistance = co.editDistance;
        if (co.addFrontClipping != 0) { *addFrontClipping = co.addFrontClipping; return 0; }
    }
    const int cigarOps = (mapped && co.kind == 2) ? co.nOps : 0;
    uint32_t qnameLen = R.idLength;
    const uint32_t fullLength = line.fullLength;
    // ---- the record (:1455-1500 / :1960-2020) ----
    uint8_t *p = out;
    const int64_t positionInContig = line.positionInContig;
    int refLength = cigarOps > 0 ? 0 : (int)fullLength;
    for (int i = 0; i < cigarOps; i++) {
        const uint32_t code = ops[i] & 0xf;
        const int consumes = (code == 0 || code == 2 || code == 3 || code == 6 || code == 7 || code == 8) ? 1 : 0;       // CigarCodeToRefBase (:270)
        refLength += consumes * (int)(ops[i] >> 4);
    }
    const int bin = mapped ? sg_bam_reg2bin((int)positionInContig - 1, (int)positionInContig - 1 + refLength) : sg_bam_reg2bin(-1, 0);
    const int seqBytes = ((int)fullLength + 1) / 2;
    const int fixed = 32 + (int)qnameLen + 1 + 4 * cigarOps + seqBytes + (int)fullLength;
    sg_put_le32(p + 4, (uint32_t)line.contig);                       // refID (-1 when unmapped)
    sg_put_le32(p + 8, (uint32_t)((int)positionInContig - 1));       // pos
    p[12] = (uint8_t)(qnameLen + 1);
    p[13] = (uint8_t)line.mapQuality;
    sg_put_le16(p + 14, (uint32_t)bin);
    sg_put_le16(p + 16, (uint32_t)cigarOps);
    sg_put_le16(p + 18, (uint32_t)line.flags);
    sg_put_le32(p + 20, fullLength);
    sg_put_le32(p + 24, (uint32_t)-1);                               // next_refID
    sg_put_le32(p + 28, (uint32_t)-1);                               // next_pos = matePositionInContig (0) - 1
    sg_put_le32(p + 32, 0);                                          // tlen
    uint8_t *q = p + 36;
    for (uint32_t i = 0; i < qnameLen; i++) *q++ = R.id[i];
    *q++ = 0;
    for (int i = 0; i < cigarOps; i++) { sg_put_le32(q, ops[i]); q += 4; }
    for (uint32_t i = 0; i + 1 < fullLength; i += 2) *q++ = (uint8_t)((sg_bam_seq_code(C.data[i]) << 4) | sg_bam_seq_code(C.data[i + 1]));
    if (fullLength % 2) *q++ = (uint8_t)(sg_bam_seq_code(C.data[fullLength - 1]) << 4);
    for (uint32_t i = 0; i < fullLength; i++) *q++ = (uint8_t)(C.quality[i] - '!');
    // buildAUX, default path: the read group line, PG:Z:SNAP, NM:C
    for (int i = 0; i < B.readGroupAuxLen; i++) *q++ = B.readGroupAux[i];
    *q++ = 'P'; *q++ = 'G'; *q++ = 'Z'; *q++ = 'S'; *q++ = 'N'; *q++ = 'A'; *q++ = 'P'; *q++ = 0;
    *q++ = 'N'; *q++ = 'M'; *q++ = 'C'; *q++ = (uint8_t)editDistance;
    const int total = (int)(q - out);
    sg_put_le32(out, (uint32_t)(total - 4));                         // block_size
    (void)fixed;
    return total;
}

// the per-result loop of SimpleReadWriter::writeReads (ReadWriter.cpp:223-310), BAM format
SG_HDN int sg_bam_write_single(const SgSamContext &C, const SgBamContext &B, SgSamRead R, SgSamResult res, uint8_t *out)
{
    const SgIndexView &ix = *C.ix;
    int addFrontClipping = 0;
    R.setAdditionalFrontClipping(res.clippingForReadAdjustment);
    int cumulativeAddFrontClipping = 0;
    int64_t finalLocation = res.status == SNAPGPU_NOT_FOUND ? SG_SAM_INVALID_LOCATION : res.location;
    if (res.status == SNAPGPU_NOT_FOUND) res.location = SG_SAM_INVALID_LOCATION;
    unsigned nAdjustments = 0;
    const bool affineGap = C.useAffineGap && (res.usedAffineGapScoring || res.score > 0);
    for (;;) {
        const int n = affineGap ? sg_bam_format(C, B, R, res.status, res.mapq, finalLocation, res.direction, true, res.score, res.basesClippedBefore, res.basesClippedAfter, out, &addFrontClipping)
                                : sg_bam_format(C, B, R, res.status, res.mapq, finalLocation, res.direction, false, 0, 0, 0, out, &addFrontClipping);
        if (n > 0) return n;
        nAdjustments++;
        if (addFrontClipping == 0) return 0;
        const int origC = res.status == SNAPGPU_NOT_FOUND ? -1 : sg_contig_at(ix, res.location);
        const int newC = res.status == SNAPGPU_NOT_FOUND ? -1 : sg_contig_at(ix, res.location + addFrontClipping);
        const int64_t endOf = (origC < 0) ? 0 : (((origC == (int)ix.nContigs - 1) ? ix.nBases : ix.contigStart[origC + 1]) - (int64_t)ix.chromosomePadding);
        if (newC < 0 || newC != origC || finalLocation + addFrontClipping > endOf || nAdjustments > R.dataLength) {
            res.status = SNAPGPU_NOT_FOUND; res.location = SG_SAM_INVALID_LOCATION; res.score = -1; res.direction = SNAPGPU_FORWARD;
            finalLocation = SG_SAM_INVALID_LOCATION;
        } else if (affineGap) {
            if (addFrontClipping < 0) {
                cumulativeAddFrontClipping += addFrontClipping;
                if (res.direction == SNAPGPU_FORWARD) R.setAdditionalFrontClipping(-cumulativeAddFrontClipping);
                else R.setAdditionalBackClipping(-cumulativeAddFrontClipping);
            } else {
                finalLocation = res.location + addFrontClipping;
            }
        } else {
            if (addFrontClipping > 0) { cumulativeAddFrontClipping += addFrontClipping; R.setAdditionalFrontClipping(cumulativeAddFrontClipping); }
            finalLocation += addFrontClipping;
        }
    }
}
