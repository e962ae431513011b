// sg_bam.h -- BAM records (unpaired: sg_bam_write_single; pairs: sg_bam_write_pair).  BAMFormat::writeRead (reference SNAPLib/Bam.cpp:1312-1509 for Landau-Vishkin results,
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
    uint32_t ops[SG_SAM_MAX_OPS];
    SgCigarOut co;
    co.kind = 1; co.editDistance = -1; co.nOps = 0;
    int editDistance = -1;
    if (affineGap && line.extraBasesClippedBefore != 0) { *addFrontClipping = (int)line.extraBasesClippedBefore; return 0; }       // (:1892-1895)
    if (mapped) {
        if (affineGap) sg_cigar_ag(ix, C.ag, C.agS, line.clippedData, line.clippedQuality, line.clippedLength, score, line.basesClippedBefore, line.extraBasesClippedBefore,
                                   line.basesClippedAfter, 0, 0, loc, C.useM, ops, SG_SAM_MAX_OPS, &co);
        else sg_cigar_lv(ix, C.lv, line.clippedData, line.clippedLength, line.basesClippedBefore, line.extraBasesClippedBefore, line.basesClippedAfter, 0, 0, loc, C.useM,
                         ops, SG_SAM_MAX_OPS, &co);
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
        if (n > 0) {
            if (C.sort) { C.sort->nRecords = 1; C.sort->bytes[0] = (uint32_t)n; C.sort->location[0] = finalLocation == SG_SAM_INVALID_LOCATION ? SG_SORT_UNALIGNED : finalLocation; }
            return n;
        }
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

// SimpleReadWriter::writePairs for one (primary) pair result with BAMFormat::writePairs (Bam.cpp:1033-1310): the control flow is
// sg_sam_write_pair's (same createSAMLine / CIGAR / fillMateInfo calls in the reference too); only the record written differs.
SG_HDN int sg_bam_write_pair(const SgSamContext &C, const SgBamContext &B, SgSamRead R0, SgSamRead R1, SgSamPairResult res, char *out)
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
            // the record (Bam.cpp:1224-1299)
            uint32_t qnameLen = idLen[w];
            const int cigarOps = (locations[w] != SG_SAM_INVALID_LOCATION && co[w].kind == 2) ? co[w].nOps : 0;
            const uint32_t fullLength = line[w].fullLength;
            int refLength = cigarOps > 0 ? 0 : (int)fullLength;
            for (int i = 0; i < cigarOps; i++) {
                const uint32_t code = ops[w][i] & 0xf;
                refLength += ((code == 0 || code == 2 || code == 3 || code == 6 || code == 7 || code == 8) ? 1 : 0) * (int)(ops[w][i] >> 4);
            }
            const int bin = locations[w] != SG_SAM_INVALID_LOCATION ? sg_bam_reg2bin((int)pos - 1, (int)pos - 1 + refLength)
                          : locations[m] != SG_SAM_INVALID_LOCATION ? sg_bam_reg2bin((int)matePos - 1, (int)matePos) : sg_bam_reg2bin(-1, 0);
            const int tl = templateLength >= 0 ? (int)(templateLength & 0x7fffffff) : -(int)((-templateLength) & 0x7fffffff);
            uint8_t *rec = (uint8_t *)p;
            sg_put_le32(rec + 4, (uint32_t)contig);
            sg_put_le32(rec + 8, (uint32_t)((int)pos - 1));
            rec[12] = (uint8_t)(qnameLen + 1);
            rec[13] = (uint8_t)line[w].mapQuality;
            sg_put_le16(rec + 14, (uint32_t)bin);
            sg_put_le16(rec + 16, (uint32_t)cigarOps);
            sg_put_le16(rec + 18, (uint32_t)flags);
            sg_put_le32(rec + 20, fullLength);
            sg_put_le32(rec + 24, (uint32_t)mateContig);
            sg_put_le32(rec + 28, (uint32_t)((int)matePos - 1));
            sg_put_le32(rec + 32, (uint32_t)tl);
            uint8_t *q = rec + 36;
            for (uint32_t i = 0; i < qnameLen; i++) *q++ = R[w].id[i];
            *q++ = 0;
            for (int i = 0; i < cigarOps; i++) { sg_put_le32(q, ops[w][i]); q += 4; }
            for (uint32_t i = 0; i + 1 < fullLength; i += 2) *q++ = (uint8_t)((sg_bam_seq_code(dataBuf[w][i]) << 4) | sg_bam_seq_code(dataBuf[w][i + 1]));
            if (fullLength % 2) *q++ = (uint8_t)(sg_bam_seq_code(dataBuf[w][fullLength - 1]) << 4);
            for (uint32_t i = 0; i < fullLength; i++) *q++ = (uint8_t)(qualBuf[w][i] - '!');
            for (int i = 0; i < B.readGroupAuxLen; i++) *q++ = B.readGroupAux[i];
            *q++ = 'P'; *q++ = 'G'; *q++ = 'Z'; *q++ = 'S'; *q++ = 'N'; *q++ = 'A'; *q++ = 'P'; *q++ = 0;
            *q++ = 'N'; *q++ = 'M'; *q++ = 'C'; *q++ = (uint8_t)editDistance[w];
            int mqs = 0;
            for (uint32_t i = 0; i < line[m].fullLength; i++) { const int qq = (int)qualBuf[m][i] - '!'; mqs += (qq >= 15) ? (qq != 255) * qq : 0; }
            *q++ = 'Q'; *q++ = 'S'; *q++ = 'i'; sg_put_le32(q, (uint32_t)mqs); q += 4;
            sg_put_le32(rec, (uint32_t)((int)(q - rec) - 4));
            p = (char *)q;
            if (C.sort) {        // ReadWriter.cpp:601-606: filed under its own location, or its mate's when it has none
                const int64_t l = locations[w] != SG_SAM_INVALID_LOCATION ? locations[w] : locations[m];
                C.sort->nRecords = 2; C.sort->location[fs] = l == SG_SAM_INVALID_LOCATION ? SG_SORT_UNALIGNED : l;
                C.sort->bytes[fs] = (uint32_t)((const uint8_t *)q - rec);
            }
        }
        n = (int)(p - out);
        int newOrder0 = (locations[0] <= locations[1]) ? 0 : 1;
        if (writeOrder[0] != newOrder0) writeOrderChanged = true;
    } while (secondReadLocationChanged || writeOrderChanged);
    return n;
}
