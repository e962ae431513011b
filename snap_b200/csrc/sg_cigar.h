// sg_cigar.h -- CIGAR of one aligned read: SAMFormat::computeCigar inside SAMFormat::computeCigarString, as BAM operations.  sg_cigar_lv:
// the LandauVishkinWithCigar overloads (reference SNAPLib/SAM.cpp:2354-2468, :2595-2671) for results that were not rescored with
// affine gap; sg_cigar_ag: the AffineGapVectorizedWithCigar overloads (:2470-2592, :2677-2766) for those that were.  Second piece of the output stage; same status as sg_lv_cigar.h: verified on the host against
// the compiled reference (tests/test_output_stage.py), no device entry point yet.
//
// What it adds to the LV routine: soft clipping of a read that hangs over the end of its contig (re-run until the clip and the
// alignment's net indel agree), the front-clipping verdict that makes the caller move the alignment start and try again, the
// soft / hard clip operations around the result, and getRefSpanFromCigar's arithmetic (:2768-2800) -- which counts a leading
// insertion and trailing clips into the reference span; reproduced as is.
#pragma once
#include "sg_seed.h"
#include "sg_lv_cigar.h"
#include "sg_ag_cigar.h"

#define SG_CIGAR_H 5u

struct SgCigarOut {
    int kind;                        // 0: no CIGAR, caller must retry with addFrontClipping (NULL in the reference); 1: "*"; 2: operations
    int editDistance;                // o_editDistance
    int addFrontClipping;            // o_addFrontClipping
    int refSpan;                     // o_refSpan (kind 2 only)
    int nOps;
};

// Genome::getContigAtLocation (Genome.cpp:573-594): index of the contig holding `location`, or -1
SG_HD int sg_contig_at(const SgIndexView &ix, int64_t location)
{
    int low = 0, high = (int)ix.nContigs - 1;
    while (low <= high) {
        const int mid = (low + high) / 2;
        const int64_t b = ix.contigStart[mid];
        if (b <= location && (mid == (int)ix.nContigs - 1 || ix.contigStart[mid + 1] > location)) return mid;
        else if (b <= location) low = mid + 1;
        else high = mid - 1;
    }
    return -1;
}

// data / dataLength: the clipped view of the read, in the alignment's direction.  ops: room for maxOps operations.
SG_HDN void sg_cigar_lv(const SgIndexView &ix, const SgLvCigarScratch &S, const uint8_t *data, int64_t dataLength, uint32_t basesClippedBefore,
                        int64_t extraBasesClippedBefore, uint32_t basesClippedAfter, uint32_t frontHardClipping, uint32_t backHardClipping,
                        int64_t genomeLocation, bool useM, uint32_t *ops, int maxOps, SgCigarOut *out)
{
    out->kind = 1; out->editDistance = 0; out->addFrontClipping = 0; out->refSpan = 0; out->nOps = 0;
    // ---- computeCigar (:2373-2466) ----
    genomeLocation += extraBasesClippedBefore;
    data += extraBasesClippedBefore;
    dataLength -= extraBasesClippedBefore;
    const int c = sg_contig_at(ix, genomeLocation);
    const int64_t contigEnd = ((c == (int)ix.nContigs - 1 || c < 0) ? ix.nBases : ix.contigStart[c + 1]) - (int64_t)ix.chromosomePadding;
    int64_t extraBasesClippedAfter = 0;
    if (genomeLocation + dataLength > contigEnd) extraBasesClippedAfter = genomeLocation + dataLength - contigEnd;
    const uint8_t *reference = sg_get_substring(ix, genomeLocation, dataLength);
    if (reference == (const uint8_t *)0) {
        // fell off the end of the contig: the reference leaves an unterminated '*' in its buffer and formats from there (:2400-2408);
        // we report "*"
        out->kind = 1; out->editDistance = 0; out->addFrontClipping = 0;
        return;
    }
    // room for: H, S, the LV operations, S, H
    uint32_t *lvOps = ops + 2;
    const int lvMax = maxOps - 4;
    SgLvCigarOut lv;
    int netIndel = 0;
    out->editDistance = sg_lv_cigar_normalized(S, reference, (int)(dataLength - extraBasesClippedAfter + SG_MAX_K), data, (int)(dataLength - extraBasesClippedAfter),
                                               SG_MAX_K - 1, lvOps, lvMax, useM, &lv, &out->addFrontClipping);
    netIndel = lv.netIndel;
    if (out->addFrontClipping != 0) { out->kind = 0; return; }
    int64_t newExtra = genomeLocation + dataLength + netIndel - contigEnd;
    if (newExtra < 0) newExtra = 0;
    for (int64_t pass = 0; pass < dataLength; pass++) {
        if (newExtra == extraBasesClippedAfter) break;
        extraBasesClippedAfter = newExtra;
        out->editDistance = sg_lv_cigar_normalized(S, reference, (int)(dataLength - extraBasesClippedAfter + SG_MAX_K), data,
                                                   (int)(dataLength - extraBasesClippedAfter), SG_MAX_K - 1, lvOps, lvMax, useM, &lv, &out->addFrontClipping);
        netIndel = lv.netIndel;
        newExtra = genomeLocation + dataLength + netIndel - contigEnd;
        if (newExtra < 0) newExtra = 0;
    }
    // ---- computeCigarString (:2625-2670) ----
    if (out->addFrontClipping != 0) { out->kind = 0; return; }
    if (out->editDistance < 0) { out->kind = 1; return; }         // -2 / -1: "*"
    int n = 0;
    if (frontHardClipping > 0) ops[n++] = (frontHardClipping << 4) | SG_CIGAR_H;
    if ((int64_t)basesClippedBefore + extraBasesClippedBefore > 0) ops[n++] = ((uint32_t)((int64_t)basesClippedBefore + extraBasesClippedBefore) << 4) | SG_CIGAR_S;
    for (int i = 0; i < lv.nOps; i++) ops[n++] = lvOps[i];          // (n <= i + 2: moving down, never overtakes the source)
    if ((int64_t)basesClippedAfter + extraBasesClippedAfter > 0) ops[n++] = ((uint32_t)((int64_t)basesClippedAfter + extraBasesClippedAfter) << 4) | SG_CIGAR_S;
    if (backHardClipping > 0) ops[n++] = (backHardClipping << 4) | SG_CIGAR_H;
    out->nOps = n;
    out->kind = 2;
    // getRefSpanFromCigar (:2768-2800): the first operation counts unless it is S or H; every later one counts unless it is I
    int span = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t code = ops[i] & 0xfu, len = ops[i] >> 4;
        if (i == 0) { if (code != SG_CIGAR_S && code != SG_CIGAR_H) span += (int)len; }
        else if (code != SG_CIGAR_I) span += (int)len;
    }
    out->refSpan = span;
}

// The AffineGapVectorizedWithCigar overloads.  Differences from sg_cigar_lv: the band is the result's score, the re-run loop stops as soon
// as the clip does not GROW (<=, :2563), and insertions at the read's tail (tailIns) are turned into soft clipping.
// data / quality: the clipped view of the read, in the alignment's direction.
SG_HDN void sg_cigar_ag(const SgIndexView &ix, const SgAgParams &P, const SgAgCigarScratch &S, const uint8_t *data, const uint8_t *quality, int64_t dataLength,
                        int score, uint32_t basesClippedBefore, int64_t extraBasesClippedBefore, uint32_t basesClippedAfter, uint32_t frontHardClipping,
                        uint32_t backHardClipping, int64_t genomeLocation, bool useM, uint32_t *ops, int maxOps, SgCigarOut *out)
{
    out->kind = 1; out->editDistance = 0; out->addFrontClipping = 0; out->refSpan = 0; out->nOps = 0;
    genomeLocation += extraBasesClippedBefore;
    data += extraBasesClippedBefore; quality += 0;                  // (the reference advances `data` only, :2502; so the qualities stay put)
    dataLength -= extraBasesClippedBefore;
    const int c = sg_contig_at(ix, genomeLocation);
    const int64_t contigEnd = ((c == (int)ix.nContigs - 1 || c < 0) ? ix.nBases : ix.contigStart[c + 1]) - (int64_t)ix.chromosomePadding;
    int64_t extraBasesClippedAfter = 0;
    if (genomeLocation + dataLength > contigEnd) extraBasesClippedAfter = genomeLocation + dataLength - contigEnd;
    const uint8_t *reference = sg_get_substring(ix, genomeLocation, dataLength);
    if (reference == (const uint8_t *)0) { out->kind = 1; return; }
    uint32_t *agOps = ops + 2;
    const int agMax = maxOps - 4;
    SgAgCigarOut ag;
    out->editDistance = sg_ag_cigar_normalized(P, S, reference, (int)(dataLength - extraBasesClippedAfter + SG_MAX_K), data, quality,
                                               (int)(dataLength - extraBasesClippedAfter), score, agOps, agMax, useM, &ag, &out->addFrontClipping);
    if (out->addFrontClipping != 0) { out->kind = 0; return; }
    int64_t newExtra = genomeLocation + dataLength + ag.netDel - contigEnd;
    if (newExtra < 0) newExtra = 0;
    for (int64_t pass = 0; pass < dataLength; pass++) {
        if (newExtra <= extraBasesClippedAfter) break;
        extraBasesClippedAfter = newExtra;
        out->editDistance = sg_ag_cigar_normalized(P, S, reference, (int)(dataLength - extraBasesClippedAfter + SG_MAX_K), data, quality,
                                                   (int)(dataLength - extraBasesClippedAfter), score, agOps, agMax, useM, &ag, &out->addFrontClipping);
        newExtra = genomeLocation + dataLength + ag.netDel - contigEnd;
        if (newExtra < 0) newExtra = 0;
    }
    if (out->addFrontClipping != 0) { out->kind = 0; return; }
    if (out->editDistance < 0) { out->kind = 1; return; }
    const int64_t clippedAfter = (int64_t)basesClippedAfter + ag.tailIns;                  // "whenever we see tail insertions, soft-clip them" (:2725)
    int n = 0;
    if (frontHardClipping > 0) ops[n++] = (frontHardClipping << 4) | SG_CIGAR_H;
    if ((int64_t)basesClippedBefore + extraBasesClippedBefore > 0) ops[n++] = ((uint32_t)((int64_t)basesClippedBefore + extraBasesClippedBefore) << 4) | SG_CIGAR_S;
    for (int i = 0; i < ag.nOps; i++) ops[n++] = agOps[i];
    if (clippedAfter + extraBasesClippedAfter > 0) ops[n++] = ((uint32_t)(clippedAfter + extraBasesClippedAfter) << 4) | SG_CIGAR_S;
    if (backHardClipping > 0) ops[n++] = (backHardClipping << 4) | SG_CIGAR_H;
    out->nOps = n;
    out->kind = 2;
    int span = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t code = ops[i] & 0xfu, len = ops[i] >> 4;
        if (i == 0) { if (code != SG_CIGAR_S && code != SG_CIGAR_H) span += (int)len; }
        else if (code != SG_CIGAR_I) span += (int)len;
    }
    out->refSpan = span;
}
