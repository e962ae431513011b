// sg_bampost.h -- what the reference's sorting writer does to a coordinate-sorted stream of BAM records after the sort:
// duplicate marking (BAMDupMarkFilter, reference SNAPLib/Bam.cpp:2619-3121) and the .bai index (BAMIndexSupplier, Bam.cpp:3229-3440).
// Host + device: the per-record field extraction and the per-key walks below are what the CUDA kernels run (snapgpu.cu) and what the
// CPU test suite runs, built for the host, against the files the reference binary writes.
//
// Duplicate marking, restated for a device.  The reference walks the sorted records ONCE, sequentially: it cuts them into overlapping
// "runs" (a run starts at a record, takes every following record whose unclipped start lies within 2*(MAX_READ_LENGTH + MAX_K) of the
// first one's, and the next run starts at the first record more than MAX_READ_LENGTH + MAX_K away), and in every run sorts the
// records by (library, 5' end and strand[, mate's 5' end and strand]), lets the members of each group of equal keys compete
// (base-quality sum, then tile / x / y of the read name, then file order), flags the losers, and carries a map of pair keys from the
// run that holds a pair's first end to the run that holds its second.  Everything that couples records is the KEY: groups of different
// keys never read each other's state.  So here (1) the runs are found by pointer jumping (every record computes where a run starting at
// it would end and where the next one would start; the runs actually visited are the orbit of record 0), (2) the records are sorted by
// key once, and (3) one thread per key replays that key's history run by run -- the same comparisons in the same order as the reference
// makes them, including what happens when a group straddles two overlapping runs.  Flags are bit-identical to the reference's for a
// stream it processes as one batch (its own results also depend on where its write buffers end, Bam.cpp:2780-2826).
#pragma once
#include "sg_common.h"

#define SG_BAM_FLAG_PAIRED 0x1
#define SG_BAM_FLAG_UNMAPPED 0x4
#define SG_BAM_FLAG_NEXT_UNMAPPED 0x8
#define SG_BAM_FLAG_RC 0x10
#define SG_BAM_FLAG_NEXT_RC 0x20
#define SG_BAM_FLAG_SECONDARY 0x100
#define SG_BAM_FLAG_DUPLICATE 0x400
#define SG_BAM_FLAG_SUPPLEMENTARY 0x800
#define SG_BAM_EXTRA_BIN 37450
#define SG_DUP_INVALID_LOCATION 0xffffffffLL          // InvalidGenomeLocation of an index with 4-byte locations (GenomeIndex.cpp:516)
#define SG_DUP_RUN_REACH (400 + 127)                  // MAX_READ_LENGTH + MAX_K (Read.h, LandauVishkin.h:11)
#define SG_DUP_BEST_ID 120                            // DuplicateMateInfo::bestReadId (Bam.cpp:2546)

SG_HD uint32_t sg_ld16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
SG_HD uint32_t sg_ld32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// The fixed part of a BAM alignment record (SAM spec 4.2; BAMAlignment, Bam.h:90-176) at an arbitrary byte address.
struct SgBamRec {
    const uint8_t *p;
    SG_HD int32_t blockSize() const { return (int32_t)sg_ld32(p); }
    SG_HD int32_t size() const { return blockSize() + 4; }
    SG_HD int32_t refID() const { return (int32_t)sg_ld32(p + 4); }
    SG_HD int32_t pos() const { return (int32_t)sg_ld32(p + 8); }
    SG_HD uint32_t lReadName() const { return p[12]; }
    SG_HD uint32_t bin() const { return sg_ld16(p + 14); }
    SG_HD uint32_t nCigar() const { return sg_ld16(p + 16); }
    SG_HD uint32_t flag() const { return sg_ld16(p + 18); }
    SG_HD int32_t lSeq() const { return (int32_t)sg_ld32(p + 20); }
    SG_HD int32_t nextRefID() const { return (int32_t)sg_ld32(p + 24); }
    SG_HD int32_t nextPos() const { return (int32_t)sg_ld32(p + 28); }
    SG_HD int32_t tlen() const { return (int32_t)sg_ld32(p + 32); }
    SG_HD const uint8_t *name() const { return p + 36; }
    SG_HD const uint8_t *cigar() const { return p + 36 + lReadName(); }
    SG_HD const uint8_t *qual() const { return cigar() + 4 * nCigar() + (lSeq() + 1) / 2; }
    SG_HD const uint8_t *aux() const { return qual() + lSeq(); }
    SG_HD const uint8_t *end() const { return p + size(); }
    SG_HD int32_t refSpan() const {           // BAMAlignment::l_ref: reference bases the CIGAR consumes (M D N = X)
        int32_t n = 0; const uint8_t *c = cigar();
        for (uint32_t k = 0; k < nCigar(); k++) { const uint32_t op = sg_ld32(c + 4 * k); const uint32_t t = op & 15u; if (t == 0 || t == 2 || t == 3 || t == 7 || t == 8) n += (int32_t)(op >> 4); }
        return n;
    }
    // getUnclippedStart / getUnclippedEnd (Bam.cpp:462-501)
    SG_HD int64_t unclippedStart(int64_t loc) const {
        if ((flag() & SG_BAM_FLAG_UNMAPPED) || nCigar() == 0) return loc;
        const uint32_t op = sg_ld32(cigar());
        return loc - (((op & 15u) == 4 || (op & 15u) == 5) ? (int64_t)(op >> 4) : 0);
    }
    SG_HD int64_t unclippedEnd(int64_t loc) const {
        if ((flag() & SG_BAM_FLAG_UNMAPPED) || nCigar() == 0) return loc;
        const uint8_t *c = cigar();
        uint32_t op = sg_ld32(c);
        int64_t len = ((op & 15u) == 4 || (op & 15u) == 5) ? 0 : (int64_t)(op >> 4);
        for (uint32_t k = 1; k < nCigar(); k++) {
            op = sg_ld32(c + 4 * k);
            const uint32_t t = op & 15u;
            if (t == 0 || (t >= 2 && t <= 8)) len += (int64_t)(op >> 4);        // op_ref = {1,0,1,1,1,1,1,1,1,0,...} (:495): S and H count here, I does not
        }
        return loc + len;
    }
};

// What duplicate marking needs of one record.
struct SgDupFields {
    int64_t logical;                 // unclipped start of the record's own location, its mate's when it is unmapped itself (:2714-2718)
    int64_t loc, nextLoc;            // getLocation / getNextLocation (Bam.h:178-185)
    uint64_t info, mateInfo;         // (5' end << 1 | strand), (mate's 5' end << 1 | mate's strand) as dupMarkBatch builds them (:2874-2878)
    uint64_t lib;                    // BamDupMarkEntry::hash of the first LB:Z tag, 0 without one (:2856-2871)
    int32_t totalQuality;            // getTotalQuality (:3132-3146)
    int32_t mateQual;                // QS:i tag, -1 without one
    int32_t tile, x, y;              // getTileXY (:3148-3192)
    uint32_t flag;
};

// genome location of (original contig number, position): contigStartByOriginal[ref] = beginningLocation of that contig
SG_HD int64_t sg_dup_location(const int64_t *contigStartByOriginal, int32_t nRef, int32_t refID, int32_t pos, bool unmapped)
{
    if (pos < 0 || refID < 0 || refID >= nRef || unmapped) return SG_DUP_INVALID_LOCATION;
    return contigStartByOriginal[refID] + pos;
}

// sscanf(":%d:%d:%d") as getTileXY uses it: fields that do not parse stay 0
SG_HD void sg_dup_scan3(const uint8_t *t, const uint8_t *tEnd, int32_t *v)
{
    for (int k = 0; k < 3; k++) {
        if (t >= tEnd || *t != ':') return;
        t++;
        while (t < tEnd && (*t == ' ' || *t == '\t' || *t == '\n' || *t == '\r' || *t == '\v' || *t == '\f')) t++;
        bool neg = false;
        if (t < tEnd && (*t == '+' || *t == '-')) { neg = *t == '-'; t++; }
        if (t >= tEnd || *t < '0' || *t > '9') return;
        int64_t val = 0;
        while (t < tEnd && *t >= '0' && *t <= '9') { val = val * 10 + (*t - '0'); if (val > 0x7fffffffLL) val = 0x7fffffffLL; t++; }
        v[k] = (int32_t)(neg ? -val : val);
    }
}

SG_HD void sg_dup_fields(const SgBamRec &r, const int64_t *contigStartByOriginal, int32_t nRef, SgDupFields *o)
{
    const uint32_t flag = r.flag();
    o->flag = flag;
    o->loc = sg_dup_location(contigStartByOriginal, nRef, r.refID(), r.pos(), (flag & SG_BAM_FLAG_UNMAPPED) != 0);
    {
        const int32_t np = r.nextPos(), nr = r.nextRefID();
        o->nextLoc = (np < 0 || nr < 0 || nr >= nRef || (flag & SG_BAM_FLAG_NEXT_UNMAPPED)) ? SG_DUP_INVALID_LOCATION : contigStartByOriginal[nr] + np;
    }
    o->logical = r.unclippedStart(o->loc != SG_DUP_INVALID_LOCATION ? o->loc : o->nextLoc);
    const bool isRC = (flag & SG_BAM_FLAG_RC) != 0, mateRC = (flag & SG_BAM_FLAG_NEXT_RC) != 0;
    const int64_t my = isRC ? r.unclippedEnd(o->loc) : r.unclippedStart(o->loc);
    o->info = ((uint64_t)my << 1) | (isRC ? 1u : 0u);
    o->mateInfo = ((uint64_t)(my + r.tlen()) << 1) | (mateRC ? 1u : 0u);
    int32_t q = 0;
    { const uint8_t *qp = r.qual(); const int32_t n = r.lSeq(); for (int32_t k = 0; k < n; k++) { const int v = qp[k]; if (v >= 15 && v != 255) q += v; } }
    o->totalQuality = q;
    // aux tags: QS:i and the first LB:Z (BAMAlignAux::next / isValidValType)
    o->mateQual = -1; o->lib = 0;
    bool foundLib = false;
    const uint8_t *a = r.aux(), *e = r.end();
    while (a + 3 <= e) {
        const uint8_t t = a[2];
        if (t == 'Z' || t == 'H') {
            const uint8_t *v = a + 3, *z = v;
            while (z < e && *z) z++;
            if (t == 'Z' && a[0] == 'L' && a[1] == 'B' && !foundLib) {
                foundLib = true;
                uint64_t h = 0x123456789abcdef0ULL;
                for (const uint8_t *c = v; c < z; c++) h = (h * 131ULL) ^ (uint64_t)(int64_t)(int8_t)*c;
                o->lib = h;
            }
            a = z + 1;
        } else if (t == 'c' || t == 'C' || t == 'A') a += 4;
        else if (t == 's' || t == 'S') a += 5;
        else if (t == 'i' || t == 'I' || t == 'f') {
            if (t == 'i' && a[0] == 'Q' && a[1] == 'S' && a + 7 <= e) o->mateQual = (int32_t)sg_ld32(a + 3);
            a += 7;
        } else if (t == 'B' && a + 8 <= e) {
            const uint8_t st = a[3]; const uint32_t cnt = sg_ld32(a + 4);
            const uint32_t w = (st == 'c' || st == 'C') ? 1u : (st == 's' || st == 'S') ? 2u : 4u;
            a += 8 + (size_t)cnt * w;
        } else break;
    }
    // tile / x / y of an Illumina read name: elements 3-5 of 5, or 5-7 of 7
    o->tile = o->x = o->y = 0;
    {
        const uint8_t *id = r.name();
        uint32_t n = r.lReadName() ? r.lReadName() - 1 : 0;
        if (n > SG_DUP_BEST_ID - 1) n = SG_DUP_BEST_ID - 1;
        uint32_t colons = 0, five = 0, seven = 0, stop = n;
        for (uint32_t k = 0; k < n; k++) {
            const uint8_t c = id[k];
            if (c == ':') { colons++; if (colons == 2) five = k; else if (colons == 4) seven = k; }
            if (c == 0 || c == ' ' || c == '/') { stop = k; break; }
        }
        (void)stop;
        int32_t v[3] = {0, 0, 0};
        if (colons == 4) sg_dup_scan3(id + five, id + n, v);
        else if (colons == 6) sg_dup_scan3(id + seven, id + n, v);
        o->tile = v[0]; o->x = v[1]; o->y = v[2];
    }
}

// readIdsMatch(best id, record name, l_read_name - 1) (SAM.cpp:39-58) with the best id as DuplicateMateInfo keeps it (strncpy into 120 zeroed bytes):
// the record's whole name must be a prefix of it.
SG_HD bool sg_dup_ids_match(const uint8_t *best, uint32_t bestLen, const uint8_t *name, uint32_t nameLen)
{
    if (bestLen > SG_DUP_BEST_ID) bestLen = SG_DUP_BEST_ID;
    for (uint32_t k = 0; k < nameLen; k++) {
        const uint8_t b = k < bestLen ? best[k] : 0;
        if (b != name[k]) return false;
    }
    return true;
}

// DuplicateMateInfo (Bam.cpp:2533-2593): the best record of a key so far.
struct SgDupBest {
    bool isMateMapped;
    int32_t quality, tile, x, y;
    long long rec;                   // record index whose name is the best id (-1: none yet, the id is empty)
    SG_HD void init() { isMateMapped = false; quality = 0; tile = x = y = 0; rec = -1; }
    SG_HD void take(long long i, int32_t q, int32_t t, int32_t xx, int32_t yy) { quality = q; rec = i; tile = t; x = xx; y = yy; }
    // checkBestRecord for a record that is not flagged a duplicate
    SG_HD void check(long long i, int32_t q, int32_t t, int32_t xx, int32_t yy) {
        if (q > quality) take(i, q, t, xx, yy);
        else if (q == quality) {
            if (t < tile) take(i, q, t, xx, yy);
            else if (t == tile) {
                if (xx < x) take(i, q, t, xx, yy);
                else if (xx == x && yy < y) take(i, q, t, xx, yy);
            }
        }
    }
};

// The tables the key walks read: per record (in stream order) its fields, its bytes, and the runs the stream was cut into.
struct SgDupView {
    long long n;
    const uint8_t *records; const unsigned long long *offsets;     // record i at records + offsets[i]
    const SgDupFields *f;
    long long nRuns; const long long *runStart, *runEnd;           // visited runs, ascending: run k = records [runStart[k], runEnd[k])
    SG_HD SgBamRec rec(long long i) const { SgBamRec r; r.p = records + offsets[i]; return r; }
    // zone of record i: the last run that starts at or before it (-1: none)
    SG_HD long long zone(long long i) const {
        long long lo = 0, hi = nRuns;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (runStart[mid] <= i) lo = mid + 1; else hi = mid; }
        return lo - 1;
    }
    SG_HD bool inRun(long long i, long long k) const { return k >= 0 && k < nRuns && runStart[k] <= i && i < runEnd[k]; }
    SG_HD bool bestMatches(const SgDupBest &b, long long i) const {
        const SgBamRec r = rec(i);
        const uint32_t nameLen = r.lReadName() ? r.lReadName() - 1 : 0;
        if (b.rec < 0) return nameLen == 0;
        const SgBamRec br = rec(b.rec);
        return sg_dup_ids_match(br.name(), br.lReadName() ? br.lReadName() - 1 : 0, r.name(), nameLen);
    }
};

// First run (ascending) after `after` that holds any of members[0, m); -1 if none.  A run ends within the zone after its own (its reach is twice the
// distance at which the next run starts), so a record lies in its zone's run and possibly the one before; a window of four is looked at.
SG_HD long long sg_dup_next_run(const SgDupView &V, const uint32_t *members, long long m, long long after)
{
    long long best = -1;
    for (long long j = 0; j < m; j++) {
        const long long i = members[j];
        const long long z = V.zone(i);
        long long kk = z - 3;
        if (kk < after + 1) kk = after + 1;
        for (; kk <= z; kk++) {
            if (V.inRun(i, kk)) { if (best < 0 || kk < best) best = kk; break; }
        }
    }
    return best;
}

// One PAIR key (library, the two ends' (5' end, strand) in ascending order): members = every record with FLAG 0x1 whose (info, mateInfo) is that
// pair either way round, in stream order.  Replays dupMarkBatch's `run` loops (:2895-2937, :3002-3080) for this key over the runs; flagRun[i] = the run in
// which record i was flagged a duplicate (untouched otherwise).
SG_HD void sg_dup_walk_pair_key(const SgDupView &V, const uint32_t *members, long long m, uint64_t lo, uint64_t hi, int32_t *flagRun)
{
    if (m < 2) return;
    SgDupBest best;
    bool present = false;
    long long k = -1;
    for (;;) {
        k = sg_dup_next_run(V, members, m, k);
        if (k < 0) break;
        if (V.runEnd[k] - V.runStart[k] < 2) continue;               // runs of one record are never marked (:2783)
        // entries of equal (library, info, mateInfo) are adjacent in the run's sorted vector: an entry takes part iff the run holds another one
        long long cnt[2] = {0, 0};
        for (long long j = 0; j < m; j++) if (V.inRun(members[j], k)) cnt[(V.f[members[j]].info == lo) ? 0 : 1]++;
        bool any = false, erase = false;
        for (int side = 0; side < 2 && !(lo == hi && side == 1); side++) {
            if (cnt[side] < 2) continue;
            for (long long j = 0; j < m; j++) {
                const long long i = members[j];
                const SgDupFields &F = V.f[i];
                if (!V.inRun(i, k) || ((F.info == lo) ? 0 : 1) != side) continue;
                if ((F.flag & SG_BAM_FLAG_UNMAPPED) || (F.flag & SG_BAM_FLAG_NEXT_UNMAPPED)) continue;
                if (!present) { best.init(); best.isMateMapped = true; present = true; }
                any = true;
                const bool flagged = (F.flag & SG_BAM_FLAG_DUPLICATE) || flagRun[i] >= 0;
                if (!flagged) best.check(i, F.totalQuality + F.mateQual, F.tile, F.x, F.y);
            }
        }
        if (!any) continue;
        for (int side = 0; side < 2 && !(lo == hi && side == 1); side++) {
            if (cnt[side] < 2) continue;
            for (long long j = 0; j < m; j++) {
                const long long i = members[j];
                const SgDupFields &F = V.f[i];
                if (!V.inRun(i, k) || ((F.info == lo) ? 0 : 1) != side) continue;
                if ((F.flag & SG_BAM_FLAG_UNMAPPED) || (F.flag & SG_BAM_FLAG_NEXT_UNMAPPED)) continue;
                if (!V.bestMatches(best, i) && flagRun[i] < 0) flagRun[i] = (int32_t)k;
                // the entry leaves the map with the record that closes the pair (:3064-3078)
                const int64_t my = (int64_t)(F.info >> 1);
                const int64_t spacing = F.nextLoc > my ? F.nextLoc - my : my - F.nextLoc;
                if (spacing > 0x7fffffffLL || F.info == hi || (F.info == lo && F.loc > F.nextLoc)) erase = true;
            }
        }
        if (erase) present = false;
    }
}

// One FRAGMENT key (library, 5' end and strand): members = every record with that key, in stream order.  Replays the `runFragment` loops (:2940-2999,
// :3083-3114).  pairFlagRun = what the pair walks left (a record flagged there counts as a duplicate from that run on); fragFlag[i] = 1 for records flagged here.
SG_HD void sg_dup_walk_fragment_key(const SgDupView &V, const uint32_t *members, long long m, const int32_t *pairFlagRun, uint8_t *fragFlag)
{
    if (m < 2) return;
    long long k = -1;
    for (;;) {
        k = sg_dup_next_run(V, members, m, k);
        if (k < 0) break;
        if (V.runEnd[k] - V.runStart[k] < 2) continue;
        long long cnt = 0;
        for (long long j = 0; j < m; j++) if (V.inRun(members[j], k)) cnt++;
        if (cnt < 2) continue;
        SgDupBest best;
        bool have = false;
        for (long long j = 0; j < m; j++) {
            const long long i = members[j];
            const SgDupFields &F = V.f[i];
            if (!V.inRun(i, k) || (F.flag & SG_BAM_FLAG_UNMAPPED)) continue;
            const bool mateMapped = (F.flag & SG_BAM_FLAG_PAIRED) && !(F.flag & SG_BAM_FLAG_NEXT_UNMAPPED);
            if (!have) { best.init(); best.isMateMapped = mateMapped; have = true; }
            const bool flagged = (F.flag & SG_BAM_FLAG_DUPLICATE) || (pairFlagRun[i] >= 0 && pairFlagRun[i] <= k) || fragFlag[i];
            if (mateMapped) {
                if (!best.isMateMapped) { best.take(i, F.totalQuality, F.tile, F.x, F.y); best.isMateMapped = true; }      // a mapped pair beats any fragment (:2977-2983)
                else if (!flagged) best.check(i, F.totalQuality, F.tile, F.x, F.y);
            } else if (!best.isMateMapped && !flagged) best.check(i, F.totalQuality, F.tile, F.x, F.y);
        }
        if (!have) continue;
        for (long long j = 0; j < m; j++) {
            const long long i = members[j];
            const SgDupFields &F = V.f[i];
            if (!V.inRun(i, k) || (F.flag & SG_BAM_FLAG_UNMAPPED)) continue;
            if ((F.flag & SG_BAM_FLAG_PAIRED) && !(F.flag & SG_BAM_FLAG_NEXT_UNMAPPED)) continue;          // only true fragments are flagged here (:3096)
            if (!V.bestMatches(best, i)) fragFlag[i] = 1;
        }
    }
}

// Run geometry (onNextBatch, :2705-2757): for a run starting at record s, where it ends (first record after s whose unclipped start is more than `reach`
// beyond s's; n if none).  reach = 2 * SG_DUP_RUN_REACH for the run's end, SG_DUP_RUN_REACH for the start of the next run.
SG_HD long long sg_dup_first_beyond(const SgDupFields *f, long long n, long long s, int64_t reach)
{
    const int64_t limit = f[s].logical + reach;
    long long i = s + 1;
    while (i < n && !(f[i].logical > limit)) i++;
    return i;
}

// ---- .bai (BAMIndexSupplier::onRead / onClosed, Bam.cpp:3313-3392) ----
// per record: is it the head of a chunk (first record of a maximal stretch of equal (refID, bin)), and which 16 Kbp window its linear-index entry falls in
SG_HD int32_t sg_bai_linear_slot(const SgBamRec &r)
{
    const int32_t end = r.pos() + r.refSpan() - 1;      // addInterval(refID, pos, pos + l_ref - 1, ...) files the record under its END's window ...
    return end <= 0 ? 0 : (end - 1) / 16384;            // ... (end - 1) / 16384 (:3433)
}

// virtual file offset of uncompressed offset u in a file made of BGZF members of SG_BGZF_PAYLOAD_BYTES payload bytes each (snapgpu_bgzf_device) followed by the
// 28-byte end-of-file member; the very end of the data translates to the end of the file, as GzipWriterFilterSupplier::translate has it.
#define SG_BGZF_PAYLOAD_BYTES 0xff00ULL
// memberOffsets (optional, nMembers + 1 entries): where each member starts in the file when the members are compressed (snapgpu_bgzf_deflate_device), the
// last entry being the end of the data members; NULL = stored members of fixed size.
SG_HD uint64_t sg_bai_virtual_offset(uint64_t u, uint64_t totalBytes, const uint64_t *memberOffsets = (const uint64_t *)0)
{
    const uint64_t nMembers = (totalBytes + SG_BGZF_PAYLOAD_BYTES - 1) / SG_BGZF_PAYLOAD_BYTES;
    if (memberOffsets) {
        if (u >= totalBytes) return (memberOffsets[nMembers] + 28ULL) << 16;
        return memberOffsets[u / SG_BGZF_PAYLOAD_BYTES] << 16 | (u % SG_BGZF_PAYLOAD_BYTES);
    }
    if (u >= totalBytes) return (nMembers * 31ULL + totalBytes + 28ULL) << 16;
    return ((u / SG_BGZF_PAYLOAD_BYTES) * (SG_BGZF_PAYLOAD_BYTES + 31ULL)) << 16 | (u % SG_BGZF_PAYLOAD_BYTES);
}

// ---- host side of the .bai: the file itself (BAMIndexSupplier::onClosed, Bam.cpp:3341-3392).  Inputs in UNCOMPRESSED offsets of the stream header ‖ records;
//      the chunks of a bin in stream order.  Bins are written in ascending order (the reference writes them in the iteration order of its hash map; readers do not care). ----
#include <vector>
#include <algorithm>
struct SgBaiChunk { int32_t ref; uint32_t bin; uint64_t start, end; };
struct SgBaiRef {
    bool any = false;
    uint64_t firstStart = 0, lastEnd = 0, mapped = 0, unmapped = 0;
    std::vector<uint64_t> intervals;     // ~0ULL = never set (written as 0, like toVirtualOffset(UINT64_MAX), GzipDataWriter.h:69-79)
};
inline void sg_bai_put64(std::vector<uint8_t> &o, uint64_t v) { for (int k = 0; k < 8; k++) o.push_back((uint8_t)(v >> (8 * k))); }
inline void sg_bai_put32(std::vector<uint8_t> &o, uint32_t v) { for (int k = 0; k < 4; k++) o.push_back((uint8_t)(v >> (8 * k))); }
// chunks: every maximal stretch of equal (refID, bin) in stream order (refID < 0 or >= nRef are dropped here, as addChunk drops them)
inline std::vector<uint8_t> sg_bai_compose(int32_t nRef, std::vector<SgBaiChunk> chunks, const std::vector<SgBaiRef> &refs, uint64_t totalBytes,
                                           const uint64_t *memberOffsets = (const uint64_t *)0)
{
    std::vector<uint8_t> o;
    o.push_back('B'); o.push_back('A'); o.push_back('I'); o.push_back(1);
    sg_bai_put32(o, (uint32_t)nRef);
    chunks.erase(std::remove_if(chunks.begin(), chunks.end(), [nRef](const SgBaiChunk &c) { return c.ref < 0 || c.ref >= nRef; }), chunks.end());
    std::stable_sort(chunks.begin(), chunks.end(), [](const SgBaiChunk &a, const SgBaiChunk &b) { return a.ref != b.ref ? a.ref < b.ref : a.bin < b.bin; });
    size_t c = 0;
    for (int32_t r = 0; r < nRef; r++) {
        size_t c1 = c;
        uint32_t nBin = 0;
        while (c1 < chunks.size() && chunks[c1].ref == r) { if (c1 == c || chunks[c1].bin != chunks[c1 - 1].bin) nBin++; c1++; }
        const SgBaiRef &R = refs[r];
        sg_bai_put32(o, nBin + (R.any ? 1u : 0u));
        while (c < c1) {
            size_t e = c;
            while (e < c1 && chunks[e].bin == chunks[c].bin) e++;
            sg_bai_put32(o, chunks[c].bin); sg_bai_put32(o, (uint32_t)(e - c));
            for (; c < e; c++) { sg_bai_put64(o, sg_bai_virtual_offset(chunks[c].start, totalBytes, memberOffsets)); sg_bai_put64(o, sg_bai_virtual_offset(chunks[c].end, totalBytes, memberOffsets)); }
        }
        if (R.any) {
            sg_bai_put32(o, SG_BAM_EXTRA_BIN); sg_bai_put32(o, 2);
            sg_bai_put64(o, sg_bai_virtual_offset(R.firstStart, totalBytes, memberOffsets)); sg_bai_put64(o, sg_bai_virtual_offset(R.lastEnd, totalBytes, memberOffsets));
            sg_bai_put64(o, R.mapped); sg_bai_put64(o, R.unmapped);
        }
        sg_bai_put32(o, (uint32_t)R.intervals.size());
        for (uint64_t v : R.intervals) sg_bai_put64(o, v == ~0ULL ? 0ULL : sg_bai_virtual_offset(v, totalBytes, memberOffsets));
    }
    return o;
}
