// sg_fastq.cuh -- FASTQ ingest on the device: text in HBM -> clipped, upper-cased reads in the layout the
// alignment kernels take (concatenated bases / qualities + offsets + lengths), plus where each read's id sits in the text.
//
// Restates, for a whole buffer at once, FASTQReader::getReadFromBuffer (reference SNAPLib/FASTQ.cpp:229-300: four lines per
// record, '\r' stripped from line ends, first-character validation of every line, id = up to the first space),
// Read::init's upper-casing / '.' -> 'N' (Read.h:465-492, Tables.cpp:94-104) and Read::clip (Read.h:567-619: quality '#'
// runs clipped from the back and/or front).  Like the reference, the quality line is read over the SEQUENCE's length.
//
// Byte-stream work, HBM-bound: K1 counts the newlines of 2 KB tiles with 16-byte loads, a scan turns the counts into
// bases, K2 writes every newline's offset in order, K3 (one thread per record) validates and clips, a scan of the clipped
// lengths gives the output offsets, K4 (one warp per record) copies.  Text is read twice (K1, K2) + once more by K4.
#pragma once
#include <stdint.h>
#include <cub/cub.cuh>

#define SG_FQ_TILE 2048              // bytes per tile
#define SG_FQ_THREADS 128            // 16 bytes per thread
#define SG_FQ_ERR_BLANK_LINE 1
#define SG_FQ_ERR_BAD_START 2
#define SG_FQ_ERR_TOO_LONG 3

// 0x80 in every byte of x that equals '\n' (exact per byte, no borrow artefacts)
__device__ __forceinline__ uint32_t sg_fq_nl_flags(uint32_t x)
{
    const uint32_t y = x ^ 0x0a0a0a0au;
    const uint32_t t = (y & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | y | 0x7f7f7f7fu);
}

__global__ void __launch_bounds__(SG_FQ_THREADS)
sg_fastq_count_kernel(const uint8_t *text, long long nBytes, uint32_t *tileCounts, long long nTiles)
{
    for (long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const long long off = tile * SG_FQ_TILE + (long long)threadIdx.x * 16;
        uint32_t c = 0;
        if (off + 16 <= nBytes) {
            const uint4 v = *(const uint4 *)(text + off);
            c = __popc(sg_fq_nl_flags(v.x)) + __popc(sg_fq_nl_flags(v.y)) + __popc(sg_fq_nl_flags(v.z)) + __popc(sg_fq_nl_flags(v.w));
        } else {
            for (long long k = off; k < nBytes && k < off + 16; k++) c += (text[k] == '\n');
        }
        typedef cub::BlockReduce<uint32_t, SG_FQ_THREADS> Reduce;
        __shared__ typename Reduce::TempStorage tmp;
        const uint32_t total = Reduce(tmp).Sum(c);
        if (threadIdx.x == 0) tileCounts[tile] = total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SG_FQ_THREADS)
sg_fastq_positions_kernel(const uint8_t *text, long long nBytes, const uint32_t *tileBase, uint32_t *nlPos, long long maxLines, long long nTiles)
{
    for (long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const long long off = tile * SG_FQ_TILE + (long long)threadIdx.x * 16;
        uint32_t f[4] = {0, 0, 0, 0};
        if (off + 16 <= nBytes) {
            const uint4 v = *(const uint4 *)(text + off);
            f[0] = sg_fq_nl_flags(v.x); f[1] = sg_fq_nl_flags(v.y); f[2] = sg_fq_nl_flags(v.z); f[3] = sg_fq_nl_flags(v.w);
        } else {
            for (long long k = off; k < nBytes && k < off + 16; k++) if (text[k] == '\n') f[(k - off) >> 2] |= 0x80u << (8 * ((k - off) & 3));
        }
        const uint32_t c = __popc(f[0]) + __popc(f[1]) + __popc(f[2]) + __popc(f[3]);
        typedef cub::BlockScan<uint32_t, SG_FQ_THREADS> Scan;
        __shared__ typename Scan::TempStorage tmp;
        uint32_t before;
        Scan(tmp).ExclusiveSum(c, before);
        long long w = (long long)tileBase[tile] + before;
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t m = f[j];
            while (m) {
                const int bit = __ffs(m) - 1;          // 7, 15, 23 or 31
                if (w < maxLines) nlPos[w] = (uint32_t)(off + 4 * j + (bit >> 3));
                w++;
                m &= m - 1;
            }
        }
        __syncthreads();
    }
}

struct SgFastqRecord {               // per record, filled by K3
    uint32_t dataStart, qualStart;   // offsets in the text of the first KEPT base / quality
    uint32_t idStart, idLen;
    uint32_t len, frontClipped;
};

__device__ __forceinline__ bool sg_fq_valid_base_start(uint8_t c)
{
    // "ACTGNURYKMSWBDHVNX." in either case (FASTQ.cpp:327-331)
    const uint8_t u = (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c;
    switch (u) {
        case 'A': case 'C': case 'T': case 'G': case 'N': case 'U': case 'R': case 'Y': case 'K': case 'M': case 'S': case 'W': case 'B': case 'D':
        case 'H': case 'V': case 'X': case '.': return true;
        default: return false;
    }
}

// clippingType: 0 none, 1 front, 2 back, 3 front and back (Read.h:88); quality in [minPhred, maxPhred] is clipped.
__global__ void sg_fastq_records_kernel(const uint8_t *text, long long nBytes, const uint32_t *nlPos, long long nRecords, int clippingType, uint8_t minPhred,
                                        uint8_t maxPhred, uint32_t maxReadLen, SgFastqRecord *rec, uint32_t *lens, unsigned long long *idOffsets,
                                        uint32_t *idLens, uint32_t *frontClipped, int *status)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRecords) return;
    uint32_t start[4], len[4];
    int err = 0;
    #pragma unroll
    for (int j = 0; j < 4; j++) {
        const long long line = 4 * r + j;
        uint32_t s = 0;
        if (line > 0) { const uint32_t p = nlPos[line - 1]; s = p + 1 + (((long long)p + 1 < nBytes && text[p + 1] == '\r') ? 1u : 0u); }     // scan = newLine + (newLine[1] == '\r' ? 2 : 1)
        const uint32_t e = nlPos[line];
        start[j] = s;
        if (e <= s) { err = err ? err : SG_FQ_ERR_BLANK_LINE; len[j] = 0; continue; }                           // (an "\n\r\n" also lands here)
        len[j] = e - s - (text[e - 1] == '\r' ? 1u : 0u);
        const uint8_t c = text[s];
        const bool ok = (j == 0) ? (c == '@') : (j == 1) ? sg_fq_valid_base_start(c) : (j == 2) ? (c == '+') : (c >= '!' && c <= '~');
        if (!ok) err = err ? err : SG_FQ_ERR_BAD_START;
    }
    uint32_t dataLength = len[1];
    if (dataLength > maxReadLen) err = err ? err : SG_FQ_ERR_TOO_LONG;
    if (err) {
        atomicCAS(status, 0, err);
        rec[r].len = 0; lens[r] = 0;
        if (idOffsets) idOffsets[r] = 0;
        if (idLens) idLens[r] = 0;
        if (frontClipped) frontClipped[r] = 0;
        return;
    }
    // (the reference reads the quality line over the sequence's length even when it is shorter; past the buffer we read 0)
    #define SG_FQ_Q(k) (((long long)start[3] + (k) < nBytes) ? text[start[3] + (k)] : (uint8_t)0)
    if (clippingType == 2 || clippingType == 3) {
        while (dataLength > 0 && SG_FQ_Q(dataLength - 1) >= minPhred && SG_FQ_Q(dataLength - 1) <= maxPhred) dataLength--;
    }
    uint32_t front = 0;
    if (clippingType == 1 || clippingType == 3) {
        while (front < dataLength && SG_FQ_Q(front) >= minPhred && SG_FQ_Q(front) <= maxPhred) front++;
    }
    #undef SG_FQ_Q
    dataLength -= front;
    // id: after '@', up to the first space (FASTQ.cpp:276-293, preserveFASTQComments = false)
    uint32_t idLen = len[0] - 1;
    for (uint32_t k = 0; k < len[0] - 1; k++) { if (text[start[0] + 1 + k] == ' ') { idLen = k; break; } }
    SgFastqRecord o;
    o.dataStart = start[1] + front; o.qualStart = start[3] + front; o.idStart = start[0] + 1; o.idLen = idLen; o.len = dataLength; o.frontClipped = front;
    rec[r] = o;
    lens[r] = dataLength;
    if (idOffsets) idOffsets[r] = o.idStart;
    if (idLens) idLens[r] = idLen;
    if (frontClipped) frontClipped[r] = front;
}

// L = number of newlines found; records = min(L / 4, maxReads); consumed = first byte after the last complete record
__global__ void sg_fastq_meta_kernel(const uint8_t *text, long long nBytes, const uint32_t *tileBase, long long nTiles, const uint32_t *nlPos, long long maxReads,
                                     long long *meta)
{
    const long long L = tileBase[nTiles];
    long long R = L / 4;
    if (R > maxReads) R = maxReads;
    long long consumed = 0;
    if (R > 0) {
        const uint32_t p = nlPos[4 * R - 1];
        consumed = (long long)p + 1 + (((long long)p + 1 < nBytes && text[p + 1] == '\r') ? 1 : 0);
    }
    meta[0] = L; meta[1] = R; meta[2] = consumed;
}

// One warp per record: upper-case + '.' -> 'N' (Tables.cpp:94-104) while copying.  Destination words: after the (up to 3) bytes that bring the
// destination to a 4-byte boundary every lane stores one 32-bit word per step, its four source bytes cut out of two aligned source words with a
// funnel shift and translated four at a time (byte-wise compares: __vcmp*4); the text is read once per array, 4 bytes per load instead of 1.
__device__ __forceinline__ uint32_t sg_fastq_upper4(uint32_t w)
{
    const uint32_t lower = __vcmpgeu4(w, 0x61616161u) & __vcmpleu4(w, 0x7a7a7a7au);
    w -= lower & 0x20202020u;
    const uint32_t dot = __vcmpeq4(w, 0x2e2e2e2eu);
    return (w & ~dot) | (0x4e4e4e4eu & dot);
}
__device__ __forceinline__ uint32_t sg_fastq_load4(const uint8_t *text, long long nBytes, long long at)      // bytes [at, at + 4) of the text, 0 past its end
{
    const long long w0 = at & ~3LL;
    if (at >= 0 && w0 + 8 <= nBytes) {
        const uint32_t a = *(const uint32_t *)(text + w0), b = *(const uint32_t *)(text + w0 + 4);
        return __funnelshift_r(a, b, (uint32_t)(at & 3) * 8u);
    }
    uint32_t v = 0;
    for (int k = 0; k < 4; k++) if (at + k >= 0 && at + k < nBytes) v |= (uint32_t)text[at + k] << (8 * k);
    return v;
}
__global__ void sg_fastq_copy_kernel(const uint8_t *text, long long nBytes, const SgFastqRecord *rec, const unsigned long long *offsets, long long nRecords,
                                     uint8_t *bases, uint8_t *quals)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nWarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const bool aligned = (((size_t)text | (size_t)bases | (size_t)quals) & 3) == 0;
    for (long long r = warp; r < nRecords; r += nWarps) {
        const SgFastqRecord o = rec[r];
        const unsigned long long dst = offsets[r];
        uint32_t head = aligned ? (uint32_t)((4 - (dst & 3)) & 3) : o.len;          // bytes before the destination's first whole word
        if (head > o.len) head = o.len;
        const uint32_t words = (o.len - head) >> 2, tailAt = head + 4 * words;
        for (uint32_t k = lane; k < head; k += 32) {
            uint8_t c = text[o.dataStart + k];
            if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
            else if (c == '.') c = 'N';
            bases[dst + k] = c;
            quals[dst + k] = ((long long)o.qualStart + k < nBytes) ? text[o.qualStart + k] : (uint8_t)0;
        }
        for (uint32_t w = lane; w < words; w += 32) {
            const uint32_t k = head + 4 * w;
            *(uint32_t *)(bases + dst + k) = sg_fastq_upper4(sg_fastq_load4(text, nBytes, (long long)o.dataStart + k));
            *(uint32_t *)(quals + dst + k) = sg_fastq_load4(text, nBytes, (long long)o.qualStart + k);
        }
        for (uint32_t k = tailAt + lane; k < o.len; k += 32) {
            uint8_t c = text[o.dataStart + k];
            if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
            else if (c == '.') c = 'N';
            bases[dst + k] = c;
            quals[dst + k] = ((long long)o.qualStart + k < nBytes) ? text[o.qualStart + k] : (uint8_t)0;
        }
    }
}
