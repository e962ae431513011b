// tests/blocksim -- TEST-ONLY: snap_b200/csrc/sg_deflate.h (the block-cooperative deflate the CUDA kernel compiles) run by a block of REAL host
// threads: threadIdx = a thread-local index, __syncthreads = a pthread barrier, shared-memory atomics = relaxed atomic builtins.  tests/hostsim runs
// the same header with one thread (every loop to its end, barriers no-ops); this build exercises what that cannot: the per-thread index ranges,
// the per-warp histograms, the strided loops and the placement of the barriers.  A missing barrier shows up here as a member that does not inflate
// (or, under -fsanitize=thread, as a reported race).
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <vector>

static thread_local uint32_t t_tid;
static uint32_t g_nt = 1;
static pthread_barrier_t g_bar;
#define SGD_CUSTOM_THREADS
#define SGD_TID t_tid
#define SGD_NT g_nt
#define SGD_WARP (t_tid >> 5)
#define SGD_SYNC() pthread_barrier_wait(&g_bar)
#define SGD_ATOMIC_OR(p, v) __atomic_fetch_or((p), (v), __ATOMIC_RELAXED)
#define SGD_ATOMIC_ADD(p, v) __atomic_fetch_add((p), (v), __ATOMIC_RELAXED)
#include "../../snap_b200/csrc/sg_deflate.h"

struct Job { SgDeflateShared *S; SgDeflateArena G; const uint8_t *src; uint32_t n; uint8_t *member; uint32_t tid; uint32_t size; };

static void *worker(void *v)
{
    Job *j = (Job *)v;
    t_tid = j->tid;
    j->size = sg_deflate_member(*j->S, j->G, j->src, j->n, j->member);
    return nullptr;
}

extern "C" {
// Same contract as hs_bgzf_deflate (tests/hostsim), with nThreads (a multiple of 32, <= 1024) threads per member.
int64_t bs_bgzf_deflate(const uint8_t *in, int64_t n, uint8_t *out, int64_t cap, uint32_t *memberSizes, int nThreads)
{
    if (nThreads < 32 || nThreads > 1024 || nThreads % 32) return -2;
    static SgDeflateShared S;
    std::vector<uint16_t> arena(SG_DEFLATE_ARENA_BYTES / 2);
    std::vector<uint8_t> member(SG_DEFLATE_MEMBER_PITCH + 64);
    std::vector<Job> jobs((size_t)nThreads);
    std::vector<pthread_t> th((size_t)nThreads);
    g_nt = (uint32_t)nThreads;
    int64_t used = 0, m = 0;
    for (int64_t off = 0; off < n; off += SG_DEFLATE_MAX_PAYLOAD, m++) {
        const uint32_t len = (uint32_t)((n - off) < (int64_t)SG_DEFLATE_MAX_PAYLOAD ? (n - off) : (int64_t)SG_DEFLATE_MAX_PAYLOAD);
        pthread_barrier_init(&g_bar, nullptr, (unsigned)nThreads);
        for (int t = 0; t < nThreads; t++) {
            Job &j = jobs[(size_t)t];
            j.S = &S; j.G.mlen = arena.data(); j.G.mdist = j.G.mlen + (SG_DEFLATE_MAX_PAYLOAD + 8); j.G.jumpA = j.G.mdist + (SG_DEFLATE_MAX_PAYLOAD + 8);
            j.G.jumpB = j.G.jumpA + (SG_DEFLATE_MAX_PAYLOAD + 8);
            j.src = in + off; j.n = len; j.member = member.data(); j.tid = (uint32_t)t; j.size = 0;
            pthread_create(&th[(size_t)t], nullptr, worker, &j);
        }
        for (int t = 0; t < nThreads; t++) pthread_join(th[(size_t)t], nullptr);
        pthread_barrier_destroy(&g_bar);
        const uint32_t sz = jobs[0].size;
        for (int t = 1; t < nThreads; t++) if (jobs[(size_t)t].size != sz) return -3;        // every thread must return the same size
        if (used + sz > cap) return -1;
        memcpy(out + used, member.data(), sz);
        if (memberSizes) memberSizes[m] = sz;
        used += sz;
    }
    return used;
}
}
