// warpsim.h -- TEST-ONLY: a 32-lane SIMT emulator for the warp-cooperative device headers (snap_b200/csrc/*.cuh).
//
// Each lane of a warp is a fiber (own stack, hand-rolled context switch) running the device function as ordinary scalar
// C++.  A warp-synchronous primitive (__shfl_sync, __ballot_sync, __reduce_*_sync, __syncwarp, ...) parks the calling lane;
// when every live lane of the warp has parked AT THE SAME KIND of primitive the exchange is carried out and the lanes resume.
// Lanes that park at different primitives (divergent code around a warp-synchronous call) abort the run: the device
// headers are written warp-convergent and this checks it.  The DPX s16x2 intrinsics are restated per CUDA's documentation.
// With it the CPU suite (`-m "not gpu"`) runs the very code the kernels run against the compiled reference.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <functional>

namespace ws {

enum Op { OP_NONE = 0, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_REDUCE_MAX_S, OP_REDUCE_MAX_U, OP_REDUCE_MIN_S, OP_REDUCE_ADD,
          OP_REDUCE_OR, OP_REDUCE_AND, OP_SYNC, OP_DONE };

struct Lane {
    void *sp;                 // saved stack pointer of the parked fiber
    uint8_t *stack;
    int op;
    unsigned mask;
    uint64_t val;             // operand
    int arg;                  // source lane / delta
    uint64_t result;
    bool done;
};

struct Warp {
    Lane lanes[32];
    void *mainSp;
    int current;
    std::function<void(int)> body;
    long long syncOps;
};

extern thread_local Warp *g_warp;

extern "C" void ws_switch(void **saveSp, void *newSp);
void run_warp(const std::function<void(int)> &body, size_t stackBytes = 512 * 1024);
uint64_t park(int op, unsigned mask, uint64_t val, int arg);
inline int lane_id() { return g_warp->current; }

} // namespace ws

// ---- CUDA spellings the device headers use ----
#define __device__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define WS_EMULATED 1

static inline int __shfl_sync(unsigned m, int v, int src) { return (int)(uint32_t)ws::park(ws::OP_SHFL, m, (uint32_t)v, src & 31); }
static inline unsigned __shfl_sync(unsigned m, unsigned v, int src) { return (unsigned)ws::park(ws::OP_SHFL, m, v, src & 31); }
static inline int __shfl_up_sync(unsigned m, int v, unsigned d) { return (int)(uint32_t)ws::park(ws::OP_SHFL_UP, m, (uint32_t)v, (int)d); }
static inline unsigned __shfl_up_sync(unsigned m, unsigned v, unsigned d) { return (unsigned)ws::park(ws::OP_SHFL_UP, m, v, (int)d); }
static inline int __shfl_down_sync(unsigned m, int v, unsigned d) { return (int)(uint32_t)ws::park(ws::OP_SHFL_DOWN, m, (uint32_t)v, (int)d); }
static inline unsigned __shfl_down_sync(unsigned m, unsigned v, unsigned d) { return (unsigned)ws::park(ws::OP_SHFL_DOWN, m, v, (int)d); }
static inline int __shfl_xor_sync(unsigned m, int v, int x) { return (int)(uint32_t)ws::park(ws::OP_SHFL_XOR, m, (uint32_t)v, x); }
static inline unsigned __shfl_xor_sync(unsigned m, unsigned v, int x) { return (unsigned)ws::park(ws::OP_SHFL_XOR, m, v, x); }
static inline unsigned __ballot_sync(unsigned m, int pred) { return (unsigned)ws::park(ws::OP_BALLOT, m, pred ? 1 : 0, 0); }
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
static inline int __reduce_max_sync(unsigned m, int v) { return (int)(uint32_t)ws::park(ws::OP_REDUCE_MAX_S, m, (uint32_t)v, 0); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)ws::park(ws::OP_REDUCE_MAX_U, m, v, 0); }
static inline int __reduce_min_sync(unsigned m, int v) { return (int)(uint32_t)ws::park(ws::OP_REDUCE_MIN_S, m, (uint32_t)v, 0); }
static inline int __reduce_add_sync(unsigned m, int v) { return (int)(uint32_t)ws::park(ws::OP_REDUCE_ADD, m, (uint32_t)v, 0); }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)ws::park(ws::OP_REDUCE_ADD, m, v, 0); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return (unsigned)ws::park(ws::OP_REDUCE_OR, m, v, 0); }
static inline unsigned __reduce_and_sync(unsigned m, unsigned v) { return (unsigned)ws::park(ws::OP_REDUCE_AND, m, v, 0); }
static inline void __syncwarp(unsigned m = 0xffffffffu) { (void)ws::park(ws::OP_SYNC, m, 0, 0); }

static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }

// ---- DPX / SIMD-in-a-word intrinsics (CUDA Math API, "SIMD intrinsics"): per 16-bit half, signed ----
static inline int ws_lo(unsigned x) { return (int)(int16_t)(x & 0xffffu); }
static inline int ws_hi(unsigned x) { return (int)(int16_t)(x >> 16); }
static inline unsigned ws_pack(int lo, int hi) { return ((unsigned)(uint16_t)(int16_t)hi << 16) | (unsigned)(uint16_t)(int16_t)lo; }
static inline int ws_max(int a, int b) { return a > b ? a : b; }
static inline unsigned __vadd2(unsigned a, unsigned b) { return ws_pack(ws_lo(a) + ws_lo(b), ws_hi(a) + ws_hi(b)); }            // wrapping
static inline unsigned __vsub2(unsigned a, unsigned b) { return ws_pack(ws_lo(a) - ws_lo(b), ws_hi(a) - ws_hi(b)); }
static inline unsigned __vimax_s16x2(unsigned a, unsigned b) { return ws_pack(ws_max(ws_lo(a), ws_lo(b)), ws_max(ws_hi(a), ws_hi(b))); }
static inline unsigned __vimax_s16x2_relu(unsigned a, unsigned b) { return ws_pack(ws_max(ws_max(ws_lo(a), ws_lo(b)), 0), ws_max(ws_max(ws_hi(a), ws_hi(b)), 0)); }
static inline unsigned __vibmax_s16x2(unsigned a, unsigned b, bool *predHi, bool *predLo)                                       // pred: a >= b
{
    *predHi = ws_hi(a) >= ws_hi(b); *predLo = ws_lo(a) >= ws_lo(b);
    return __vimax_s16x2(a, b);
}
static inline unsigned __vmaxs2(unsigned a, unsigned b) { return __vimax_s16x2(a, b); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    const uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int k = 0; k < 4; k++) r |= (unsigned)((v >> (8 * ((s >> (4 * k)) & 7))) & 0xff) << (8 * k);
    return r;
}
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) { return __vimax_s16x2(__vadd2(a, b), c); }
static inline unsigned __viaddmax_s16x2_relu(unsigned a, unsigned b, unsigned c) { return __vimax_s16x2_relu(__vadd2(a, b), c); }
