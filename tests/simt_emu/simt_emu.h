// simt_emu.h -- a small SIMT emulator for CPU-side tests of the fused CUDA kernels.  TEST INFRASTRUCTURE ONLY.
//
// It is NOT a CPU fallback of the product: nothing under socp.jl_b200/ includes or links it, libsocp_b200.so has no
// CPU code path, and the library built from it (tests/_build/libsocp_emu.so) is loaded by tests/ only.  What it is
// for: the whole-solve kernels are long state machines (barriers, warp shuffles, mma fragments, named barriers,
// tables in shared memory) and the GPU is a scarce, remote resource; compiling the SAME kernel source for the host
// against this header lets `pytest -m "not gpu"` run the kernel's logic against the oracle, under gdb / ASan /
// UBSan, before any GPU time is spent.
//
// Model: one fibre (ucontext) per CUDA thread of one CTA; CTAs of a launch run one after the other.  A fibre runs
// until it blocks in a collective (CTA barrier, named barrier, warp shuffle / vote / mma.sync, __syncwarp) and then
// yields; the scheduler walks the fibres round-robin (order selectable: forward, reverse, random -- a missing
// barrier shows up as a wrong answer under at least one of them).  Divergent collectives (a FULL_MASK shuffle that
// not every lane reaches) dead-lock; the scheduler detects that and aborts with the positions of the fibres.
//
// Implemented CUDA surface (what the kernels use): threadIdx/blockIdx/blockDim/gridDim (.x), __syncthreads,
// __syncwarp, __shfl_xor_sync, __shfl_sync, __any_sync, atomicAdd(int*), clock64, dynamic + static __shared__,
// bar.sync / bar.arrive (emu_bar_sync / emu_bar_arrive), mma.sync.m8n8k4.f64 (emu_dmma884).
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <random>
#include <vector>

namespace simt_emu {

struct Idx3 { unsigned x = 0, y = 0, z = 0; };

struct Warp {
    uint64_t xbuf[2][32];
    double ma[2][32], mb[2][32];
    int bar_cnt = 0;
    unsigned bar_gen = 0;
    int nlanes = 32;               // lanes the warp was launched with
    int nlive = 32;                // ... that have not returned yet: what a warp collective waits for
    unsigned xseq[2][32] = {};     // which collective (per-lane running count) wrote xbuf[par][lane]: a lane that
                                   // returned before this collective holds a stale entry and does not take part
};

struct Block;
struct Thread {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    Idx3 tidx;
    int lane = 0, warp = 0;
    unsigned xpar = 0;             // parity of the warp-collective exchange buffers
    const char* where = "running";
    Block* blk = nullptr;
};
struct Block {
    std::vector<Thread> th;
    std::vector<Warp> warps;
    int bar_cnt[16];
    unsigned bar_gen[16];
    char* smem = nullptr;
    Idx3 bidx, bdim, gdim;
    ucontext_t sched;
    unsigned long progress = 0;    // bumped whenever a barrier completes or a fibre ends
};

inline Block*& cur_block() { static Block* b = nullptr; return b; }
inline Thread*& cur_thread() { static Thread* t = nullptr; return t; }

inline void yield(const char* where) {
    Thread* t = cur_thread();
    t->where = where;
    swapcontext(&t->ctx, &t->blk->sched);
    t->where = "running";
}

inline void warp_barrier(const char* where) {
    Thread* t = cur_thread();
    Warp& w = t->blk->warps[t->warp];
    const unsigned gen = w.bar_gen;
    if (++w.bar_cnt == w.nlive) {
        w.bar_cnt = 0;
        w.bar_gen++;
        t->blk->progress++;
    } else {
        while (w.bar_gen == gen) yield(where);
    }
}
inline void block_bar(int id, int count, bool wait, const char* where) {
    Block* b = cur_block();
    const unsigned gen = b->bar_gen[id];
    if (++b->bar_cnt[id] == count) {
        b->bar_cnt[id] = 0;
        b->bar_gen[id]++;
        b->progress++;
    } else if (wait) {
        while (b->bar_gen[id] == gen) yield(where);
    }
}

template <class T>
inline T exchange(T v, int src_lane_xor, int src_lane_abs, const char* where) {
    static_assert(sizeof(T) <= 8, "exchange of at most 8 bytes");
    Thread* t = cur_thread();
    Warp& w = t->blk->warps[t->warp];
    const unsigned par = (t->xpar++) & 1u;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.xbuf[par][t->lane] = bits;
    w.xseq[par][t->lane] = t->xpar;
    warp_barrier(where);
    const int src = src_lane_abs >= 0 ? (src_lane_abs & 31) : (t->lane ^ src_lane_xor);
    T r;
    const uint64_t got = (src < w.nlanes && w.xseq[par][src] == t->xpar) ? w.xbuf[par][src] : bits;
    memcpy(&r, &got, sizeof(T));
    return r;
}
inline int vote_any(int pred) {
    Thread* t = cur_thread();
    Warp& w = t->blk->warps[t->warp];
    const unsigned par = (t->xpar++) & 1u;
    w.xbuf[par][t->lane] = pred ? 1 : 0;
    w.xseq[par][t->lane] = t->xpar;
    warp_barrier("__any_sync");
    int r = 0;
    for (int l = 0; l < w.nlanes; ++l)
        if (w.xseq[par][l] == t->xpar) r |= (int)w.xbuf[par][l];
    return r;
}
// D(8x8) += A(8x4, row) * B(4x8, col): lane l holds A[l>>2][l&3], B[l&3][l>>2], C[l>>2][2*(l&3) + {0,1}]
inline void dmma884(double& c0, double& c1, double a, double b) {
    Thread* t = cur_thread();
    Warp& w = t->blk->warps[t->warp];
    const unsigned par = (t->xpar++) & 1u;
    w.ma[par][t->lane] = a;
    w.mb[par][t->lane] = b;
    warp_barrier("mma.sync");
    const int fr = t->lane >> 2, fk = t->lane & 3;
    double acc0 = c0, acc1 = c1;
    for (int k = 0; k < 4; ++k) {
        acc0 = std::fma(w.ma[par][fr * 4 + k], w.mb[par][(2 * fk) * 4 + k], acc0);
        acc1 = std::fma(w.ma[par][fr * 4 + k], w.mb[par][(2 * fk + 1) * 4 + k], acc1);
    }
    c0 = acc0;
    c1 = acc1;
}

struct LaunchCfg {
    unsigned grid = 1, block = 32;
    size_t smem = 0;
    int order = 0;                 // 0 forward, 1 reverse, 2 random (seeded)
    size_t stack_bytes = 512 * 1024;
};

template <class Fn>
struct Tramp {
    static Fn*& fn() { static Fn* f = nullptr; return f; }
    static void entry() {
        (*fn())();
        Thread* t = cur_thread();
        t->done = true;
        t->blk->progress++;
        // a thread that has returned no longer takes part in the warp's collectives (kernels that retire lanes early)
        // (lanes that have passed their last collective return one after the other while their neighbours may not
        // even have read the result of that collective yet: the launch width stays what it was)
        Warp& w = t->blk->warps[t->warp];
        if (--w.nlive > 0 && w.bar_cnt == w.nlive) { w.bar_cnt = 0; w.bar_gen++; }
        swapcontext(&t->ctx, &t->blk->sched);
    }
};

// Runs `body` (a callable taking no arguments: the kernel call with its arguments bound) once per CUDA thread.
template <class Fn>
inline void launch(const LaunchCfg& cfg, Fn body) {
    Tramp<Fn>::fn() = &body;
    std::mt19937 rng(12345);
    for (unsigned bx = 0; bx < cfg.grid; ++bx) {
        Block blk;
        blk.bidx.x = bx;
        blk.bdim.x = cfg.block; blk.bdim.y = blk.bdim.z = 1;
        blk.gdim.x = cfg.grid; blk.gdim.y = blk.gdim.z = 1;
        memset(blk.bar_cnt, 0, sizeof blk.bar_cnt);
        memset(blk.bar_gen, 0, sizeof blk.bar_gen);
        std::vector<char> smem(cfg.smem + 64);
        // poison: uninitialised shared memory must not look like zeros
        for (size_t i = 0; i + 8 <= smem.size(); i += 8) { const double nan = std::nan(""); memcpy(&smem[i], &nan, 8); }
        blk.smem = smem.data() + (16 - ((uintptr_t)smem.data() & 15)) % 16;
        const int nwarp = (cfg.block + 31) / 32;
        blk.warps.resize(nwarp);
        for (int w = 0; w < nwarp; ++w) blk.warps[w].nlanes = blk.warps[w].nlive = (int)std::min<unsigned>(32, cfg.block - 32 * w);
        blk.th.resize(cfg.block);
        cur_block() = &blk;
        for (unsigned i = 0; i < cfg.block; ++i) {
            Thread& t = blk.th[i];
            t.tidx.x = i;
            t.lane = i & 31; t.warp = i >> 5;
            t.blk = &blk;
            t.stack = (char*)malloc(cfg.stack_bytes);
            getcontext(&t.ctx);
            t.ctx.uc_stack.ss_sp = t.stack;
            t.ctx.uc_stack.ss_size = cfg.stack_bytes;
            t.ctx.uc_link = &blk.sched;
            makecontext(&t.ctx, (void (*)())&Tramp<Fn>::entry, 0);
        }
        std::vector<int> order(cfg.block);
        for (unsigned i = 0; i < cfg.block; ++i) order[i] = cfg.order == 1 ? (int)(cfg.block - 1 - i) : (int)i;
        unsigned long last_progress = 0;
        int idle_passes = 0;
        for (;;) {
            if (cfg.order == 2) std::shuffle(order.begin(), order.end(), rng);
            bool alive = false;
            for (int i : order) {
                Thread& t = blk.th[i];
                if (t.done) continue;
                alive = true;
                cur_thread() = &t;
                swapcontext(&blk.sched, &t.ctx);
            }
            if (!alive) break;
            if (blk.progress == last_progress) {
                if (++idle_passes > 4) {
                    fprintf(stderr, "simt_emu: dead-lock in block %u (divergent collective or missing arrival):\n", bx);
                    for (unsigned i = 0; i < cfg.block; ++i)
                        if (!blk.th[i].done && (i % 32 == 0 || strcmp(blk.th[i].where, blk.th[i - 1].where)))
                            fprintf(stderr, "  thread %u (warp %d lane %d): %s\n", i, blk.th[i].warp, blk.th[i].lane, blk.th[i].where);
                    abort();
                }
            } else {
                idle_passes = 0;
                last_progress = blk.progress;
            }
        }
        for (auto& t : blk.th) free(t.stack);
        cur_block() = nullptr;
        cur_thread() = nullptr;
    }
}

}  // namespace simt_emu

// ------------------------------------------------------------------------------------------------ the CUDA surface
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static          /* CTAs run one after the other: one static instance is the CTA's */
#define __align__(n) alignas(n)

struct EmuIdxProxy {
    int which;
    struct X { int which; operator unsigned() const {
        simt_emu::Block* b = simt_emu::cur_block();
        switch (which) { case 0: return simt_emu::cur_thread()->tidx.x; case 1: return b->bidx.x; case 2: return b->bdim.x; default: return b->gdim.x; } } };
};
struct EmuIdx {
    struct C { int which, comp; operator unsigned() const {
        simt_emu::Block* b = simt_emu::cur_block();
        const simt_emu::Idx3& v = which == 0 ? simt_emu::cur_thread()->tidx : which == 1 ? b->bidx : which == 2 ? b->bdim : b->gdim;
        return comp == 0 ? v.x : (which >= 2 && comp > 0 ? 1u : 0u); } };
    C x, y, z;
    constexpr EmuIdx(int w) : x{w, 0}, y{w, 1}, z{w, 2} {}
};
static const EmuIdx threadIdx(0), blockIdx(1), blockDim(2), gridDim(3);

inline void __syncthreads() { simt_emu::block_bar(0, (int)simt_emu::cur_block()->bdim.x, true, "__syncthreads"); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt_emu::warp_barrier("__syncwarp"); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return simt_emu::exchange<T>(v, m, -1, "__shfl_xor_sync"); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return simt_emu::exchange<T>(v, 0, src, "__shfl_sync"); }
inline int __any_sync(unsigned, int p) { return simt_emu::vote_any(p); }
inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const auto o = *p; *p = o + v; return o; }
inline long long clock64() { return 0; }
inline void emu_bar_sync(int id, int count) { simt_emu::block_bar(id, count, true, "bar.sync"); }
inline void emu_bar_arrive(int id, int count) { simt_emu::block_bar(id, count, false, "bar.arrive"); }
inline void* emu_dyn_smem() { return simt_emu::cur_block()->smem; }
using std::fma;
using std::fmax;
using std::fmin;
using std::fabs;
using std::isfinite;
using std::sqrt;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct alignas(16) double2 { double x, y; };
