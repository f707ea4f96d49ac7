// simt.h — a small single-threaded SIMT emulator for CPU-side LOGIC TESTS of the warp-cooperative CUDA kernels
// (test infrastructure, not a product path: nothing under zra_b200/ includes it in a product build).
//
// Every CUDA thread of a block is a fiber (ucontext). Fibers run until they reach a collective operation
// (__syncthreads, __syncwarp, __ballot_sync, __shfl_*_sync, __reduce_*_sync, ...), where they yield until every
// participant has arrived — so warp-synchronous code runs with exactly the data flow it has on the GPU, and a
// collective that not every named lane reaches shows up as a reported deadlock instead of a hang. Blocks run one
// after another. The kernel source is compiled unchanged by g++ through the macros at the end of this file.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

namespace simt {

struct Dim3 {
  unsigned x{1}, y{1}, z{1};
  Dim3() {}
  Dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

struct WarpSync {
  uint64_t vals[32], snap[32];
  uint32_t arrived{0}, leaving{0}, alive{0};
  bool releasing{false};
};

struct Block {
  std::vector<ucontext_t> ctx;
  std::vector<std::vector<char>> stacks;
  std::vector<char> done;
  std::vector<WarpSync> warps;
  ucontext_t sched;
  unsigned cur{0}, nThreads{0};
  // __syncthreads
  unsigned barArrived{0}, barGen{0};
  // named barriers (bar.sync / bar.arrive id, count)
  unsigned nbArrived[16] = {}, nbGen[16] = {};
  unsigned long progress{0};
  std::vector<unsigned char> dynSmem;
  const std::function<void()>* body{nullptr};
};

inline Block*& cur_block() { static Block* b = nullptr; return b; }
inline Dim3& v_threadIdx() { static Dim3 d; return d; }
inline Dim3& v_blockIdx() { static Dim3 d; return d; }
inline Dim3& v_blockDim() { static Dim3 d; return d; }
inline Dim3& v_gridDim() { static Dim3 d; return d; }

inline void yield() {
  Block* b = cur_block();
  swapcontext(&b->ctx[b->cur], &b->sched);
}

inline void fiber_entry() {
  Block* b = cur_block();
  (*b->body)();
  b->done[b->cur] = 1;
  unsigned w = b->cur >> 5, l = b->cur & 31;
  b->warps[w].alive &= ~(1u << l);
  b->progress++;
  swapcontext(&b->ctx[b->cur], &b->sched);
}

// Runs `body` once per thread of a grid x block launch, blocks one after another.
inline void launch(Dim3 grid, Dim3 block, size_t dynSmemBytes, const std::function<void()>& body, size_t stackBytes = 256 << 10) {
  Block blk;
  const unsigned n = block.x * block.y * block.z;
  blk.nThreads = n;
  blk.ctx.resize(n);
  blk.stacks.resize(n);
  blk.done.assign(n, 0);
  blk.body = &body;
  for (auto& s : blk.stacks) s.resize(stackBytes);
  v_blockDim() = block;
  v_gridDim() = grid;
  Block* saved = cur_block();
  cur_block() = &blk;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        blk.warps.assign((n + 31) / 32, WarpSync());
        blk.dynSmem.assign(dynSmemBytes + 64, 0xCD);
        blk.barArrived = 0;
        blk.barGen = 0;
        for (int q = 0; q < 16; q++) { blk.nbArrived[q] = 0; blk.nbGen[q] = 0; }
        for (unsigned t = 0; t < n; t++) {
          blk.done[t] = 0;
          blk.warps[t >> 5].alive |= 1u << (t & 31);
          getcontext(&blk.ctx[t]);
          blk.ctx[t].uc_stack.ss_sp = blk.stacks[t].data();
          blk.ctx[t].uc_stack.ss_size = blk.stacks[t].size();
          blk.ctx[t].uc_link = &blk.sched;
          makecontext(&blk.ctx[t], (void (*)())fiber_entry, 0);
        }
        unsigned live = n;
        while (live) {
          unsigned long before = blk.progress;
          live = 0;
          for (unsigned t = 0; t < n; t++) {
            if (blk.done[t]) continue;
            blk.cur = t;
            v_threadIdx() = Dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            v_blockIdx() = Dim3(bx, by, bz);
            swapcontext(&blk.sched, &blk.ctx[t]);
            if (!blk.done[t]) live++;
          }
          if (live && blk.progress == before) {
            fprintf(stderr, "simt: deadlock in block (%u,%u,%u): %u threads wait at a collective not everyone reaches\n", bx, by, bz, live);
            abort();
          }
        }
      }
  cur_block() = saved;
}

inline unsigned lane_id() { return cur_block()->cur & 31; }

// One warp collective: every lane named in `mask` contributes `val`; returns the snapshot of all contributions.
inline const uint64_t* collective(uint32_t mask, uint64_t val) {
  Block* b = cur_block();
  WarpSync& w = b->warps[b->cur >> 5];
  const uint32_t bit = 1u << (b->cur & 31);
  if (!(mask & bit)) { fprintf(stderr, "simt: lane not in its own collective mask\n"); abort(); }
  while (w.releasing) yield();
  w.vals[b->cur & 31] = val;
  w.arrived |= bit;
  b->progress++;
  if (w.arrived == mask) {
    memcpy(w.snap, w.vals, sizeof(w.snap));
    w.releasing = true;
    w.leaving = mask;
    w.arrived = 0;
  } else {
    while (!(w.releasing && (w.leaving & bit))) yield();
  }
  return w.snap;
}
inline void collective_done() {
  Block* b = cur_block();
  WarpSync& w = b->warps[b->cur >> 5];
  w.leaving &= ~(1u << (b->cur & 31));
  if (!w.leaving) w.releasing = false;
  b->progress++;
}

inline void syncthreads() {
  Block* b = cur_block();
  unsigned gen = b->barGen;
  unsigned need = 0;
  for (unsigned t = 0; t < b->nThreads; t++) need += !b->done[t];
  b->barArrived++;
  b->progress++;
  if (b->barArrived >= need) { b->barArrived = 0; b->barGen++; return; }
  while (b->barGen == gen) {
    yield();
    // threads may have exited since: re-check
    if (b->barGen == gen) {
      unsigned n2 = 0;
      for (unsigned t = 0; t < b->nThreads; t++) n2 += !b->done[t];
      if (b->barArrived >= n2) { b->barArrived = 0; b->barGen++; b->progress++; }
    }
  }
}

// bar.sync id, count (wait = true) / bar.arrive id, count (wait = false)
inline void named_barrier(unsigned id, unsigned count, bool wait) {
  Block* b = cur_block();
  const unsigned gen = b->nbGen[id];
  b->progress++;
  if (++b->nbArrived[id] >= count) { b->nbArrived[id] = 0; b->nbGen[id]++; return; }
  if (wait) while (b->nbGen[id] == gen) yield();
}

inline unsigned char* dyn_smem() {
  Block* b = cur_block();
  uintptr_t p = reinterpret_cast<uintptr_t>(b->dynSmem.data());
  return reinterpret_cast<unsigned char*>((p + 15) & ~(uintptr_t)15);
}

}  // namespace simt

// ---------------------------------------------------------------- CUDA spellings
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static
#define threadIdx (simt::v_threadIdx())
#define blockIdx (simt::v_blockIdx())
#define blockDim (simt::v_blockDim())
#define gridDim (simt::v_gridDim())

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }

inline void __syncthreads() { simt::syncthreads(); }
inline void __syncwarp(uint32_t mask = 0xFFFFFFFFu) { simt::collective(mask, 0); simt::collective_done(); }
inline uint32_t __ballot_sync(uint32_t mask, int pred) {
  const uint64_t* s = simt::collective(mask, pred ? 1 : 0);
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) if ((mask >> i & 1) && s[i]) r |= 1u << i;
  simt::collective_done();
  return r;
}
inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <class T> inline T __shfl_sync(uint32_t mask, T v, int src, int width = 32) {
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  const uint64_t* s = simt::collective(mask, raw);
  int lane = (int)simt::lane_id();
  int from = (lane & ~(width - 1)) | (src & (width - 1));
  uint64_t got = s[from];
  simt::collective_done();
  T r; memcpy(&r, &got, sizeof(T));
  return r;
}
template <class T> inline T __shfl_up_sync(uint32_t mask, T v, unsigned delta, int width = 32) {
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  const uint64_t* s = simt::collective(mask, raw);
  int lane = (int)simt::lane_id();
  int from = lane - (int)delta;
  if (from < (lane & ~(width - 1))) from = lane;
  uint64_t got = s[from];
  simt::collective_done();
  T r; memcpy(&r, &got, sizeof(T));
  return r;
}
template <class T> inline T __shfl_down_sync(uint32_t mask, T v, unsigned delta, int width = 32) {
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  const uint64_t* s = simt::collective(mask, raw);
  int lane = (int)simt::lane_id();
  int from = lane + (int)delta;
  if (from > (lane | (width - 1))) from = lane;
  uint64_t got = s[from];
  simt::collective_done();
  T r; memcpy(&r, &got, sizeof(T));
  return r;
}
template <class T> inline T __shfl_xor_sync(uint32_t mask, T v, int x, int width = 32) {
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  const uint64_t* s = simt::collective(mask, raw);
  int from = (int)simt::lane_id() ^ x;
  uint64_t got = s[from];
  simt::collective_done();
  T r; memcpy(&r, &got, sizeof(T));
  return r;
}
#define SIMT_REDUCE(name, init, expr)                                      \
  inline uint32_t name(uint32_t mask, uint32_t v) {                        \
    const uint64_t* s = simt::collective(mask, v);                         \
    uint32_t r = init;                                                     \
    for (int i = 0; i < 32; i++) if (mask >> i & 1) { uint32_t x = (uint32_t)s[i]; r = (expr); } \
    simt::collective_done();                                               \
    return r;                                                              \
  }
SIMT_REDUCE(__reduce_min_sync, 0xFFFFFFFFu, x < r ? x : r)
SIMT_REDUCE(__reduce_max_sync, 0u, x > r ? x : r)
SIMT_REDUCE(__reduce_or_sync, 0u, r | x)
SIMT_REDUCE(__reduce_and_sync, 0xFFFFFFFFu, r & x)
SIMT_REDUCE(__reduce_add_sync, 0u, r + x)

inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((uint32_t)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline uint32_t __brev(uint32_t v) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t n) { n &= 31; return n ? (hi << n) | (lo >> (32 - n)) : hi; }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t n) { n &= 31; return n ? (lo >> n) | (hi << (32 - n)) : lo; }
inline uint32_t __funnelshift_lc(uint32_t lo, uint32_t hi, uint32_t n) { return n >= 32 ? lo : (n ? (hi << n) | (lo >> (32 - n)) : hi); }
inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t n) { return n >= 32 ? hi : (n ? (lo >> n) | (hi << (32 - n)) : lo); }
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
  uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
  return r;
}
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
