// warp_emul.h - TEST-ONLY shim that lets g++ compile warp-synchronous device code (hbtplus_b200/csrc/walk_masked.cuh)
// and run it as 32 cooperative fibers (ucontext), one per lane.  Warp collectives are rendez-vous points that also check
// that every lane arrived at the same kind of collective.  Never linked into the product library.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __device__
#define __forceinline__ inline
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(8))) int2 { int x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
using std::max;
using std::min;

namespace wemu
{
static constexpr int kLanes = 32;
static ucontext_t g_main, g_ctx[kLanes];
static bool g_done[kLanes];
static int g_cur = 0;
static uint64_t g_slot[kLanes];
static int g_op[kLanes];
static int g_arrived = 0;
static uint64_t g_gen = 0;
static uint64_t g_ncollectives = 0;
static uint64_t g_shuffle = 0; // 0: lanes run in lane order; otherwise seed/state of the random lane order
static std::function<void(int)> g_body;

static inline void yield() { swapcontext(&g_ctx[g_cur], &g_main); }
static inline void barrier()
{
  const uint64_t gen = g_gen;
  if (++g_arrived == kLanes) { g_arrived = 0; g_gen++; }
  else while (g_gen == gen) yield();
}
// every lane deposits (op, value); returns after all lanes did; values stay readable until the closing barrier
static inline void exchange(int op, uint64_t v)
{
  g_slot[g_cur] = v;
  g_op[g_cur] = op;
  barrier();
  for (int l = 0; l < kLanes; l++)
    if (g_op[l] != op) { fprintf(stderr, "warp_emul: lanes diverged at a collective (lane %d op %d, lane %d op %d)\n", g_cur, op, l, g_op[l]); abort(); }
  g_ncollectives++;
}
static void trampoline()
{
  g_body(g_cur);
  g_done[g_cur] = true;
  swapcontext(&g_ctx[g_cur], &g_main);
}
// run body(lane) for 32 lanes to completion
static inline void run_warp(const std::function<void(int)> &body)
{
  static std::vector<char> stacks;
  const size_t kStack = 256 << 10;
  if (stacks.empty()) stacks.resize(kStack * kLanes);
  g_body = body;
  g_arrived = 0;
  for (int l = 0; l < kLanes; l++)
  {
    g_done[l] = false;
    getcontext(&g_ctx[l]);
    g_ctx[l].uc_stack.ss_sp = stacks.data() + kStack * l;
    g_ctx[l].uc_stack.ss_size = kStack;
    g_ctx[l].uc_link = &g_main;
    makecontext(&g_ctx[l], trampoline, 0);
  }
  // Between two rendez-vous the lanes run one after the other.  With g_shuffle != 0 the order is a fresh pseudo-random
  // permutation on every pass, so that code whose result depends on the order in which lanes execute between two
  // synchronisation points (a missing __syncwarp around a shared-memory hand-over) gives different answers.
  bool any = true;
  int order[kLanes];
  for (int l = 0; l < kLanes; l++) order[l] = l;
  while (any)
  {
    any = false;
    if (g_shuffle)
      for (int l = kLanes - 1; l > 0; l--)
      {
        g_shuffle = g_shuffle * 6364136223846793005ull + 1442695040888963407ull;
        std::swap(order[l], order[(g_shuffle >> 33) % (unsigned)(l + 1)]);
      }
    for (int q = 0; q < kLanes; q++)
    {
      const int l = order[q];
      if (!g_done[l]) { any = true; g_cur = l; swapcontext(&g_main, &g_ctx[l]); }
    }
  }
  if (g_arrived != 0) { fprintf(stderr, "warp_emul: %d lanes left waiting at a collective\n", g_arrived); abort(); }
}
} // namespace wemu

static constexpr unsigned kFull = 0xffffffffu;
static inline unsigned __ballot_sync(unsigned, bool p)
{
  wemu::exchange(1, p ? 1 : 0);
  unsigned r = 0;
  for (int l = 0; l < 32; l++) r |= (unsigned)(wemu::g_slot[l] & 1) << l;
  wemu::barrier();
  return r;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
static inline unsigned __reduce_min_sync(unsigned, unsigned v)
{
  wemu::exchange(2, v);
  unsigned r = 0xffffffffu;
  for (int l = 0; l < 32; l++) r = std::min(r, (unsigned)wemu::g_slot[l]);
  wemu::barrier();
  return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
  wemu::exchange(3, v);
  unsigned r = 0;
  for (int l = 0; l < 32; l++) r = std::max(r, (unsigned)wemu::g_slot[l]);
  wemu::barrier();
  return r;
}
static inline unsigned __shfl_sync(unsigned, unsigned v, int src)
{
  wemu::exchange(5, v);
  const unsigned r = (unsigned)wemu::g_slot[src & 31];
  wemu::barrier();
  return r;
}
static inline void __syncwarp()
{
  wemu::exchange(4, 0);
  wemu::barrier();
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float rsqrt_raw(float x) { return 1.0f / std::sqrt(x); }
static inline float2 f2_add(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 f2_mul(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 f2_fma(float2 a, float2 b, float2 c) { return float2{std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)}; }
template <bool COUNT>
static inline unsigned decide_half(unsigned mask, unsigned lanebit, float lenq, float r2, float w, float rinv, float &acc, unsigned &n_acc)
{ // host twin of walk_common.cuh::decide_half
  const bool in = (mask & lanebit) != 0u, open = lenq > r2;
  if (in && !open)
  {
    acc = std::fma(w, rinv, acc);
    if (COUNT) n_acc++;
  }
  return __ballot_sync(kFull, in && open);
}
template <bool COUNT>
static inline unsigned decide_half_rsq(unsigned mask, unsigned lanebit, float lenq, float r2, float w, float &acc, unsigned &n_acc)
{ // host twin of walk_common.cuh::decide_half_rsq
  return decide_half<COUNT>(mask, lanebit, lenq, r2, w, rsqrt_raw(r2), acc, n_acc);
}
template <bool COUNT>
static inline void accept_half_rsq(unsigned mask, unsigned lanebit, float r2, float w, float &acc, unsigned &n_acc)
{ // host twin of walk_common.cuh::accept_half_rsq
  if ((mask & lanebit) != 0u)
  {
    acc = std::fma(w, rsqrt_raw(r2), acc);
    if (COUNT) n_acc++;
  }
}
namespace hbt
{
static inline float nearest_f(float x, float box, float half) { return x > half ? x - box : (x < -half ? x + box : x); }
} // namespace hbt
