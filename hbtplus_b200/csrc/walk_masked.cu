// walk_masked.cu - the masked group walk kernel (sm_100a): GravityTree_t::EvaluatePotential / BindingEnergy
// (src/gravity_tree.cpp:79-175) for segments of at least HBTU_WALK_GROUP_MIN targets.  The algorithm is in
// walk_masked.cuh; this file is the kernel around it: segment lookup, target load (periodic: un-wrapped towards the
// warp's first target), the per-lane fallback when the chain stack runs out, interaction counters, energy epilogue.
// Roofline: FP32 issue.  Tensor cores are deliberately not used.
#include <atomic>
#include <cstdlib>

#include "walk_common.cuh"
#include "walk_masked.cuh"

namespace hbt
{

static constexpr int kMW = 4; // warps per CTA (warps are independent)
#ifndef HBT_MASKED_MINBLOCKS
#define HBT_MASKED_MINBLOCKS 5 // default resident CTAs per SM the register allocation allows (HBTU_WALK_MASKED_BLOCKS = 4, 5, 6 selects)
#endif

template <int STACK>
union MaskedWarpSmem
{
  MaskedSmemT<STACK> m;
  TileNode tile[32]; // per-lane fallback only (the group restarts from scratch, so the lists are dead by then)
};
// chain-stack entries per warp for a given number of resident CTAs per SM: what fits into 228 KB of shared memory
template <int MINB> struct MaskedStack { static constexpr int value = MINB >= 6 ? 104 : (MINB == 5 ? 184 : 216); };

template <bool PERIODIC, bool COUNT, int MINB>
__global__ void __launch_bounds__(kMW * 32, MINB) walk_masked_kernel(const WalkArgs a, const DevConfig cfg)
{
  constexpr int T = 4;
  __shared__ MaskedWarpSmem<MaskedStack<MINB>::value> s_all[kMW];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kMW + w;
  if (warp >= a.nwarps) return;
  MaskedWarpSmem<MaskedStack<MINB>::value> &sm = s_all[w];
  const int seg = segment_of_warp(a.warp_off, a.nseg, warp);
  const Segment sg = a.segs[seg];
  const int j0 = (warp - a.warp_off[seg]) * (32 * T) + lane;
  const int n0 = min(32 * T, sg.tgt_n - (warp - a.warp_off[seg]) * (32 * T)); // valid targets of the warp
  float px[T], py[T], pz[T];
  bool valid[T];
  {
    float rx = 0.f, ry = 0.f, rz = 0.f;
    if (PERIODIC)
    {
      const float4 r = a.tgt_pm[sg.tgt_off + j0 - lane];
      rx = r.x; ry = r.y; rz = r.z;
    }
#pragma unroll
    for (int k = 0; k < T; k++)
    {
      const int j = j0 + 32 * k;
      valid[k] = j < sg.tgt_n;
      const float4 tp = a.tgt_pm[sg.tgt_off + (valid[k] ? j : j0 - lane)];
      px[k] = tp.x; py[k] = tp.y; pz[k] = tp.z;
      if (PERIODIC)
      {
        const float ax = tp.x - rx, ay = tp.y - ry, az = tp.z - rz;
        if (ax > cfg.box_half) px[k] = tp.x - cfg.box_size; else if (ax < -cfg.box_half) px[k] = tp.x + cfg.box_size;
        if (ay > cfg.box_half) py[k] = tp.y - cfg.box_size; else if (ay < -cfg.box_half) py[k] = tp.y + cfg.box_size;
        if (az > cfg.box_half) pz[k] = tp.z - cfg.box_size; else if (az < -cfg.box_half) pz[k] = tp.z + cfg.box_size;
      }
    }
  }
  const int t0 = a.tree_off[seg], t1 = a.tree_off[seg + 1];
  const int node_begin = t0 + (t0 > 0 ? a.cellcount[t0 - 1] : 0);
  const int node_end = t1 > t0 ? t1 + a.cellcount[t1 - 1] : node_begin;

  double accd[T];
#pragma unroll
  for (int k = 0; k < T; k++) accd[k] = 0.0;
  unsigned long long nacc = 0;   // warp-uniform part of the interaction count (dense ring x valid targets)
  unsigned n_acc = 0, n_vis = 0; // per-lane part, warp node visits
  const bool ok = masked_group_walk<PERIODIC, COUNT>(sm.m, lane, a.node_xm, a.node_aux, node_begin, node_end, px, py, pz, valid, n0, cfg.box_size,
                                                     cfg.box_half, cfg.softening, accd, nacc, n_acc, n_vis);
  if (!ok)
  { // chain stack exhausted: redo this group with the per-lane walk from scratch
    __syncwarp();
    const float h = 2.8f * cfg.softening, h2 = h * h, hinv = 1.0f / h;
    int skip[T];
#pragma unroll
    for (int k = 0; k < T; k++) { accd[k] = 0.0; skip[k] = valid[k] ? node_begin : 0x7fffffff; }
    nacc = 0;
    n_acc = 0;
    walk_range<T, PERIODIC, COUNT>(a.node_xm, a.node_aux, sm.tile, node_begin, node_end, px, py, pz, skip, accd, cfg, h2, hinv, n_acc, n_vis);
  }
  if (COUNT)
  {
    unsigned long long tot = n_acc;
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
    if (lane == 0)
    {
      atomicAdd(&a.counters[0], tot + nacc);
      atomicAdd(&a.counters[1], (unsigned long long)n_vis);
      if (!ok) atomicAdd(&a.counters[2], 1ull);
    }
  }
  float rxp[T], ryp[T], rzp[T], pm[T];
#pragma unroll
  for (int k = 0; k < T; k++)
  { // raw positions + self mass for the energy epilogue
    const float4 tp = a.tgt_pm[sg.tgt_off + (valid[k] ? j0 + 32 * k : j0 - lane)];
    rxp[k] = tp.x; ryp[k] = tp.y; rzp[k] = tp.z; pm[k] = tp.w;
  }
  walk_epilogue<T>(a, cfg, sg, j0, valid, rxp, ryp, rzp, pm, accd);
}

template <int MINB>
static void launch_masked_b(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream)
{
  const int grid = div_up(a.nwarps, kMW);
  const bool count = a.counters != nullptr;
  // MINB CTAs of 32-48 KB static shared memory only fit with the largest shared-memory carve-out; the attribute is per device
  // (one host thread and context per device when a rank shards over several GPUs)
  static std::atomic<unsigned long long> carved{0ull};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(carved.load(std::memory_order_acquire) & bit))
  {
    for (auto *k : {walk_masked_kernel<true, true, MINB>, walk_masked_kernel<true, false, MINB>, walk_masked_kernel<false, true, MINB>,
                    walk_masked_kernel<false, false, MINB>})
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carved.fetch_or(bit, std::memory_order_release);
  }
  if (cfg.periodic)
  {
    if (count) walk_masked_kernel<true, true, MINB><<<grid, kMW * 32, 0, stream>>>(a, cfg);
    else walk_masked_kernel<true, false, MINB><<<grid, kMW * 32, 0, stream>>>(a, cfg);
  }
  else
  {
    if (count) walk_masked_kernel<false, true, MINB><<<grid, kMW * 32, 0, stream>>>(a, cfg);
    else walk_masked_kernel<false, false, MINB><<<grid, kMW * 32, 0, stream>>>(a, cfg);
  }
}

void launch_walk_masked(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls)
{
  if (a.nwarps <= 0) return;
  static const int blocks = [] { const char *e = getenv("HBTU_WALK_MASKED_BLOCKS"); return e ? atoi(e) : HBT_MASKED_MINBLOCKS; }();
  if (blocks == 4) launch_masked_b<4>(a, cfg, stream);
  else if (blocks == 6) launch_masked_b<6>(a, cfg, stream);
  else launch_masked_b<5>(a, cfg, stream);
  HBT_CHECK_LAUNCH();
  ls.launches++;
}

} // namespace hbt
