// walk_masked_kernel.cuh - the kernel around walk_masked.cuh: segment lookup, target load (periodic: un-wrapped towards the
// warp's first target), the per-lane fallback when the chain stack runs out, interaction counters, energy epilogue.
// Included by the translation units that instantiate the (resident CTAs per SM, slice pairs per warp) variants.
#pragma once
#include <atomic>
#include <cstdlib>

#include "walk_common.cuh"
#include "walk_masked.cuh"

namespace hbt
{

static constexpr int kMW = 4; // warps per CTA (warps are independent)

#ifndef HBT_MASKED_SMEM_TOTAL
#define HBT_MASKED_SMEM_TOTAL 233472 // shared memory of an SM the resident CTAs may fill (228 KB = the largest carve-out) ...
#endif
#ifndef HBT_MASKED_CARVEOUT
#define HBT_MASKED_CARVEOUT cudaSharedmemCarveoutMaxShared // ... and the matching carve-out preference (percent of the maximum)
#endif

// chain-stack entries per warp: what fits into the 228 KB of shared memory of an SM with MINB resident CTAs of kMW warps
// (1 KB per CTA is reserved by the system; static shared memory is limited to 48 KB per CTA)
template <int MINB, int NP>
struct MaskedStack
{
  static constexpr int kPerWarp = ((HBT_MASKED_SMEM_TOTAL / MINB - 1024) / kMW) & ~15;
  static constexpr int kFixed = (int)sizeof(MaskedSmemT<8, NP>) - 8 * (int)sizeof(ChainEntryT<NP>);
  static constexpr int kBySm = (kPerWarp - kFixed) / (int)sizeof(ChainEntryT<NP>);
  static constexpr int kBy48K = ((48 * 1024 - 64) / kMW - kFixed) / (int)sizeof(ChainEntryT<NP>);
  static constexpr int value = ((kBySm < kBy48K ? kBySm : kBy48K) / 8) * 8;
  static constexpr bool ok = value >= 96; // a shallower stack would send real trees to the per-lane fallback
};

template <int STACK, int NP>
union MaskedWarpSmem
{
  MaskedSmemT<STACK, NP> m;
  TileNode tile[32]; // per-lane fallback only (the group restarts from scratch, so the lists are dead by then)
};

template <bool PERIODIC, bool COUNT, int MINB, int NP>
__global__ void __launch_bounds__(kMW * 32, MINB) walk_masked_kernel(const WalkArgs a, const DevConfig cfg)
{
  constexpr int T = 2 * NP;
  static_assert(MaskedStack<MINB, NP>::ok, "too little shared memory left for the chain stack");
  typedef MaskedWarpSmem<MaskedStack<MINB, NP>::value, NP> Smem;
  __shared__ Smem s_all[kMW];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = walk_cta(a) * kMW + w;
  if (warp >= a.nwarps) return;
  Smem &sm = s_all[w];
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

template <int MINB, int NP>
void launch_masked_variant(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream)
{
  if constexpr (!MaskedStack<MINB, NP>::ok)
  { // the element lists of this build leave too little shared memory for MINB resident CTAs: use one CTA fewer
    launch_masked_variant<MINB - 1, NP>(a, cfg, stream);
    return;
  }
  else
  {
  const int grid = walk_grid(a, div_up(a.nwarps, kMW));
  if (grid <= 0) return; // target split: none of the 16-CTA chunks of this launch is this context's
  const bool count = a.counters != nullptr;
  // MINB CTAs of up to 48 KB static shared memory only fit with the largest shared-memory carve-out; the attribute is per device
  // (one host thread and context per device when a rank shards over several GPUs)
  static std::atomic<unsigned long long> carved{0ull};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(carved.load(std::memory_order_acquire) & bit))
  {
    for (auto *k : {walk_masked_kernel<true, true, MINB, NP>, walk_masked_kernel<true, false, MINB, NP>, walk_masked_kernel<false, true, MINB, NP>,
                    walk_masked_kernel<false, false, MINB, NP>})
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, HBT_MASKED_CARVEOUT);
    carved.fetch_or(bit, std::memory_order_release);
  }
  if (cfg.periodic)
  {
    if (count) walk_masked_kernel<true, true, MINB, NP><<<grid, kMW * 32, 0, stream>>>(a, cfg);
    else walk_masked_kernel<true, false, MINB, NP><<<grid, kMW * 32, 0, stream>>>(a, cfg);
  }
  else
  {
    if (count) walk_masked_kernel<false, true, MINB, NP><<<grid, kMW * 32, 0, stream>>>(a, cfg);
    else walk_masked_kernel<false, false, MINB, NP><<<grid, kMW * 32, 0, stream>>>(a, cfg);
  }
  }
}

} // namespace hbt
