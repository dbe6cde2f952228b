// walk.cu - warp-cooperative Barnes-Hut potential walk with the reference's per-target opening
// criterion (sm_100a).  Replaces GravityTree_t::EvaluatePotential / BindingEnergy
// (src/gravity_tree.cpp:79-175) and the two OpenMP target loops of Subhalo_t::Unbind
// (src/subhalo_unbind.cpp:319-328 and :341-354).
//
// A warp owns 32 consecutive targets of ONE subhalo (targets are in octal-key order, so they are
// spatial neighbours) and scans that subhalo's pre-order node array front to back:
//
//   * the current node index `no` is warp-uniform; 32 nodes at a time are staged in shared memory by
//     one coalesced 16 B + 8 B load per lane and then read back as broadcasts (no bank conflicts);
//   * every lane applies the REFERENCE criterion to ITS OWN target: open iff len^2 > r^2 theta^2
//     (src/gravity_tree.cpp:135).  A lane that accepts a cell adds its monopole and sets its private
//     resume index to the cell's `end`; it then idles while other lanes descend into that cell.  The warp
//     advances to no+1 if any lane opened the node, else jumps to `end` (the reference's `sibling`).
//     Decisions are therefore identical to the reference's scalar walk (up to fp32 vs fp64 rounding of
//     r^2), which a group-level criterion would not give; tests/test_tree_core.py checks that the
//     per-target accepted-interaction counts equal the reference's.
//   * pair arithmetic is fp32 (3 FADD + FMUL + 2 FFMA + MUFU.RSQ + FFMA per interaction); partial sums are
//     flushed into an fp64 accumulator every 64 nodes; the spline-softened branch (r < 2.8 eps,
//     src/gravity_tree.cpp:146-160) is taken only when some lane needs it (warp vote).
//
// Roofline: FP32 issue.  Tensor cores are deliberately not used.
#include <cstdlib>

#include "walk_common.cuh"

namespace hbt
{

static constexpr int kWalkWarps = 4;

// T targets per lane: a warp owns 32*T consecutive targets (lane l holds targets l, l+32, ...).  T=1 for small
// subhaloes; T=4 for large ones, where it quarters the dependent tile loads and the control instructions per
// interaction at the price of a ~1.3x larger node union.
template <int T, bool PERIODIC, bool COUNT>
__global__ void __launch_bounds__(kWalkWarps * 32) walk_kernel(const WalkArgs a, const DevConfig cfg)
{
  __shared__ TileNode s_tile[kWalkWarps][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = walk_cta(a) * kWalkWarps + w;
  if (warp >= a.nwarps) return;
  // segment of this warp: largest s with warp_off[s] <= warp (segments of the other class have no warps)
  int lo = 0, hi = a.nseg;
  while (hi - lo > 1)
  {
    int mid = (lo + hi) >> 1;
    if (a.warp_off[mid] <= warp) lo = mid; else hi = mid;
  }
  const Segment sg = a.segs[lo];
  const int j0 = (warp - a.warp_off[lo]) * (32 * T) + lane;
  float px[T], py[T], pz[T], pm[T];
  int skip[T];
  bool valid[T];
#pragma unroll
  for (int k = 0; k < T; k++)
  {
    const int j = j0 + 32 * k;
    valid[k] = j < sg.tgt_n;
    const float4 tp = a.tgt_pm[sg.tgt_off + (valid[k] ? j : 0)];
    px[k] = tp.x; py[k] = tp.y; pz[k] = tp.z; pm[k] = tp.w;
    skip[k] = valid[k] ? 0 : 0x7fffffff; // resume index of this target
  }
  const int t0 = a.tree_off[lo], t1 = a.tree_off[lo + 1];
  const int node_begin = t0 + (t0 > 0 ? a.cellcount[t0 - 1] : 0);
  const int node_end = t1 + a.cellcount[t1 - 1];

  const float h = 2.8f * cfg.softening, h2 = h * h, hinv = 1.0f / h;
  int no = node_begin;
  double accd[T];
#pragma unroll
  for (int k = 0; k < T; k++) accd[k] = 0.0;
  unsigned n_acc = 0, n_vis = 0;
  TileNode *const tile = s_tile[w];

  walk_range<T, PERIODIC, COUNT>(a.node_xm, a.node_aux, tile, no, node_end, px, py, pz, skip, accd, cfg, h2, hinv, n_acc, n_vis);
  if (COUNT)
  {
    unsigned long long na = n_acc;
    for (int o = 16; o > 0; o >>= 1) na += __shfl_xor_sync(kFull, na, o);
    if (lane == 0)
    {
      atomicAdd(&a.counters[0], na);
      atomicAdd(&a.counters[1], (unsigned long long)n_vis);
    }
  }
  walk_epilogue<T>(a, cfg, sg, j0, valid, px, py, pz, pm, accd);
}

template <int T>
static void launch_t(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream)
{
  const int grid = walk_grid(a, div_up(a.nwarps, kWalkWarps));
  if (grid <= 0) return; // target split: none of the 16-CTA chunks of this launch is this context's
  const bool count = a.counters != nullptr;
  if (cfg.periodic)
  {
    if (count) walk_kernel<T, true, true><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
    else walk_kernel<T, true, false><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
  }
  else
  {
    if (count) walk_kernel<T, false, true><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
    else walk_kernel<T, false, false><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
  }
}

// Kernel routing.  Defaults are the measured best (profiles/); the HBTU_* environment variables and hbtu_set_tuning (a
// diagnostics entry point: tests and the A/B harness tools/ab_walk.py) override them process-wide.
WalkTuning &walk_tuning()
{
  static WalkTuning t = [] {
    auto env = [](const char *n, int d) { const char *e = getenv(n); return e ? atoi(e) : d; };
    WalkTuning v;
    v.forced_tpl = env("HBTU_WALK_TPL", 0);
    v.big4 = env("HBTU_WALK_BIG4", 1 << 20);
    v.big2 = env("HBTU_WALK_BIG2", 1 << 19);
    v.group_min = env("HBTU_WALK_GROUP_MIN", 256); // segments with at least this many targets use the masked group walk (0 = never).
    // Swept on B200 with the 64-target kernel (profiles/r02_ab_group_min.jsonl): 8192 -> 256 takes the walk of cfg 3 from 36.4 to
    // 31.4 ms, of a cfg-4 shard from 733 to 689 ms, of cfg 2 from 918 to 909 ms; flat between 512 and 64
    v.masked_pairs = env("HBTU_WALK_MASKED_PAIRS", HBT_MASKED_DEFAULT_PAIRS) == 1 ? 1 : 2;
    v.masked_blocks = env("HBTU_WALK_MASKED_BLOCKS", v.masked_pairs == 1 ? HBT_MASKED_DEFAULT_BLOCKS_NP1 : HBT_MASKED_DEFAULT_BLOCKS_NP2);
    v.small_max = env("HBTU_WALK_SMALL_MAX", HBT_SMALL_DEFAULT_MAX);
    return v;
  }();
  return t;
}

void launch_walk(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls)
{
  if (a.nwarps <= 0) return;
  if (a.targets_per_lane == kWalkGroup4 || a.targets_per_lane == kWalkGroup2)
  {
    launch_walk_masked(a, cfg, stream, ls);
    return;
  }
  if (a.targets_per_lane == kWalkSmall)
  {
    launch_walk_small(a, cfg, stream, ls);
    return;
  }
  if (a.targets_per_lane == 4) launch_t<4>(a, cfg, stream);
  else if (a.targets_per_lane == 2) launch_t<2>(a, cfg, stream);
  else launch_t<1>(a, cfg, stream);
  HBT_CHECK_LAUNCH();
  ls.launches++;
}

int walk_class_tpl(int index)
{
  if (index == 4) return kWalkSmall;
  return index == 3 ? (walk_tuning().masked_pairs == 1 ? kWalkGroup2 : kWalkGroup4) : (1 << index);
}

WalkClass walk_class(int tgt_n, int tree_n)
{ // measured on B200 (profiles/): the masked group walk wins for large segments; below it T=1 per-lane walks win, where the
  // warp count, not the instruction count, limits throughput; trees of a few hundred sources are swept densely (walk_small.cu)
  const WalkTuning &p = walk_tuning();
  const bool forced = p.forced_tpl == 1 || p.forced_tpl == 2 || p.forced_tpl == 4;
  if (p.small_max > 0 && tree_n <= p.small_max && !forced) return WalkClass{4, kWalkSmall, 32};
  if (p.group_min > 0 && tgt_n >= p.group_min && !forced) return WalkClass{3, walk_class_tpl(3), p.masked_pairs == 1 ? 64 : 128};
  int t;
  if (forced) t = tgt_n >= 32 * p.forced_tpl ? p.forced_tpl : 1;
  else t = tgt_n >= p.big4 ? 4 : (tgt_n >= p.big2 ? 2 : 1);
  return WalkClass{t == 4 ? 2 : (t == 2 ? 1 : 0), t, 32 * t};
}

} // namespace hbt
