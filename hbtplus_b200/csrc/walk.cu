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
  const int warp = blockIdx.x * kWalkWarps + w;
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

// Variant without shared-memory staging: every step loads the (warp-uniform) node with two broadcast loads that hit
// L1 for runs of consecutive nodes.  One REDUX.OR per step carries both warp decisions: bit 0 = some lane opens the
// node, bit 1 = some lane accepted a spline-softened pair (then this node alone is accumulated with the exact kernel).
template <int T, bool PERIODIC, bool COUNT>
__global__ void __launch_bounds__(kWalkWarps * 32) walk_direct_kernel(const WalkArgs a, const DevConfig cfg)
{
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kWalkWarps + w;
  if (warp >= a.nwarps) return;
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
    skip[k] = valid[k] ? 0 : 0x7fffffff;
  }
  const int t0 = a.tree_off[lo], t1 = a.tree_off[lo + 1];
  const int node_begin = t0 + (t0 > 0 ? a.cellcount[t0 - 1] : 0);
  const int node_end = t1 + a.cellcount[t1 - 1];
  const float h = 2.8f * cfg.softening, h2 = h * h, hinv = 1.0f / h;
  int no = node_begin;
  double accd[T];
  float accf[T];
#pragma unroll
  for (int k = 0; k < T; k++) { accd[k] = 0.0; accf[k] = 0.f; }
  unsigned n_acc = 0, n_vis = 0, it = 0;

  while (no < node_end)
  {
    const float4 n = __ldg(&a.node_xm[no]);
    const float2 ax = __ldg(&a.node_aux[no]);
    const float lenq = ax.x;
    const int nend = __float_as_int(ax.y);
    float r2[T], rinv[T];
    bool acc[T];
    unsigned code = 0;
#pragma unroll
    for (int k = 0; k < T; k++)
    {
      float dx = n.x - px[k], dy = n.y - py[k], dz = n.z - pz[k];
      if (PERIODIC)
      {
        dx = nearest_f(dx, cfg.box_size, cfg.box_half);
        dy = nearest_f(dy, cfg.box_size, cfg.box_half);
        dz = nearest_f(dz, cfg.box_size, cfg.box_half);
      }
      r2[k] = dx * dx + dy * dy + dz * dz;
      const bool active = no >= skip[k];
      const bool open = active && (lenq > r2[k]);
      acc[k] = active && !(lenq > r2[k]);
      rinv[k] = rsqrt_raw(r2[k]);
      if (open) code |= 1u;
      if (acc[k] && r2[k] < h2) code |= 2u;
    }
    const unsigned red = __reduce_or_sync(kFull, code);
    if (red & 2u)
    { // exact kernel for this node (src/gravity_tree.cpp:141-161)
#pragma unroll
      for (int k = 0; k < T; k++)
      {
        float contrib = -n.w * rinv[k];
        if (r2[k] < h2)
        {
          float u = sqrtf(r2[k]) * hinv, wp;
          if (u < 0.5f)
            wp = -2.8f + u * u * (5.333333333333f + u * u * (6.4f * u - 9.6f));
          else
            wp = -3.2f + 0.066666666667f / u + u * u * (10.666666666667f + u * (-16.0f + u * (9.6f - 2.133333333333f * u)));
          contrib = n.w * hinv * wp;
        }
        if (acc[k]) accf[k] += contrib;
      }
    }
    else
    {
#pragma unroll
      for (int k = 0; k < T; k++)
        if (acc[k]) accf[k] = fmaf(-n.w, rinv[k], accf[k]);
    }
#pragma unroll
    for (int k = 0; k < T; k++)
      if (acc[k])
      {
        skip[k] = nend;
        if (COUNT) n_acc++;
      }
    if (COUNT) n_vis++;
    no = (red & 1u) ? no + 1 : nend;
    if (((++it) & 63u) == 0u)
    {
#pragma unroll
      for (int k = 0; k < T; k++) { accd[k] += (double)accf[k]; accf[k] = 0.f; }
    }
  }
#pragma unroll
  for (int k = 0; k < T; k++) accd[k] += (double)accf[k];
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

// 0 = shared-memory tiles (default), 1 = direct broadcast loads, 2 = direct for T=1 only (experiments: HBTU_WALK_DIRECT)
static int walk_direct()
{
  static int v = -1;
  if (v < 0)
  {
    const char *e = getenv("HBTU_WALK_DIRECT");
    v = e ? atoi(e) : 0; // default: shared-memory tiles; the direct variant lost in the bench workload (profiles/r01_walk_notes.md)
  }
  return v;
}

template <int T>
static void launch_t(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream)
{
  const int grid = div_up(a.nwarps, kWalkWarps);
  const bool count = a.counters != nullptr;
  const int pol = walk_direct();
  if (pol == 1 || (pol == 2 && T == 1))
  {
    if (cfg.periodic)
    {
      if (count) walk_direct_kernel<T, true, true><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
      else walk_direct_kernel<T, true, false><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
    }
    else
    {
      if (count) walk_direct_kernel<T, false, true><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
      else walk_direct_kernel<T, false, false><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
    }
    return;
  }
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

namespace
{
struct WalkPolicy
{
  int forced, big4, big2, group_min, group_t, masked;
  WalkPolicy()
  {
    auto env = [](const char *n, int d) { const char *e = getenv(n); return e ? atoi(e) : d; };
    forced = env("HBTU_WALK_TPL", 0);
    big4 = env("HBTU_WALK_BIG4", 1 << 20);
    big2 = env("HBTU_WALK_BIG2", 1 << 19);
    group_min = env("HBTU_WALK_GROUP_MIN", 1 << 13); // segments with at least this many targets use the group walk (0 = never)
    group_t = env("HBTU_WALK_GROUP_T", 4) == 8 ? 8 : 4;
    masked = env("HBTU_WALK_MASKED", 1); // 128-target groups: masked group walk (walk_masked.cu); 0 = walk_group.cu (measured: profiles/r01_walk_notes.md)
  }
};
const WalkPolicy &policy()
{
  static WalkPolicy p;
  return p;
}
} // namespace

void launch_walk(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls)
{
  if (a.nwarps <= 0) return;
  if (a.targets_per_lane == kWalkGroup4 || a.targets_per_lane == kWalkGroup8)
  {
    if (a.targets_per_lane == kWalkGroup4 && policy().masked) launch_walk_masked(a, cfg, stream, ls);
    else launch_walk_group(a, cfg, stream, ls);
    return;
  }
  if (a.targets_per_lane == 4) launch_t<4>(a, cfg, stream);
  else if (a.targets_per_lane == 2) launch_t<2>(a, cfg, stream);
  else launch_t<1>(a, cfg, stream);
  HBT_CHECK_LAUNCH();
  ls.launches++;
}


int walk_class_tpl(int index) { return index == 3 ? (policy().group_t == 8 ? kWalkGroup8 : kWalkGroup4) : (1 << index); }

WalkClass walk_class(int tgt_n)
{ // measured on B200 (profiles/): the group walk wins for large segments; below it T=1 per-lane walks win, where the
  // warp count, not the instruction count, limits throughput
  const WalkPolicy &p = policy();
  if (p.group_min > 0 && tgt_n >= p.group_min && !(p.forced == 1 || p.forced == 2 || p.forced == 4))
    return WalkClass{3, walk_class_tpl(3), 32 * p.group_t};
  int t;
  if (p.forced == 1 || p.forced == 2 || p.forced == 4) t = tgt_n >= 32 * p.forced ? p.forced : 1;
  else t = tgt_n >= p.big4 ? 4 : (tgt_n >= p.big2 ? 2 : 1);
  return WalkClass{t == 4 ? 2 : (t == 2 ? 1 : 0), t, 32 * t};
}

} // namespace hbt
