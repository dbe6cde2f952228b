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

#include "device_tree.cuh"

namespace hbt
{

static constexpr int kWalkWarps = 4;
static constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void relative_velocity(const float tp[3], const float tv[3], const float rp[3], const float rv[3],
                                                  const DevConfig &cfg, float dv[3])
{ // Snapshot_t::RelativeVelocity, src/snapshot.h:100-111, in HBTReal=float with no FMA contraction
#pragma unroll
  for (int j = 0; j < 3; j++)
  {
    float dx = __fsub_rn(tp[j], rp[j]);
    if (cfg.periodic) dx = nearest_f(dx, cfg.box_size, cfg.box_half);
    float d = __fsub_rn(tv[j], rv[j]);
    dv[j] = __fadd_rn(d, __fmul_rn(__fmul_rn(cfg.hz, cfg.scale_factor), dx));
  }
}
__device__ __forceinline__ float dot3_rn(const float a[3], const float b[3])
{ // VecDot macro, src/mymath.h:20, float arithmetic left to right
  return __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
}

struct __align__(16) TileNode
{ // one staged node: 32 B so that a warp's tile is 1 KB and both loads are broadcasts
  float4 xm;  // x, y, z, mass
  float lenq; // len^2/theta^2 (0 for particles)
  int end;    // index of the first node after this node's subtree
  int pad0, pad1;
};

__device__ __forceinline__ float rsqrt_raw(float x)
{ // single MUFU.RSQ; r2 == 0 (self / co-located pair) gives +inf and is replaced by the spline branch
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One pass over the staged tile [tile_base, tile_lim) starting at node `no`, for the T targets of every lane.
//   CAREFUL=false: every accepted source is added as -m/r and the smallest accepted r^2 is tracked; the caller
//                  redoes the tile with CAREFUL=true if any lane met r < 2.8 eps (spline-softened pair, or r = 0).
//   CAREFUL=true : the reference's full kernel (src/gravity_tree.cpp:141-161), branch taken on a warp vote.
// Returns the node index at which the warp left the tile.
template <int T, bool PERIODIC, bool COUNT, bool CAREFUL>
__device__ __forceinline__ int walk_tile(const TileNode *__restrict__ tile, int tile_base, int tile_lim, int no, const float (&px)[T],
                                         const float (&py)[T], const float (&pz)[T], int (&skip)[T], float (&accf)[T], float &minr2,
                                         double (&accs)[T], const DevConfig &cfg, float h2, float hinv, unsigned &n_acc, unsigned &n_vis)
{
  do
  {
    const TileNode *nd = &tile[no - tile_base];
    const float4 n = nd->xm;
    const float lenq = nd->lenq;
    const int nend = nd->end;
    bool any_open = false;
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
      const float r2 = dx * dx + dy * dy + dz * dz;
      const bool active = no >= skip[k];
      const bool open = active && (lenq > r2); // reference criterion, per target (src/gravity_tree.cpp:135)
      const bool acc = active && !(lenq > r2);
      const float rinv = rsqrt_raw(r2);
      any_open |= open;
      if (CAREFUL)
      {
        if (__any_sync(kFull, acc && (r2 < h2)))
        { // Gadget spline kernel in double, like the reference (src/gravity_tree.cpp:146-160): the self term
          // -m*h_inv*2.8 then cancels targetMass/eps to the reference's own residual instead of fp32 round-off
          if (acc)
          {
            if (r2 < h2)
            {
              const double hd = 2.8 * (double)cfg.softening, hinv_d = 1.0 / hd;
              const double u = sqrt((double)r2) * hinv_d;
              double wp;
              if (u < 0.5)
                wp = -2.8 + u * u * (5.333333333333 + u * u * (6.4 * u - 9.6));
              else
                wp = -3.2 + 0.066666666667 / u + u * u * (10.666666666667 + u * (-16.0 + u * (9.6 - 2.133333333333 * u)));
              accs[k] += (double)n.w * hinv_d * wp;
            }
            else
              accf[k] = fmaf(-n.w, rinv, accf[k]);
          }
        }
        else if (acc)
          accf[k] = fmaf(-n.w, rinv, accf[k]);
      }
      else if (acc)
      {
        accf[k] = fmaf(-n.w, rinv, accf[k]);
        minr2 = fminf(minr2, r2);
      }
      if (acc)
      {
        skip[k] = nend; // resume after this subtree (a particle's end is no+1)
        if (COUNT) n_acc++;
      }
    }
    if (COUNT) n_vis++;
    no = __any_sync(kFull, any_open) ? no + 1 : nend;
  } while (no < tile_lim);
  return no;
}

// E / potential epilogue shared by the walk variants
template <int T>
__device__ __forceinline__ void walk_epilogue(const WalkArgs &a, const DevConfig &cfg, const Segment &sg, int j0, const bool (&valid)[T],
                                              const float (&px)[T], const float (&py)[T], const float (&pz)[T], const float (&pm)[T],
                                              const double (&accd)[T])
{
  const int MODE = sg.mode; // warp-uniform: one segment per warp
#pragma unroll
  for (int k = 0; k < T; k++)
  {
    if (!valid[k]) continue;
    const int t = sg.tgt_off + j0 + 32 * k;
    // pot = targetMass/eps + sum ; return pot*G/a   (src/gravity_tree.cpp:98,163)
    double pot = accd[k] + (double)__fdiv_rn(pm[k], cfg.softening);
    pot = pot * (double)cfg.G / (double)cfg.scale_factor;
    if (MODE == kWalkPotential)
    {
      a.out[t] = pot;
      continue;
    }
    const float x[3] = {px[k], py[k], pz[k]};
    if (MODE == kWalkBindingEnergy)
    {
      float4 v4 = a.vel[t];
      const float v[3] = {v4.x, v4.y, v4.z};
      float dv[3];
      relative_velocity(x, v, a.ref_pos, a.ref_vel, cfg, dv);
      a.out[t] = (double)dot3_rn(dv, dv) * 0.5 + pot;
      continue;
    }
    const int64_t slot = a.tgt_slot[t];
    const SubState &st = a.subs[sg.sub];
    float4 v4 = a.vel[a.ids[slot]];
    const float v[3] = {v4.x, v4.y, v4.z};
    if (MODE == kWalkRefine)
    { // Einner = BindingEnergy among the most-bound sample, current frame (src/subhalo_unbind.cpp:247)
      float dv[3];
      relative_velocity(x, v, st.ref_pos, st.ref_vel, cfg, dv);
      a.out_f[t] = (float)((double)dot3_rn(dv, dv) * 0.5 + pot);
      continue;
    }
    if (MODE == kWalkUnbindFull)
    { // E = VecNorm(dv)*0.5 + pot, stored as float  (src/gravity_tree.cpp:174, src/subhalo_unbind.cpp:350)
      float dv[3];
      relative_velocity(x, v, st.ref_pos, st.ref_vel, cfg, dv);
      a.E[slot] = (float)((double)dot3_rn(dv, dv) * 0.5 + pot);
    }
    else
    { // E += VecDot(OldVel, RefVelDiff) + dK - pot_removed  (src/subhalo_unbind.cpp:326-327)
      float ov[3];
      relative_velocity(x, v, st.old_ref_pos, st.old_ref_vel, cfg, ov);
      float s = __fadd_rn(dot3_rn(ov, st.ref_diff), st.dK);
      a.E[slot] = (float)((double)a.E[slot] + ((double)s - pot));
    }
  }
}

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

  while (no < node_end)
  {
    // stage nodes [no, no+32): one coalesced 16 B + 8 B load per lane (arrays are padded by 32 nodes)
    const int tile_base = no;
    const int tile_lim = min(no + 32, node_end);
    {
      const float4 xm = __ldg(&a.node_xm[no + lane]);
      const float2 ax = __ldg(&a.node_aux[no + lane]);
      __syncwarp();
      tile[lane].xm = xm;
      *reinterpret_cast<float2 *>(&tile[lane].lenq) = ax;
      __syncwarp();
    }
    int skip0[T];
    float accf[T];
#pragma unroll
    for (int k = 0; k < T; k++) { skip0[k] = skip[k]; accf[k] = 0.f; }
    float minr2 = INFINITY;
    const unsigned c0 = n_acc, c1 = n_vis;
    int nx = walk_tile<T, PERIODIC, COUNT, false>(tile, tile_base, tile_lim, no, px, py, pz, skip, accf, minr2, accd, cfg, h2, hinv, n_acc, n_vis);
    if (__any_sync(kFull, minr2 < h2))
    { // some lane met a softened pair in this tile: redo the tile exactly
#pragma unroll
      for (int k = 0; k < T; k++) { skip[k] = skip0[k]; accf[k] = 0.f; }
      n_acc = c0;
      n_vis = c1;
      nx = walk_tile<T, PERIODIC, COUNT, true>(tile, tile_base, tile_lim, no, px, py, pz, skip, accf, minr2, accd, cfg, h2, hinv, n_acc, n_vis);
    }
    no = nx;
#pragma unroll
    for (int k = 0; k < T; k++) accd[k] += (double)accf[k]; // <= 32 fp32 terms per flush
  }
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
  const int MODE = sg.mode; // warp-uniform: one segment per warp
#pragma unroll
  for (int k = 0; k < T; k++)
  {
    if (!valid[k]) continue;
    const int t = sg.tgt_off + j0 + 32 * k;
    // pot = targetMass/eps + sum ; return pot*G/a   (src/gravity_tree.cpp:98,163)
    double pot = accd[k] + (double)__fdiv_rn(pm[k], cfg.softening);
    pot = pot * (double)cfg.G / (double)cfg.scale_factor;
    if (MODE == kWalkPotential)
    {
      a.out[t] = pot;
      continue;
    }
    const float x[3] = {px[k], py[k], pz[k]};
    if (MODE == kWalkBindingEnergy)
    {
      float4 v4 = a.vel[t];
      const float v[3] = {v4.x, v4.y, v4.z};
      float dv[3];
      relative_velocity(x, v, a.ref_pos, a.ref_vel, cfg, dv);
      a.out[t] = (double)dot3_rn(dv, dv) * 0.5 + pot;
      continue;
    }
    const int64_t slot = a.tgt_slot[t];
    const SubState &st = a.subs[sg.sub];
    float4 v4 = a.vel[a.ids[slot]];
    const float v[3] = {v4.x, v4.y, v4.z};
    if (MODE == kWalkRefine)
    { // Einner = BindingEnergy among the most-bound sample, current frame (src/subhalo_unbind.cpp:247)
      float dv[3];
      relative_velocity(x, v, st.ref_pos, st.ref_vel, cfg, dv);
      a.out_f[t] = (float)((double)dot3_rn(dv, dv) * 0.5 + pot);
      continue;
    }
    if (MODE == kWalkUnbindFull)
    { // E = VecNorm(dv)*0.5 + pot, stored as float  (src/gravity_tree.cpp:174, src/subhalo_unbind.cpp:350)
      float dv[3];
      relative_velocity(x, v, st.ref_pos, st.ref_vel, cfg, dv);
      a.E[slot] = (float)((double)dot3_rn(dv, dv) * 0.5 + pot);
    }
    else
    { // E += VecDot(OldVel, RefVelDiff) + dK - pot_removed  (src/subhalo_unbind.cpp:326-327)
      float ov[3];
      relative_velocity(x, v, st.old_ref_pos, st.old_ref_vel, cfg, ov);
      float s = __fadd_rn(dot3_rn(ov, st.ref_diff), st.dK);
      a.E[slot] = (float)((double)a.E[slot] + ((double)s - pot));
    }
  }
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

void launch_walk(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls)
{
  if (a.nwarps <= 0) return;
  if (a.targets_per_lane == 4) launch_t<4>(a, cfg, stream);
  else if (a.targets_per_lane == 2) launch_t<2>(a, cfg, stream);
  else launch_t<1>(a, cfg, stream);
  HBT_CHECK_LAUNCH();
  ls.launches++;
}

} // namespace hbt

namespace hbt
{
int walk_targets_per_lane(int tgt_n)
{ // measured on B200 (profiles/r01_walk_notes.md): T=4 wins once a segment alone fills the GPU (~1e6 targets),
  // T=1 wins below ~5e5 where the warp count, not the instruction count, limits throughput
  static int forced = -1, big4 = 0, big2 = 0;
  if (forced < 0)
  {
    const char *e = getenv("HBTU_WALK_TPL");
    forced = e ? atoi(e) : 0;
    const char *b4 = getenv("HBTU_WALK_BIG4"), *b2 = getenv("HBTU_WALK_BIG2");
    big4 = b4 ? atoi(b4) : (1 << 20);
    big2 = b2 ? atoi(b2) : (1 << 19);
  }
  if (forced == 1 || forced == 2 || forced == 4) return tgt_n >= 32 * forced ? forced : 1;
  return tgt_n >= big4 ? 4 : (tgt_n >= big2 ? 2 : 1);
}
} // namespace hbt
