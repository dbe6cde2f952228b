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

template <bool PERIODIC, bool COUNT>
__global__ void __launch_bounds__(kWalkWarps * 32) walk_kernel(const WalkArgs a, const DevConfig cfg)
{
  __shared__ float4 s_xm[kWalkWarps][32];
  __shared__ float2 s_aux[kWalkWarps][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kWalkWarps + w;
  if (warp >= a.nwarps) return;
  // segment of this warp: largest s with warp_off[s] <= warp
  int lo = 0, hi = a.nseg;
  while (hi - lo > 1)
  {
    int mid = (lo + hi) >> 1;
    if (a.warp_off[mid] <= warp) lo = mid; else hi = mid;
  }
  const Segment sg = a.segs[lo];
  const int j = (warp - a.warp_off[lo]) * 32 + lane;
  const bool valid = j < sg.tgt_n;
  const int t = sg.tgt_off + (valid ? j : 0);
  const float4 tp = a.tgt_pm[t];
  const int t0 = a.tree_off[lo], t1 = a.tree_off[lo + 1];
  const int node_begin = t0 + (t0 > 0 ? a.cellcount[t0 - 1] : 0);
  const int node_end = t1 + a.cellcount[t1 - 1];

  const float h = 2.8f * cfg.softening, h2 = h * h, hinv = 1.0f / h;
  int skip = valid ? 0 : 0x7fffffff; // resume index of this lane
  int no = node_begin, tile_base = -0x40000000;
  float accf = 0.f;
  double accd = 0.0;
  unsigned it = 0;
  unsigned long long n_acc = 0, n_vis = 0;

  while (no < node_end)
  {
    int jj = no - tile_base;
    if ((unsigned)jj >= 32u)
    {
      tile_base = no;
      jj = 0;
      int idx = no + lane;
      float4 xm = make_float4(0.f, 0.f, 0.f, 0.f);
      float2 ax = make_float2(0.f, 0.f);
      if (idx < node_end)
      {
        xm = __ldg(&a.node_xm[idx]);
        ax = __ldg(&a.node_aux[idx]);
      }
      __syncwarp();
      s_xm[w][lane] = xm;
      s_aux[w][lane] = ax;
      __syncwarp();
    }
    const float4 n = s_xm[w][jj];
    const float2 ax = s_aux[w][jj];
    float dx = n.x - tp.x, dy = n.y - tp.y, dz = n.z - tp.z;
    if (PERIODIC)
    {
      dx = nearest_f(dx, cfg.box_size, cfg.box_half);
      dy = nearest_f(dy, cfg.box_size, cfg.box_half);
      dz = nearest_f(dz, cfg.box_size, cfg.box_half);
    }
    const float r2 = dx * dx + dy * dy + dz * dz;
    const bool active = no >= skip;
    const bool open = active && (ax.x > r2);
    const bool acc = active && !open;
    const int nend = __float_as_int(ax.y);
    float contrib = -n.w * rsqrtf(r2);
    if (__any_sync(kFull, acc && (r2 < h2)))
    {
      if (r2 < h2)
      { // Gadget spline kernel, src/gravity_tree.cpp:146-160
        float u = sqrtf(r2) * hinv, wp;
        if (u < 0.5f)
          wp = -2.8f + u * u * (5.333333333333f + u * u * (6.4f * u - 9.6f));
        else
          wp = -3.2f + 0.066666666667f / u + u * u * (10.666666666667f + u * (-16.0f + u * (9.6f - 2.133333333333f * u)));
        contrib = n.w * hinv * wp;
      }
    }
    if (acc)
    {
      accf += contrib;
      skip = nend;
      if (COUNT) n_acc++;
    }
    if (COUNT) n_vis++;
    no = __any_sync(kFull, open) ? no + 1 : nend;
    if (((++it) & 63u) == 0u)
    {
      accd += (double)accf;
      accf = 0.f;
    }
  }
  if (COUNT)
  {
    for (int o = 16; o > 0; o >>= 1) n_acc += __shfl_xor_sync(kFull, n_acc, o);
    if (lane == 0)
    {
      atomicAdd(&a.counters[0], n_acc);
      atomicAdd(&a.counters[1], n_vis);
    }
  }
  if (!valid) return;
  // pot = targetMass/eps + sum ; return pot*G/a   (src/gravity_tree.cpp:98,163)
  double pot = accd + (double)accf + (double)__fdiv_rn(tp.w, cfg.softening);
  pot = pot * (double)cfg.G / (double)cfg.scale_factor;

  const int MODE = sg.mode; // warp-uniform: one segment per warp
  if (MODE == kWalkPotential)
  {
    a.out[t] = pot;
    return;
  }
  const float x[3] = {tp.x, tp.y, tp.z};
  if (MODE == kWalkBindingEnergy)
  {
    float4 v4 = a.vel[t];
    const float v[3] = {v4.x, v4.y, v4.z};
    float dv[3];
    relative_velocity(x, v, a.ref_pos, a.ref_vel, cfg, dv);
    a.out[t] = (double)dot3_rn(dv, dv) * 0.5 + pot;
    return;
  }
  const int64_t slot = a.tgt_slot[t];
  const SubState &st = a.subs[sg.sub];
  float4 v4 = a.vel[a.ids[slot]];
  const float v[3] = {v4.x, v4.y, v4.z};
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

void launch_walk(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls)
{
  if (a.nwarps <= 0) return;
  const int grid = div_up(a.nwarps, kWalkWarps);
  const bool count = a.counters != nullptr;
  if (cfg.periodic)
  {
    if (count) walk_kernel<true, true><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
    else walk_kernel<true, false><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
  }
  else
  {
    if (count) walk_kernel<false, true><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
    else walk_kernel<false, false><<<grid, kWalkWarps * 32, 0, stream>>>(a, cfg);
  }
  HBT_CHECK_LAUNCH();
  ls.launches++;
}

} // namespace hbt
