// walk_small.cu - the small-subhalo potential kernel (sm_100a): GravityTree_t::EvaluatePotential / BindingEnergy
// (src/gravity_tree.cpp:79-175) for segments whose TREE has at most `small_max` sources - the thousands of 20..200-particle
// subhaloes of a snapshot (SURVEY.md 8(d) cfg 5: ~1e5 of them) and the correction rounds whose tree is the handful of
// particles removed in the previous iteration (src/subhalo_unbind.cpp:312-330).
//
// The north star asks for a direct-sum tile kernel here.  A direct sum cannot meet the parity gates: the reference's monopole
// tree differs from the exact sum by up to 3e-3 per particle at n ~ 100 (tests/test_oracle.py::test_direct_sum_vs_reference_tree),
// the gate is 1e-3.  So this is the direct-sum SHAPE with the reference's DECISIONS: a warp owns 32 targets and sweeps ALL
// nodes of its subhalo's pre-order array front to back in one dense, branch-free loop - no tile staging, no warp-level jump
// logic, no votes.  Lane state is the resume index `skip` of the stackless walk: a node takes part iff no >= skip; an accepted
// node (!(len^2/theta^2 > r^2), src/gravity_tree.cpp:135) adds its monopole and sets skip = end(node), an opened one falls
// through to its first child (= no + 1).  Every target therefore accepts exactly the reference's node set.  Node loads are
// warp-uniform (one broadcast transaction each, L1-resident: the whole tree is a few KB) and the loop index is
// unconditional, so the loads pipeline ahead of the arithmetic.
//
// Spline-softened pairs (r < 2.8 eps) are the RULE in these small dense blobs, not the exception, so they are not handled by
// a warp-voted slow path inside the loop: a lane that accepts a softened pair pushes (r^2, m) onto its private ring in shared
// memory, and every 16 node steps all lanes evaluate their queued pairs together with the reference's kernel in double
// (src/gravity_tree.cpp:146-160) - the fp64 sequence is issued max-queue-length times per 16 steps instead of once per step.
//
// Roofline: FP32 issue; the unit is an accepted pair interaction as everywhere else.
#include "walk_common.cuh"
#include "walk_masked.cuh" // spline_wp

namespace hbt
{

static constexpr int kSmallWarps = 4;
static constexpr int kSoftFlush = 16; // node steps between two flushes of the softened-pair rings (= ring capacity per lane)

template <bool PERIODIC, bool COUNT>
__global__ void __launch_bounds__(kSmallWarps * 32) walk_small_kernel(const WalkArgs a, const DevConfig cfg)
{
  __shared__ float2 s_soft[kSmallWarps][kSoftFlush][32]; // [warp][entry][lane]: conflict-free (bank = lane)
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = walk_cta(a) * kSmallWarps + w;
  if (warp >= a.nwarps) return;
  const int seg = segment_of_warp(a.warp_off, a.nseg, warp);
  const Segment sg = a.segs[seg];
  const int j = (warp - a.warp_off[seg]) * 32 + lane;
  const bool valid = j < sg.tgt_n;
  const float4 tp = a.tgt_pm[sg.tgt_off + (valid ? j : 0)];
  const float px = tp.x, py = tp.y, pz = tp.z;
  const int t0 = a.tree_off[seg], t1 = a.tree_off[seg + 1];
  const int node_begin = t0 + (t0 > 0 ? a.cellcount[t0 - 1] : 0);
  const int node_end = t1 > t0 ? t1 + a.cellcount[t1 - 1] : node_begin;
  const float h = 2.8f * cfg.softening, h2 = h * h;
  const double hinv_d = 1.0 / (2.8 * (double)cfg.softening);
  float2(*ring)[32] = s_soft[w];

  int skip = valid ? node_begin : 0x7fffffff; // resume index of this lane's target
  float accf = 0.f;
  double accd = 0.0;
  int nsoft = 0;
  unsigned n_acc = 0;
  for (int base = node_begin; base < node_end; base += kSoftFlush)
  {
    const int lim = min(base + kSoftFlush, node_end);
#pragma unroll 4
    for (int no = base; no < lim; no++)
    {
      const float4 n = __ldg(&a.node_xm[no]);
      const float2 ax = __ldg(&a.node_aux[no]);
      float dx = n.x - px, dy = n.y - py, dz = n.z - pz;
      if (PERIODIC)
      {
        dx = nearest_f(dx, cfg.box_size, cfg.box_half);
        dy = nearest_f(dy, cfg.box_size, cfg.box_half);
        dz = nearest_f(dz, cfg.box_size, cfg.box_half);
      }
      const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx)); // FMUL, FFMA, FFMA like every other walk of the library
      const bool acc = (no >= skip) && !(ax.x > r2); // reference criterion, per target (src/gravity_tree.cpp:135)
      const bool soft = acc && (r2 < h2);
      const float rinv = rsqrt_raw(r2);
      if (acc && !soft) accf = fmaf(-n.w, rinv, accf);
      if (soft)
      {
        ring[nsoft][lane] = make_float2(r2, n.w);
        nsoft++;
      }
      if (acc)
      {
        skip = __float_as_int(ax.y); // resume after this subtree (a particle's end is no + 1)
        if (COUNT) n_acc++;
      }
    }
    accd += (double)accf; // <= 16 fp32 terms per flush
    accf = 0.f;
    const int mx = __reduce_max_sync(kFull, (unsigned)nsoft);
    for (int i = 0; i < mx; i++)
      if (i < nsoft)
      { // Gadget spline kernel in double, like the reference (src/gravity_tree.cpp:146-160); r = 0 (the self term) gives
        // -2.8 m / h exactly and cancels targetMass/eps to the reference's own residual
        const float2 e = ring[i][lane];
        accd += (double)e.y * hinv_d * spline_wp(e.x, hinv_d);
      }
    nsoft = 0;
  }
  if (COUNT)
  {
    unsigned long long na = n_acc;
    for (int o = 16; o > 0; o >>= 1) na += __shfl_xor_sync(kFull, na, o);
    if (lane == 0)
    {
      atomicAdd(&a.counters[0], na);
      atomicAdd(&a.counters[1], (unsigned long long)(node_end - node_begin));
    }
  }
  const bool v1[1] = {valid};
  const float x1[1] = {px}, y1[1] = {py}, z1[1] = {pz}, m1[1] = {tp.w};
  const double d1[1] = {accd};
  walk_epilogue<1>(a, cfg, sg, (warp - a.warp_off[seg]) * 32 + lane, v1, x1, y1, z1, m1, d1);
}

void launch_walk_small(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls)
{
  if (a.nwarps <= 0) return;
  const int grid = walk_grid(a, div_up(a.nwarps, kSmallWarps));
  if (grid <= 0) return; // target split: none of the 16-CTA chunks of this launch is this context's
  const bool count = a.counters != nullptr;
  if (cfg.periodic)
  {
    if (count) walk_small_kernel<true, true><<<grid, kSmallWarps * 32, 0, stream>>>(a, cfg);
    else walk_small_kernel<true, false><<<grid, kSmallWarps * 32, 0, stream>>>(a, cfg);
  }
  else
  {
    if (count) walk_small_kernel<false, true><<<grid, kSmallWarps * 32, 0, stream>>>(a, cfg);
    else walk_small_kernel<false, false><<<grid, kSmallWarps * 32, 0, stream>>>(a, cfg);
  }
  HBT_CHECK_LAUNCH();
  ls.launches++;
}

} // namespace hbt
