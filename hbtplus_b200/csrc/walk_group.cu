// walk_group.cu - group walk: the Barnes-Hut potential walk for large subhaloes (sm_100a).
//
// Same results as walk.cu (every target applies the REFERENCE opening criterion len^2 > r^2 theta^2 to every
// node it meets, src/gravity_tree.cpp:135, so each target accepts exactly the reference's set of nodes), but the
// far field - the nodes on which all targets of the warp agree - no longer goes through the per-lane walk:
//
//   * a warp owns a GROUP of 32*T consecutive targets (key order => a compact box);
//   * the children of opened cells are classified NODE-PARALLEL (one node per lane, lanes follow the sibling
//     links of up to 32 opened cells at once) against the group's bounding box:
//          FAR    every target accepts it   (len^2/theta^2 <= min r^2, min r > 2.8 eps, one periodic image)
//          OPEN   every target opens it     (len^2/theta^2 >  max r^2)   -> its children are queued
//          NEAR   a particle (always accepted) that may be softened or wrapped differently per target
//          MIXED  the targets disagree (or the bounds cannot tell);
//     the bounds are conservative (box inflated by 1e-5, decisions taken with a 2e-5 margin), so FAR / OPEN are
//     exactly what each target's own fp32 test would say;
//   * FAR nodes go to a shared-memory ring that all lanes then consume densely: 3 FADD + FMUL + 2 FFMA + MUFU.RSQ +
//     FFMA per pair and nothing else (no criterion, no resume index, no softening test) - the roofline's 8 slots;
//   * the subtree of a MIXED node is walked per lane exactly like walk.cu does (walk_range: staged tiles, per-target
//     criterion and resume index, exact spline redo), starting at the MIXED node itself.
//
// Partial sums: fp32 over <= 32 nodes, then fp64 per target (the reference accumulates in double).  The order of
// the sums is fixed, so results are run-to-run reproducible.
// Roofline: FP32 issue.  Tensor cores are deliberately not used.
#include <cstdlib>

#include "walk_common.cuh"

namespace hbt
{

static constexpr int kGW = 4;      // warps per CTA (warps are independent)
#ifndef HBT_GROUP_MINBLOCKS
#define HBT_GROUP_MINBLOCKS 7 // resident CTAs per SM the register allocation must allow (measured, profiles/r01_walk_notes.md)
#endif
static constexpr int kWork = 768;  // ints per warp: chain stack (bottom) + MIXED list (top, grows down), both (node, end) pairs

struct GroupSmem
{
  float4 alist[64]; // ring of FAR nodes (periodic: already shifted to the group's image)
  float4 clist[64]; // ring of NEAR nodes (raw)
  TileNode tile[32];
  int work[kWork];
};

__device__ __forceinline__ double spline_d(float r2, double hinv_d)
{ // Gadget spline kernel in double (src/gravity_tree.cpp:146-160); returns wp
  const double u = sqrt((double)r2) * hinv_d;
  if (u < 0.5) return -2.8 + u * u * (5.333333333333 + u * u * (6.4 * u - 9.6));
  return -3.2 + 0.066666666667 / u + u * u * (10.666666666667 + u * (-16.0 + u * (9.6 - 2.133333333333 * u)));
}

__device__ __forceinline__ unsigned lanemask_lt()
{
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// dense evaluation of `cnt` FAR nodes for all targets of the warp: the pair kernel and nothing else
template <int T>
__device__ __forceinline__ void eval_far(const float4 *__restrict__ ring, int base, int cnt, const float (&px)[T], const float (&py)[T],
                                         const float (&pz)[T], double (&accd)[T])
{
  // two targets per FADD2 / FMUL2 / FFMA2: 7 packed + 2 MUFU.RSQ per pair of interactions
  float2 accf[T / 2];
#pragma unroll
  for (int k = 0; k < T / 2; k++) accf[k] = make_float2(0.f, 0.f);
#pragma unroll 2
  for (int i = 0; i < cnt; i++)
  {
    const float4 nd = ring[(base + i) & 63];
    const float2 nx = make_float2(-nd.x, -nd.x), ny = make_float2(-nd.y, -nd.y), nz = make_float2(-nd.z, -nd.z), nw = make_float2(-nd.w, -nd.w);
#pragma unroll
    for (int k = 0; k < T; k += 2)
    {
      const float2 dx = f2_add(make_float2(px[k], px[k + 1]), nx);
      const float2 dy = f2_add(make_float2(py[k], py[k + 1]), ny);
      const float2 dz = f2_add(make_float2(pz[k], pz[k + 1]), nz);
      const float2 r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
      accf[k / 2] = f2_fma(nw, make_float2(rsqrt_raw(r2.x), rsqrt_raw(r2.y)), accf[k / 2]);
    }
  }
#pragma unroll
  for (int k = 0; k < T; k += 2)
  {
    accd[k] += (double)accf[k / 2].x;
    accd[k + 1] += (double)accf[k / 2].y;
  }
}

// NEAR nodes: particles accepted by every target, but the pair may be softened (r < 2.8 eps, incl. the self pair
// r = 0) or wrapped differently per target: the reference's full kernel per target (src/gravity_tree.cpp:141-161)
template <int T, bool PERIODIC>
__device__ __forceinline__ void eval_near(const float4 *__restrict__ ring, int base, int cnt, const float (&px)[T], const float (&py)[T],
                                          const float (&pz)[T], const bool (&valid)[T], double (&accd)[T], const DevConfig &cfg, float h2,
                                          double hinv_d)
{
  float accf[T];
#pragma unroll
  for (int k = 0; k < T; k++) accf[k] = 0.f;
  for (int i = 0; i < cnt; i++)
  {
    const float4 nd = ring[(base + i) & 63];
#pragma unroll
    for (int k = 0; k < T; k++)
    {
      float dx = nd.x - px[k], dy = nd.y - py[k], dz = nd.z - pz[k];
      if (PERIODIC)
      {
        dx = nearest_f(dx, cfg.box_size, cfg.box_half);
        dy = nearest_f(dy, cfg.box_size, cfg.box_half);
        dz = nearest_f(dz, cfg.box_size, cfg.box_half);
      }
      const float r2 = dx * dx + dy * dy + dz * dz;
      const bool soft = valid[k] && r2 < h2;
      if (__any_sync(kFull, soft))
      {
        if (soft)
          accd[k] += (double)nd.w * hinv_d * spline_d(r2, hinv_d);
        else
          accf[k] = fmaf(-nd.w, rsqrt_raw(r2), accf[k]);
      }
      else
        accf[k] = fmaf(-nd.w, rsqrt_raw(r2), accf[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < T; k++) accd[k] += (double)accf[k];
}

template <int T, bool PERIODIC, bool COUNT>
__global__ void __launch_bounds__(kGW * 32, HBT_GROUP_MINBLOCKS) walk_group_kernel(const WalkArgs a, const DevConfig cfg)
{
  __shared__ GroupSmem s_all[kGW];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kGW + w;
  if (warp >= a.nwarps) return;
  GroupSmem &sm = s_all[w];
  const unsigned lt = lanemask_lt();
  const int seg = segment_of_warp(a.warp_off, a.nseg, warp);
  const Segment sg = a.segs[seg];
  const int j0 = (warp - a.warp_off[seg]) * (32 * T) + lane;
  const int n0 = min(32 * T, sg.tgt_n - (warp - a.warp_off[seg]) * (32 * T)); // valid targets of the warp
  // px/py/pz: positions used by the walk (periodic: unwrapped towards the warp's first target, so that the group
  // has one bounding box and one image per FAR node); the epilogue gets the raw positions again from tgt_pm.
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
  const float h = 2.8f * cfg.softening, h2 = h * h, hinv = 1.0f / h;
  const double hinv_d = 1.0 / (2.8 * (double)cfg.softening);

  // bounding box of the group (ordered-uint REDUX; invalid slots repeat the first target), centre + inflated half widths
  float cx, cy, cz, hx, hy, hz;
  {
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < T; k++)
    {
      const unsigned ux = float_to_ordered(px[k]), uy = float_to_ordered(py[k]), uz = float_to_ordered(pz[k]);
      lo[0] = min(lo[0], ux); lo[1] = min(lo[1], uy); lo[2] = min(lo[2], uz);
      hi[0] = max(hi[0], ux); hi[1] = max(hi[1], uy); hi[2] = max(hi[2], uz);
    }
    float l[3], hh[3];
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
      l[j] = ordered_to_float(__reduce_min_sync(kFull, lo[j]));
      hh[j] = ordered_to_float(__reduce_max_sync(kFull, hi[j]));
    }
    cx = 0.5f * (l[0] + hh[0]); cy = 0.5f * (l[1] + hh[1]); cz = 0.5f * (l[2] + hh[2]);
    hx = fmaxf(hh[0] - cx, cx - l[0]) * 1.00001f + 1e-30f;
    hy = fmaxf(hh[1] - cy, cy - l[1]) * 1.00001f + 1e-30f;
    hz = fmaxf(hh[2] - cz, cz - l[2]) * 1.00001f + 1e-30f;
  }

  double accd[T];
#pragma unroll
  for (int k = 0; k < T; k++) accd[k] = 0.0;
  unsigned long long nacc = 0; // warp-uniform part of the interaction count (FAR / NEAR nodes x valid targets)
  unsigned n_acc = 0, n_vis = 0; // per-lane count of the MIXED walks, warp node visits
  int ncs = 0, mtop = kWork;     // chains on the stack (pairs at work[0 .. 2*ncs)), lowest used int of the MIXED list
  int na = 0, ab = 0, nc = 0, cb = 0;
  bool overflow = false;
  if (node_end > node_begin)
  {
    if (lane == 0)
    {
      sm.work[0] = node_begin;
      sm.work[1] = node_end;
    }
    ncs = 1;
  }
  __syncwarp();

  while (true)
  {
    // MIXED subtrees: when the list is long enough to threaten the chain stack, and at the end
    const int room = mtop - 2 * ncs;
    if (mtop < kWork && (ncs == 0 || room < 16 * 32))
    {
      // entries are (node, end) pairs; the first tile of the next entry is prefetched while the current one is walked
      float4 nxm = __ldg(&a.node_xm[sm.work[kWork - 2] + lane]);
      float2 nax = __ldg(&a.node_aux[sm.work[kWork - 2] + lane]);
      for (int i = kWork - 2; i >= mtop; i -= 2)
      {
        const int no = sm.work[i], ne = sm.work[i + 1];
        const float4 xm0 = nxm;
        const float2 ax0 = nax;
        if (i - 2 >= mtop)
        {
          const int nn = sm.work[i - 2];
          nxm = __ldg(&a.node_xm[nn + lane]);
          nax = __ldg(&a.node_aux[nn + lane]);
        }
        int skip[T];
#pragma unroll
        for (int k = 0; k < T; k++) skip[k] = valid[k] ? no : 0x7fffffff;
        walk_range<T, PERIODIC, COUNT, true>(a.node_xm, a.node_aux, sm.tile, no, ne, px, py, pz, skip, accd, cfg, h2, hinv, n_acc, n_vis, xm0, ax0);
      }
      mtop = kWork;
      __syncwarp();
    }
    if (ncs == 0) break;
    const int take = min(min(32, ncs), max(1, (mtop - 2 * ncs) >> 4)); // a chain has <= 8 children, each queues one pair
    ncs -= take;
    int cur = 0, pend = 0;
    if (lane < take)
    {
      cur = sm.work[2 * (ncs + lane)];
      pend = sm.work[2 * (ncs + lane) + 1];
    }
    __syncwarp();
    while (true)
    {
      const bool act = cur < pend;
      if (!__any_sync(kFull, act)) break;
      int cls = 0; // 1 FAR, 2 NEAR, 3 OPEN, 4 MIXED
      float4 xm = make_float4(0.f, 0.f, 0.f, 0.f), xs = xm;
      int kend = 0;
      if (act)
      {
        xm = __ldg(&a.node_xm[cur]);
        const float2 ax = __ldg(&a.node_aux[cur]);
        const float lenq = ax.x;
        kend = __float_as_int(ax.y);
        float dx = xm.x - cx, dy = xm.y - cy, dz = xm.z - cz;
        xs = xm;
        bool wrap_ok = true;
        if (PERIODIC)
        {
          if (dx > cfg.box_half) { dx -= cfg.box_size; xs.x -= cfg.box_size; } else if (dx < -cfg.box_half) { dx += cfg.box_size; xs.x += cfg.box_size; }
          if (dy > cfg.box_half) { dy -= cfg.box_size; xs.y -= cfg.box_size; } else if (dy < -cfg.box_half) { dy += cfg.box_size; xs.y += cfg.box_size; }
          if (dz > cfg.box_half) { dz -= cfg.box_size; xs.z -= cfg.box_size; } else if (dz < -cfg.box_half) { dz += cfg.box_size; xs.z += cfg.box_size; }
          const float lim = cfg.box_half * 0.9999f;
          wrap_ok = (fabsf(dx) + hx < lim) && (fabsf(dy) + hy < lim) && (fabsf(dz) + hz < lim);
        }
        const float adx = fabsf(dx), ady = fabsf(dy), adz = fabsf(dz);
        const float nx = fmaxf(adx - hx, 0.f), ny = fmaxf(ady - hy, 0.f), nz = fmaxf(adz - hz, 0.f);
        const float fx = adx + hx, fy = ady + hy, fz = adz + hz;
        const float r2min = (nx * nx + ny * ny + nz * nz) * 0.99998f;
        const float r2max = (fx * fx + fy * fy + fz * fz) * 1.00002f;
        const bool far_ok = wrap_ok && r2min >= h2;
        if (lenq == 0.f) cls = far_ok ? 1 : 2; // a particle is accepted by everyone
        else if (!wrap_ok) cls = 4;
        else if (lenq > r2max) cls = 3;
        else if (!(lenq > r2min) && far_ok) cls = 1;
        else cls = 4;
      }
      if (COUNT) n_vis++;
      const unsigned mA = __ballot_sync(kFull, cls == 1), mC = __ballot_sync(kFull, cls == 2);
      const unsigned mO = __ballot_sync(kFull, cls == 3), mM = __ballot_sync(kFull, cls == 4);
      if (cls == 1) sm.alist[(ab + na + __popc(mA & lt)) & 63] = xs;
      if (cls == 2) sm.clist[(cb + nc + __popc(mC & lt)) & 63] = xm;
      na += __popc(mA);
      nc += __popc(mC);
      const int cO = __popc(mO), cM = __popc(mM);
      if (2 * (ncs + cO) > mtop - 2 * cM)
      { // stacks exhausted (pathologically deep tree): redo this group with the per-lane walk from scratch
        overflow = true;
        break;
      }
      if (cls == 3)
      {
        const int s = ncs + __popc(mO & lt);
        sm.work[2 * s] = cur + 1;
        sm.work[2 * s + 1] = kend;
      }
      if (cls == 4)
      {
        const int s = mtop - 2 - 2 * __popc(mM & lt);
        sm.work[s] = cur;
        sm.work[s + 1] = kend;
      }
      ncs += cO;
      mtop -= 2 * cM;
      __syncwarp();
      if (na >= 32)
      {
        eval_far<T>(sm.alist, ab, 32, px, py, pz, accd);
        if (COUNT) nacc += 32ull * n0;
        ab = (ab + 32) & 63;
        na -= 32;
      }
      if (nc >= 32)
      {
        eval_near<T, PERIODIC>(sm.clist, cb, 32, px, py, pz, valid, accd, cfg, h2, hinv_d);
        if (COUNT) nacc += 32ull * n0;
        cb = (cb + 32) & 63;
        nc -= 32;
      }
      if (act) cur = kend;
      __syncwarp();
    }
    if (overflow) break;
  }
  if (!overflow)
  {
    if (na > 0) eval_far<T>(sm.alist, ab, na, px, py, pz, accd);
    if (nc > 0) eval_near<T, PERIODIC>(sm.clist, cb, nc, px, py, pz, valid, accd, cfg, h2, hinv_d);
    if (COUNT) nacc += (unsigned long long)(na + nc) * n0;
  }
  else
  {
    __syncwarp();
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

template <int T, bool PERIODIC, bool COUNT>
static void launch_one(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream)
{
  walk_group_kernel<T, PERIODIC, COUNT><<<div_up(a.nwarps, kGW), kGW * 32, 0, stream>>>(a, cfg);
}

template <int T>
static void launch_group_t(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream)
{
  const bool count = a.counters != nullptr;
  if (cfg.periodic)
  {
    if (count) launch_one<T, true, true>(a, cfg, stream);
    else launch_one<T, true, false>(a, cfg, stream);
  }
  else
  {
    if (count) launch_one<T, false, true>(a, cfg, stream);
    else launch_one<T, false, false>(a, cfg, stream);
  }
}

void launch_walk_group(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls)
{
  if (a.nwarps <= 0) return;
  if (a.targets_per_lane == kWalkGroup8) launch_group_t<8>(a, cfg, stream);
  else launch_group_t<4>(a, cfg, stream);
  HBT_CHECK_LAUNCH();
  ls.launches++;
}

} // namespace hbt
