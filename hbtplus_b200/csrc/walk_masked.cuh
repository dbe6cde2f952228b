// walk_masked.cuh - core of the masked group walk (sm_100a): the whole Barnes-Hut walk of a group of 128 targets
// without any per-lane tree traversal.
//
// Same decisions as every other walk of this library: each target applies the REFERENCE opening criterion
// len^2 > r^2 theta^2 to every node it meets (src/gravity_tree.cpp:135) and so accepts exactly the reference's nodes.
// What changes against walk_group.cu is what happens to the nodes on which the targets of the group DISAGREE.  There
// the subtree of such a node went through the lock-step per-lane walk, in which every lane steps through the union of
// the nodes any target visits (31 % useful lane-steps, a long dependent chain per step).  Here every sibling chain on
// the stack carries a 128-bit OPENER MASK - the targets that opened all ancestors of the chain - and
//
//   * chains are classified node-parallel against the bounding box of the whole group, one lane per chain, exactly as
//     in walk_group.cu: FAR (all targets accept), OPEN (all open), NEAR (particle, maybe softened), MIXED;
//   * FAR nodes of full-mask chains go to the dense ring (the bare 8-slot pair kernel, two targets per FADD2/FFMA2);
//   * OPEN cells push their children as a new chain with the SAME mask;
//   * every other node becomes a MASKED ENTRY (node, len^2/theta^2, mask): the warp evaluates it for the targets of the
//     mask only - slice pairs whose mask words are empty are skipped -, each target deciding for itself; the targets
//     that open the node form the mask of the chain of its children.  A node everyone in the mask accepts is the same
//     entry with len^2/theta^2 = 0.  Entries whose box bound cannot exclude a softened pair or a second periodic image
//     are evaluated with the reference's full kernel per target (spline in double, NEAREST per target).
//
// All loops are dense and free of walk-order dependencies, partial sums are fp32 over at most kMCap entries and then
// fp64 per target in a fixed order (run-to-run reproducible).  The chain stack is bounded; running out of it makes the
// function return false and the caller redoes the group with the per-lane walk (never seen with real trees).
//
// This header has no includes on purpose: walk_masked.cu includes it after walk_common.cuh; tests/host_emul/
// masked_emul.cpp includes it after a warp-emulation shim (32 fibers) to unit-test the logic on the CPU.
#pragma once

namespace hbt
{

static constexpr int kMStack = 176; // chain entries per warp
static constexpr int kMCap = 48;    // ring of pending masked entries
static constexpr int kMPend = 16;   // evaluated whenever more than this many are pending (<= 32 arrive per iteration)

struct __align__(16) MEntry
{
  float4 nxm;    // -x, -y, -z, -m: operands of the packed adds and of the accumulate
  float lenq;    // len^2/theta^2; 0 = every target of the mask accepts
  int cur1, kend; // children of the node: [cur1, kend)
  int exact;     // 1: raw coordinates, per-target NEAREST and spline test (the box bound could not exclude them)
  unsigned m[4]; // targets taking part (bit = lane, word = slice)
};
struct __align__(8) ChainEntry
{
  int cur, pend; // siblings still to classify: cur, end(cur), ... < pend
  unsigned m[4];
};
struct MaskedSmem
{
  float4 alist[64]; // ring of FAR nodes of full-mask chains (periodic: shifted to the group's image)
  MEntry mlist[kMCap];
  ChainEntry stack[kMStack];
};

__device__ __forceinline__ double spline_wp(float r2, double hinv_d)
{ // Gadget spline kernel in double (src/gravity_tree.cpp:146-160)
  const double u = sqrt((double)r2) * hinv_d;
  if (u < 0.5) return -2.8 + u * u * (5.333333333333 + u * u * (6.4 * u - 9.6));
  return -3.2 + 0.066666666667 / u + u * u * (10.666666666667 + u * (-16.0 + u * (9.6 - 2.133333333333 * u)));
}

// dense evaluation of `cnt` FAR nodes for all 128 targets: 7 packed + 2 MUFU.RSQ per two interactions
__device__ __forceinline__ void masked_eval_far(const float4 *__restrict__ ring, int base, int cnt, const float (&px)[4], const float (&py)[4],
                                                const float (&pz)[4], double (&accd)[4])
{
  float2 accf[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll 2
  for (int i = 0; i < cnt; i++)
  {
    const float4 nd = ring[(base + i) & 63];
    const float2 nx = make_float2(-nd.x, -nd.x), ny = make_float2(-nd.y, -nd.y), nz = make_float2(-nd.z, -nd.z), nw = make_float2(-nd.w, -nd.w);
#pragma unroll
    for (int k = 0; k < 4; k += 2)
    {
      const float2 dx = f2_add(make_float2(px[k], px[k + 1]), nx);
      const float2 dy = f2_add(make_float2(py[k], py[k + 1]), ny);
      const float2 dz = f2_add(make_float2(pz[k], pz[k + 1]), nz);
      const float2 r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
      accf[k / 2] = f2_fma(nw, make_float2(rsqrt_raw(r2.x), rsqrt_raw(r2.y)), accf[k / 2]);
    }
  }
  accd[0] += (double)accf[0].x;
  accd[1] += (double)accf[0].y;
  accd[2] += (double)accf[1].x;
  accd[3] += (double)accf[1].y;
}

// one slice pair (K, K+1) of a masked entry whose box bound already excluded softening and a second image
template <int K, bool COUNT>
__device__ __forceinline__ void masked_pair(const float4 &n, float lenq, unsigned ma, unsigned mb, unsigned lanebit, const float (&px)[4],
                                            const float (&py)[4], const float (&pz)[4], float (&accf)[4], unsigned &opa, unsigned &opb,
                                            unsigned &n_acc)
{
  const float2 dx = f2_add(make_float2(px[K], px[K + 1]), make_float2(n.x, n.x));
  const float2 dy = f2_add(make_float2(py[K], py[K + 1]), make_float2(n.y, n.y));
  const float2 dz = f2_add(make_float2(pz[K], pz[K + 1]), make_float2(n.z, n.z));
  const float2 r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
  const float ra = rsqrt_raw(r2.x), rb = rsqrt_raw(r2.y);
  const bool ina = (ma & lanebit) != 0u, inb = (mb & lanebit) != 0u;
  const bool opena = lenq > r2.x, openb = lenq > r2.y; // reference criterion, per target (src/gravity_tree.cpp:135)
  if (ina && !opena) accf[K] = fmaf(n.w, ra, accf[K]);
  if (inb && !openb) accf[K + 1] = fmaf(n.w, rb, accf[K + 1]);
  opa = __ballot_sync(kFull, ina && opena);
  opb = __ballot_sync(kFull, inb && openb);
  if (COUNT) n_acc += (unsigned)(ina && !opena) + (unsigned)(inb && !openb);
}

// one slice of an entry that needs the reference's full kernel per target (src/gravity_tree.cpp:141-161)
template <bool PERIODIC, bool COUNT>
__device__ __forceinline__ void masked_exact(const float4 &n, float lenq, unsigned m, unsigned lanebit, float pxk, float pyk, float pzk, float &accf,
                                             double &accd, unsigned &op, float box_size, float box_half, float h2, double hinv_d,
                                             unsigned &n_acc)
{
  float dx = pxk + n.x, dy = pyk + n.y, dz = pzk + n.z;
  if (PERIODIC)
  {
    dx = nearest_f(dx, box_size, box_half);
    dy = nearest_f(dy, box_size, box_half);
    dz = nearest_f(dz, box_size, box_half);
  }
  const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx)); // FMUL, FFMA, FFMA like the packed path
  const bool in = (m & lanebit) != 0u;
  const bool open = lenq > r2;
  const bool acc = in && !open;
  const bool soft = acc && r2 < h2;
  if (__any_sync(kFull, soft))
  {
    if (soft)
      accd += (double)(-n.w) * hinv_d * spline_wp(r2, hinv_d);
    else if (acc)
      accf = fmaf(n.w, rsqrt_raw(r2), accf);
  }
  else if (acc)
    accf = fmaf(n.w, rsqrt_raw(r2), accf);
  op = __ballot_sync(kFull, in && open);
  if (COUNT) n_acc += (unsigned)acc;
}

// evaluate the pending masked entries [mb, mb+cnt) of the ring; openers push the chain of the node's children.
// Returns false when the chain stack is full.
template <bool PERIODIC, bool COUNT>
__device__ __forceinline__ bool masked_eval(MaskedSmem &sm, int mb, int cnt, int lane, unsigned lanebit, const float (&px)[4], const float (&py)[4],
                                            const float (&pz)[4], double (&accd)[4], float box_size, float box_half, float h2, double hinv_d,
                                            int &ncs, unsigned &n_acc)
{
  float accf[4] = {0.f, 0.f, 0.f, 0.f};
  bool ok = true;
  for (int i = 0; i < cnt; i++)
  {
    int idx = mb + i;
    if (idx >= kMCap) idx -= kMCap;
    const MEntry &e = sm.mlist[idx];
    const float4 n = e.nxm;
    const float lenq = e.lenq;
    const int exact = e.exact;
    const unsigned m0 = e.m[0], m1 = e.m[1], m2 = e.m[2], m3 = e.m[3];
    unsigned o0 = 0u, o1 = 0u, o2 = 0u, o3 = 0u;
    if (!exact)
    {
      if ((m0 | m1) != 0u) masked_pair<0, COUNT>(n, lenq, m0, m1, lanebit, px, py, pz, accf, o0, o1, n_acc);
      if ((m2 | m3) != 0u) masked_pair<2, COUNT>(n, lenq, m2, m3, lanebit, px, py, pz, accf, o2, o3, n_acc);
    }
    else
    {
      if (m0 != 0u) masked_exact<PERIODIC, COUNT>(n, lenq, m0, lanebit, px[0], py[0], pz[0], accf[0], accd[0], o0, box_size, box_half, h2, hinv_d, n_acc);
      if (m1 != 0u) masked_exact<PERIODIC, COUNT>(n, lenq, m1, lanebit, px[1], py[1], pz[1], accf[1], accd[1], o1, box_size, box_half, h2, hinv_d, n_acc);
      if (m2 != 0u) masked_exact<PERIODIC, COUNT>(n, lenq, m2, lanebit, px[2], py[2], pz[2], accf[2], accd[2], o2, box_size, box_half, h2, hinv_d, n_acc);
      if (m3 != 0u) masked_exact<PERIODIC, COUNT>(n, lenq, m3, lanebit, px[3], py[3], pz[3], accf[3], accd[3], o3, box_size, box_half, h2, hinv_d, n_acc);
    }
    if ((o0 | o1 | o2 | o3) != 0u)
    { // the targets that opened this node walk its children
      if (ncs >= kMStack)
      {
        ok = false;
        break;
      }
      if (lane == 0)
      {
        ChainEntry &c = sm.stack[ncs];
        c.cur = e.cur1;
        c.pend = e.kend;
        c.m[0] = o0; c.m[1] = o1; c.m[2] = o2; c.m[3] = o3;
      }
      ncs++;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) accd[k] += (double)accf[k];
  __syncwarp();
  return ok;
}

// The walk of one group: targets px/py/pz (4 per lane: slice k = targets 32k .. 32k+31 of the group; periodic: already
// un-wrapped towards one common image; invalid slots repeat a valid position) over the pre-order nodes
// [node_begin, node_end).  accd[k] receives sum(-m/r) (softened pairs: the spline term) of target (lane, k).
// nacc: warp-uniform part of the accepted-interaction count; n_acc: per-lane part; n_vis: node-parallel iterations.
template <bool PERIODIC, bool COUNT>
__device__ __forceinline__ bool masked_group_walk(MaskedSmem &sm, int lane, const float4 *__restrict__ node_xm, const float2 *__restrict__ node_aux,
                                                  int node_begin, int node_end, const float (&px)[4], const float (&py)[4], const float (&pz)[4],
                                                  const bool (&valid)[4], int n0, float box_size, float box_half, float softening, double (&accd)[4],
                                                  unsigned long long &nacc, unsigned &n_acc, unsigned &n_vis)
{
  const unsigned lt = (1u << lane) - 1u, lanebit = 1u << lane;
  const float h = 2.8f * softening, h2 = h * h;
  const double hinv_d = 1.0 / (2.8 * (double)softening);
  // bounding box of the group (ordered-uint REDUX), centre + inflated half widths
  float cx, cy, cz, hx, hy, hz;
  {
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 4; k++)
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
  // mask of the whole group (n0 bits)
  const unsigned vm0 = __ballot_sync(kFull, valid[0]), vm1 = __ballot_sync(kFull, valid[1]);
  const unsigned vm2 = __ballot_sync(kFull, valid[2]), vm3 = __ballot_sync(kFull, valid[3]);

  int ncs = 0;         // chains on the stack
  int na = 0, ab = 0;  // dense ring: pending, base
  int nm = 0, mb = 0;  // masked ring: pending, base
  if (node_end > node_begin)
  {
    if (lane == 0)
    {
      ChainEntry &c = sm.stack[0];
      c.cur = node_begin;
      c.pend = node_end;
      c.m[0] = vm0; c.m[1] = vm1; c.m[2] = vm2; c.m[3] = vm3;
    }
    ncs = 1;
  }
  __syncwarp();

  // One loop: every iteration classifies the node each active lane stands on and moves the lane to the node's sibling.
  // When no lane has a node left, up to 32 chains are taken off the stack; pending masked entries are evaluated at one
  // place, before they could overflow their ring or when nothing else is left.
  int cur = 0, pend = 0;
  unsigned c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u; // mask of the lane's chain
  bool full = false;                           // ... it is the whole group
  while (true)
  {
    const bool anyact = __any_sync(kFull, cur < pend);
    if (nm > kMPend || (!anyact && ncs == 0 && nm > 0))
    {
      if (!masked_eval<PERIODIC, COUNT>(sm, mb, nm, lane, lanebit, px, py, pz, accd, box_size, box_half, h2, hinv_d, ncs, n_acc)) return false;
      mb += nm;
      if (mb >= kMCap) mb -= kMCap;
      nm = 0;
    }
    if (!anyact)
    {
      if (ncs == 0) break;
      // a chain pushes at most one entry per child; most children are accepted, so a quarter of the worst case is reserved
      const int take = min(min(32, ncs), max(1, (kMStack - ncs - nm) >> 2));
      ncs -= take;
      if (lane < take)
      {
        const ChainEntry &c = sm.stack[ncs + lane];
        cur = c.cur;
        pend = c.pend;
        c0 = c.m[0]; c1 = c.m[1]; c2 = c.m[2]; c3 = c.m[3];
      }
      full = __popc(c0) + __popc(c1) + __popc(c2) + __popc(c3) == n0;
      __syncwarp();
    }
    const bool act = cur < pend;
    int cls = 0; // 1 FAR, 2 NEAR, 3 OPEN, 4 MIXED
    float4 xm = make_float4(0.f, 0.f, 0.f, 0.f), xs = xm;
    float lenq = 0.f;
    int kend = 0;
    bool bare = false; // box bound excludes softened pairs and a second periodic image
    if (act)
    {
      xm = __ldg(&node_xm[cur]);
      const float2 ax = __ldg(&node_aux[cur]);
      lenq = ax.x;
      kend = __float_as_int(ax.y);
      float dx = xm.x - cx, dy = xm.y - cy, dz = xm.z - cz;
      xs = xm;
      bool wrap_ok = true;
      if (PERIODIC)
      {
        if (dx > box_half) { dx -= box_size; xs.x -= box_size; } else if (dx < -box_half) { dx += box_size; xs.x += box_size; }
        if (dy > box_half) { dy -= box_size; xs.y -= box_size; } else if (dy < -box_half) { dy += box_size; xs.y += box_size; }
        if (dz > box_half) { dz -= box_size; xs.z -= box_size; } else if (dz < -box_half) { dz += box_size; xs.z += box_size; }
        const float lim = box_half * 0.9999f;
        wrap_ok = (fabsf(dx) + hx < lim) && (fabsf(dy) + hy < lim) && (fabsf(dz) + hz < lim);
      }
      const float adx = fabsf(dx), ady = fabsf(dy), adz = fabsf(dz);
      const float nx = fmaxf(adx - hx, 0.f), ny = fmaxf(ady - hy, 0.f), nz = fmaxf(adz - hz, 0.f);
      const float fx = adx + hx, fy = ady + hy, fz = adz + hz;
      const float r2min = (nx * nx + ny * ny + nz * nz) * 0.99998f;
      const float r2max = (fx * fx + fy * fy + fz * fz) * 1.00002f;
      bare = wrap_ok && r2min >= h2;
      if (lenq == 0.f) cls = bare ? 1 : 2; // a particle is accepted by everyone
      else if (!wrap_ok) cls = 4;
      else if (lenq > r2max) cls = 3;
      else if (!(lenq > r2min) && bare) cls = 1;
      else cls = 4;
    }
    if (COUNT) n_vis++;
    const bool toA = (cls == 1) && full, toO = (cls == 3), toM = act && !toA && !toO;
    const unsigned mA = __ballot_sync(kFull, toA), mO = __ballot_sync(kFull, toO), mM = __ballot_sync(kFull, toM);
    const int cO = __popc(mO);
    if (ncs + cO > kMStack) return false; // stack exhausted (pathologically deep tree): the caller redoes the group per lane
    if (toA) sm.alist[(ab + na + __popc(mA & lt)) & 63] = xs;
    if (toO)
    {
      ChainEntry &c = sm.stack[ncs + __popc(mO & lt)];
      c.cur = cur + 1;
      c.pend = kend;
      c.m[0] = c0; c.m[1] = c1; c.m[2] = c2; c.m[3] = c3;
    }
    if (toM)
    {
      int idx = mb + nm + __popc(mM & lt);
      if (idx >= kMCap) idx -= kMCap;
      MEntry &e = sm.mlist[idx];
      const float4 p = bare ? xs : xm;
      e.nxm = make_float4(-p.x, -p.y, -p.z, -p.w);
      e.lenq = (cls == 1) ? 0.f : lenq; // FAR for the whole group: accepted by every target of the mask
      e.cur1 = cur + 1;
      e.kend = kend;
      e.exact = bare ? 0 : 1;
      e.m[0] = c0; e.m[1] = c1; e.m[2] = c2; e.m[3] = c3;
    }
    na += __popc(mA);
    ncs += cO;
    nm += __popc(mM);
    __syncwarp();
    if (na >= 32)
    {
      masked_eval_far(sm.alist, ab, 32, px, py, pz, accd);
      if (COUNT) nacc += 32ull * n0;
      ab = (ab + 32) & 63;
      na -= 32;
    }
    if (act) cur = kend;
  }
  if (na > 0)
  {
    masked_eval_far(sm.alist, ab, na, px, py, pz, accd);
    if (COUNT) nacc += (unsigned long long)na * n0;
  }
  return true;
}

} // namespace hbt
