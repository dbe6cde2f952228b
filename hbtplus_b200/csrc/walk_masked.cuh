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
//   * every other node becomes a MASKED ELEMENT (node, len^2/theta^2, mask) in the ring of each slice pair (64 targets)
//     that has targets in the mask: the warp evaluates it with one packed fp32x2 pair computation, each target of the
//     mask deciding for itself; the targets that open the node form the mask of the chain of its children, which from
//     there on belongs to that slice pair.  A node everyone in the mask accepts is the same element with
//     len^2/theta^2 = 0.  Elements for which a softened accepted pair or a second periodic image cannot be excluded
//     are evaluated with the reference's full kernel per target (spline in double, NEAREST per target).
//
// All loops are dense and free of walk-order dependencies, partial sums are fp32 over at most kMCap entries and then
// fp64 per target in a fixed order (run-to-run reproducible).  The chain stack is bounded; running out of it makes the
// function return false and the caller redoes the group with the per-lane walk (never seen with real trees).
//
// This header has no includes on purpose: walk_masked.cu includes it after walk_common.cuh; tests/host_emul/
// masked_emul.cpp includes it after a warp-emulation shim (32 fibers) to unit-test the logic on the CPU.
#pragma once

#ifndef HBT_MASKED_ROOM_PER_LANE
#define HBT_MASKED_ROOM_PER_LANE 3 // free stack entries demanded per walking lane (a chain pushes <= 8, typically 2)
#endif
#ifndef HBT_M_SPLITX
#define HBT_M_SPLITX 1 // deciding lists: exact elements queued from the far end of the list and evaluated in their own loop (0: one loop with a branch; measured 1093 -> 1045 ms)
#endif
#ifndef HBT_M_UNROLL
#define HBT_M_UNROLL 4 // unroll factor of the dense / accept-all list loops (2 -> 4: 1045 -> 992 ms with SPLITX)
#endif
#ifndef HBT_M_UNROLL_D
#define HBT_M_UNROLL_D 2 // unroll factor of the branch-free deciding loop
#endif
#define HBT_M_PRAGMA_(x) _Pragma(#x)
#define HBT_M_PRAGMA_UNROLL(n) HBT_M_PRAGMA_(unroll n)
#ifndef HBT_MASKED_STAT
#define HBT_MASKED_STAT(what, n) // test hook of the CPU emulation (element counts)
#endif
#ifndef HBT_MASKED_TRACK
#define HBT_MASKED_TRACK(ncs) // test hook of the CPU emulation (stack high-water mark)
#endif

namespace hbt
{

static constexpr int kMPend = 8;          // deciding elements / pending chains are drained when more than this many wait ...
static constexpr int kMCap = kMPend + 32; // ... and at most 32 arrive per iteration
static constexpr int kAPend = 16;         // accept-all elements are evaluated when more than this many wait
static constexpr int kACap = kAPend + 32;

// Masked elements are kept per SLICE PAIR (slices 0,1 = targets 0..63 of the group, slices 2,3 = targets 64..127): an
// element sits in the list of every pair that has targets in its mask, so its evaluation is one packed fp32x2 pair
// computation without any per-word branching.  Two kinds of lists per pair:
//   accept-all (a_*): a node that is FAR for the whole group met by a chain with a partial mask - the bare pair kernel
//                     under the mask, nothing else (72 % of the masked elements of the bench);
//   deciding (d):     each target of the mask applies the criterion (or needs the exact kernel).  The chain of the
//                     children waits in `pending` with empty masks while the evaluation of each pair fills in its two
//                     opener words; a DRAIN evaluates both deciding lists and moves the pending chains somebody opened
//                     onto the stack (the others are dropped).
struct __align__(8) ChainEntry
{
  int cur, pend; // siblings still to classify: cur, end(cur), ... < pend
  unsigned m[4]; // targets walking the chain (bit = lane, word = slice)
};
struct __align__(16) DecidingElem
{
  float4 nxm;      // -x, -y, -z, -m: operands of the packed adds and of the accumulate
  float lenq;      // len^2/theta^2 (0 for a particle: every target of the mask accepts)
  int slot;        // index of the children's chain in `pending` (kMCap = none: scratch entry); bit-complemented when the
                   // element needs the exact kernel (softened pair or second periodic image not excluded)
  unsigned ma, mb; // targets taking part (bit = lane; first / second slice of the pair)
};
template <int STACK> // chain entries per warp
struct MaskedSmemT
{
  static constexpr int kStack = STACK;
  float4 alist[64];          // ring of FAR nodes of whole-group chains (periodic: shifted to the group's image)
  float box[8];              // centre [0..2] and inflated half widths [4..6] of the group's bounding box (read as two float4)
  DecidingElem d[2][kMCap];  // deciding elements per pair
  float4 a_xm[2][kACap];     // accept-all elements per pair: -x, -y, -z, -m
  uint2 a_m[2][kACap];       //   their masks
  ChainEntry pending[kMCap + 1];
  ChainEntry stack[STACK];
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
  HBT_M_PRAGMA_UNROLL(HBT_M_UNROLL)
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

// one slice of an element that needs the reference's full kernel per target (src/gravity_tree.cpp:141-161)
template <bool PERIODIC, bool COUNT>
__device__ __forceinline__ void masked_exact(const float4 &n, float lenq, bool in, float pxk, float pyk, float pzk, float &accf, double &accd,
                                             unsigned &op, float box_size, float box_half, float h2, float softening, unsigned &n_acc)
{
  float dx = pxk + n.x, dy = pyk + n.y, dz = pzk + n.z;
  if (PERIODIC)
  {
    dx = nearest_f(dx, box_size, box_half);
    dy = nearest_f(dy, box_size, box_half);
    dz = nearest_f(dz, box_size, box_half);
  }
  const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx)); // FMUL, FFMA, FFMA like the packed path
  const bool open = lenq > r2;
  const bool acc = in && !open;
  const bool soft = acc && r2 < h2;
  if (__any_sync(kFull, soft))
  {
    if (soft)
    {
      const double hinv_d = 1.0 / (2.8 * (double)softening);
      accd += (double)(-n.w) * hinv_d * spline_wp(r2, hinv_d);
    }
    else if (acc)
      accf = fmaf(n.w, rsqrt_raw(r2), accf);
  }
  else if (acc)
    accf = fmaf(n.w, rsqrt_raw(r2), accf);
  op = __ballot_sync(kFull, in && open);
  if (COUNT) n_acc += (unsigned)acc;
}

// evaluate the `cnt` accept-all elements of slice pair (K, K+1): the bare pair kernel under the mask
template <int K, bool COUNT, class MaskedSmem>
__device__ __forceinline__ void masked_eval_accept(const MaskedSmem &sm, int cnt, unsigned lanebit, const float (&px)[4], const float (&py)[4],
                                                   const float (&pz)[4], double (&accd)[4], unsigned &n_acc)
{
  constexpr int R = K / 2;
  float acca = 0.f, accb = 0.f;
  const float2 pxx = make_float2(px[K], px[K + 1]), pyy = make_float2(py[K], py[K + 1]), pzz = make_float2(pz[K], pz[K + 1]);
  HBT_M_PRAGMA_UNROLL(HBT_M_UNROLL)
  for (int i = 0; i < cnt; i++)
  {
    const float4 n = sm.a_xm[R][i];
    const uint2 m = sm.a_m[R][i];
    const float2 dx = f2_add(pxx, make_float2(n.x, n.x));
    const float2 dy = f2_add(pyy, make_float2(n.y, n.y));
    const float2 dz = f2_add(pzz, make_float2(n.z, n.z));
    const float2 r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
    const float ra = rsqrt_raw(r2.x), rb = rsqrt_raw(r2.y);
    const bool ina = (m.x & lanebit) != 0u, inb = (m.y & lanebit) != 0u;
    if (ina) acca = fmaf(n.w, ra, acca);
    if (inb) accb = fmaf(n.w, rb, accb);
    if (COUNT) n_acc += (unsigned)ina + (unsigned)inb;
  }
  accd[K] += (double)acca;
  accd[K + 1] += (double)accb;
}

// evaluate the `cnt` deciding elements of slice pair (K, K+1); the targets that open an element are written into the
// pending chain of the node's children
template <int K, bool PERIODIC, bool COUNT, class MaskedSmem>
__device__ __forceinline__ void masked_eval(MaskedSmem &sm, int cnt, int lane, unsigned lanebit, const float (&px)[4], const float (&py)[4],
                                            const float (&pz)[4], double (&accd)[4], float box_size, float box_half, float h2, float softening,
                                            unsigned &n_acc)
{
  constexpr int R = K / 2;
  float acca = 0.f, accb = 0.f;
  const float2 pxx = make_float2(px[K], px[K + 1]), pyy = make_float2(py[K], py[K + 1]), pzz = make_float2(pz[K], pz[K + 1]);
  for (int i = 0; i < cnt; i++)
  {
    const DecidingElem &e = sm.d[R][i];
    const float4 n = e.nxm;
    const float lenq = e.lenq;
    int slot = e.slot;
    const bool ina = (e.ma & lanebit) != 0u, inb = (e.mb & lanebit) != 0u;
    unsigned oa, ob;
    if (slot >= 0)
    { // no accepted pair can be softened, one periodic image: the bare pair kernel + the criterion
      const float2 dx = f2_add(pxx, make_float2(n.x, n.x));
      const float2 dy = f2_add(pyy, make_float2(n.y, n.y));
      const float2 dz = f2_add(pzz, make_float2(n.z, n.z));
      const float2 r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
      const float ra = rsqrt_raw(r2.x), rb = rsqrt_raw(r2.y);
      const bool opena = lenq > r2.x, openb = lenq > r2.y; // reference criterion, per target (src/gravity_tree.cpp:135)
      if (ina && !opena) acca = fmaf(n.w, ra, acca);
      if (inb && !openb) accb = fmaf(n.w, rb, accb);
      oa = __ballot_sync(kFull, ina && opena);
      ob = __ballot_sync(kFull, inb && openb);
      if (COUNT) n_acc += (unsigned)(ina && !opena) + (unsigned)(inb && !openb);
    }
    else
    {
      slot = ~slot;
      masked_exact<PERIODIC, COUNT>(n, lenq, ina, px[K], py[K], pz[K], acca, accd[K], oa, box_size, box_half, h2, softening, n_acc);
      masked_exact<PERIODIC, COUNT>(n, lenq, inb, px[K + 1], py[K + 1], pz[K + 1], accb, accd[K + 1], ob, box_size, box_half, h2, softening, n_acc);
    }
    if (lane == 0) *reinterpret_cast<uint2 *>(&sm.pending[slot].m[K]) = make_uint2(oa, ob); // the openers walk the node's children
  }
  accd[K] += (double)acca;
  accd[K + 1] += (double)accb;
}

#if HBT_M_SPLITX
// split layout of a deciding list: elements [0, cnt) take the bare pair kernel + the criterion in a branch-free loop, the
// elements that need the exact kernel were queued from the far end, [kMCap - cntx, kMCap)
template <int K, bool PERIODIC, bool COUNT, class MaskedSmem>
__device__ __forceinline__ void masked_eval_split(MaskedSmem &sm, int cnt, int cntx, int lane, unsigned lanebit, const float (&px)[4],
                                                  const float (&py)[4], const float (&pz)[4], double (&accd)[4], float box_size, float box_half,
                                                  float h2, float softening, unsigned &n_acc)
{
  constexpr int R = K / 2;
  float acca = 0.f, accb = 0.f;
  const float2 pxx = make_float2(px[K], px[K + 1]), pyy = make_float2(py[K], py[K + 1]), pzz = make_float2(pz[K], pz[K + 1]);
  HBT_M_PRAGMA_UNROLL(HBT_M_UNROLL_D)
  for (int i = 0; i < cnt; i++)
  {
    const DecidingElem &e = sm.d[R][i];
    const float4 n = e.nxm;
    const float lenq = e.lenq;
    const bool ina = (e.ma & lanebit) != 0u, inb = (e.mb & lanebit) != 0u;
    const float2 dx = f2_add(pxx, make_float2(n.x, n.x));
    const float2 dy = f2_add(pyy, make_float2(n.y, n.y));
    const float2 dz = f2_add(pzz, make_float2(n.z, n.z));
    const float2 r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
    const float ra = rsqrt_raw(r2.x), rb = rsqrt_raw(r2.y);
    const bool opena = lenq > r2.x, openb = lenq > r2.y; // reference criterion, per target (src/gravity_tree.cpp:135)
    if (ina && !opena) acca = fmaf(n.w, ra, acca);
    if (inb && !openb) accb = fmaf(n.w, rb, accb);
    const unsigned oa = __ballot_sync(kFull, ina && opena);
    const unsigned ob = __ballot_sync(kFull, inb && openb);
    if (COUNT) n_acc += (unsigned)(ina && !opena) + (unsigned)(inb && !openb);
    if (lane == 0) *reinterpret_cast<uint2 *>(&sm.pending[e.slot].m[K]) = make_uint2(oa, ob); // the openers walk the node's children
  }
  for (int i = kMCap - cntx; i < kMCap; i++)
  {
    const DecidingElem &e = sm.d[R][i];
    const float4 n = e.nxm;
    const bool ina = (e.ma & lanebit) != 0u, inb = (e.mb & lanebit) != 0u;
    unsigned oa, ob;
    masked_exact<PERIODIC, COUNT>(n, e.lenq, ina, px[K], py[K], pz[K], acca, accd[K], oa, box_size, box_half, h2, softening, n_acc);
    masked_exact<PERIODIC, COUNT>(n, e.lenq, inb, px[K + 1], py[K + 1], pz[K + 1], accb, accd[K + 1], ob, box_size, box_half, h2, softening, n_acc);
    if (lane == 0) *reinterpret_cast<uint2 *>(&sm.pending[e.slot].m[K]) = make_uint2(oa, ob);
  }
  accd[K] += (double)acca;
  accd[K + 1] += (double)accb;
}
#endif

// The walk of one group: targets px/py/pz (4 per lane: slice k = targets 32k .. 32k+31 of the group; periodic: already
// un-wrapped towards one common image; invalid slots repeat a valid position) over the pre-order nodes
// [node_begin, node_end).  accd[k] receives sum(-m/r) (softened pairs: the spline term) of target (lane, k).
// nacc: warp-uniform part of the accepted-interaction count; n_acc: per-lane part; n_vis: node-parallel iterations.
template <bool PERIODIC, bool COUNT, class MaskedSmem>
__device__ __forceinline__ bool masked_group_walk(MaskedSmem &sm, int lane, const float4 *__restrict__ node_xm, const float2 *__restrict__ node_aux,
                                                  int node_begin, int node_end, const float (&px)[4], const float (&py)[4], const float (&pz)[4],
                                                  const bool (&valid)[4], int n0, float box_size, float box_half, float softening, double (&accd)[4],
                                                  unsigned long long &nacc, unsigned &n_acc, unsigned &n_vis)
{
  constexpr int kMStack = MaskedSmem::kStack;
  const unsigned lt = (1u << lane) - 1u, lanebit = 1u << lane;
  const float h = 2.8f * softening, h2 = h * h;
  // bounding box of the group (ordered-uint REDUX), centre + inflated half widths
  {
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
      const unsigned ux = float_to_ordered(px[k]), uy = float_to_ordered(py[k]), uz = float_to_ordered(pz[k]);
      lo[0] = min(lo[0], ux); lo[1] = min(lo[1], uy); lo[2] = min(lo[2], uz);
      hi[0] = max(hi[0], ux); hi[1] = max(hi[1], uy); hi[2] = max(hi[2], uz);
    }
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
      const float l = ordered_to_float(__reduce_min_sync(kFull, lo[j]));
      const float hh = ordered_to_float(__reduce_max_sync(kFull, hi[j]));
      const float c = 0.5f * (l + hh);
      if (lane == 0)
      {
        sm.box[j] = c;
        sm.box[4 + j] = fmaxf(hh - c, c - l) * 1.00001f + 1e-30f;
      }
    }
  }
  // masks of the whole group (n0 bits)
  const unsigned vm0 = __ballot_sync(kFull, valid[0]), vm1 = __ballot_sync(kFull, valid[1]);
  const unsigned vm2 = __ballot_sync(kFull, valid[2]), vm3 = __ballot_sync(kFull, valid[3]);

  int ncs = 0;        // chains on the stack
  int na = 0, ab = 0; // dense ring: pending, base
  int nd0 = 0, nd1 = 0, np = 0; // deciding elements of slices 0,1 / 2,3; pending chains
  int na0 = 0, na1 = 0;         // accept-all elements of slices 0,1 / 2,3
#if HBT_M_SPLITX
  int nx0 = 0, nx1 = 0;         // deciding elements that need the exact kernel (queued from the far end of the lists)
#else
  constexpr int nx0 = 0, nx1 = 0;
#endif
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

  // One loop.  Every iteration a lane without a node takes a chain off the stack, every lane classifies the node it
  // stands on and moves to the node's sibling.
  int cur = 0, pend = 0;
  unsigned c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u; // mask of the lane's chain
  bool full = false;                           // ... it is the whole group
  while (true)
  {
    const unsigned mI = __ballot_sync(kFull, !(cur < pend));
    if (mI != 0u && ncs > 0)
    { // idle lanes take chains; the fuller the stack, the fewer (every walking lane can push one entry per iteration)
      const int room = kMStack - ncs;
      const int lim = max(mI == kFull ? 1 : 0, room / HBT_MASKED_ROOM_PER_LANE - (32 - __popc(mI))); // walking lanes <= room / 3
      const int t = min(min(__popc(mI), ncs), lim);
      const int r = __popc(mI & lt);
      if (!(cur < pend) && r < t)
      {
        const ChainEntry &c = sm.stack[ncs - 1 - r];
        cur = c.cur;
        pend = c.pend;
        c0 = c.m[0]; c1 = c.m[1]; c2 = c.m[2]; c3 = c.m[3];
        full = __popc(c0) + __popc(c1) + __popc(c2) + __popc(c3) == n0;
      }
      ncs -= t;
      __syncwarp();
    }
    HBT_MASKED_TRACK(ncs);
    const bool act = cur < pend;
    const bool anyact = __any_sync(kFull, act);
    if (!anyact && ncs == 0 && np == 0 && nd0 + nx0 == 0 && nd1 + nx1 == 0 && na0 == 0 && na1 == 0) break;
    int cls = 0; // 1 FAR, 2 NEAR, 3 OPEN, 4 MIXED
    float4 xm = make_float4(0.f, 0.f, 0.f, 0.f), xs = xm;
    float lenq = 0.f;
    int kend = 0;
    bool bare = false; // no accepted pair can be softened and the group sees one periodic image of the node
    if (act)
    {
      xm = __ldg(&node_xm[cur]);
      const float2 ax = __ldg(&node_aux[cur]);
      lenq = ax.x;
      kend = __float_as_int(ax.y);
      const float4 bc = *reinterpret_cast<const float4 *>(&sm.box[0]), bh = *reinterpret_cast<const float4 *>(&sm.box[4]);
      const float hx = bh.x, hy = bh.y, hz = bh.z;
      float dx = xm.x - bc.x, dy = xm.y - bc.y, dz = xm.z - bc.z;
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
      const bool far_ok = wrap_ok && r2min >= h2;
      // a cell is only accepted at r^2 >= len^2/theta^2: with len^2/theta^2 >= h^2 no accepted pair is softened
      bare = wrap_ok && (r2min >= h2 || lenq >= h2);
      if (lenq == 0.f) cls = far_ok ? 1 : 2; // a particle is accepted by everyone
      else if (!wrap_ok) cls = 4;
      else if (lenq > r2max) cls = 3;
      else if (!(lenq > r2min) && far_ok) cls = 1;
      else cls = 4;
    }
    if (COUNT) n_vis++;
    const bool toA = (cls == 1) && full, toO = (cls == 3), toM = act && !toA && !toO;
    const bool toP = toM && cls == 4; // the targets decide: the chain of the children waits for their answer
    const bool toAcc = toM && cls == 1; // FAR for the whole group, partial mask: accepted by every target of the mask
    const bool toD = toM && cls != 1;
    const bool h0 = (c0 | c1) != 0u, h1 = (c2 | c3) != 0u;
    const unsigned mA = __ballot_sync(kFull, toA), mO = __ballot_sync(kFull, toO), mP = __ballot_sync(kFull, toP);
#if HBT_M_SPLITX
    const unsigned mD0 = __ballot_sync(kFull, toD && h0 && bare), mD1 = __ballot_sync(kFull, toD && h1 && bare);
    const unsigned mX0 = __ballot_sync(kFull, toD && h0 && !bare), mX1 = __ballot_sync(kFull, toD && h1 && !bare);
#else
    const unsigned mD0 = __ballot_sync(kFull, toD && h0), mD1 = __ballot_sync(kFull, toD && h1);
#endif
    const unsigned mA0 = __ballot_sync(kFull, toAcc && h0), mA1 = __ballot_sync(kFull, toAcc && h1);
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
    int slot = kMCap; // scratch entry: particles have no children chain
    if (toP)
    {
      slot = np + __popc(mP & lt);
      ChainEntry &c = sm.pending[slot];
      c.cur = cur + 1;
      c.pend = kend;
      c.m[0] = 0u; c.m[1] = 0u; c.m[2] = 0u; c.m[3] = 0u; // filled in by the evaluation of the element
    }
    if (toAcc)
    {
      const float4 np4 = make_float4(-xs.x, -xs.y, -xs.z, -xs.w);
      if (h0)
      {
        const int idx = na0 + __popc(mA0 & lt);
        sm.a_xm[0][idx] = np4;
        sm.a_m[0][idx] = make_uint2(c0, c1);
      }
      if (h1)
      {
        const int idx = na1 + __popc(mA1 & lt);
        sm.a_xm[1][idx] = np4;
        sm.a_m[1][idx] = make_uint2(c2, c3);
      }
    }
    if (toD)
    {
      const float4 p = bare ? xs : xm;
      DecidingElem e;
      e.nxm = make_float4(-p.x, -p.y, -p.z, -p.w);
      e.lenq = lenq;
#if HBT_M_SPLITX
      e.slot = slot;
      if (h0)
      {
        e.ma = c0; e.mb = c1;
        sm.d[0][bare ? nd0 + __popc(mD0 & lt) : kMCap - 1 - nx0 - __popc(mX0 & lt)] = e;
      }
      if (h1)
      {
        e.ma = c2; e.mb = c3;
        sm.d[1][bare ? nd1 + __popc(mD1 & lt) : kMCap - 1 - nx1 - __popc(mX1 & lt)] = e;
      }
    }
    nx0 += __popc(mX0);
    nx1 += __popc(mX1);
#else
      e.slot = bare ? slot : ~slot;
      if (h0)
      {
        e.ma = c0; e.mb = c1;
        sm.d[0][nd0 + __popc(mD0 & lt)] = e;
      }
      if (h1)
      {
        e.ma = c2; e.mb = c3;
        sm.d[1][nd1 + __popc(mD1 & lt)] = e;
      }
    }
#endif
    HBT_MASKED_STAT(0, __popc(mA)); HBT_MASKED_STAT(1, __popc(mA0) + __popc(mA1)); HBT_MASKED_STAT(2, __popc(mD0) + __popc(mD1)); HBT_MASKED_STAT(3, __popc(mP)); HBT_MASKED_STAT(4, cO);
    na += __popc(mA);
    ncs += cO;
    np += __popc(mP);
    nd0 += __popc(mD0);
    nd1 += __popc(mD1);
    na0 += __popc(mA0);
    na1 += __popc(mA1);
    if (act) cur = kend;
    __syncwarp();
    if (na >= 32)
    {
      masked_eval_far(sm.alist, ab, 32, px, py, pz, accd);
      if (COUNT) nacc += 32ull * n0;
      ab = (ab + 32) & 63;
      na -= 32;
    }
    const bool idle_all = !anyact && ncs == 0; // nothing walking, nothing on the stack: flush everything
    if (na0 > kAPend || (idle_all && na0 > 0))
    {
      masked_eval_accept<0, COUNT, MaskedSmem>(sm, na0, lanebit, px, py, pz, accd, n_acc);
      na0 = 0;
    }
    if (na1 > kAPend || (idle_all && na1 > 0))
    {
      masked_eval_accept<2, COUNT, MaskedSmem>(sm, na1, lanebit, px, py, pz, accd, n_acc);
      na1 = 0;
    }
    if (nd0 + nx0 > kMPend || nd1 + nx1 > kMPend || np > kMPend || idle_all || (mI == kFull && ncs < 32))
    { // DRAIN: evaluate both deciding lists, then move the pending chains somebody opened onto the stack
#if HBT_M_SPLITX
      if (nd0 + nx0 > 0) masked_eval_split<0, PERIODIC, COUNT, MaskedSmem>(sm, nd0, nx0, lane, lanebit, px, py, pz, accd, box_size, box_half, h2, softening, n_acc);
      if (nd1 + nx1 > 0) masked_eval_split<2, PERIODIC, COUNT, MaskedSmem>(sm, nd1, nx1, lane, lanebit, px, py, pz, accd, box_size, box_half, h2, softening, n_acc);
      nx0 = 0;
      nx1 = 0;
#else
      if (nd0 > 0) masked_eval<0, PERIODIC, COUNT, MaskedSmem>(sm, nd0, lane, lanebit, px, py, pz, accd, box_size, box_half, h2, softening, n_acc);
      if (nd1 > 0) masked_eval<2, PERIODIC, COUNT, MaskedSmem>(sm, nd1, lane, lanebit, px, py, pz, accd, box_size, box_half, h2, softening, n_acc);
#endif
      nd0 = 0;
      nd1 = 0;
      __syncwarp();
      for (int b = 0; b < np; b += 32)
      {
        const int i = b + lane;
        ChainEntry c;
        c.cur = 0; c.pend = 0; c.m[0] = 0u; c.m[1] = 0u; c.m[2] = 0u; c.m[3] = 0u;
        if (i < np) c = sm.pending[i];
        const bool live = (c.m[0] | c.m[1] | c.m[2] | c.m[3]) != 0u;
        const unsigned mL = __ballot_sync(kFull, live);
        const int cL = __popc(mL);
        if (ncs + cL > kMStack) return false;
        if (live) sm.stack[ncs + __popc(mL & lt)] = c;
        ncs += cL;
      }
      np = 0;
    }
    __syncwarp();
  }
  if (na > 0)
  {
    masked_eval_far(sm.alist, ab, na, px, py, pz, accd);
    if (COUNT) nacc += (unsigned long long)na * n0;
  }
  return true;
}

} // namespace hbt
