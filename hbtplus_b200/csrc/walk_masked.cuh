// walk_masked.cuh - core of the masked group walk (sm_100a): the whole Barnes-Hut walk of a group of 64 or 128 targets
// without any per-lane tree traversal.
//
// Same decisions as every other walk of this library: each target applies the REFERENCE opening criterion
// len^2 > r^2 theta^2 to every node it meets (src/gravity_tree.cpp:135) and so accepts exactly the reference's nodes.
// A warp owns NP slice pairs (NP = 1: 64 targets, 2 per lane; NP = 2: 128 targets, 4 per lane; slice k = targets
// 32k .. 32k+31 of the group).  Every sibling chain on the per-warp stack carries an OPENER MASK - the targets that opened
// all ancestors of the chain - and
//
//   * chains are classified node-parallel against the bounding box of the whole group, one lane per chain:
//     FAR (all targets accept), OPEN (all open), NEAR (particle, maybe softened), MIXED;
//   * FAR nodes of full-mask chains go to the dense ring (the bare 8-slot pair kernel, two targets per FADD2/FFMA2);
//   * OPEN cells push their children as a new chain with the SAME mask;
//   * every other node becomes a MASKED ELEMENT (node, len^2/theta^2, mask) in the list of each slice pair (64 targets)
//     that has targets in the mask: the warp evaluates it with one packed fp32x2 pair computation, each target of the
//     mask deciding for itself; the targets that open the node form the mask of the chain of its children.  A node
//     everyone in the mask accepts is an ACCEPT-ALL element (bare pair kernel under the mask).  Elements for which a
//     softened accepted pair or a second periodic image cannot be excluded are evaluated with the reference's full
//     kernel per target (spline in double, NEAREST per target).
//
// All loops are dense and free of walk-order dependencies, partial sums are fp32 over at most a list's capacity and then
// fp64 per target in a fixed order (run-to-run reproducible).  The chain stack is bounded; running out of it makes the
// function return false and the caller redoes the group with the per-lane walk (never seen with real trees).
//
// This header has no includes on purpose: walk_masked.cu includes it after walk_common.cuh; tests/host_emul/
// masked_emul.cpp includes it after a warp-emulation shim (32 fibers) to unit-test the logic on the CPU.
#pragma once

#ifndef HBT_MASKED_ROOM_PER_LANE
#define HBT_MASKED_ROOM_PER_LANE 3 // free stack entries demanded per walking lane (a chain pushes <= 8, typically 2)
#endif
#ifndef HBT_M_UNROLL
#define HBT_M_UNROLL 4 // unroll factor of the dense / accept-all list loops (2 -> 4: 1045 -> 992 ms, profiles/r01_walk_notes.md)
#endif
#ifndef HBT_M_UNROLL_D
#define HBT_M_UNROLL_D 4 // unroll factor of the branch-free deciding loop (2 -> 4: 956 -> 926 ms, profiles/r02_walk_notes.md)
#endif
#ifndef HBT_M_PEND
#define HBT_M_PEND 8 // deciding elements / pending chains are drained when more than this many wait ...
#endif
#ifndef HBT_A_PEND
#define HBT_A_PEND 16 // accept-all elements are evaluated when more than this many wait
#endif
#ifndef HBT_M_PREFETCH
#define HBT_M_PREFETCH 0 // 1: a lane that stays on its chain loads its NEXT node right after classifying the current one.  ncu puts
                         // 8 % of the stall samples on the first use of the loaded node, but the six extra live registers cost more than
                         // the hidden latency returns: 925 -> 982 ms on the bench (profiles/r02_walk_notes.md).  Kept for the record.
#endif
#ifndef HBT_M_PRED_RSQ
#define HBT_M_PRED_RSQ 0 // 1: the masked lists issue MUFU.RSQ under the accept predicate (meant to save the XU work of empty slices and
                         // all-open nodes).  ptxas 12.9 hoists the predicated rsqrt into an unconditional MUFU.RSQ anyway and the longer
                         // predicate chains spill (28 -> 140 bytes at 72 registers): kept for the record, off
#endif
#ifndef HBT_M_POPC_SEL
#define HBT_M_POPC_SEL 0 // 1: ONE prefix popcount per lane for its (exclusive) queue instead of one per queue (POPC shares the XU pipe with
                         // MUFU.RSQ: 16 -> 10 POPC per iteration).  Measured: 904 -> 916 ms on the bench - the select chain in front of the
                         // popcount costs more latency than the XU cycles return (profiles/r02_ab_cfg2_v4.jsonl).  Off.
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

static constexpr int kMPend = HBT_M_PEND;
static constexpr int kMCap = kMPend + 32; // ... and at most 32 arrive per iteration
static constexpr int kAPend = HBT_A_PEND;
static constexpr int kACap = kAPend + 32;

// Masked elements are kept per SLICE PAIR (slices 2R, 2R+1 = targets 64R .. 64R+63 of the group): an element sits in the
// list of every pair that has targets in its mask, so its evaluation is one packed fp32x2 pair computation without any
// per-word branching.  Two kinds of lists per pair:
//   accept-all (a_*): a node that is FAR for the whole group met by a chain with a partial mask - the bare pair kernel
//                     under the mask, nothing else (72 % of the masked elements of the bench);
//   deciding (d):     each target of the mask applies the criterion (or needs the exact kernel: those are queued from the
//                     far end of the list and evaluated in their own loop).  The chain of the children waits in `pending`
//                     with empty masks while the evaluation of each pair fills in its two opener words; a DRAIN evaluates
//                     the deciding lists and moves the pending chains somebody opened onto the stack (the others are dropped).
template <int NP>
struct ChainEntryT;
template <>
struct __align__(8) ChainEntryT<2>
{
  int cur, pend; // siblings still to classify: cur, end(cur), ... < pend
  unsigned m[4]; // targets walking the chain (bit = lane, word = slice)
};
template <>
struct __align__(16) ChainEntryT<1>
{
  int cur, pend;
  unsigned m[2];
};
struct __align__(16) DecidingElem
{
  float4 nxm;      // -x, -y, -z, -m: operands of the packed adds and of the accumulate
  float lenq;      // len^2/theta^2 (0 for a particle: every target of the mask accepts)
  int pad;
  unsigned ma, mb; // in: targets taking part (bit = lane; first / second slice of the pair); out: the targets that OPENED the
                   // node (written by the evaluation; the drain copies them into the pending chain of the node's children)
};
template <int STACK, int NP = 2> // chain entries per warp, slice pairs per warp
struct MaskedSmemT
{
  static constexpr int kStack = STACK;
  static constexpr int kPairs = NP;
  float4 alist[64];           // ring of FAR nodes of whole-group chains (periodic: shifted to the group's image)
  float box[8];               // centre [0..2] and inflated half widths [4..6] of the group's bounding box (read as two float4)
  DecidingElem d[NP][kMCap];  // deciding elements per pair
  float4 a_xm[NP][kACap];     // accept-all elements per pair: -x, -y, -z, -m
  uint2 a_m[NP][kACap];       //   their masks
  ChainEntryT<NP> pending[kMCap];
  ChainEntryT<NP> stack[STACK];
};

__device__ __forceinline__ double spline_wp(float r2, double hinv_d)
{ // Gadget spline kernel in double (src/gravity_tree.cpp:146-160)
  if (r2 == 0.f) return -2.8; // the self term (and co-located pairs): no square root (DSQRT takes its slow path for 0)
  const double u = sqrt((double)r2) * hinv_d;
  if (u < 0.5) return -2.8 + u * u * (5.333333333333 + u * u * (6.4 * u - 9.6));
  return -3.2 + 0.066666666667 / u + u * u * (10.666666666667 + u * (-16.0 + u * (9.6 - 2.133333333333 * u)));
}

// dense evaluation of `cnt` FAR nodes for all targets of the group: 7 packed + 2 MUFU.RSQ per two interactions
template <int T>
__device__ __forceinline__ void masked_eval_far(const float4 *__restrict__ ring, int base, int cnt, const float (&px)[T], const float (&py)[T],
                                                const float (&pz)[T], double (&accd)[T])
{
  float2 accf[T / 2];
#pragma unroll
  for (int k = 0; k < T / 2; k++) accf[k] = make_float2(0.f, 0.f);
  HBT_M_PRAGMA_UNROLL(HBT_M_UNROLL)
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

// both slices of an element that needs the reference's full kernel per target (src/gravity_tree.cpp:141-161): NEAREST per
// target in periodic runs, the spline in double for the pairs inside the softening.  The fp64 sequence is issued once for the
// lanes with a softened pair in either slice, and a second time only if some lane has one in both.
template <bool PERIODIC, bool COUNT>
__device__ __forceinline__ void masked_exact_pair(const float4 &n, float lenq, bool ina, bool inb, float pxa, float pya, float pza, float pxb, float pyb,
                                                  float pzb, float &accfa, float &accfb, double &accda, double &accdb, unsigned &oa, unsigned &ob,
                                                  float box_size, float box_half, float h2, float softening, unsigned &n_acc)
{
  float dxa = pxa + n.x, dya = pya + n.y, dza = pza + n.z;
  float dxb = pxb + n.x, dyb = pyb + n.y, dzb = pzb + n.z;
  if (PERIODIC)
  {
    dxa = nearest_f(dxa, box_size, box_half); dya = nearest_f(dya, box_size, box_half); dza = nearest_f(dza, box_size, box_half);
    dxb = nearest_f(dxb, box_size, box_half); dyb = nearest_f(dyb, box_size, box_half); dzb = nearest_f(dzb, box_size, box_half);
  }
  const float r2a = fmaf(dza, dza, fmaf(dya, dya, dxa * dxa)); // FMUL, FFMA, FFMA like the packed path
  const float r2b = fmaf(dzb, dzb, fmaf(dyb, dyb, dxb * dxb));
  const bool opena = lenq > r2a, openb = lenq > r2b; // reference criterion, per target (src/gravity_tree.cpp:135)
  const bool acca = ina && !opena, accb = inb && !openb;
  const bool softa = acca && r2a < h2, softb = accb && r2b < h2;
  if (acca && !softa) accfa = fmaf(n.w, rsqrt_raw(r2a), accfa);
  if (accb && !softb) accfb = fmaf(n.w, rsqrt_raw(r2b), accfb);
  if (__any_sync(kFull, softa || softb))
  {
    const double hinv_d = 1.0 / (2.8 * (double)softening);
    if (softa || softb)
    {
      const double t = (double)(-n.w) * hinv_d * spline_wp(softa ? r2a : r2b, hinv_d);
      if (softa) accda += t; else accdb += t;
    }
    if (__any_sync(kFull, softa && softb))
      if (softa && softb) accdb += (double)(-n.w) * hinv_d * spline_wp(r2b, hinv_d);
  }
  oa = __ballot_sync(kFull, ina && opena);
  ob = __ballot_sync(kFull, inb && openb);
  if (COUNT) n_acc += (unsigned)acca + (unsigned)accb;
}

// evaluate the `cnt` accept-all elements of slice pair R (slices K = 2R, K+1): the bare pair kernel under the mask
template <int K, bool COUNT, int T, class MaskedSmem>
__device__ __forceinline__ void masked_eval_accept(const MaskedSmem &sm, int cnt, unsigned lanebit, const float (&px)[T], const float (&py)[T],
                                                   const float (&pz)[T], double (&accd)[T], unsigned &n_acc)
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
#if HBT_M_PRED_RSQ
    accept_half_rsq<COUNT>(m.x, lanebit, r2.x, n.w, acca, n_acc);
    accept_half_rsq<COUNT>(m.y, lanebit, r2.y, n.w, accb, n_acc);
#else
    const float ra = rsqrt_raw(r2.x), rb = rsqrt_raw(r2.y);
    const bool ina = (m.x & lanebit) != 0u, inb = (m.y & lanebit) != 0u;
    if (ina) acca = fmaf(n.w, ra, acca);
    if (inb) accb = fmaf(n.w, rb, accb);
    if (COUNT) n_acc += (unsigned)ina + (unsigned)inb;
#endif
  }
  accd[K] += (double)acca;
  accd[K + 1] += (double)accb;
}

// evaluate the deciding list of slice pair R (slices K = 2R, K+1): elements [0, cnt) take the bare pair kernel + the
// criterion in a branch-free loop, the elements that need the exact kernel were queued from the far end,
// [kMCap - cntx, kMCap).  The targets that open an element replace its masks in place (one 8-byte store at a fixed offset);
// the drain hands them to the pending chain of the node's children.
template <int K, bool PERIODIC, bool COUNT, int T, class MaskedSmem>
__device__ __forceinline__ void masked_eval_split(MaskedSmem &sm, int cnt, int cntx, int lane, unsigned lanebit, const float (&px)[T],
                                                  const float (&py)[T], const float (&pz)[T], double (&accd)[T], float box_size, float box_half,
                                                  float h2, float softening, unsigned &n_acc)
{
  constexpr int R = K / 2;
  float acca = 0.f, accb = 0.f;
  const float2 pxx = make_float2(px[K], px[K + 1]), pyy = make_float2(py[K], py[K + 1]), pzz = make_float2(pz[K], pz[K + 1]);
  HBT_M_PRAGMA_UNROLL(HBT_M_UNROLL_D)
  for (int i = 0; i < cnt; i++)
  {
    DecidingElem &e = sm.d[R][i];
    const float4 n = e.nxm;
    const float lenq = e.lenq;
    const uint2 m = *reinterpret_cast<const uint2 *>(&e.ma);
    const float2 dx = f2_add(pxx, make_float2(n.x, n.x));
    const float2 dy = f2_add(pyy, make_float2(n.y, n.y));
    const float2 dz = f2_add(pzz, make_float2(n.z, n.z));
    const float2 r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
    // reference criterion, per target (src/gravity_tree.cpp:135): in the mask and lenq > r2 -> opener, else accumulate
#if HBT_M_PRED_RSQ
    const unsigned oa = decide_half_rsq<COUNT>(m.x, lanebit, lenq, r2.x, n.w, acca, n_acc);
    const unsigned ob = decide_half_rsq<COUNT>(m.y, lanebit, lenq, r2.y, n.w, accb, n_acc);
#else
    const float ra = rsqrt_raw(r2.x), rb = rsqrt_raw(r2.y);
    const unsigned oa = decide_half<COUNT>(m.x, lanebit, lenq, r2.x, n.w, ra, acca, n_acc);
    const unsigned ob = decide_half<COUNT>(m.y, lanebit, lenq, r2.y, n.w, rb, accb, n_acc);
#endif
    if (lane == 0) *reinterpret_cast<uint2 *>(&e.ma) = make_uint2(oa, ob); // the openers walk the node's children
  }
  for (int i = kMCap - cntx; i < kMCap; i++)
  {
    DecidingElem &e = sm.d[R][i];
    const float4 n = e.nxm;
    const bool ina = (e.ma & lanebit) != 0u, inb = (e.mb & lanebit) != 0u;
    unsigned oa, ob;
    masked_exact_pair<PERIODIC, COUNT>(n, e.lenq, ina, inb, px[K], py[K], pz[K], px[K + 1], py[K + 1], pz[K + 1], acca, accb, accd[K], accd[K + 1], oa, ob,
                                       box_size, box_half, h2, softening, n_acc);
    if (lane == 0) *reinterpret_cast<uint2 *>(&e.ma) = make_uint2(oa, ob);
  }
  accd[K] += (double)acca;
  accd[K + 1] += (double)accb;
}

// The walk of one group: targets px/py/pz (T = 2 NP per lane: slice k = targets 32k .. 32k+31 of the group; periodic:
// already un-wrapped towards one common image; invalid slots repeat a valid position) over the pre-order nodes
// [node_begin, node_end).  accd[k] receives sum(-m/r) (softened pairs: the spline term) of target (lane, k).
// nacc: warp-uniform part of the accepted-interaction count; n_acc: per-lane part; n_vis: node-parallel iterations.
template <bool PERIODIC, bool COUNT, class MaskedSmem, int T>
__device__ __forceinline__ bool masked_group_walk(MaskedSmem &sm, int lane, const float4 *__restrict__ node_xm, const float2 *__restrict__ node_aux,
                                                  int node_begin, int node_end, const float (&px)[T], const float (&py)[T], const float (&pz)[T],
                                                  const bool (&valid)[T], int n0, float box_size, float box_half, float softening, double (&accd)[T],
                                                  unsigned long long &nacc, unsigned &n_acc, unsigned &n_vis)
{
  constexpr int kMStack = MaskedSmem::kStack;
  constexpr int NP = MaskedSmem::kPairs;
  static_assert(T == 2 * NP, "two slices per pair");
  typedef ChainEntryT<NP> ChainEntry;
  const unsigned lt = (1u << lane) - 1u, lanebit = 1u << lane;
  const float h = 2.8f * softening, h2 = h * h;
  // bounding box of the group (ordered-uint REDUX), centre + inflated half widths
  {
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < T; k++)
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
  unsigned vm[T];
#pragma unroll
  for (int k = 0; k < T; k++) vm[k] = __ballot_sync(kFull, valid[k]);

  int ncs = 0;        // chains on the stack
  int na = 0, ab = 0; // dense ring: pending, base
  int np = 0;         // pending chains
  int nd[NP], nx[NP], nac[NP]; // per pair: deciding elements (bare), deciding elements that need the exact kernel, accept-all elements
#pragma unroll
  for (int r = 0; r < NP; r++) nd[r] = nx[r] = nac[r] = 0;
  if (node_end > node_begin)
  {
    if (lane == 0)
    {
      ChainEntry &c = sm.stack[0];
      c.cur = node_begin;
      c.pend = node_end;
#pragma unroll
      for (int k = 0; k < T; k++) c.m[k] = vm[k];
    }
    ncs = 1;
  }
  __syncwarp();

  // One loop.  Every iteration a lane without a node takes a chain off the stack, every lane classifies the node it
  // stands on and moves to the node's sibling.
  int cur = 0, pend = 0;
  unsigned cm[T]; // mask of the lane's chain
#pragma unroll
  for (int k = 0; k < T; k++) cm[k] = 0u;
  bool full = false; // ... it is the whole group
#if HBT_M_PREFETCH
  float4 pre_xm = make_float4(0.f, 0.f, 0.f, 0.f); // the node the lane will stand on next, loaded ahead
  float2 pre_ax = make_float2(0.f, 0.f);
  bool pre = false;
#endif
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
        full = true; // chain masks are subsets of the group's: the whole group walks the chain iff they are equal
#pragma unroll
        for (int k = 0; k < T; k++)
        {
          cm[k] = c.m[k];
          full = full && cm[k] == vm[k];
        }
#if HBT_M_PREFETCH
        pre = false;
#endif
      }
      ncs -= t;
      __syncwarp();
    }
    HBT_MASKED_TRACK(ncs);
    const bool act = cur < pend;
    const bool anyact = __any_sync(kFull, act);
    bool lists_empty = np == 0;
#pragma unroll
    for (int r = 0; r < NP; r++) lists_empty = lists_empty && (nd[r] + nx[r] == 0) && (nac[r] == 0);
    if (!anyact && ncs == 0 && lists_empty) break;
    int cls = 0; // 1 FAR, 2 NEAR, 3 OPEN, 4 MIXED
    float4 xm = make_float4(0.f, 0.f, 0.f, 0.f), xs = xm;
    float lenq = 0.f;
    int kend = 0;
    bool bare = false; // no accepted pair can be softened and the group sees one periodic image of the node
    if (act)
    {
#if HBT_M_PREFETCH
      float2 ax;
      if (pre) { xm = pre_xm; ax = pre_ax; }
      else { xm = __ldg(&node_xm[cur]); ax = __ldg(&node_aux[cur]); }
#else
      xm = __ldg(&node_xm[cur]);
      const float2 ax = __ldg(&node_aux[cur]);
#endif
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
      const float nx_ = fmaxf(adx - hx, 0.f), ny_ = fmaxf(ady - hy, 0.f), nz_ = fmaxf(adz - hz, 0.f);
      const float fx = adx + hx, fy = ady + hy, fz = adz + hz;
      const float r2min = (nx_ * nx_ + ny_ * ny_ + nz_ * nz_) * 0.99998f;
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
    const bool toP = toM && cls == 4;   // the targets decide: the chain of the children waits for their answer
    const bool toAcc = toM && cls == 1; // FAR for the whole group, partial mask: accepted by every target of the mask
    const bool toD = toM && cls != 1;
    bool hp[NP]; // the lane's chain has targets in pair r
#pragma unroll
    for (int r = 0; r < NP; r++) hp[r] = (cm[2 * r] | cm[2 * r + 1]) != 0u;
    const unsigned mA = __ballot_sync(kFull, toA), mO = __ballot_sync(kFull, toO), mP = __ballot_sync(kFull, toP);
    unsigned mD[NP], mX[NP], mAc[NP];
#pragma unroll
    for (int r = 0; r < NP; r++)
    {
      mD[r] = __ballot_sync(kFull, toD && hp[r] && bare);
      mX[r] = __ballot_sync(kFull, toD && hp[r] && !bare);
      mAc[r] = __ballot_sync(kFull, toAcc && hp[r]);
    }
    const int cO = __popc(mO);
    if (ncs + cO > kMStack) return false; // stack exhausted (pathologically deep tree): the caller redoes the group per lane
#if HBT_M_POPC_SEL
    // a lane queues its node in exactly one of: dense ring, stack, and per pair accept-all / deciding (bare) / deciding (exact);
    // one pair per warp: ONE prefix popcount over the ballot of the lane's own queue gives its position there
    int psel = 0;
    if constexpr (NP == 1) psel = __popc((toA ? mA : toO ? mO : toAcc ? mAc[0] : toD ? (bare ? mD[0] : mX[0]) : 0u) & lt);
    const int posA = NP == 1 ? psel : __popc(mA & lt), posO = NP == 1 ? psel : __popc(mO & lt);
#else
    const int posA = __popc(mA & lt), posO = __popc(mO & lt);
#endif
    if (toA) sm.alist[(ab + na + posA) & 63] = xs;
    if (toO)
    {
      ChainEntry &c = sm.stack[ncs + posO];
      c.cur = cur + 1;
      c.pend = kend;
#pragma unroll
      for (int k = 0; k < T; k++) c.m[k] = cm[k];
    }
    int eidx[NP]; // position of the lane's deciding element in the list of pair r
#pragma unroll
    for (int r = 0; r < NP; r++)
    {
#if HBT_M_POPC_SEL
      if constexpr (NP == 1) eidx[r] = bare ? nd[r] + psel : kMCap - 1 - nx[r] - psel;
      else
#endif
        eidx[r] = bare ? nd[r] + __popc(mD[r] & lt) : kMCap - 1 - nx[r] - __popc(mX[r] & lt);
    }
    if (toP)
    { // the chain of the children waits for the openers: it remembers where its element sits in every pair's list
      ChainEntry &c = sm.pending[np + __popc(mP & lt)];
      c.cur = cur + 1;
      c.pend = kend;
#pragma unroll
      for (int r = 0; r < NP; r++)
      {
        c.m[2 * r] = hp[r] ? (unsigned)eidx[r] : 0xffffffffu;
        c.m[2 * r + 1] = 0u;
      }
    }
    if (toAcc)
    {
      const float4 np4 = make_float4(-xs.x, -xs.y, -xs.z, -xs.w);
#pragma unroll
      for (int r = 0; r < NP; r++)
        if (hp[r])
        {
#if HBT_M_POPC_SEL
          const int idx = nac[r] + (NP == 1 ? psel : __popc(mAc[r] & lt));
#else
          const int idx = nac[r] + __popc(mAc[r] & lt);
#endif
          sm.a_xm[r][idx] = np4;
          sm.a_m[r][idx] = make_uint2(cm[2 * r], cm[2 * r + 1]);
        }
    }
    if (toD)
    {
      const float4 p = bare ? xs : xm;
      DecidingElem e;
      e.nxm = make_float4(-p.x, -p.y, -p.z, -p.w);
      e.lenq = lenq;
      e.pad = 0;
#pragma unroll
      for (int r = 0; r < NP; r++)
        if (hp[r])
        {
          e.ma = cm[2 * r];
          e.mb = cm[2 * r + 1];
          sm.d[r][eidx[r]] = e;
        }
    }
#ifdef HBT_MASKED_STAT_ON
    {
      int sa = 0, sd = 0;
      for (int r = 0; r < NP; r++) { sa += __popc(mAc[r]); sd += __popc(mD[r]); }
      HBT_MASKED_STAT(0, __popc(mA)); HBT_MASKED_STAT(1, sa); HBT_MASKED_STAT(2, sd); HBT_MASKED_STAT(3, __popc(mP)); HBT_MASKED_STAT(4, cO);
      // pair-elements with targets in only ONE slice of the pair, and the lanes (of 64 lane-slots) inside the masks
      for (int r = 0; r < NP; r++)
      {
        const bool single = hp[r] && (cm[2 * r] == 0u || cm[2 * r + 1] == 0u);
        const unsigned sA = __ballot_sync(kFull, toAcc && single), sD = __ballot_sync(kFull, toD && bare && single);
        HBT_MASKED_STAT(5, __popc(sA)); HBT_MASKED_STAT(6, __popc(sD));
        int bitsA = 0, bitsD = 0;
        for (int l = 0; l < 32; l++)
        {
          const unsigned b0 = __shfl_sync(kFull, cm[2 * r], l), b1 = __shfl_sync(kFull, cm[2 * r + 1], l);
          if ((mAc[r] >> l) & 1u) bitsA += __popc(b0) + __popc(b1);
          if ((mD[r] >> l) & 1u) bitsD += __popc(b0) + __popc(b1);
        }
        HBT_MASKED_STAT(7, bitsA); HBT_MASKED_STAT(8, bitsD);
      }
    }
#endif
    na += __popc(mA);
    ncs += cO;
    np += __popc(mP);
#pragma unroll
    for (int r = 0; r < NP; r++)
    {
      nd[r] += __popc(mD[r]);
      nx[r] += __popc(mX[r]);
      nac[r] += __popc(mAc[r]);
    }
    if (act) cur = kend;
#if HBT_M_PREFETCH
    pre = act && cur < pend;
    if (pre)
    { // the sibling the lane moves to: its loads overlap the list evaluations below and the next iteration's bookkeeping
      pre_xm = __ldg(&node_xm[cur]);
      pre_ax = __ldg(&node_aux[cur]);
    }
#endif
    __syncwarp();
    if (na >= 32)
    {
      masked_eval_far<T>(sm.alist, ab, 32, px, py, pz, accd);
      if (COUNT) nacc += 32ull * n0;
      ab = (ab + 32) & 63;
      na -= 32;
    }
    const bool idle_all = !anyact && ncs == 0; // nothing walking, nothing on the stack: flush everything
    if (nac[0] > kAPend || (idle_all && nac[0] > 0))
    {
      masked_eval_accept<0, COUNT, T, MaskedSmem>(sm, nac[0], lanebit, px, py, pz, accd, n_acc);
      nac[0] = 0;
    }
    if constexpr (NP == 2)
    {
      if (nac[1] > kAPend || (idle_all && nac[1] > 0))
      {
        masked_eval_accept<2, COUNT, T, MaskedSmem>(sm, nac[1], lanebit, px, py, pz, accd, n_acc);
        nac[1] = 0;
      }
    }
    bool drain = np > kMPend || idle_all || (mI == kFull && ncs < 32);
#pragma unroll
    for (int r = 0; r < NP; r++) drain = drain || (nd[r] + nx[r] > kMPend);
    if (drain)
    { // DRAIN: evaluate the deciding lists, then move the pending chains somebody opened onto the stack
      if (nd[0] + nx[0] > 0)
        masked_eval_split<0, PERIODIC, COUNT, T, MaskedSmem>(sm, nd[0], nx[0], lane, lanebit, px, py, pz, accd, box_size, box_half, h2, softening, n_acc);
      if constexpr (NP == 2)
      {
        if (nd[1] + nx[1] > 0)
          masked_eval_split<2, PERIODIC, COUNT, T, MaskedSmem>(sm, nd[1], nx[1], lane, lanebit, px, py, pz, accd, box_size, box_half, h2, softening, n_acc);
      }
#pragma unroll
      for (int r = 0; r < NP; r++) nd[r] = nx[r] = 0;
      __syncwarp();
      for (int b = 0; b < np; b += 32)
      {
        const int i = b + lane;
        ChainEntry c;
        c.cur = 0;
        c.pend = 0;
#pragma unroll
        for (int k = 0; k < T; k++) c.m[k] = 0u;
        if (i < np)
        {
          c = sm.pending[i];
#pragma unroll
          for (int r = 0; r < NP; r++)
          { // the openers of the chain's element in pair r (none if the chain had no targets there)
            const unsigned idx = c.m[2 * r];
            uint2 o = make_uint2(0u, 0u);
            if (idx != 0xffffffffu) o = *reinterpret_cast<const uint2 *>(&sm.d[r][idx].ma);
            c.m[2 * r] = o.x;
            c.m[2 * r + 1] = o.y;
          }
        }
        unsigned any = 0u;
#pragma unroll
        for (int k = 0; k < T; k++) any |= c.m[k];
        const bool live = any != 0u;
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
    masked_eval_far<T>(sm.alist, ab, na, px, py, pz, accd);
    if (COUNT) nacc += (unsigned long long)na * n0;
  }
  return true;
}

} // namespace hbt
