// walk_common.cuh - helpers shared by the walk kernels (walk.cu: per-lane pre-order walk, walk_masked.cu: masked group walk).
#pragma once
#include "device_tree.cuh"

namespace hbt
{

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

// Packed fp32x2 arithmetic (sm_100a: FADD2 / FMUL2 / FFMA2, one issue slot for two targets).  Each half is an IEEE
// round-to-nearest operation, so results are bit-identical to the scalar FADD / FMUL / FFMA sequence.
__device__ __forceinline__ float2 f2_add(float2 a, float2 b)
{
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
  return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b)
{
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
  return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c)
{
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)), "l"(*reinterpret_cast<unsigned long long *>(&c)));
  return *reinterpret_cast<float2 *>(&d);
}

// One half of a deciding pair-element of the masked walk: the target (lane, slice) takes part iff its bit is in `mask`; it OPENS
// the node iff lenq > r2 (reference criterion, src/gravity_tree.cpp:135), otherwise it accepts it: acc += w * rinv.  Returns the
// ballot of the opening lanes.  Inline PTX so that the accept is ONE predicated FFMA fed by the same two compares that feed the
// vote (the compiler's own sequence: FFMA + FSEL + two more FSETP).  r2 and lenq are finite, so !(lenq > r2) == (lenq <= r2).
template <bool COUNT>
__device__ __forceinline__ unsigned decide_half(unsigned mask, unsigned lanebit, float lenq, float r2, float w, float rinv, float &acc,
                                                unsigned &n_acc)
{
  unsigned openers;
  if (COUNT)
  {
    unsigned accepted;
    asm volatile("{\n\t.reg .pred pin, pop, pac;\n\t.reg .b32 t;\n\t"
                 "and.b32 t, %3, %4;\n\tsetp.ne.u32 pin, t, 0;\n\t"
                 "setp.gt.and.f32 pop, %5, %6, pin;\n\tsetp.le.and.f32 pac, %5, %6, pin;\n\t"
                 "@pac fma.rn.f32 %0, %7, %8, %0;\n\tselp.u32 %2, 1, 0, pac;\n\t"
                 "vote.sync.ballot.b32 %1, pop, 0xffffffff;\n\t}"
                 : "+f"(acc), "=r"(openers), "=r"(accepted)
                 : "r"(mask), "r"(lanebit), "f"(lenq), "f"(r2), "f"(w), "f"(rinv));
    n_acc += accepted;
  }
  else
    asm volatile("{\n\t.reg .pred pin, pop, pac;\n\t.reg .b32 t;\n\t"
                 "and.b32 t, %2, %3;\n\tsetp.ne.u32 pin, t, 0;\n\t"
                 "setp.gt.and.f32 pop, %4, %5, pin;\n\tsetp.le.and.f32 pac, %4, %5, pin;\n\t"
                 "@pac fma.rn.f32 %0, %6, %7, %0;\n\t"
                 "vote.sync.ballot.b32 %1, pop, 0xffffffff;\n\t}"
                 : "+f"(acc), "=r"(openers)
                 : "r"(mask), "r"(lanebit), "f"(lenq), "f"(r2), "f"(w), "f"(rinv));
  return openers;
}

// The same with the reciprocal square root INSIDE the predicate: only the lanes that accept need it, and a slice whose mask is
// empty (half of the accept-all pair-elements have targets in one slice only) or whose targets all open the node issues no
// MUFU.RSQ work at all - the XU pipe co-limits the masked walk (profiles/r02_walk_notes.md).
template <bool COUNT>
__device__ __forceinline__ unsigned decide_half_rsq(unsigned mask, unsigned lanebit, float lenq, float r2, float w, float &acc, unsigned &n_acc)
{
  unsigned openers;
  if (COUNT)
  {
    unsigned accepted;
    asm volatile("{\n\t.reg .pred pin, pop, pac;\n\t.reg .b32 t;\n\t.reg .f32 ri;\n\t"
                 "and.b32 t, %3, %4;\n\tsetp.ne.u32 pin, t, 0;\n\t"
                 "setp.gt.and.f32 pop, %5, %6, pin;\n\tsetp.le.and.f32 pac, %5, %6, pin;\n\t"
                 "@pac rsqrt.approx.ftz.f32 ri, %6;\n\t@pac fma.rn.f32 %0, %7, ri, %0;\n\tselp.u32 %2, 1, 0, pac;\n\t"
                 "vote.sync.ballot.b32 %1, pop, 0xffffffff;\n\t}"
                 : "+f"(acc), "=r"(openers), "=r"(accepted)
                 : "r"(mask), "r"(lanebit), "f"(lenq), "f"(r2), "f"(w));
    n_acc += accepted;
  }
  else
    asm volatile("{\n\t.reg .pred pin, pop, pac;\n\t.reg .b32 t;\n\t.reg .f32 ri;\n\t"
                 "and.b32 t, %2, %3;\n\tsetp.ne.u32 pin, t, 0;\n\t"
                 "setp.gt.and.f32 pop, %4, %5, pin;\n\tsetp.le.and.f32 pac, %4, %5, pin;\n\t"
                 "@pac rsqrt.approx.ftz.f32 ri, %5;\n\t@pac fma.rn.f32 %0, %6, ri, %0;\n\t"
                 "vote.sync.ballot.b32 %1, pop, 0xffffffff;\n\t}"
                 : "+f"(acc), "=r"(openers)
                 : "r"(mask), "r"(lanebit), "f"(lenq), "f"(r2), "f"(w));
  return openers;
}
// accept-all element, one slice: acc += w / sqrt(r2) for the lanes of the mask, MUFU.RSQ and FFMA under the same predicate
template <bool COUNT>
__device__ __forceinline__ void accept_half_rsq(unsigned mask, unsigned lanebit, float r2, float w, float &acc, unsigned &n_acc)
{
  if (COUNT)
  {
    unsigned accepted;
    asm volatile("{\n\t.reg .pred pin;\n\t.reg .b32 t;\n\t.reg .f32 ri;\n\t"
                 "and.b32 t, %2, %3;\n\tsetp.ne.u32 pin, t, 0;\n\t"
                 "@pin rsqrt.approx.ftz.f32 ri, %4;\n\t@pin fma.rn.f32 %0, %5, ri, %0;\n\tselp.u32 %1, 1, 0, pin;\n\t}"
                 : "+f"(acc), "=r"(accepted)
                 : "r"(mask), "r"(lanebit), "f"(r2), "f"(w));
    n_acc += accepted;
  }
  else
    asm volatile("{\n\t.reg .pred pin;\n\t.reg .b32 t;\n\t.reg .f32 ri;\n\t"
                 "and.b32 t, %1, %2;\n\tsetp.ne.u32 pin, t, 0;\n\t"
                 "@pin rsqrt.approx.ftz.f32 ri, %3;\n\t@pin fma.rn.f32 %0, %4, ri, %0;\n\t}"
                 : "+f"(acc)
                 : "r"(mask), "r"(lanebit), "f"(r2), "f"(w));
}

// squared distances of node n to the T targets of a lane: (dx*dx + dy*dy) + dz*dz as FMUL, FFMA, FFMA - two targets per
// instruction where the layout allows it (non-periodic, even T; the sign of dx is irrelevant)
template <int T, bool PERIODIC>
__device__ __forceinline__ void pair_r2(const float4 &n, const float (&px)[T], const float (&py)[T], const float (&pz)[T], const DevConfig &cfg,
                                        float (&r2)[T])
{
  if constexpr (!PERIODIC && (T % 2 == 0))
  {
    const float2 nx = make_float2(-n.x, -n.x), ny = make_float2(-n.y, -n.y), nz = make_float2(-n.z, -n.z);
#pragma unroll
    for (int k = 0; k < T; k += 2)
    {
      const float2 dx = f2_add(make_float2(px[k], px[k + 1]), nx);
      const float2 dy = f2_add(make_float2(py[k], py[k + 1]), ny);
      const float2 dz = f2_add(make_float2(pz[k], pz[k + 1]), nz);
      const float2 r = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
      r2[k] = r.x;
      r2[k + 1] = r.y;
    }
  }
  else
  {
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
    }
  }
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
    float r2s[T];
    pair_r2<T, PERIODIC>(n, px, py, pz, cfg, r2s);
#pragma unroll
    for (int k = 0; k < T; k++)
    {
      const float r2 = r2s[k];
      const bool active = no >= skip[k];
      const bool acc = active && !(lenq > r2); // reference criterion, per target (src/gravity_tree.cpp:135); else the node is opened
      const float rinv = rsqrt_raw(r2);
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
        accf[k] = fmaf(-n.w, rinv, accf[k]);
      if (acc)
      {
        skip[k] = nend; // resume after this subtree (a particle's end is no+1)
        if (COUNT) n_acc++;
      }
    }
    if (COUNT) n_vis++;
    if (!CAREFUL)
    { // closest pair of the step whether accepted or not (FMNMX3): a conservative trigger for the exact redo of the tile
      float m = r2s[0];
#pragma unroll
      for (int k = 1; k < T; k++) m = fminf(m, r2s[k]);
      minr2 = fminf(minr2, m);
    }
    // A target that opened this node still has skip <= no; one that accepted it resumes at nend; the others resume at or
    // after nend.  The warp goes on at the first node some target still needs (one VIMNMX3 tree + one CREDUX.MIN instead
    // of T compares + a vote, and runs of nodes nobody needs are jumped over in one step).
    int ms = skip[0];
#pragma unroll
    for (int k = 1; k < T; k++) ms = min(ms, skip[k]);
    no = max(no + 1, __reduce_min_sync(kFull, ms));
  } while (no < tile_lim);
  return no;
}

// Per-lane pre-order walk of the node range [first, pend): 32 nodes at a time are staged in the warp's shared-memory
// tile by one coalesced 16 B + 8 B load per lane (the node arrays are padded by 32 entries); a tile is walked with the
// fast pass and redone exactly if some lane met a softened pair.  Target k of a lane takes part from node skip[k] on.
template <int T, bool PERIODIC, bool COUNT, bool PRELOADED = false>
__device__ __forceinline__ void walk_range(const float4 *__restrict__ node_xm, const float2 *__restrict__ node_aux, TileNode *tile, int first,
                                           int pend, const float (&px)[T], const float (&py)[T], const float (&pz)[T], int (&skip)[T],
                                           double (&accd)[T], const DevConfig &cfg, float h2, float hinv, unsigned &n_acc, unsigned &n_vis,
                                           float4 xm0 = float4(), float2 ax0 = float2())
{ // PRELOADED: the caller already loaded node first+lane into (xm0, ax0) (prefetched while it did other work)
  const int lane = threadIdx.x & 31;
  int no = first;
  bool pre = PRELOADED;
  while (no < pend)
  {
    const int tile_base = no;
    const int tile_lim = min(no + 32, pend);
    {
      float4 xm;
      float2 ax;
      if (pre) { xm = xm0; ax = ax0; pre = false; }
      else { xm = __ldg(&node_xm[no + lane]); ax = __ldg(&node_aux[no + lane]); }
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
      float e = (float)((double)dot3_rn(dv, dv) * 0.5 + pot);
      if (cfg.thermal_energy) e = __fadd_rn(e, v4.w); // UNBIND_WITH_THERMAL_ENERGY: E += InternalEnergy (src/subhalo_unbind.cpp:351-353)
      if (a.E_stage) a.E_stage[t] = e; else a.E[slot] = e;
    }
    else
    { // E += VecDot(OldVel, RefVelDiff) + dK - pot_removed  (src/subhalo_unbind.cpp:326-327)
      float ov[3];
      relative_velocity(x, v, st.old_ref_pos, st.old_ref_vel, cfg, ov);
      float s = __fadd_rn(dot3_rn(ov, st.ref_diff), st.dK);
      const float e = (float)((double)a.E[slot] + ((double)s - pot));
      if (a.E_stage) a.E_stage[t] = e; else a.E[slot] = e;
    }
  }
}


// find the segment of a walk warp: largest s with warp_off[s] <= warp (segments of another class have no warps)
__device__ __forceinline__ int segment_of_warp(const int *__restrict__ warp_off, int nseg, int warp)
{
  int lo = 0, hi = nseg;
  while (hi - lo > 1)
  {
    int mid = (lo + hi) >> 1;
    if (warp_off[mid] <= warp) lo = mid; else hi = mid;
  }
  return lo;
}

// walk_small.cu
void launch_walk_small(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls);
// walk_masked.cu
void launch_walk_masked(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls);

} // namespace hbt
