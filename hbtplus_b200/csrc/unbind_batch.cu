// unbind_batch.cu - the batched, level-synchronous unbinding driver and its bookkeeping kernels (sm_100a).
//
// Replaces SubhaloSnapshot_t::RefineParticles / Subhalo_t::RecursiveUnbind / Subhalo_t::Unbind /
// Subhalo_t::TruncateSource (src/subhalo_unbind.cpp:263-516).  The reference iterates one subhalo at a
// time on one OpenMP thread; here ALL subhaloes of one nesting level iterate together, one "round" per
// potential evaluation:
//
//   plan (host)  -> gather sources -> bbox -> build_trees (tree_build.cu) -> walk (walk.cu)
//                -> count E<0 -> state1 (disruption / convergence / CorrectionLoop, :357-395)
//                -> ONE radix sort that is at once the bound/unbound partition (:21-58), the E-sort of the
//                   freshly removed tail (:382) and, for converged subhaloes, the final E-sort of the bound
//                   part (:405)
//                -> mass-weighted frame reductions in fp64 (:108-188) -> state2
//                -> kinematics of converged subhaloes (:189-232) -> finalize
//
// The Elist of a subhalo (ParticleEnergy_t{pid,E}, :12-16) lives in two flat arrays ids[]/E[] at
// [slot_base, slot_base+n_src); its layout after every round is the reference's: bound | removed this
// round (E ascending) | removed earlier.  The host keeps a 16-byte mirror per active subhalo (status,
// Nbound, Nlast, CorrectionLoop) read back once per round to plan the next one.
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>

#include "context.cuh"
#include "seg_reduce.cuh"

namespace hbt
{

static constexpr int kBlock = 256;
static inline int grid_for(int64_t n, int per_block = kBlock) { return n > 0 ? div_up(n, per_block) : 1; }

// largest a in [0,n) with off[a] <= k
__device__ __forceinline__ int find_seg(const int *__restrict__ off, int n, int k)
{
  int lo = 0, hi = n;
  while (hi - lo > 1)
  {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= k) lo = mid; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int find_seg64(const int64_t *__restrict__ off, int n, int64_t k)
{
  int lo = 0, hi = n;
  while (hi - lo > 1)
  {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= k) lo = mid; else hi = mid;
  }
  return lo;
}

// The same for a whole warp whose lanes hold NON-DECREASING keys (consecutive elements of a concatenation): lane 0 searches,
// the others start from its segment and step forward - one load per lane while the warp stays inside one segment instead of
// log2(n) dependent loads each (with 5e5 segments per batch the per-element searches were a measurable part of every
// streaming kernel).  Every lane of the warp must call it; lanes past the end pass the last valid key.
template <class OffT, class KeyT>
__device__ __forceinline__ int find_seg_warp(const OffT *__restrict__ off, int n, KeyT k)
{
  const KeyT k0 = __shfl_sync(0xffffffffu, k, 0);
  int s = 0;
  if ((threadIdx.x & 31) == 0)
  {
    int lo = 0, hi = n;
    while (hi - lo > 1)
    {
      int mid = (lo + hi) >> 1;
      if (off[mid] <= k0) lo = mid; else hi = mid;
    }
    s = lo;
  }
  s = __shfl_sync(0xffffffffu, s, 0);
  while (s + 1 < n && off[s + 1] <= k) s++;
  return s;
}

// ---------------------------------------------------------------------------------------------------
// batch set-up kernels
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) init_ids_kernel(const int64_t *__restrict__ part_offset, const int64_t *__restrict__ slot_base,
                                                           int nsub, int64_t N, int *__restrict__ ids)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < N;
  if (!valid) i = N - 1;
  int s = find_seg_warp(part_offset, nsub, i);
  if (valid) ids[slot_base[s] + (i - part_offset[s])] = (int)i;
}

struct CopyJob
{
  int64_t dst, src; // slot offsets
};
// generic segmented copy between slot arrays: job j copies count = job_off[j+1]-job_off[j] ints
__global__ void __launch_bounds__(kBlock) seg_copy_kernel(const CopyJob *__restrict__ jobs, const int64_t *__restrict__ job_off, int njobs,
                                                           int64_t total, const int *__restrict__ src, int *__restrict__ dst)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < total;
  if (!valid) i = total - 1;
  int j = find_seg_warp(job_off, njobs, i);
  int64_t o = i - job_off[j];
  if (valid) dst[jobs[j].dst + o] = src[jobs[j].src + o];
}

struct LevelInit
{
  int sub, n_src, activate, shuffled;
};
// runs BEFORE the sampling shuffle: Particles[0] on entry is the old most-bound particle (src/subhalo_unbind.cpp:298)
__global__ void level_init_kernel(const LevelInit *__restrict__ li, int n, SubState *__restrict__ subs, const int *__restrict__ ids)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SubState &st = subs[li[i].sub];
  st.n_src = li[i].n_src;
  st.first_id = li[i].n_src > 0 ? ids[st.slot_base] : -1;
  st.shuffled = li[i].shuffled;
  if (li[i].activate) st.status = kActive;
}

// sampled mode: 64-bit keys (job << 40 | shuffle_key) over the concatenated sources that are larger than the sample
struct ShuffleJob
{
  int64_t slot_base;
  int sub, n;
};
__global__ void __launch_bounds__(kBlock) shuffle_keys_kernel(const ShuffleJob *__restrict__ jobs, const int64_t *__restrict__ job_off, int njobs,
                                                               int64_t total, uint64_t seed, uint64_t *__restrict__ key, int *__restrict__ val)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int j = find_seg64(job_off, njobs, i);
  int64_t o = i - job_off[j];
  key[i] = ((uint64_t)j << 40) | shuffle_key(seed, (uint64_t)jobs[j].sub, (uint64_t)o);
  val[i] = (int)i;
}
__global__ void __launch_bounds__(kBlock) shuffle_read_kernel(const ShuffleJob *__restrict__ jobs, const int64_t *__restrict__ job_off, int njobs,
                                                               int64_t total, const int *__restrict__ order, const int *__restrict__ ids,
                                                               int *__restrict__ tmp)
{
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  int64_t src = order[p];
  int j = find_seg64(job_off, njobs, src);
  tmp[p] = ids[jobs[j].slot_base + (src - job_off[j])];
}
__global__ void __launch_bounds__(kBlock) shuffle_write_kernel(const ShuffleJob *__restrict__ jobs, const int64_t *__restrict__ job_off, int njobs,
                                                                int64_t total, const int *__restrict__ tmp, int *__restrict__ ids)
{
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  int j = find_seg64(job_off, njobs, p);
  ids[jobs[j].slot_base + (p - job_off[j])] = tmp[p];
}
// disrupted + shuffled: the old most-bound particle is swapped back to the front (src/subhalo_unbind.cpp:367-374)
__global__ void __launch_bounds__(kBlock) restore_front_kernel(const CopyJob *__restrict__ jobs, const int64_t *__restrict__ job_off, int njobs,
                                                                int64_t total, const int *__restrict__ job_sub, const SubState *__restrict__ subs,
                                                                const int *__restrict__ ids_orig, int *__restrict__ ids)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int j = find_seg64(job_off, njobs, i);
  const SubState &st = subs[job_sub[j]];
  if (!st.shuffled || st.status != kDone || st.nbound != 1) return;
  int64_t slot = jobs[j].dst + (i - job_off[j]);
  if (ids_orig[slot] == st.first_id && slot != st.slot_base)
  {
    ids[st.slot_base] = st.first_id;
    ids[slot] = ids_orig[st.slot_base];
  }
}

// subhaloes whose source is too small to iterate (src/subhalo_unbind.cpp:269-293 and the disruption
// branch :361-379, which every source with 2 <= n < MinNumPartOfSub necessarily takes)
__global__ void trivial_kernel(const int *__restrict__ list, int n, SubState *__restrict__ subs, const int *__restrict__ ids,
                               const float4 *__restrict__ pos, DevConfig cfg, float *__restrict__ E)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SubState &st = subs[list[i]];
  // the one energy such a subhalo reports (Energies[0]): 0 as in the reference's n == 1 branch (:285-287); for 2 <= n <
  // MinNumPartOfSub the reference reports the E of whatever its partition left in Elist[0], which is not Particles[0]
  if ((st.is_orphan ? st.n_own : st.n_src) > 0) E[st.slot_base] = 0.f;
  const int nu = st.is_orphan ? st.n_own : st.n_src;
  if (nu < cfg.min_num_part && st.death == -1) st.death = cfg.snapshot_index;
  st.iterations = 0;
  if (nu == 0)
  {
    st.nbound = 0;
    st.mbound = 0.f;
  }
  else if (nu == 1)
  {
    st.nbound = 1;
    st.mbound = pos[ids[st.slot_base]].w;
  }
  else
  { // disruption without iterating
    st.nbound = 1;
    for (int j = 0; j < 3; j++) { st.ref_pos[j] = st.mb_pos[j]; st.ref_vel[j] = st.mb_vel[j]; }
    st.mbound = pos[ids[st.slot_base]].w; // no shuffle below MinNumPartOfSub <= MaxSampleSize... see classify
    st.spec_pot = st.spec_kin = 0.f;
    st.am[0] = st.am[1] = st.am[2] = 0.f;
  }
  st.nlast = st.nbound;
  st.status = kDone;
}

// ---------------------------------------------------------------------------------------------------
// round kernels
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) gather_src_kernel(const Segment *__restrict__ segs, const int *__restrict__ tree_off, int nseg, int S,
                                                             const int *__restrict__ ids, const float4 *__restrict__ pos,
                                                             float4 *__restrict__ tpos, int *__restrict__ ts_seg)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = k < S;
  if (!valid) k = S - 1;
  int a = find_seg_warp(tree_off, nseg, k);
  if (!valid) return;
  const Segment sg = segs[a];
  int64_t slot = sg.slot_base + sg.tree_first + (k - sg.tree_off);
  float4 p = pos[ids[slot]];
  p.w = __fmul_rn(p.w, sg.mass_factor); // MassFactor of a sampled tree (src/subhalo_unbind.cpp:96-99), 1 otherwise
  tpos[k] = p;
  ts_seg[k] = a;
}

// ids of the sorted sources (read phase), then written back in key order for full-evaluation segments:
// the Elist order inside [0,Nlast) is free in exact mode, and key order makes walk targets coherent.
__global__ void __launch_bounds__(kBlock) sorted_ids_read_kernel(const Segment *__restrict__ segs, const int *__restrict__ ts_seg,
                                                                  const int *__restrict__ sperm, int S, const int *__restrict__ ids,
                                                                  int *__restrict__ tmp, const int *__restrict__ rho, int *__restrict__ tmp_rho)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  const Segment sg = segs[ts_seg[k]];
  int src = sperm[k]; // S-index before sorting (same segment)
  const int64_t slot = sg.slot_base + sg.tree_first + (src - sg.tree_off);
  tmp[k] = ids[slot];
  if (rho) tmp_rho[k] = rho[slot];
}
__global__ void __launch_bounds__(kBlock) sorted_ids_write_kernel(const Segment *__restrict__ segs, const int *__restrict__ ts_seg, int S,
                                                                   const int *__restrict__ tmp, int *__restrict__ ids,
                                                                   const int *__restrict__ tmp_rho, int *__restrict__ rho)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  const Segment sg = segs[ts_seg[k]];
  if (sg.mode == kWalkUnbindFull && !sg.keep_order)
  {
    const int64_t slot = sg.slot_base + sg.tree_first + (k - sg.tree_off);
    ids[slot] = tmp[k];
    if (rho) rho[slot] = tmp_rho[k];
  }
}

// Periodic runs: rho[slot] = index of that Elist entry in the REFERENCE's Elist.  The reference's AveragePosition
// measures NEAREST offsets from Elist[0] (src/subhalo_unbind.cpp:152-154), and which particle that is follows from
// its hole-based partition (:21-58) - while this library keeps the bound part in key order.  So the reference order is
// carried as one int per entry: identity when a subhalo starts (its Particles order: own list + children's tails),
// moved with every permutation, and advanced by the Hoare rule each round (rho_update_kernel).
__global__ void __launch_bounds__(kBlock) rho_init_kernel(const Segment *__restrict__ segs, const int *__restrict__ tgt_seg, int T,
                                                           const SubState *__restrict__ subs, int *__restrict__ rho)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const Segment sg = segs[tgt_seg[t]];
  if (subs[sg.sub].iterations == 0) rho[sg.slot_base + (t - sg.tgt_off)] = t - sg.tgt_off;
}

__global__ void __launch_bounds__(kBlock) targets_kernel(const Segment *__restrict__ segs, const int *__restrict__ tgt_off, int nseg, int T,
                                                          const float4 *__restrict__ spos, const int *__restrict__ ids,
                                                          const float4 *__restrict__ pos, float4 *__restrict__ tgt_pm,
                                                          int64_t *__restrict__ tgt_slot, int *__restrict__ tgt_seg)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < T;
  if (!valid) t = T - 1;
  int a = find_seg_warp(tgt_off, nseg, t);
  if (!valid) return;
  const Segment sg = segs[a];
  int j = t - sg.tgt_off;
  int64_t slot = sg.slot_base + j;
  float4 p;
  if (sg.mode == kWalkUnbindFull && !sg.keep_order)
    p = spos[sg.tree_off + j]; // target j is sorted source j; w = its own mass (self term)
  else
  {
    p = pos[ids[slot]];
    // self term only for targets that are tree sources, with the tree's (scaled) mass (src/subhalo_unbind.cpp:345-349);
    // correction targets are never in the tree (:327)
    p.w = (sg.mode != kWalkUnbindCorrect && j < sg.tree_n) ? __fmul_rn(p.w, sg.mass_factor) : 0.f;
  }
  tgt_pm[t] = p;
  tgt_slot[t] = slot;
  tgt_seg[t] = a;
}

// Sampled mode: the Elist order of a sampled subhalo IS its sample (shuffled on entry, src/subhalo_unbind.cpp:302), so consecutive
// targets are spatially unrelated and a walk group's bounding box is the whole subhalo.  The WALK therefore visits the targets in
// key order - 30 key bits in the segment's root cube - through a permuted copy of (position, slot); nothing else changes order.
__global__ void __launch_bounds__(kBlock) walk_keys_kernel(const float4 *__restrict__ tgt_pm, const int *__restrict__ tgt_seg, int T,
                                                            const SegRoot *__restrict__ roots, uint64_t *__restrict__ key, int *__restrict__ val)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float4 p = tgt_pm[t];
  const int a = tgt_seg[t];
  key[t] = ((uint64_t)a << 30) | (morton_key(p.x, p.y, p.z, roots[a]) >> 33);
  val[t] = t;
}
__global__ void __launch_bounds__(kBlock) walk_gather_kernel(const int *__restrict__ order, const float4 *__restrict__ tgt_pm,
                                                              const int64_t *__restrict__ tgt_slot, int T, float4 *__restrict__ wpm,
                                                              int64_t *__restrict__ wslot)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int src = order[t];
  wpm[t] = tgt_pm[src];
  wslot[t] = tgt_slot[src];
}

__global__ void __launch_bounds__(kBlock) fill_tgt_seg_kernel(const int *__restrict__ tgt_off, int nseg, int T, int *__restrict__ tgt_seg)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < T;
  const int a = find_seg_warp(tgt_off, nseg, valid ? t : T - 1);
  if (valid) tgt_seg[t] = a;
}

__global__ void pre_walk_kernel(const Segment *__restrict__ segs, int nseg, SubState *__restrict__ subs, DevConfig cfg)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nseg) return;
  SubState &st = subs[segs[a].sub];
  if (segs[a].mode != kWalkRefine) st.iterations++;
  st.count_bound = 0;
  st.hoare_last = 0;
  st.hoare_first_bound = 0;
  for (int j = 0; j < 8; j++) st.sums[j] = 0.0;
  if (segs[a].mode == kWalkUnbindCorrect)
  { // RefVelDiff = RelativeVelocity(OldRef -> Ref), dK = 0.5*|RefVelDiff|^2  (src/subhalo_unbind.cpp:314-316)
    float d[3];
    for (int j = 0; j < 3; j++)
    {
      float dx = __fsub_rn(st.old_ref_pos[j], st.ref_pos[j]);
      if (cfg.periodic) dx = nearest_f(dx, cfg.box_size, cfg.box_half);
      float dv = __fsub_rn(st.old_ref_vel[j], st.ref_vel[j]);
      d[j] = __fadd_rn(dv, __fmul_rn(__fmul_rn(cfg.hz, cfg.scale_factor), dx));
      st.ref_diff[j] = d[j];
    }
    float vn = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
    st.dK = (float)(0.5 * (double)vn);
  }
}

__global__ void __launch_bounds__(kBlock) count_bound_kernel(const Segment *__restrict__ segs, const int *__restrict__ tgt_seg,
                                                              const int64_t *__restrict__ tgt_slot, int T, const float *__restrict__ E,
                                                              SubState *__restrict__ subs)
{ // 4 targets per thread; a block that lies inside one segment (the rule for large subhaloes) issues ONE atomic
  __shared__ int s_cnt[kBlock / 32];
  const int base = blockIdx.x * (kBlock * 4);
  const int last = min(base + kBlock * 4, T) - 1;
  const int a_first = tgt_seg[base], a_last = tgt_seg[last];
  if (a_first == a_last)
  {
    int c = 0;
#pragma unroll
    for (int it = 0; it < 4; it++)
    {
      const int t = base + it * kBlock + threadIdx.x;
      if (t < T) c += (E[tgt_slot[t]] < 0.f) ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0)
    {
      int tot = 0;
#pragma unroll
      for (int w = 0; w < kBlock / 32; w++) tot += s_cnt[w];
      if (tot) atomicAdd(&subs[segs[a_first].sub].count_bound, tot);
    }
    return;
  }
  for (int it = 0; it < 4; it++)
  { // the block straddles segment boundaries: warp-uniform or per-lane atomics
    const int t = base + it * kBlock + threadIdx.x;
    const bool valid = t < T;
    const int a = valid ? tgt_seg[t] : -1;
    const bool bound = valid && (E[tgt_slot[t]] < 0.f);
    const int a0 = __shfl_sync(0xffffffffu, a, 0);
    if (__all_sync(0xffffffffu, a == a0))
    {
      const unsigned m = __ballot_sync(0xffffffffu, bound);
      if ((threadIdx.x & 31) == 0 && m && a0 >= 0) atomicAdd(&subs[segs[a0].sub].count_bound, __popc(m));
    }
    else if (bound)
      atomicAdd(&subs[segs[a].sub].count_bound, 1);
  }
}

// PartitionBindingEnergy result -> disruption / CorrectionLoop / convergence (src/subhalo_unbind.cpp:357-403)
__global__ void state1_kernel(const Segment *__restrict__ segs, int nseg, SubState *__restrict__ subs, const float4 *__restrict__ pos, DevConfig cfg)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nseg) return;
  SubState &st = subs[segs[a].sub];
  const int Nlast = segs[a].tgt_n;
  st.hoare_nb = st.count_bound; // what the partition saw
  const int Nbound = cfg.no_stripping ? Nlast : st.count_bound; // NO_STRIPPING: Nbound=Nlast (src/subhalo_unbind.cpp:358-360)
  st.nlast = Nlast;
  if (Nbound < cfg.min_num_part)
  {
    st.nbound = 1;
    st.nlast = 1;
    if (st.death == -1) st.death = cfg.snapshot_index;
    for (int j = 0; j < 3; j++) { st.ref_pos[j] = st.mb_pos[j]; st.ref_vel[j] = st.mb_vel[j]; }
    st.mbound = pos[st.first_id].w; // Particles[0] after the old most-bound particle is swapped back (:367-377)
    st.spec_pot = st.spec_kin = 0.f;
    st.am[0] = st.am[1] = st.am[2] = 0.f;
    st.status = kDisrupted;
    return;
  }
  st.nbound = Nbound;
  const int Ndiff = Nlast - Nbound;
  if (Ndiff < Nbound && (cfg.max_sample <= 0 || Ndiff < cfg.max_sample))
  {
    st.correction = 1;
    for (int j = 0; j < 3; j++) { st.old_ref_pos[j] = st.ref_pos[j]; st.old_ref_vel[j] = st.ref_vel[j]; }
  }
  if ((float)Nbound >= __fmul_rn((float)Nlast, cfg.bound_mass_precision))
  {
    st.status = kConverged;
    if (st.death != -1) st.death = -1;
    if (st.sinktrack != -1) { st.sink = -1; st.sinktrack = -1; }
  }
}

// Sampled mode, full-evaluation rounds with Nlast > MaxSampleSize: the order of the bound part after the
// partition selects the next sample, so the reference's hole-based Hoare partition (src/subhalo_unbind.cpp:21-58)
// is reproduced exactly.  With Nb bound elements, f_1<f_2<.. the unbound elements at indices [1,Nb) and
// b_1>b_2>.. the bound elements at indices >= Nb:  new[0]=old[b_1], new[f_k]=old[b_(k+1)], old[0] (if bound)
// ends at f_last; bound elements already in [1,Nb) stay.  hoare_flags marks the two misplaced sets, two scans rank
// them, hoare_fpos inverts the front ranks.
// `rho` == nullptr (non-periodic): only the sampled full-evaluation segments take part and the reference index of an
// entry is its position.  Periodic: every active segment takes part, indexed through rho (flags live in rho space).
__device__ __forceinline__ bool hoare_segment(const Segment &sg, const int *rho) { return rho ? true : (sg.keep_order && sg.mode == kWalkUnbindFull); }

__global__ void __launch_bounds__(kBlock) hoare_clear_kernel(int T, int *__restrict__ uflag, int *__restrict__ bflag)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uflag[t] = 0;
  bflag[t] = 0;
}
__global__ void __launch_bounds__(kBlock) hoare_flags_kernel(const Segment *__restrict__ segs, const int *__restrict__ tgt_seg,
                                                              const int64_t *__restrict__ tgt_slot, int T, const float *__restrict__ E,
                                                              SubState *__restrict__ subs, const int *__restrict__ rho,
                                                              int *__restrict__ uflag, int *__restrict__ bflag)
{ // uflag/bflag are pre-cleared; every entry writes the flags of ITS reference index.  hoare_last (the largest reference index
  // of a bound entry inside [1, Nb)) is a max over up to 1e8 entries of ONE subhalo: reduced per block before the atomic
  __shared__ int s_max[kBlock / 32];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int t_first = blockIdx.x * blockDim.x, t_last = min(t_first + (int)blockDim.x, T) - 1;
  const bool uniform = tgt_seg[t_first] == tgt_seg[t_last]; // tgt_seg is non-decreasing: the whole block is one segment
  int cand = 0;                                             // this thread's candidate for hoare_last (0 = none)
  int sub = -1;
  if (t < T)
  {
    const Segment sg = segs[tgt_seg[t]];
    if (hoare_segment(sg, rho))
    {
      SubState &st = subs[sg.sub];
      if (st.status != kDisrupted)
      {
        const int r = rho ? rho[tgt_slot[t]] : t - sg.tgt_off;
        const bool bound = E[tgt_slot[t]] < 0.f;
        uflag[sg.tgt_off + r] = (!bound && r >= 1 && r < st.hoare_nb);
        bflag[sg.tgt_off + r] = (bound && r >= st.hoare_nb);
        if (bound && r == 0) st.hoare_first_bound = 1;
        if (bound && r >= 1 && r < st.hoare_nb)
        {
          cand = r;
          sub = sg.sub;
        }
      }
    }
  }
  if (!uniform)
  {
    if (cand > 0) atomicMax(&subs[sub].hoare_last, cand);
    return;
  }
  const int wmax = __reduce_max_sync(0xffffffffu, cand);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = wmax;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    int m = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; w++) m = max(m, s_max[w]);
    if (m > 0) atomicMax(&subs[segs[tgt_seg[t_first]].sub].hoare_last, m);
  }
}
__global__ void __launch_bounds__(kBlock) hoare_fpos_kernel(const Segment *__restrict__ segs, const int *__restrict__ tgt_seg, int T,
                                                             const int *__restrict__ uflag, const int *__restrict__ uscan, int *__restrict__ fpos)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T || !uflag[t]) return;
  const Segment sg = segs[tgt_seg[t]];
  int before = sg.tgt_off > 0 ? uscan[sg.tgt_off - 1] : 0;
  fpos[sg.tgt_off + (uscan[t] - before) - 1] = t - sg.tgt_off; // k-th misplaced unbound element of the front
}

// Reference index of a BOUND entry after the hole-based partition (src/subhalo_unbind.cpp:21-58), from its index r
// before it.  Elist[0] is lifted out (Etmp) and the hole travels: with f_1<f_2<..<f_K the unbound entries at indices
// [1,Nb) and b_1>b_2>.. the bound entries at indices >= Nb,
//     new[0] = old[b_1],  new[f_k] = old[b_(k+1)]  (while such b exist),
// and when Elist[0] itself is bound (then there are exactly K such b) the backward scan goes on below Nb: the LAST bound
// entry L of [1,Nb) moves into the hole f_K if L > f_K (f_0 = 0), and old[0] lands in the last hole (L, else f_K).
// bscan = inclusive scan of the misplaced-bound flags, fpos = the holes f_k in ascending order.
__device__ __forceinline__ int hoare_dest(const Segment &sg, const SubState &st, int r, const int *__restrict__ bscan,
                                          const int *__restrict__ fpos)
{
  const int nb = st.hoare_nb;
  const int b0 = sg.tgt_off > 0 ? bscan[sg.tgt_off - 1] : 0;
  const int mis = bscan[sg.tgt_off + sg.tgt_n - 1] - b0; // bound entries at indices >= Nb
  if (r >= nb)
  {
    const int k = mis - (bscan[sg.tgt_off + r] - b0 - 1); // 1 = last misplaced bound element of the segment
    return (k == 1) ? 0 : fpos[sg.tgt_off + k - 2];
  }
  if (!st.hoare_first_bound) return r; // Elist[0] unbound: K = mis - 1 holes, all filled by the b's
  const int fK = mis >= 1 ? fpos[sg.tgt_off + mis - 1] : 0;
  const int L = st.hoare_last;
  const bool moveL = L > fK;
  if (r == 0) return moveL ? L : fK;
  if (r == L && moveL) return fK;
  return r;
}

// periodic runs: advance rho of the bound entries, remember which particle is now the reference's Elist[0]
__global__ void __launch_bounds__(kBlock) rho_update_kernel(const Segment *__restrict__ segs, const int *__restrict__ tgt_seg,
                                                             const int64_t *__restrict__ tgt_slot, int T, const float *__restrict__ E,
                                                             SubState *__restrict__ subs, const int *__restrict__ ids, const int *__restrict__ bscan,
                                                             const int *__restrict__ fpos, const int *__restrict__ rho, int *__restrict__ rho_new)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const Segment sg = segs[tgt_seg[t]];
  SubState &st = subs[sg.sub];
  int r = -1; // unbound entries get their index from the sort (position in the E-sorted tail)
  if (st.status != kDisrupted && E[tgt_slot[t]] < 0.f)
  {
    r = hoare_dest(sg, st, rho[tgt_slot[t]], bscan, fpos);
    if (r == 0) st.origin_id = ids[tgt_slot[t]];
  }
  else if (st.status != kDisrupted && st.hoare_nb == 0 && rho[tgt_slot[t]] == 0)
    st.origin_id = ids[tgt_slot[t]]; // NO_STRIPPING with nothing bound: the partition leaves Elist[0] where it is

  rho_new[t] = r;
}

// One key per target: (segment, bound|unbound) major, then E (where the reference sorts) or the current
// position (where it does not): a single radix sort = partition + tail sort + final bound sort.
__global__ void __launch_bounds__(kBlock) sort_keys_kernel(const Segment *__restrict__ segs, const int *__restrict__ tgt_seg,
                                                            const int64_t *__restrict__ tgt_slot, int T, const float *__restrict__ E,
                                                            const SubState *__restrict__ subs, const int *__restrict__ bscan,
                                                            const int *__restrict__ fpos, const int *__restrict__ rho, uint64_t *__restrict__ key,
                                                            int *__restrict__ val)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int a = tgt_seg[t];
  const Segment sg = segs[a];
  int status = subs[sg.sub].status;
  float e = E[tgt_slot[t]];
  uint32_t j = (uint32_t)(t - sg.tgt_off);
  uint64_t hi, low;
  if (status == kDisrupted) { hi = 2ull * a; low = j; }
  else if (!(e < 0.f)) { hi = 2ull * a + 1; low = float_to_ordered(e); }
  else
  {
    hi = 2ull * a;
    if (status == kConverged)
      low = float_to_ordered(e);
    else if (bscan && sg.keep_order && sg.mode == kWalkUnbindFull)
      low = (uint32_t)hoare_dest(sg, subs[sg.sub], rho ? rho[tgt_slot[t]] : (int)j, bscan, fpos); // Hoare destination of a bound element
    else
      low = j;
  }
  key[t] = (hi << 32) | low;
  val[t] = t;
}
__global__ void __launch_bounds__(kBlock) permute_read_kernel(const int *__restrict__ order, const int64_t *__restrict__ tgt_slot, int T,
                                                               const int *__restrict__ ids, const float *__restrict__ E,
                                                               int *__restrict__ tmp_id, float *__restrict__ tmp_E,
                                                               const int *__restrict__ rho_new = nullptr, int *__restrict__ tmp_rho = nullptr)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= T) return;
  int64_t slot = tgt_slot[order[p]];
  tmp_id[p] = ids[slot];
  tmp_E[p] = E[slot];
  if (rho_new) tmp_rho[p] = rho_new[order[p]];
}
__global__ void __launch_bounds__(kBlock) permute_write_kernel(const int64_t *__restrict__ tgt_slot, int T, const int *__restrict__ tmp_id,
                                                                const float *__restrict__ tmp_E, int *__restrict__ ids, float *__restrict__ E,
                                                                const Segment *__restrict__ segs = nullptr, const int *__restrict__ tgt_seg = nullptr,
                                                                const int *__restrict__ tmp_rho = nullptr, int *__restrict__ rho = nullptr)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= T) return;
  int64_t slot = tgt_slot[p];
  ids[slot] = tmp_id[p];
  E[slot] = tmp_E[p];
  if (rho)
  { // bound entries keep the Hoare index; the unbound tail is E-sorted in the reference too (:382): index = position.
    // Sampled full-evaluation segments were physically arranged by the Hoare rule: index = position as well.
    const Segment sg = segs[tgt_seg[p]];
    const int r = tmp_rho[p];
    rho[slot] = (r < 0 || (sg.keep_order && sg.mode == kWalkUnbindFull)) ? p - sg.tgt_off : r;
  }
}

__device__ __forceinline__ double warp_sum_d(double v)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NV fp64 sums per target into subs[sub].sums with a fixed summation tree aligned to each subhalo (seg_reduce.cuh): pass A
// inside the producing kernel (block = one 256-target chunk of one segment), pass B = seg_finish_kernel.  The same subhalo
// therefore gives the same bits on every run and in every batch.
template <int NV>
struct SumsDone
{
  const Segment *segs;
  SubState *subs;
  __device__ void operator()(int a, const double (&s)[NV]) const
  {
    SubState &st = subs[segs[a].sub];
#pragma unroll
    for (int i = 0; i < NV; i++) st.sums[i] = s[i];
  }
};
template <int NV>
__global__ void __launch_bounds__(kBlock) seg_finish_kernel(const Segment *__restrict__ segs, int nseg, const int *__restrict__ chunk_off,
                                                             const double *__restrict__ partial, SubState *__restrict__ subs)
{
  const int a = blockIdx.x;
  if (a >= nseg) return;
  seg_reduce_finish_block<NV>(a, chunk_off, partial, SumsDone<NV>{segs, subs});
}

// EnergySnapshot_t::AverageVelocity / AveragePosition over the first Nbound (src/subhalo_unbind.cpp:108-188)
__global__ void __launch_bounds__(kBlock) frame_reduce_kernel(const Segment *__restrict__ segs, int nseg, const int *__restrict__ chunk_off,
                                                               const int *__restrict__ ids, const float4 *__restrict__ pos,
                                                               const float4 *__restrict__ vel, SubState *__restrict__ subs, DevConfig cfg,
                                                               double *__restrict__ partial)
{
  int a, c;
  seg_chunk_of_block(chunk_off, nseg, blockIdx.x, a, c);
  const Segment sg = segs[a];
  const int j = c * kBlock + threadIdx.x; // Elist index inside the subhalo
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  if (j < sg.tgt_n)
  {
    const SubState &st = subs[sg.sub];
    if (st.status != kDisrupted && j < st.nbound)
    {
      int id = ids[sg.slot_base + j];
      float4 x = pos[id], u = vel[id];
      float m = x.w;
      v[0] = (double)m;
      v[1] = (double)__fmul_rn(u.x, m); // float product, double accumulate (:129-131)
      v[2] = (double)__fmul_rn(u.y, m);
      v[3] = (double)__fmul_rn(u.z, m);
      if (cfg.periodic)
      {
        float4 o = pos[st.origin_id]; // origin = the reference's Elist[0] (:152-154)
        v[4] = nearest_d((double)x.x - (double)o.x, (double)cfg.box_size, (double)cfg.box_half) * (double)m;
        v[5] = nearest_d((double)x.y - (double)o.y, (double)cfg.box_size, (double)cfg.box_half) * (double)m;
        v[6] = nearest_d((double)x.z - (double)o.z, (double)cfg.box_size, (double)cfg.box_half) * (double)m;
      }
      else
      {
        v[4] = (double)__fmul_rn(x.x, m);
        v[5] = (double)__fmul_rn(x.y, m);
        v[6] = (double)__fmul_rn(x.z, m);
      }
    }
  }
  seg_reduce_chunk<7>(v, partial);
}

__global__ void state2_kernel(const Segment *__restrict__ segs, int nseg, SubState *__restrict__ subs, const int *__restrict__ ids,
                              const float4 *__restrict__ pos, const float4 *__restrict__ vel, DevConfig cfg)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nseg) return;
  SubState &st = subs[segs[a].sub];
  if (st.status == kDisrupted) return;
  int id0 = ids[st.slot_base];
  float4 x0 = pos[id0], v0 = vel[id0];
  const float4 xo = pos[cfg.periodic ? st.origin_id : id0];
  if (st.nbound == 1)
  {
    st.ref_vel[0] = v0.x; st.ref_vel[1] = v0.y; st.ref_vel[2] = v0.z;
    st.ref_pos[0] = x0.x; st.ref_pos[1] = x0.y; st.ref_pos[2] = x0.z;
    st.mbound = x0.w;
  }
  else
  {
    double msum = st.sums[0];
    st.mbound = (float)msum;
    for (int j = 0; j < 3; j++) st.ref_vel[j] = (float)(st.sums[1 + j] / msum);
    double o[3] = {(double)xo.x, (double)xo.y, (double)xo.z};
    for (int j = 0; j < 3; j++)
    {
      double s = st.sums[4 + j] / msum;
      if (cfg.periodic) s += o[j];
      st.ref_pos[j] = (float)s;
    }
  }
  for (int j = 0; j < 8; j++) st.sums[j] = 0.0;
}

// EnergySnapshot_t::AverageKinematics (src/subhalo_unbind.cpp:189-232) for subhaloes that converged this round
__global__ void __launch_bounds__(kBlock) kinematics_kernel(const Segment *__restrict__ segs, int nseg, const int *__restrict__ chunk_off,
                                                             const int *__restrict__ ids, const float *__restrict__ E,
                                                             const float4 *__restrict__ pos, const float4 *__restrict__ vel,
                                                             SubState *__restrict__ subs, DevConfig cfg, double *__restrict__ partial)
{
  int a, c;
  seg_chunk_of_block(chunk_off, nseg, blockIdx.x, a, c);
  const Segment sg = segs[a];
  const int j = c * kBlock + threadIdx.x;
  double v[6] = {0, 0, 0, 0, 0, 0};
  if (j < sg.tgt_n)
  {
    const SubState &st = subs[sg.sub];
    if (st.status == kConverged && j < st.nbound)
    {
      int64_t slot = sg.slot_base + j;
      int id = ids[slot];
      float4 x = pos[id], u = vel[id];
      float m = x.w;
      const float xs[3] = {x.x, x.y, x.z}, us[3] = {u.x, u.y, u.z};
      double dx[3], dv[3], K = 0.0;
      for (int q = 0; q < 3; q++)
      {
        dx[q] = (double)__fsub_rn(xs[q], st.ref_pos[q]);
        if (cfg.periodic) dx[q] = nearest_d(dx[q], (double)cfg.box_size, (double)cfg.box_half);
        dx[q] *= (double)cfg.scale_factor;
        dv[q] = (double)__fsub_rn(us[q], st.ref_vel[q]) + (double)cfg.hz * dx[q];
        K += dv[q] * dv[q] * (double)m;
      }
      v[0] = (double)__fmul_rn(E[slot], m);
      v[1] = K;
      v[2] = (dx[1] * dv[2] - dx[2] * dv[1]) * (double)m;
      v[3] = (dx[2] * dv[0] - dx[0] * dv[2]) * (double)m;
      v[4] = (dx[0] * dv[1] - dx[1] * dv[0]) * (double)m;
      v[5] = (double)m;
    }
  }
  seg_reduce_chunk<6>(v, partial);
}

// target split: E_stage (target order, summed over the cooperating contexts) -> E[slot]
__global__ void __launch_bounds__(kBlock) scatter_stage_kernel(const int64_t *__restrict__ tgt_slot, int T, const float *__restrict__ stage,
                                                                float *__restrict__ E)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) E[tgt_slot[t]] = stage[t];
}

struct RoundResult
{
  int status, nbound, nlast, correction;
};

__global__ void finalize_kernel(const Segment *__restrict__ segs, int nseg, SubState *__restrict__ subs, const int *__restrict__ ids,
                                const float4 *__restrict__ pos, const float4 *__restrict__ vel, RoundResult *__restrict__ res)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nseg) return;
  SubState &st = subs[segs[a].sub];
  if (st.status == kConverged)
  {
    if (st.nbound > 1)
    {
      double M = st.sums[5], Eav = st.sums[0] / M, K = st.sums[1] * (0.5 / M);
      st.spec_pot = (float)(Eav - K);
      st.spec_kin = (float)K;
      st.am[0] = (float)(st.sums[2] / M);
      st.am[1] = (float)(st.sums[3] / M);
      st.am[2] = (float)(st.sums[4] / M);
    }
    else
    {
      st.spec_pot = st.spec_kin = 0.f;
      st.am[0] = st.am[1] = st.am[2] = 0.f;
    }
    int id0 = ids[st.slot_base]; // Particles[0] after the final E-sort (:417-418)
    float4 x0 = pos[id0], v0 = vel[id0];
    st.mb_pos[0] = x0.x; st.mb_pos[1] = x0.y; st.mb_pos[2] = x0.z;
    st.mb_vel[0] = v0.x; st.mb_vel[1] = v0.y; st.mb_vel[2] = v0.z;
  }
  res[a] = RoundResult{st.status, st.nbound, st.nlast, st.correction};
  if (st.status == kConverged || st.status == kDisrupted) st.status = kDone;
}

// final packing of Subhalo_t::Particles (and SAVE_BINDING_ENERGY energies) into the caller's layout
__global__ void __launch_bounds__(kBlock) pack_output_kernel(const int64_t *__restrict__ out_off, const int64_t *__restrict__ slot_base,
                                                              const int *__restrict__ nbound, int nsub, int64_t total,
                                                              const int *__restrict__ ids, const float *__restrict__ E,
                                                              int *__restrict__ out_ids, float *__restrict__ out_E, int id_base)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < total;
  if (!valid) i = total - 1;
  int s = find_seg_warp(out_off, nsub, i);
  if (!valid) return;
  int64_t o = i - out_off[s];
  out_ids[i] = ids[slot_base[s] + o] + id_base; // id_base: this batch is a part of a pipelined hbtu_unbind_batch
  if (out_E) out_E[i] = (o < nbound[s]) ? E[slot_base[s] + o] : 0.f;
}

// ---------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) ring_copy_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

char *ring_reserve(Context &c, size_t bytes)
{
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (c.ring_used + need > c.ring_cap)
  { // wrap: everything staged so far must have been consumed
    HBT_CUDA(cudaStreamSynchronize(c.stream));
    c.ring_used = 0;
    if (need > c.ring_cap)
    {
      if (c.h_ring) cudaFreeHost(c.h_ring);
      c.h_ring = c.d_ring = nullptr;
      c.ring_cap = 0;
      const size_t cap = std::max<size_t>(need * 2, (size_t)16 << 20);
      HBT_CUDA(cudaHostAlloc((void **)&c.h_ring, cap, cudaHostAllocMapped));
      HBT_CUDA(cudaHostGetDevicePointer((void **)&c.d_ring, c.h_ring, 0));
      c.ring_cap = cap;
    }
  }
  char *p = c.h_ring + c.ring_used;
  c.ring_used += need;
  return p;
}

void ring_commit(Context &c, void *dst, const char *staged, size_t bytes)
{
  if (bytes == 0) return;
  const size_t n16 = ((bytes + 15) & ~(size_t)15) / 16;
  const int grid = (int)std::min<size_t>((n16 + kBlock - 1) / kBlock, 148 * 4);
  ring_copy_kernel<<<grid, kBlock, 0, c.stream>>>(reinterpret_cast<uint4 *>(dst), reinterpret_cast<const uint4 *>(c.d_ring + (staged - c.h_ring)), n16);
  HBT_CHECK_LAUNCH();
}

void upload_bytes(Context &c, void *dst, const void *src, size_t bytes)
{
  if (bytes == 0) return;
  char *p = ring_reserve(c, bytes);
  const char *q = static_cast<const char *>(src);
  parallel_ranges((int64_t)bytes, 4 << 20, [=](int64_t b, int64_t e) { std::memcpy(p + b, q + b, (size_t)(e - b)); });
  ring_commit(c, dst, p, bytes);
}

void *readback_buffer(Context &c, size_t bytes)
{
  if (bytes > c.back_cap)
  {
    if (c.h_back) cudaFreeHost(c.h_back);
    c.h_back = nullptr;
    c.back_cap = 0;
    const size_t cap = std::max<size_t>(bytes + bytes / 4, (size_t)1 << 20);
    HBT_CUDA(cudaHostAlloc((void **)&c.h_back, cap, cudaHostAllocDefault));
    c.back_cap = cap;
  }
  return c.h_back;
}

template <class T>
static T *upload(Context &c, const std::vector<T> &v)
{ // arena allocations are 256-byte aligned and padded, so the 16-byte granularity of the copy kernel stays inside them
  T *d = c.arena.alloc<T>((int64_t)v.size() + 16 / sizeof(T) + 1);
  if (!v.empty()) upload_bytes(c, d, v.data(), sizeof(T) * v.size());
  return d;
}

// The job tables of the per-level set-up kernels are tiny; they live in the round arena, which is dead between rounds
// (run_round / run_refine reset it when they start).  No cudaMalloc / cudaFree / stream synchronisation per level: the
// copies are stream-ordered and cudaMemcpyAsync from pageable host memory returns once the source has been staged.
static void run_seg_copy(Context &c, const std::vector<CopyJob> &jobs, const std::vector<int64_t> &job_off, const int *src, int *dst)
{
  if (jobs.empty() || job_off.back() == 0) return;
  CopyJob *d_jobs = upload(c, jobs);
  int64_t *d_off = upload(c, job_off);
  int64_t total = job_off.back();
  seg_copy_kernel<<<grid_for(total), kBlock, 0, c.stream>>>(d_jobs, d_off, (int)jobs.size(), total, src, dst);
  HBT_CHECK_LAUNCH();
  c.ls.launches++;
}

// one potential evaluation for every active subhalo of the level
static void run_round(Context &c, std::vector<int> &active)
{
  const int nseg = (int)active.size();
  std::vector<Segment> segs(nseg);
  std::vector<int> tree_off(nseg + 1), tgt_off(nseg + 1);
  // walk warps per walk class (device_tree.cuh: T=1, 2, 4 per-lane walks and the group walk); a segment belongs to one class
  std::vector<int> warp_off[kWalkClasses];
  for (auto &v : warp_off) v.resize(nseg + 1);
  // planned on several host threads (the GPU is idle meanwhile): every chunk of segments fills its records and sums its sources,
  // targets and warps per walk class; the chunk totals are scanned; every chunk then writes its offsets
  struct ChunkSum
  {
    int64_t S = 0, T = 0, W[kWalkClasses] = {};
    bool hoare = false;
  };
  const auto chunks = chunk_ranges(nseg, 1 << 13);
  std::vector<ChunkSum> csum(chunks.size());
  const int64_t max_sample = c.cfg.max_sample;
  const SubHost *subs = c.subs.data();
  const int *act = active.data();
  Segment *segp = segs.data();
  run_chunks(chunks, [&, subs, act, segp](int k, int64_t a0, int64_t a1) {
    ChunkSum cs;
    for (int64_t a = a0; a < a1; a++)
    {
      const SubHost &h = subs[act[a]];
      Segment &sg = segp[a];
      sg.slot_base = h.slot_base;
      sg.sub = act[a];
      sg.mode = h.correction ? kWalkUnbindCorrect : kWalkUnbindFull;
      if (h.correction)
      {
        sg.tree_first = h.nbound;
        sg.tree_n = h.nlast - h.nbound;
      }
      else
      {
        sg.tree_first = 0;
        sg.tree_n = h.nbound;
      }
      sg.keep_order = 0;
      sg.mass_factor = 1.f;
      if (max_sample > 0 && h.nbound > max_sample)
      { // sampled potential: tree of the first MaxSampleSize entries, masses scaled by Nlast/MaxSampleSize (:333-339)
        sg.keep_order = 1;
        if (!h.correction)
        {
          sg.tree_n = (int)max_sample;
          sg.mass_factor = (float)h.nbound / (float)max_sample;
          cs.hoare = true;
        }
      }
      sg.tgt_n = h.nbound;
      const WalkClass wcl = walk_class(sg.tgt_n, sg.tree_n);
      cs.S += sg.tree_n;
      cs.T += sg.tgt_n;
      cs.W[wcl.index] += (sg.tgt_n + wcl.targets_per_warp - 1) / wcl.targets_per_warp;
    }
    csum[k] = cs;
  });
  int64_t S = 0, T = 0, W[kWalkClasses] = {};
  bool any_hoare = false;
  std::vector<ChunkSum> cbase(chunks.size());
  for (size_t k = 0; k < chunks.size(); k++)
  {
    cbase[k].S = S;
    cbase[k].T = T;
    for (int q = 0; q < kWalkClasses; q++) cbase[k].W[q] = W[q];
    S += csum[k].S;
    T += csum[k].T;
    for (int q = 0; q < kWalkClasses; q++) W[q] += csum[k].W[q];
    any_hoare = any_hoare || csum[k].hoare;
  }
  if (S > 0x3fffffff || T > 0x3fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "round larger than 2^30 particles"};
  {
    int *toff = tree_off.data(), *goff = tgt_off.data();
    int *woff[kWalkClasses];
    for (int q = 0; q < kWalkClasses; q++) woff[q] = warp_off[q].data();
    run_chunks(chunks, [&, segp, toff, goff](int k, int64_t a0, int64_t a1) {
      int64_t s = cbase[k].S, t = cbase[k].T, w[kWalkClasses];
      for (int q = 0; q < kWalkClasses; q++) w[q] = cbase[k].W[q];
      for (int64_t a = a0; a < a1; a++)
      {
        Segment &sg = segp[a];
        const WalkClass wcl = walk_class(sg.tgt_n, sg.tree_n);
        sg.tree_off = (int)s;
        sg.tgt_off = (int)t;
        sg.warp_off = (int)w[wcl.index];
        toff[a] = (int)s;
        goff[a] = (int)t;
        for (int q = 0; q < kWalkClasses; q++) woff[q][a] = (int)w[q];
        s += sg.tree_n;
        t += sg.tgt_n;
        w[wcl.index] += (sg.tgt_n + wcl.targets_per_warp - 1) / wcl.targets_per_warp;
      }
    });
  }
  tree_off[nseg] = (int)S;
  tgt_off[nseg] = (int)T;
  for (int q = 0; q < kWalkClasses; q++) warp_off[q][nseg] = (int)W[q];

  Arena &ar = c.arena;
  ar.reset();
  ar.reserve(tree_arena_bytes(S, nseg) + T * (c.cfg.max_sample > 0 ? 132 : 68) + (int64_t)nseg * 128);
  cudaStream_t st = c.stream;
  Segment *d_segs = upload(c, segs);
  int *d_tree_off = upload(c, tree_off), *d_tgt_off = upload(c, tgt_off);
  int *d_warp_off[kWalkClasses];
  for (int q = 0; q < kWalkClasses; q++) d_warp_off[q] = upload(c, warp_off[q]);

  HBT_CUDA(cudaEventRecord(c.ev[0], st));
  HBT_CUDA(cudaEventRecord(c.ev_ph[0], st));
  TreeArrays tr;
  tr.S = (int)S;
  tr.nseg = nseg;
  tr.tree_off = d_tree_off;
  tr.h_tree_off = tree_off.data();
  tr.tpos = ar.alloc<float4>(S);
  tr.ts_seg = ar.alloc<int>(S);
  tr.bbox = ar.alloc<uint32_t>(6 * (int64_t)nseg);
  gather_src_kernel<<<grid_for(S), kBlock, 0, st>>>(d_segs, d_tree_off, nseg, (int)S, c.d_ids, c.d_pos, tr.tpos, tr.ts_seg);
  HBT_CHECK_LAUNCH();
  c.ls.launches++;
  launch_init_bbox(tr.bbox, nseg, st, c.ls);
  launch_bbox(tr.tpos, tr.ts_seg, (int)S, tr.bbox, st, c.ls);
  HBT_CUDA(cudaEventRecord(c.ev_ph[1], st));
  build_trees(tr, ar, c.cfg, st, c.ls);
  HBT_CUDA(cudaEventRecord(c.ev_ph[2], st));
  int *const rho = c.cfg.periodic ? c.d_rho : nullptr; // reference Elist order, tracked in periodic runs only
  float4 *tgt_pm = ar.alloc<float4>(T);
  int64_t *tgt_slot = ar.alloc<int64_t>(T);
  int *tgt_seg = ar.alloc<int>(T);
  if (rho)
  { // a subhalo's first round sees its whole source in the reference's Particles order: rho = identity
    fill_tgt_seg_kernel<<<grid_for(T), kBlock, 0, st>>>(d_tgt_off, nseg, (int)T, tgt_seg);
    HBT_CHECK_LAUNCH();
    rho_init_kernel<<<grid_for(T), kBlock, 0, st>>>(d_segs, tgt_seg, (int)T, c.d_subs, rho);
    HBT_CHECK_LAUNCH();
    c.ls.launches += 2;
  }
  int *tmp_ids = ar.alloc<int>(S), *tmp_rho_s = rho ? ar.alloc<int>(S) : nullptr;
  sorted_ids_read_kernel<<<grid_for(S), kBlock, 0, st>>>(d_segs, tr.ts_seg, tr.sperm, (int)S, c.d_ids, tmp_ids, rho, tmp_rho_s);
  HBT_CHECK_LAUNCH();
  sorted_ids_write_kernel<<<grid_for(S), kBlock, 0, st>>>(d_segs, tr.ts_seg, (int)S, tmp_ids, c.d_ids, tmp_rho_s, rho);
  HBT_CHECK_LAUNCH();
  targets_kernel<<<grid_for(T), kBlock, 0, st>>>(d_segs, d_tgt_off, nseg, (int)T, tr.spos, c.d_ids, c.d_pos, tgt_pm, tgt_slot, tgt_seg);
  HBT_CHECK_LAUNCH();
  pre_walk_kernel<<<grid_for(nseg), kBlock, 0, st>>>(d_segs, nseg, c.d_subs, c.cfg);
  HBT_CHECK_LAUNCH();
  c.ls.launches += 4;
  const float4 *walk_pm = tgt_pm;
  const int64_t *walk_slot = tgt_slot;
  if (any_hoare || (c.cfg.max_sample > 0 && std::any_of(segs.begin(), segs.end(), [](const Segment &g) { return g.keep_order != 0; })))
  { // sampled subhaloes in this round: walk their (shuffled) targets in key order
    uint64_t *wk_a = ar.alloc<uint64_t>(T), *wk_b = ar.alloc<uint64_t>(T);
    int *wv_a = ar.alloc<int>(T), *wv_b = ar.alloc<int>(T);
    walk_keys_kernel<<<grid_for(T), kBlock, 0, st>>>(tgt_pm, tgt_seg, (int)T, tr.roots, wk_a, wv_a);
    HBT_CHECK_LAUNCH();
    int bits = 31;
    while ((1ll << (bits - 30)) < nseg) bits++;
    cub::DoubleBuffer<uint64_t> dk(wk_a, wk_b);
    cub::DoubleBuffer<int> dv(wv_a, wv_b);
    size_t tb = 0;
    HBT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)T, 0, bits, st));
    void *tmp = ar.alloc<char>((int64_t)tb);
    HBT_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (int)T, 0, bits, st));
    float4 *wpm = ar.alloc<float4>(T);
    int64_t *wslot = ar.alloc<int64_t>(T);
    walk_gather_kernel<<<grid_for(T), kBlock, 0, st>>>(dv.Current(), tgt_pm, tgt_slot, (int)T, wpm, wslot);
    HBT_CHECK_LAUNCH();
    c.ls.launches += 3 + (bits + 7) / 8;
    walk_pm = wpm;
    walk_slot = wslot;
  }
  HBT_CUDA(cudaEventRecord(c.ev[1], st));
  HBT_CUDA(cudaEventRecord(c.ev_ph[3], st));

  WalkArgs wa{};
  wa.node_xm = tr.node_xm;
  wa.node_aux = tr.node_aux;
  wa.cellcount = tr.cellcount;
  wa.tree_off = d_tree_off;
  wa.segs = d_segs;
  wa.nseg = nseg;
  wa.tgt_pm = walk_pm;
  wa.tgt_slot = walk_slot;
  wa.ids = c.d_ids;
  wa.vel = c.d_vel;
  wa.E = c.d_E;
  wa.subs = c.d_subs;
  wa.out = nullptr;
  wa.counters = c.count_interactions ? c.d_counters : nullptr;
  const bool split = c.split_n > 1 && c.split_fn != nullptr;
  float *e_stage = nullptr;
  if (split)
  { // this context walks its share of the CTAs and writes E in target order; the others' entries stay zero until the all-reduce
    e_stage = ar.alloc<float>(T);
    HBT_CUDA(cudaMemsetAsync(e_stage, 0, sizeof(float) * (size_t)T, st));
    wa.split_rank = c.split_rank;
    wa.split_n = c.split_n;
    wa.E_stage = e_stage;
  }
  for (int q = kWalkClasses - 1; q >= 0; q--)
  { // largest segments first: their long-running warps start while the small classes fill the tail
    wa.warp_off = d_warp_off[q];
    wa.nwarps = (int)W[q];
    wa.targets_per_lane = walk_class_tpl(q);
    launch_walk(wa, c.cfg, st, c.ls);
  }
  if (split)
  {
    HBT_CUDA(cudaStreamSynchronize(st)); // the exchange may run on another stream (NCCL through the caller's runtime)
    if (c.split_fn(c.split_user, e_stage, T, (void *)st) != 0) throw CudaError{HBTU_ERR_CUDA, "walk split: the all-reduce callback failed"};
    scatter_stage_kernel<<<grid_for(T), kBlock, 0, st>>>(walk_slot, (int)T, e_stage, c.d_E);
    HBT_CHECK_LAUNCH();
    c.ls.launches++;
  }
  HBT_CUDA(cudaEventRecord(c.ev[2], st));
  HBT_CUDA(cudaEventRecord(c.ev_ph[4], st));

  count_bound_kernel<<<grid_for(T, kBlock * 4), kBlock, 0, st>>>(d_segs, tgt_seg, tgt_slot, (int)T, c.d_E, c.d_subs);
  HBT_CHECK_LAUNCH();
  state1_kernel<<<grid_for(nseg), kBlock, 0, st>>>(d_segs, nseg, c.d_subs, c.d_pos, c.cfg);
  HBT_CHECK_LAUNCH();
  // partition + E-sorts in one radix sort
  uint64_t *key_a = ar.alloc<uint64_t>(T), *key_b = ar.alloc<uint64_t>(T);
  int *val_a = ar.alloc<int>(T), *val_b = ar.alloc<int>(T);
  int *bscan = nullptr, *fpos = nullptr, *rho_new = nullptr;
  if (any_hoare || rho)
  {
    int *uflag = ar.alloc<int>(T), *bflag = ar.alloc<int>(T), *uscan = ar.alloc<int>(T);
    bscan = ar.alloc<int>(T);
    fpos = ar.alloc<int>(T);
    hoare_clear_kernel<<<grid_for(T), kBlock, 0, st>>>((int)T, uflag, bflag);
    HBT_CHECK_LAUNCH();
    hoare_flags_kernel<<<grid_for(T), kBlock, 0, st>>>(d_segs, tgt_seg, tgt_slot, (int)T, c.d_E, c.d_subs, rho, uflag, bflag);
    HBT_CHECK_LAUNCH();
    size_t sb = 0;
    HBT_CUDA(cub::DeviceScan::InclusiveSum(nullptr, sb, uflag, uscan, (int)T, st));
    void *stmp = ar.alloc<char>((int64_t)sb);
    HBT_CUDA(cub::DeviceScan::InclusiveSum(stmp, sb, uflag, uscan, (int)T, st));
    HBT_CUDA(cub::DeviceScan::InclusiveSum(stmp, sb, bflag, bscan, (int)T, st));
    hoare_fpos_kernel<<<grid_for(T), kBlock, 0, st>>>(d_segs, tgt_seg, (int)T, uflag, uscan, fpos);
    HBT_CHECK_LAUNCH();
    c.ls.launches += 7;
    if (rho)
    {
      rho_new = ar.alloc<int>(T);
      rho_update_kernel<<<grid_for(T), kBlock, 0, st>>>(d_segs, tgt_seg, tgt_slot, (int)T, c.d_E, c.d_subs, c.d_ids, bscan, fpos, rho, rho_new);
      HBT_CHECK_LAUNCH();
      c.ls.launches++;
    }
  }
  HBT_CUDA(cudaEventRecord(c.ev_ph[5], st));
  sort_keys_kernel<<<grid_for(T), kBlock, 0, st>>>(d_segs, tgt_seg, tgt_slot, (int)T, c.d_E, c.d_subs, bscan, fpos, rho, key_a, val_a);
  HBT_CHECK_LAUNCH();
  c.ls.launches += 3;
  {
    int bits = 33;
    while ((1ll << (bits - 32)) < 2ll * nseg) bits++;
    cub::DoubleBuffer<uint64_t> dk(key_a, key_b);
    cub::DoubleBuffer<int> dv(val_a, val_b);
    size_t tb = 0;
    HBT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)T, 0, bits, st));
    void *tmp = ar.alloc<char>((int64_t)tb);
    HBT_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (int)T, 0, bits, st));
    c.ls.launches += 1 + (bits + 7) / 8;
    int *tmp_id = ar.alloc<int>(T), *tmp_rho = rho ? ar.alloc<int>(T) : nullptr;
    float *tmp_E = ar.alloc<float>(T);
    permute_read_kernel<<<grid_for(T), kBlock, 0, st>>>(dv.Current(), tgt_slot, (int)T, c.d_ids, c.d_E, tmp_id, tmp_E, rho_new, tmp_rho);
    HBT_CHECK_LAUNCH();
    permute_write_kernel<<<grid_for(T), kBlock, 0, st>>>(tgt_slot, (int)T, tmp_id, tmp_E, c.d_ids, c.d_E, d_segs, tgt_seg, tmp_rho, rho);
    HBT_CHECK_LAUNCH();
    c.ls.launches += 2;
  }
  HBT_CUDA(cudaEventRecord(c.ev_ph[6], st));
  static_assert(kBlock == kSegBlock, "seg_reduce.cuh blocks");
  std::vector<int> chunk_off;
  const int nchunk = seg_chunk_table(nseg, [&](int a) { return segs[a].tgt_n; }, chunk_off);
  int *d_chunk_off = upload(c, chunk_off);
  double *partial = ar.alloc<double>((int64_t)nchunk * 7);
  if (nchunk > 0)
  {
    frame_reduce_kernel<<<nchunk, kBlock, 0, st>>>(d_segs, nseg, d_chunk_off, c.d_ids, c.d_pos, c.d_vel, c.d_subs, c.cfg, partial);
    HBT_CHECK_LAUNCH();
  }
  seg_finish_kernel<7><<<nseg, kBlock, 0, st>>>(d_segs, nseg, d_chunk_off, partial, c.d_subs);
  HBT_CHECK_LAUNCH();
  state2_kernel<<<grid_for(nseg), kBlock, 0, st>>>(d_segs, nseg, c.d_subs, c.d_ids, c.d_pos, c.d_vel, c.cfg);
  HBT_CHECK_LAUNCH();
  HBT_CUDA(cudaEventRecord(c.ev_ph[7], st));
  if (nchunk > 0)
  { // `partial` is reused: state2 has consumed the frame sums
    kinematics_kernel<<<nchunk, kBlock, 0, st>>>(d_segs, nseg, d_chunk_off, c.d_ids, c.d_E, c.d_pos, c.d_vel, c.d_subs, c.cfg, partial);
    HBT_CHECK_LAUNCH();
  }
  seg_finish_kernel<6><<<nseg, kBlock, 0, st>>>(d_segs, nseg, d_chunk_off, partial, c.d_subs);
  HBT_CHECK_LAUNCH();
  c.ls.launches += 2;
  RoundResult *d_res = ar.alloc<RoundResult>(nseg);
  finalize_kernel<<<grid_for(nseg), kBlock, 0, st>>>(d_segs, nseg, c.d_subs, c.d_ids, c.d_pos, c.d_vel, d_res);
  HBT_CHECK_LAUNCH();
  c.ls.launches += 4;
  HBT_CUDA(cudaEventRecord(c.ev[3], st));
  HBT_CUDA(cudaEventRecord(c.ev_ph[8], st));
  const RoundResult *res = static_cast<const RoundResult *>(readback_buffer(c, sizeof(RoundResult) * nseg));
  HBT_CUDA(cudaMemcpyAsync(const_cast<RoundResult *>(res), d_res, sizeof(RoundResult) * nseg, cudaMemcpyDeviceToHost, st));
  HBT_CUDA(cudaStreamSynchronize(st));
  c.ring_used = 0; // every staged table of this round has been consumed
  float ms = 0;
  cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]);
  c.stats.build_ms += ms;
  cudaEventElapsedTime(&ms, c.ev[1], c.ev[2]);
  c.stats.walk_ms += ms;
  cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]);
  c.stats.other_ms += ms;
  for (int k = 0; k < 8; k++)
    if (cudaEventElapsedTime(&ms, c.ev_ph[k], c.ev_ph[k + 1]) == cudaSuccess) c.stats.phase_ms[k] += ms;
  c.stats.rounds++;
  c.stats.tree_builds += nseg;
  c.stats.walk_targets += T;
  c.stats.tree_sources += S;

  // results back into the host records (chunks of segments on several threads), then the next round's list in order
  std::vector<uint8_t> fate(nseg); // 0 finished, 1 active in the next round, 2 finished and to be refined
  {
    SubHost *subs_w = c.subs.data();
    const bool refine = c.params.refine_mostbound_particle && c.cfg.max_sample > 0;
    uint8_t *fp = fate.data();
    run_chunks(chunks, [=](int, int64_t a0, int64_t a1) {
      for (int64_t a = a0; a < a1; a++)
      {
        SubHost &h = subs_w[act[a]];
        h.nbound = res[a].nbound;
        h.nlast = res[a].nlast;
        h.correction = res[a].correction;
        h.iterations++;
        uint8_t f = 1;
        if (res[a].status != kActive)
        {
          h.done = true;
          h.disrupted = res[a].status == kDisrupted;
          f = (res[a].status == kConverged && refine && h.nbound > max_sample) ? 2 : 0;
        }
        fp[a] = f;
      }
    });
  }
  std::vector<int> next;
  for (int a = 0; a < nseg; a++)
  {
    if (fate[a] == 1) next.push_back(active[a]);
    else if (fate[a] == 2) c.refine_list.push_back(active[a]);
  }
  active.swap(next);
}

__global__ void refine_keys_kernel(const Segment *__restrict__ segs, const int *__restrict__ tgt_seg, int T, const float *__restrict__ einner,
                                   uint64_t *__restrict__ key, int *__restrict__ val)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  key[t] = ((uint64_t)tgt_seg[t] << 32) | float_to_ordered(einner[t]);
  val[t] = t;
}
__global__ void refine_mostbound_kernel(const Segment *__restrict__ segs, int nseg, SubState *__restrict__ subs, const int *__restrict__ ids,
                                        const float4 *__restrict__ pos, const float4 *__restrict__ vel)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nseg) return;
  SubState &st = subs[segs[a].sub];
  int id0 = ids[st.slot_base];
  float4 x0 = pos[id0], v0 = vel[id0];
  st.mb_pos[0] = x0.x; st.mb_pos[1] = x0.y; st.mb_pos[2] = x0.z;
  st.mb_vel[0] = v0.x; st.mb_vel[1] = v0.y; st.mb_vel[2] = v0.z;
}

// RefineBindingEnergyOrder (src/subhalo_unbind.cpp:234-262) for every subhalo of the level that converged with
// Nbound > MaxSampleSize: tree of its MaxSampleSize most-bound particles (unscaled masses), their binding energies among
// themselves in the final frame, and the first MaxSampleSize Elist entries (pid AND the original E) re-ordered by that.
static void run_refine(Context &c, const std::vector<int> &list)
{
  const int nseg = (int)list.size(), M = (int)c.cfg.max_sample;
  std::vector<Segment> segs(nseg);
  std::vector<int> off(nseg + 1), woff(nseg + 1);
  for (int a = 0; a < nseg; a++)
  {
    Segment &sg = segs[a];
    std::memset(&sg, 0, sizeof(sg));
    sg.slot_base = c.subs[list[a]].slot_base;
    sg.sub = list[a];
    sg.mode = kWalkRefine;
    sg.tree_n = sg.tgt_n = M;
    sg.tree_off = sg.tgt_off = a * M;
    sg.warp_off = a * ((M + 31) / 32);
    sg.keep_order = 1;
    sg.mass_factor = 1.f;
    off[a] = a * M;
    woff[a] = sg.warp_off;
  }
  const int64_t S = (int64_t)nseg * M;
  if (S > 0x3fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "refine round larger than 2^30 particles"};
  off[nseg] = (int)S;
  woff[nseg] = nseg * ((M + 31) / 32);
  Arena &ar = c.arena;
  ar.reset();
  ar.reserve(tree_arena_bytes(S, nseg) + S * 96 + (int64_t)nseg * 128);
  cudaStream_t st = c.stream;
  Segment *d_segs = upload(c, segs);
  int *d_off = upload(c, off), *d_woff = upload(c, woff);
  TreeArrays tr;
  tr.S = (int)S;
  tr.nseg = nseg;
  tr.tree_off = d_off;
  tr.h_tree_off = off.data();
  tr.tpos = ar.alloc<float4>(S);
  tr.ts_seg = ar.alloc<int>(S);
  tr.bbox = ar.alloc<uint32_t>(6 * (int64_t)nseg);
  gather_src_kernel<<<grid_for(S), kBlock, 0, st>>>(d_segs, d_off, nseg, (int)S, c.d_ids, c.d_pos, tr.tpos, tr.ts_seg);
  HBT_CHECK_LAUNCH();
  launch_init_bbox(tr.bbox, nseg, st, c.ls);
  launch_bbox(tr.tpos, tr.ts_seg, (int)S, tr.bbox, st, c.ls);
  build_trees(tr, ar, c.cfg, st, c.ls);
  float4 *tgt_pm = ar.alloc<float4>(S);
  int64_t *tgt_slot = ar.alloc<int64_t>(S);
  int *tgt_seg = ar.alloc<int>(S);
  targets_kernel<<<grid_for(S), kBlock, 0, st>>>(d_segs, d_off, nseg, (int)S, tr.spos, c.d_ids, c.d_pos, tgt_pm, tgt_slot, tgt_seg);
  HBT_CHECK_LAUNCH();
  float *einner = ar.alloc<float>(S);
  WalkArgs wa{};
  wa.node_xm = tr.node_xm;
  wa.node_aux = tr.node_aux;
  wa.cellcount = tr.cellcount;
  wa.tree_off = d_off;
  wa.segs = d_segs;
  wa.warp_off = d_woff;
  wa.nseg = nseg;
  wa.nwarps = woff[nseg];
  wa.targets_per_lane = 1;
  wa.tgt_pm = tgt_pm;
  wa.tgt_slot = tgt_slot;
  wa.ids = c.d_ids;
  wa.vel = c.d_vel;
  wa.E = c.d_E;
  wa.subs = c.d_subs;
  wa.out_f = einner;
  wa.counters = c.count_interactions ? c.d_counters : nullptr;
  launch_walk(wa, c.cfg, st, c.ls);
  uint64_t *ka = ar.alloc<uint64_t>(S), *kb = ar.alloc<uint64_t>(S);
  int *va = ar.alloc<int>(S), *vb = ar.alloc<int>(S);
  refine_keys_kernel<<<grid_for(S), kBlock, 0, st>>>(d_segs, tgt_seg, (int)S, einner, ka, va);
  HBT_CHECK_LAUNCH();
  int bits = 33;
  while ((1ll << (bits - 32)) < nseg) bits++;
  cub::DoubleBuffer<uint64_t> dk(ka, kb);
  cub::DoubleBuffer<int> dv(va, vb);
  size_t tb = 0;
  HBT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)S, 0, bits, st));
  void *tmp = ar.alloc<char>((int64_t)tb);
  HBT_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (int)S, 0, bits, st));
  int *tmp_id = ar.alloc<int>(S);
  float *tmp_E = ar.alloc<float>(S);
  permute_read_kernel<<<grid_for(S), kBlock, 0, st>>>(dv.Current(), tgt_slot, (int)S, c.d_ids, c.d_E, tmp_id, tmp_E);
  HBT_CHECK_LAUNCH();
  permute_write_kernel<<<grid_for(S), kBlock, 0, st>>>(tgt_slot, (int)S, tmp_id, tmp_E, c.d_ids, c.d_E);
  HBT_CHECK_LAUNCH();
  refine_mostbound_kernel<<<grid_for(nseg), kBlock, 0, st>>>(d_segs, nseg, c.d_subs, c.d_ids, c.d_pos, c.d_vel);
  HBT_CHECK_LAUNCH();
  c.ls.launches += 8 + (bits + 7) / 8;
  c.stats.tree_builds += nseg;
  c.stats.walk_targets += S;
  HBT_CUDA(cudaStreamSynchronize(st));
}

// HBTU_TRACE=1: host timestamps (ms since the start of hbtu_execute) at the stations of a step, on stderr
static bool trace_on()
{
  static const bool on = getenv("HBTU_TRACE") != nullptr;
  return on;
}
#define HBT_TRACE(t0, ...)                                                                                                  \
  do                                                                                                                        \
  {                                                                                                                         \
    if (trace_on())                                                                                                         \
    {                                                                                                                       \
      fprintf(stderr, "[hbtu %8.2f ms] ", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - (t0)).count()); \
      fprintf(stderr, __VA_ARGS__);                                                                                         \
      fprintf(stderr, "\n");                                                                                                \
    }                                                                                                                       \
  } while (0)

void wait_upload_wave(Context &c, int wave)
{
  std::unique_lock<std::mutex> lk(c.up_m);
  c.up_cv.wait(lk, [&] { return c.up_wave_done >= wave || !c.uploader.joinable(); });
  if (!c.up_error.empty()) throw CudaError{HBTU_ERR_CUDA, c.up_error};
}
void finish_upload(Context &c)
{
  if (c.uploader.joinable()) c.uploader.join();
}

void execute_batch(Context &c)
{
  const auto trace_t0 = std::chrono::steady_clock::now();
  if (!c.staged) throw CudaError{HBTU_ERR_INVALID, "hbtu_execute before hbtu_stage"};
  cudaStream_t st = c.stream;
  const int nsub = (int)c.nsub;
  c.ls.launches = 0;
  std::memset(&c.stats, 0, offsetof(hbtu_stats, h2d_ms)); // the copy statistics of the staging call survive
  c.stats.tree_sources = 0;
  c.stats.walk_fallbacks = 0;
  for (double &x : c.stats.phase_ms) x = 0.0;
  HBT_CUDA(cudaEventRecord(c.ev_exec[0], st));
  if (c.count_interactions) HBT_CUDA(cudaMemsetAsync(c.d_counters, 0, kWalkCounters * sizeof(unsigned long long), st));

  // (re)initialise per-subhalo state from the staged inputs: written straight into the staging ring, on several host threads
  // (one thread needs ~0.3 us per subhalo for this loop: 150 ms of idle GPU at the 5e5 subhaloes of an EAGLE-shaped shard)
  if (nsub > 0)
  {
    char *staged = ring_reserve(c, sizeof(SubState) * (size_t)nsub);
    SubState *init = reinterpret_cast<SubState *>(staged);
    SubHost *subs = c.subs.data();
    const hbtu_sub_io *io_in = c.io_in.data();
    parallel_ranges(nsub, 1 << 14, [=](int64_t s0, int64_t s1) {
      for (int64_t s = s0; s < s1; s++)
      {
        SubHost &h = subs[s];
        const hbtu_sub_io &io = io_in[s];
        SubState z;
        std::memset(&z, 0, sizeof(z));
        z.slot_base = h.slot_base;
        z.part_begin = h.part_begin;
        z.n_own = h.n_own;
        z.n_src = h.n_own;
        z.status = kPending;
        z.death = io.snapshot_index_of_death;
        z.sink = io.snapshot_index_of_sink;
        z.sinktrack = io.sink_track_id;
        // the orphan rule belongs to RecursiveUnbind (src/subhalo_unbind.cpp:434-446); a subhalo the caller enters through plain
        // Unbind (HBTU_SUB_PLAIN_UNBIND) is unbound like any other whatever its entry Nbound
        const bool orphan = io.nbound <= 1 && !(io.flags & HBTU_SUB_PLAIN_UNBIND);
        z.is_orphan = orphan;
        for (int j = 0; j < 3; j++)
        {
          z.ref_pos[j] = (float)io.avg_pos[j];
          z.ref_vel[j] = (float)io.avg_vel[j];
          z.mb_pos[j] = (float)io.mostbound_pos[j];
          z.mb_vel[j] = (float)io.mostbound_vel[j];
          z.am[j] = io.specific_angular_momentum[j];
        }
        z.spec_pot = io.specific_self_potential_energy;
        z.spec_kin = io.specific_self_kinetic_energy;
        z.mbound = io.mbound;
        z.nbound = (int)io.nbound;
        init[s] = z;
        h.n_src = h.n_own;
        h.nbound = h.nlast = 0;
        h.correction = 0;
        h.done = h.disrupted = false;
        h.iterations = 0;
        h.is_orphan = orphan;
      }
    });
    ring_commit(c, c.d_subs, staged, sizeof(SubState) * (size_t)nsub); // d_subs holds nsub + 1 records: room for the 16-byte granularity
  }
  if (c.N > 0)
  {
    init_ids_kernel<<<grid_for(c.N), kBlock, 0, st>>>(c.d_part_offset, c.d_slot_base, nsub, c.N, c.d_ids);
    HBT_CHECK_LAUNCH();
    c.ls.launches++;
  }
  HBT_TRACE(trace_t0, "state + ids enqueued");
  HBT_CUDA(cudaStreamSynchronize(st));
  HBT_TRACE(trace_t0, "state + ids on the device (copy stream %s)", cudaStreamQuery(c.copy_stream) == cudaSuccess ? "idle" : "busy");

  // asynchronous staging (hbtu_unbind_batch): the first wave of the upload is needed by the first round, the dominant root only
  // by level 0; a re-execution of the same staged batch finds both events completed
  if (c.waves_pending) wait_upload_wave(c, 1);
  HBT_TRACE(trace_t0, "upload wave 1 landed");
  for (int level = c.max_depth; level >= 0; level--)
  {
    if (level == 0 && c.waves_pending)
    {
      wait_upload_wave(c, 2);
      c.waves_pending = false;
      HBT_TRACE(trace_t0, "upload wave 2 landed");
    }
    const std::vector<int> &lv = c.levels[level];
    const int64_t M = c.cfg.max_sample;
    HBT_TRACE(trace_t0, "level %d: %zu subhaloes, %lld rounds so far (copy stream %s)", level, lv.size(), (long long)c.stats.rounds,
              cudaStreamQuery(c.copy_stream) == cudaSuccess ? "idle" : "busy");
    c.arena.reset(); // the previous level's rounds are over: the arena holds this level's small job tables until its first round
    // 1. feed children's unbound tails into this level's sources (src/subhalo_unbind.cpp:437-443)
    {
      std::vector<CopyJob> jobs;
      std::vector<int64_t> job_off{0};
      for (int s : lv)
      {
        SubHost &h = c.subs[s];
        int64_t n = h.n_own;
        for (int ch : h.children)
        {
          SubHost &k = c.subs[ch];
          int64_t tail = k.n_src - k.nbound;
          if (tail > 0)
          {
            jobs.push_back(CopyJob{h.slot_base + n, k.slot_base + k.nbound});
            job_off.push_back(job_off.back() + tail);
            n += tail;
          }
        }
        h.n_src = (int)n;
      }
      run_seg_copy(c, jobs, job_off, c.d_ids, c.d_ids);
    }
    // 2. classify
    std::vector<int> active, trivial;
    std::vector<ShuffleJob> shuf;
    std::vector<int64_t> shuf_off{0};
    std::vector<LevelInit> li(lv.size());
    for (size_t i = 0; i < lv.size(); i++)
    {
      int s = lv[i];
      SubHost &h = c.subs[s];
      int nu = h.is_orphan ? h.n_own : h.n_src;
      bool shuffled = false;
      if (nu < 2 || nu < c.cfg.min_num_part)
      {
        trivial.push_back(s);
        h.nbound = nu == 0 ? 0 : 1;
        h.done = true;
        h.disrupted = nu >= 2;
      }
      else
      {
        active.push_back(s);
        h.nbound = h.nlast = nu;
        if (M > 0 && nu > M)
        { // random_shuffle of the source before sampling (src/subhalo_unbind.cpp:302)
          shuffled = true;
          shuf.push_back(ShuffleJob{h.slot_base, s + (int)c.sub_index_base, nu}); // the key uses the subhalo's index in the CALLER's batch
          shuf_off.push_back(shuf_off.back() + nu);
        }
      }
      li[i] = LevelInit{s, h.n_src, h.done ? 0 : 1, shuffled ? 1 : 0};
    }
    { // orphans keep their list in the order it had BEFORE the sampling shuffle: RecursiveUnbind swaps the untouched
      // extended list back in after Unbind worked on the backup (src/subhalo_unbind.cpp:436,444-446)
      std::vector<CopyJob> snap;
      std::vector<int64_t> snap_off{0};
      for (int s : lv)
        if (c.subs[s].is_orphan && c.subs[s].n_src > 0)
        {
          snap.push_back(CopyJob{c.subs[s].slot_base, c.subs[s].slot_base});
          snap_off.push_back(snap_off.back() + c.subs[s].n_src);
        }
      run_seg_copy(c, snap, snap_off, c.d_ids, c.d_ids_orig);
    }
    { // n_src, the entry value of Particles[0] and the active flag on the device
      LevelInit *d_li = upload(c, li);
      level_init_kernel<<<grid_for((int64_t)li.size()), kBlock, 0, st>>>(d_li, (int)li.size(), c.d_subs, c.d_ids);
      HBT_CHECK_LAUNCH();
      c.ls.launches++;
    }
    if (!shuf.empty())
    {
      const int64_t total = shuf_off.back();
      if (total > 0x3fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "shuffle larger than 2^30 particles"};
      Arena &ar = c.arena;
      ar.reset();
      ar.reserve(total * 40 + (int64_t)shuf.size() * 64 + (64 << 20));
      ShuffleJob *d_jobs = upload(c, shuf);
      int64_t *d_off = upload(c, shuf_off);
      uint64_t *ka = ar.alloc<uint64_t>(total), *kb = ar.alloc<uint64_t>(total);
      int *va = ar.alloc<int>(total), *vb = ar.alloc<int>(total), *tmp = ar.alloc<int>(total);
      shuffle_keys_kernel<<<grid_for(total), kBlock, 0, st>>>(d_jobs, d_off, (int)shuf.size(), total, (uint64_t)c.params.shuffle_seed, ka, va);
      HBT_CHECK_LAUNCH();
      int bits = 41;
      while ((1ll << (bits - 40)) < (int64_t)shuf.size()) bits++;
      cub::DoubleBuffer<uint64_t> dk(ka, kb);
      cub::DoubleBuffer<int> dv(va, vb);
      size_t tb = 0;
      HBT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)total, 0, bits, st));
      void *tmpb = ar.alloc<char>((int64_t)tb);
      HBT_CUDA(cub::DeviceRadixSort::SortPairs(tmpb, tb, dk, dv, (int)total, 0, bits, st));
      shuffle_read_kernel<<<grid_for(total), kBlock, 0, st>>>(d_jobs, d_off, (int)shuf.size(), total, dv.Current(), c.d_ids, tmp);
      HBT_CHECK_LAUNCH();
      shuffle_write_kernel<<<grid_for(total), kBlock, 0, st>>>(d_jobs, d_off, (int)shuf.size(), total, tmp, c.d_ids);
      HBT_CHECK_LAUNCH();
      c.ls.launches += 4 + (bits + 7) / 8;
      ar.reset();
    }
    { // snapshot of the (shuffled) input order: what Particles holds when a disrupted Unbind does not reorder it (:361-379)
      std::vector<CopyJob> snap;
      std::vector<int64_t> snap_off{0};
      for (int s : lv)
        if (!c.subs[s].is_orphan && c.subs[s].n_src > 0)
        {
          snap.push_back(CopyJob{c.subs[s].slot_base, c.subs[s].slot_base});
          snap_off.push_back(snap_off.back() + c.subs[s].n_src);
        }
      run_seg_copy(c, snap, snap_off, c.d_ids, c.d_ids_orig);
    }
    if (!trivial.empty())
    {
      int *d_list = upload(c, trivial);
      trivial_kernel<<<grid_for((int64_t)trivial.size()), kBlock, 0, st>>>(d_list, (int)trivial.size(), c.d_subs, c.d_ids, c.d_pos, c.cfg, c.d_E);
      HBT_CHECK_LAUNCH();
      c.ls.launches++;
    }
    // 3. iterate; then re-rank the most-bound sample of converged sampled subhaloes (RefineBindingEnergyOrder)
    c.refine_list.clear();
    while (!active.empty()) run_round(c, active);
    if (!c.refine_list.empty()) run_refine(c, c.refine_list);
    // 4. restore the input order where the reference leaves Particles untouched: disrupted subhaloes
    //    (no permutation copy, :361-379) and orphans (the unbound backup is discarded, :444-446)
    {
      std::vector<CopyJob> jobs;
      std::vector<int64_t> job_off{0};
      std::vector<int> job_sub;
      c.arena.reset(); // the level's rounds are over
      for (int s : lv)
      {
        SubHost &h = c.subs[s];
        if ((h.disrupted || h.is_orphan) && h.n_src > 1)
        {
          jobs.push_back(CopyJob{h.slot_base, h.slot_base});
          job_off.push_back(job_off.back() + h.n_src);
          job_sub.push_back(s);
        }
      }
      run_seg_copy(c, jobs, job_off, c.d_ids_orig, c.d_ids);
      if (M > 0 && !jobs.empty())
      {
        Arena &ar = c.arena;
        CopyJob *d_jobs = upload(c, jobs);
        int64_t *d_off = upload(c, job_off);
        int *d_sub = upload(c, job_sub);
        restore_front_kernel<<<grid_for(job_off.back()), kBlock, 0, st>>>(d_jobs, d_off, (int)jobs.size(), job_off.back(), d_sub, c.d_subs,
                                                                           c.d_ids_orig, c.d_ids);
        HBT_CHECK_LAUNCH();
        c.ls.launches++;
      }
    }
  }
  if (c.count_interactions)
  {
    unsigned long long cnt[kWalkCounters];
    HBT_CUDA(cudaMemcpy(cnt, c.d_counters, sizeof(cnt), cudaMemcpyDeviceToHost));
    c.stats.pair_interactions = (int64_t)cnt[0];
    c.stats.nodes_visited = (int64_t)cnt[1];
    c.stats.walk_fallbacks = (int64_t)cnt[2];
  }
  HBT_CUDA(cudaEventRecord(c.ev_exec[1], st));
  HBT_CUDA(cudaEventSynchronize(c.ev_exec[1]));
  HBT_TRACE(trace_t0, "done");
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev_exec[0], c.ev_exec[1]);
    c.stats.execute_ms = ms;
  }
  if (c.staged_async)
  { // duration of the asynchronous upload (it ran behind the kernels of the deeper levels)
    finish_upload(c);
    c.stats.h2d_ms = c.up_ms;
  }
  c.stats.kernel_launches = c.ls.launches;
  c.executed = true;
}

void fetch_batch(Context &c, hbtu_sub_io *io, int64_t order_capacity, int64_t *order_offset, int32_t *order_out, float *energy_out)
{
  if (!c.executed) throw CudaError{HBTU_ERR_INVALID, "hbtu_fetch before hbtu_execute"};
  const auto trace_t0 = std::chrono::steady_clock::now();
  cudaStream_t st = c.stream;
  const int nsub = (int)c.nsub;
  cudaEvent_t e0 = c.ev[0], e1 = c.ev[1];
  HBT_CUDA(cudaEventRecord(e0, st));
  const SubState *fin = static_cast<const SubState *>(readback_buffer(c, sizeof(SubState) * (size_t)std::max(nsub, 1)));
  HBT_TRACE(trace_t0, "fetch: readback buffer ready");
  HBT_CUDA(cudaMemcpyAsync(const_cast<SubState *>(fin), c.d_subs, sizeof(SubState) * nsub, cudaMemcpyDeviceToHost, st));
  HBT_CUDA(cudaStreamSynchronize(st));
  HBT_TRACE(trace_t0, "fetch: %zu bytes of SubState records on the host", sizeof(SubState) * (size_t)nsub);
  std::vector<int64_t> out_off(nsub + 1, 0), slot_base(nsub);
  std::vector<int> nb(nsub);
  const bool truncate = (c.flags & HBTU_FLAG_TRUNCATE_SOURCE) != 0;
  const float relax = c.cfg.relax_factor;
  const SubHost *subs = c.subs.data();
  const hbtu_sub_io *io_in = c.io_in.data();
  auto convert = [&](int s0, int s1) {
    for (int s = s0; s < s1; s++)
    {
      const SubState &z = fin[s];
      const SubHost &h = subs[s];
      int64_t full = h.n_src, ns = full;
      if (truncate)
      { // Subhalo_t::TruncateSource, src/subhalo_unbind.cpp:449-458 (int*float -> float -> int)
        int64_t nsrc = z.nbound <= 1 ? z.nbound : (int64_t)((float)z.nbound * relax);
        if (nsrc > full) nsrc = full;
        ns = nsrc;
      }
      out_off[s + 1] = ns; // lengths here, offsets after the prefix sum below
      slot_base[s] = h.slot_base;
      nb[s] = z.nbound;
      hbtu_sub_io &o = io[s];
      for (int j = 0; j < 3; j++)
      {
        o.avg_pos[j] = z.ref_pos[j];
        o.avg_vel[j] = z.ref_vel[j];
        o.mostbound_pos[j] = z.mb_pos[j];
        o.mostbound_vel[j] = z.mb_vel[j];
        o.specific_angular_momentum[j] = z.am[j];
      }
      o.nbound = z.nbound;
      o.sink_track_id = z.sinktrack;
      o.snapshot_index_of_death = z.death;
      o.snapshot_index_of_sink = z.sink;
      o.mbound = z.mbound;
      o.specific_self_potential_energy = z.spec_pot;
      o.specific_self_kinetic_energy = z.spec_kin;
      o.nsource_full = full;
      o.nsource = ns;
      o.iterations = z.iterations;
      o.flags = io_in[s].flags;
    }
  };
  // ~700 bytes of host memory traffic per subhalo: a snapshot with 5e5 subhaloes per GPU (EAGLE-shaped shard) spent 165 ms
  // here on one thread
  parallel_ranges(nsub, 1 << 14, [&](int64_t s0, int64_t s1) { convert((int)s0, (int)s1); });
  for (int s = 0; s < nsub; s++) out_off[s + 1] += out_off[s];
  const int64_t total = out_off[nsub];
  HBT_TRACE(trace_t0, "fetch: records read back and converted (%d subhaloes)", nsub);
  if (total > order_capacity) throw CudaError{HBTU_ERR_CAPACITY, "order_out too small"};
  std::memcpy(order_offset, out_off.data(), sizeof(int64_t) * (nsub + 1));
  if (total > 0)
  {
    c.arena.reset();
    c.arena.reserve(total * 8 + (int64_t)nsub * 24 + (1 << 20));
    int64_t *d_off = upload(c, out_off), *d_sb = upload(c, slot_base);
    int *d_nb = upload(c, nb);
    int *d_out = c.arena.alloc<int>(total);
    float *d_oe = energy_out ? c.arena.alloc<float>(total) : nullptr;
    pack_output_kernel<<<grid_for(total), kBlock, 0, st>>>(d_off, d_sb, d_nb, nsub, total, c.d_ids, c.d_E, d_out, d_oe, (int)c.order_index_base);
    HBT_CHECK_LAUNCH();
    if (trace_on())
    {
      HBT_CUDA(cudaStreamSynchronize(st));
      HBT_TRACE(trace_t0, "fetch: particle orders packed on the device (%lld entries)", (long long)total);
    }
    HBT_CUDA(cudaMemcpyAsync(order_out, d_out, sizeof(int) * total, cudaMemcpyDeviceToHost, st));
    if (energy_out) HBT_CUDA(cudaMemcpyAsync(energy_out, d_oe, sizeof(float) * total, cudaMemcpyDeviceToHost, st));
    c.stats.d2h_bytes = total * (energy_out ? 8 : 4) + (int64_t)nsub * sizeof(SubState);
  }
  HBT_CUDA(cudaEventRecord(e1, st));
  HBT_CUDA(cudaStreamSynchronize(st));
  HBT_TRACE(trace_t0, "fetch: done");
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  c.stats.d2h_ms = ms;
}

} // namespace hbt
