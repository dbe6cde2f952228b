// device_tree.cuh - data layout in HBM of one unbinding round and the kernel launch API shared by
// tree_build.cu / walk.cu / unbind_batch.cu.
//
// One ROUND processes every still-active subhalo of the current nesting level at once.  All per-round
// arrays are concatenations over the active subhaloes ("segments"):
//
//   S-arrays (tree sources, length S = sum tree_n):  tpos  float4 (x,y,z,m) gathered source particles
//                                                    key   uint64 63-bit octal key (tree_core.cuh)
//                                                    spos  float4 sources in key order
//                                                    cells (l,r,depth) per adjacent pair, depthmask, cellcount
//                                                    msum  double4 segmented prefix sums of m, m*(x-centre)
//   node arrays (length <= 2S-1):                    node_xm  float4 (x,y,z,m)  particle or cell CoM/mass
//                                                    node_aux float2 (len^2/theta^2 [0 for particles], end index bits)
//   T-arrays (walk targets, length T = sum tgt_n):   tgt_pm float4 (x,y,z,self mass), tgt_slot int64
//
// Algorithmic bytes per unit: a source particle costs 16 B (gather) + 16+12 B (key, sort
// payload) + 16 B (sorted copy) + 32 B (moment scan) + 24 B (node) per build; a walk target reads 16 B and
// writes 4 B; everything else is per-interaction fp32 issue work.
#pragma once
#include "common.cuh"

namespace hbt
{

struct double4s
{
  double m, x, y, z;
};

// per active subhalo of a round; uploaded by the host planner every round
struct Segment
{
  int64_t slot_base; // first slot of this subhalo in ids/E
  int sub;           // subhalo index in the batch
  int mode;          // 0 = full evaluation (tree == targets), 1 = correction (tree = removed, targets = bound)
  int tree_first;    // first Elist index of the tree sources
  int tree_n;
  int tgt_n;
  int tree_off; // offset in the S-concatenation
  int tgt_off;  // offset in the T-concatenation
  int warp_off; // first walk warp of this segment
  int keep_order;    // sampled mode with Nlast > MaxSampleSize: the Elist order is the sample, do not re-order it
  float mass_factor; // MassFactor of the sampled tree: Nlast/MaxSampleSize (src/subhalo_unbind.cpp:335-339), else 1
};

enum WalkMode
{
  kWalkUnbindFull = 0,   // E = 0.5|dv|^2 + pot            (src/subhalo_unbind.cpp:341-354)
  kWalkUnbindCorrect = 1, // E += v_old.dv + dK - pot_removed (src/subhalo_unbind.cpp:312-330)
  kWalkPotential = 2,    // out = pot                      (GravityTree_t::EvaluatePotential)
  kWalkBindingEnergy = 3, // out = 0.5|dv|^2 + pot          (GravityTree_t::BindingEnergy)
  kWalkRefine = 4         // out_f = self-binding energy among the MaxSampleSize most bound (RefineBindingEnergyOrder, :234-262)
};

// Per-subhalo iteration state kept on the device for the whole batch.
struct SubState
{
  int64_t slot_base;  // offset of this subhalo's Elist in ids/E
  int64_t part_begin; // first input particle
  int n_own;          // input particles
  int n_src;          // Particles.size() seen by Unbind (own + children's unbound tails)
  int nbound, nlast;
  int status;     // SubStatus
  int correction; // sticky CorrectionLoop flag
  int iterations;
  int death, sink;
  int is_orphan;
  int first_id; // input index of Particles[0] on entry (OldMostboundParticle, src/subhalo_unbind.cpp:298)
  int shuffled; // the source was permuted for sampling
  int64_t sinktrack;
  float ref_pos[3], ref_vel[3];         // ComovingAveragePosition / PhysicalAverageVelocity (RefPos/RefVel)
  float old_ref_pos[3], old_ref_vel[3]; // OldRefPos / OldRefVel
  float mb_pos[3], mb_vel[3];           // ComovingMostBoundPosition / PhysicalMostBoundVelocity
  float mbound, spec_pot, spec_kin, am[3];
  float ref_diff[3], dK; // RefVelDiff and dK of the current correction round
  int count_bound;       // scratch: number of E<0 among this round's targets
  int origin_id;         // periodic: particle that is Elist[0] in the reference's order (origin of AveragePosition, :152-154)
  int hoare_last;        // scratch: largest reference index in [1,Nbound) of a bound entry (0 = none)
  int hoare_first_bound; // scratch: the reference's Elist[0] is bound
  int hoare_nb;          // scratch: entries with E<0 this round (= Nbound unless NO_STRIPPING forces Nbound = Nlast)
  double sums[8]; // scratch: reduction accumulators
};

enum SubStatus
{
  kPending = 0,   // waiting for its nesting level
  kActive = 1,    // iterating
  kConverged = 2, // converged this round (final sort + kinematics pending)
  kDisrupted = 3, // Nbound < MinNumPartOfSub
  kDone = 4       // finished
};

struct TreeArrays
{ // device pointers of one round (arena-owned)
  int S = 0, nseg = 0;
  float4 *tpos = nullptr;
  int *ts_seg = nullptr;
  uint64_t *skey = nullptr;
  int *sperm = nullptr;
  float4 *spos = nullptr;
  int2 *cell_lr = nullptr;
  int8_t *cell_depth = nullptr;
  uint32_t *depthmask = nullptr;
  int *cellcount = nullptr;
  float4 *node_xm = nullptr;
  float2 *node_aux = nullptr;
  SegRoot *roots = nullptr;
  uint32_t *bbox = nullptr; // 6 ordered-uint per segment
  const int *tree_off = nullptr; // [nseg+1] device
  const int *h_tree_off = nullptr; // the same on the host
};

struct LaunchStats
{
  int64_t launches = 0;
};

// tree_build.cu -----------------------------------------------------------------------------------------
// Build the pre-order node arrays for all segments from tpos/ts_seg (already filled, bbox accumulated).
void build_trees(TreeArrays &t, Arena &arena, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls);
int64_t tree_arena_bytes(int64_t S, int64_t nseg);
// accumulate per-segment bounding boxes of tpos into t.bbox (must be pre-initialised by init_bbox)
void launch_init_bbox(uint32_t *bbox, int nseg, cudaStream_t stream, LaunchStats &ls);
void launch_bbox(const float4 *tpos, const int *ts_seg, int S, uint32_t *bbox, cudaStream_t stream, LaunchStats &ls);

// walk.cu -----------------------------------------------------------------------------------------------
struct WalkArgs
{
  const float4 *node_xm;
  const float2 *node_aux;
  const int *cellcount;  // inclusive cell counts (node range of a segment derives from it)
  const int *tree_off;   // [nseg+1]
  const Segment *segs;   // [nseg]
  const int *warp_off;   // [nseg+1]
  int nseg, nwarps;
  int targets_per_lane;  // walk class: 1, 2 or 4 (walk.cu, warps own 32*T consecutive targets) or kWalkGroup2/4 (walk_masked.cu)
  const float4 *tgt_pm;  // [T] x,y,z,self mass
  const int64_t *tgt_slot; // [T] slot in ids/E (unbind modes)
  const int *ids;        // Elist pid per slot
  const float4 *vel;     // input velocities (by particle id) or per-target velocities (kWalkBindingEnergy)
  float *E;              // per slot
  const SubState *subs;
  double *out;           // kWalkPotential / kWalkBindingEnergy
  float *out_f;          // kWalkRefine: per target
  float ref_pos[3], ref_vel[3]; // kWalkBindingEnergy frame
  unsigned long long *counters; // [kWalkCounters]: accepted interactions, warp node visits, groups redone per lane (nullptr = do not count)
  // target split of the walk over cooperating contexts (the non-natural case): this context walks the CTAs
  // whose 16-CTA chunk index is congruent to split_rank modulo split_n and writes the new energies to E_stage[t] (T-order,
  // zero elsewhere) instead of E[slot]; a sum all-reduce over the contexts then completes E_stage everywhere
  int split_rank, split_n;
  float *E_stage;
};
// global CTA index of a walk kernel's block under the target split (identity without it), and the grid that covers `ctas` of them
__device__ __forceinline__ int walk_cta(const WalkArgs &a)
{
  return a.split_n > 1 ? ((((int)blockIdx.x >> 4) * a.split_n + a.split_rank) << 4) + ((int)blockIdx.x & 15) : (int)blockIdx.x;
}
inline int walk_grid(const WalkArgs &a, int ctas)
{
  if (a.split_n <= 1) return ctas;
  const int chunks = (ctas + 15) >> 4;                       // 16-CTA chunks in all
  const int mine = (chunks - a.split_rank + a.split_n - 1) / a.split_n; // chunks split_rank, split_rank + split_n, ...
  return mine << 4;
}
constexpr int kWalkCounters = 4;
void launch_walk(const WalkArgs &a, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls);
// Walk classes.  A segment belongs to one class; every class has its own warp numbering (warp_off) and launch.
constexpr int kWalkGroup2 = 102, kWalkGroup4 = 104; // masked group walk with 64 / 128 targets per warp
constexpr int kWalkSmall = 201;                     // small-subhalo kernel (walk_small.cu): 32 targets per warp, dense sweep of the whole tree
constexpr int kWalkClasses = 5;
struct WalkClass
{
  int index;            // 0..kWalkClasses-1
  int targets_per_lane; // value for WalkArgs::targets_per_lane
  int targets_per_warp;
};
// class used for a segment with tgt_n targets over a tree of tree_n sources (walk_tuning() decides)
WalkClass walk_class(int tgt_n, int tree_n);
int walk_class_tpl(int index); // targets_per_lane value of class `index` (the group class depends on the tuning)

// kernel routing knobs (process-wide; defaults = measured best, HBTU_* environment variables and hbtu_set_tuning override)
#ifndef HBT_MASKED_DEFAULT_PAIRS
#define HBT_MASKED_DEFAULT_PAIRS 1 // measured on the bench (profiles/r02_walk_notes.md): 64-target groups at 7 CTAs/SM beat 128-target groups at 5
#endif
#ifndef HBT_MASKED_DEFAULT_BLOCKS_NP2
#define HBT_MASKED_DEFAULT_BLOCKS_NP2 5
#endif
#ifndef HBT_MASKED_DEFAULT_BLOCKS_NP1
#define HBT_MASKED_DEFAULT_BLOCKS_NP1 7
#endif
#ifndef HBT_SMALL_DEFAULT_MAX
#define HBT_SMALL_DEFAULT_MAX 0
#endif
struct WalkTuning
{
  int forced_tpl;    // HBTU_WALK_TPL: force the per-lane walk with 1, 2 or 4 targets per lane for every segment (0 = off)
  int big4, big2;    // per-lane walk: segments with at least this many targets take 4 / 2 targets per lane
  int group_min;     // HBTU_WALK_GROUP_MIN: segments with at least this many targets use the masked group walk (0 = never)
  int masked_pairs;  // HBTU_WALK_MASKED_PAIRS: slice pairs per warp of the masked walk (1: 64-target groups, 2: 128-target groups)
  int masked_blocks; // HBTU_WALK_MASKED_BLOCKS: resident CTAs per SM the masked kernel variant is compiled for
  int small_max;     // HBTU_WALK_SMALL_MAX: segments with at most this many tree sources use the small-subhalo kernel (0 = never)
};
WalkTuning &walk_tuning();

} // namespace hbt
