/* hbt_unbind.h - C-ABI of the B200-native HBT+ unbinding path.
 *
 * This is the drop-in boundary.  The reference has no
 * FFI for this path; its seam is three C++ member functions, which a host
 * shim (integration/subhalo_unbind_b200.cpp) re-implements by pack -> call ->
 * unpack on top of the entry points declared here:
 *
 *   void SubhaloSnapshot_t::RefineParticles()        src/subhalo.h:245, src/subhalo_unbind.cpp:460-516
 *   void Subhalo_t::Unbind(const Snapshot_t&)        src/subhalo.h:111, src/subhalo_unbind.cpp:263-431
 *   void Subhalo_t::RecursiveUnbind(...)             src/subhalo.h:112, src/subhalo_unbind.cpp:432-447
 *   void Subhalo_t::TruncateSource()                 src/subhalo.h:114, src/subhalo_unbind.cpp:449-458
 *   double GravityTree_t::EvaluatePotential(...)     src/gravity_tree.h:14,  src/gravity_tree.cpp:79-164
 *   double GravityTree_t::BindingEnergy(...)         src/gravity_tree.h:15,  src/gravity_tree.cpp:166-175
 *
 * Plain C, POD only: no STL, no torch types, caller owns every host buffer.
 * Every function returns HBTU_OK (0) or a negative HBTU_ERR_* code; the text of
 * the last error of a context is available from hbtu_last_error().  There is
 * NO CPU fallback: without a CUDA device hbtu_create() fails.
 *
 * The same structs and the same batch signature are implemented by the two
 * CPU checkers used only by tests/bench baselines:
 *   oracle/_ref/libhbtref_v32.so   (the unmodified reference sources, hbtref_*)
 *   oracle/libhbtoracle.so         (plain-C restatement, hbto_*)
 */
#ifndef HBT_UNBIND_H
#define HBT_UNBIND_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HBTU_ABI_VERSION 1

#define HBTU_OK 0
#define HBTU_ERR_INVALID (-1)     /* bad argument / malformed nest forest            */
#define HBTU_ERR_CUDA (-2)        /* a CUDA runtime call or kernel failed            */
#define HBTU_ERR_NOMEM (-3)       /* device or pinned-host allocation failed         */
#define HBTU_ERR_NODEVICE (-4)    /* no usable sm_100 device: there is no fallback   */
#define HBTU_ERR_UNSUPPORTED (-5) /* ABI variant not built (e.g. HBT_REAL8)          */
#define HBTU_ERR_CAPACITY (-6)    /* caller-provided output buffer too small         */

/* POD copy of the Parameter_t fields the path reads (src/config_parser.h:22-123,
 * derived ones from src/config_parser.cpp:95-103) plus PhysicalConst::G.
 * Values are passed as double but must hold the HBTReal-rounded value of the
 * caller's build (float when real_bytes==4), because the reference evaluates
 * e.g. Nlast*BoundMassPrecision in HBTReal arithmetic (src/subhalo_unbind.cpp:395). */
typedef struct hbtu_params
{
  int32_t struct_size;                 /* = sizeof(hbtu_params), ABI check                  */
  int32_t real_bytes;                  /* sizeof(HBTReal) of the caller: 4 (8 unsupported)  */
  int32_t min_num_part_of_sub;         /* MinNumPartOfSub                                   */
  int32_t periodic_boundary_on;        /* PeriodicBoundaryOn                                */
  int32_t refine_mostbound_particle;   /* RefineMostboundParticle                           */
  int32_t device;                      /* CUDA device ordinal this context drives           */
  int64_t max_sample_size;             /* MaxSampleSizeOfPotentialEstimate (0 = exact)      */
  double bound_mass_precision;         /* BoundMassPrecision                                */
  double source_sub_relax_factor;      /* SourceSubRelaxFactor                              */
  double box_size;                     /* BoxSize                                           */
  double box_half;                     /* BoxHalf                                           */
  double softening_halo;               /* SofteningHalo                                     */
  double tree_node_open_angle_square;  /* TreeNodeOpenAngleSquare                           */
  double tree_node_resolution;         /* TreeNodeResolution                                */
  double tree_node_resolution_half;    /* TreeNodeResolutionHalf                            */
  double tree_alloc_factor;            /* TreeAllocFactor   (accepted; device tree self-sizes) */
  int64_t tree_min_num_of_cells;       /* TreeMinNumOfCells (accepted; unused)              */
  double G;                            /* PhysicalConst::G                                  */
  int64_t direct_sum_max;              /* reserved: a direct-sum path cannot meet the parity gates against the
                                          reference's monopole tree and is not enabled */
  int64_t shuffle_seed;                /* sampled mode only (max_sample_size > 0): seed of the counter-based permutation
                                          that replaces the reference's random_shuffle/rand() (subhalo_unbind.cpp:302),
                                          whose stream depends on libc state and OpenMP scheduling */
} hbtu_params;

/* What Unbind reads from `epoch` (src/snapshot.h:17-39, src/snapshot_number.h). */
typedef struct hbtu_epoch
{
  double scale_factor; /* Cosmology.ScaleFactor */
  double hz;           /* Cosmology.Hz          */
  int32_t snapshot_index;
  int32_t reserved;
} hbtu_epoch;

/* Per-subhalo scalar state: the Subhalo_t fields Unbind reads and writes
 * (src/subhalo.h:24-146).  [io] = read on entry, written on exit. */
typedef struct hbtu_sub_io
{
  double avg_pos[3];       /* [io] ComovingAveragePosition (initial frame must be set)  */
  double avg_vel[3];       /* [io] PhysicalAverageVelocity                              */
  double mostbound_pos[3]; /* [io] ComovingMostBoundPosition                            */
  double mostbound_vel[3]; /* [io] PhysicalMostBoundVelocity                            */
  int64_t nbound;          /* [io] Nbound (entry value drives the orphan rule, :434)    */
  int64_t sink_track_id;   /* [io] SinkTrackId                                          */
  int32_t snapshot_index_of_death; /* [io] */
  int32_t snapshot_index_of_sink;  /* [io] */
  float mbound;                         /* [out] Mbound                                 */
  float specific_self_potential_energy; /* [out] */
  float specific_self_kinetic_energy;   /* [out] */
  float specific_angular_momentum[3];   /* [out] */
  int64_t nsource_full;    /* [out] Particles.size() after unbinding, before TruncateSource */
  int64_t nsource;         /* [out] entries written to order_out for this subhalo        */
  int32_t iterations;      /* [out] potential evaluations performed (diagnostic)         */
  int32_t flags;           /* [in]  HBTU_SUB_* (0 = a subhalo reached through RecursiveUnbind)    */
} hbtu_sub_io;

/* hbtu_sub_io.flags */
#define HBTU_SUB_PLAIN_UNBIND 1 /* the caller enters this subhalo through plain Subhalo_t::Unbind, not RecursiveUnbind
                                   (field and new-born subhaloes, src/subhalo_unbind.cpp:498-510; the merge path,
                                   src/subhalo_merge.cpp:210; INCLUSIVE_MASS builds, :470-476): the orphan rule of
                                   RecursiveUnbind (:434-446, entry Nbound <= 1: the list is not reordered) does not
                                   apply.  Only valid for a subhalo without parent and without nested subhaloes. */

/* flags for hbtu_unbind_batch */
#define HBTU_FLAG_TRUNCATE_SOURCE 1 /* apply Subhalo_t::TruncateSource to every subhalo at the end
                                       (what RefineParticles does, src/subhalo_unbind.cpp:511-513) */

/* the reference's compile-time physics variants, selected per batch: */
#define HBTU_FLAG_NO_STRIPPING 2    /* -DNO_STRIPPING: Nbound = Nlast after the partition (src/subhalo_unbind.cpp:358-360) */
#define HBTU_FLAG_THERMAL_ENERGY 4  /* -DUNBIND_WITH_THERMAL_ENERGY: E += Particle_t::InternalEnergy in full evaluations
                                       (src/subhalo_unbind.cpp:351-353); the internal energy travels in vel[4*i+3] */

typedef struct hbtu_ctx hbtu_ctx;

/* Create / destroy a context bound to one CUDA device.  One context per host
 * thread / MPI rank / GPU; a context is not re-entrant. */
int hbtu_create(const hbtu_params *params, hbtu_ctx **out);
void hbtu_destroy(hbtu_ctx *ctx);
const char *hbtu_last_error(const hbtu_ctx *ctx); /* ctx may be NULL: error of the failed create */
int hbtu_abi_version(void);

/* Upper bound of the number of entries hbtu_unbind_batch writes to order_out
 * for this forest (every subhalo can receive all particles of its descendants). */
int64_t hbtu_order_capacity(int64_t nsub, const int64_t *part_offset, const int64_t *nest_offset,
                            const int32_t *nest_list);

/* Unbind a batch of subhaloes (replaces RefineParticles / RecursiveUnbind / Unbind).
 *
 *  part_offset[nsub+1]  subhalo s owns input particles [part_offset[s], part_offset[s+1])
 *  pos_mass[4*N]        x,y,z (comoving), mass      per input particle (HOST memory)
 *  vel[4*N]             vx,vy,vz (physical), InternalEnergy (read with HBTU_FLAG_THERMAL_ENERGY only) per input particle (HOST memory)
 *  nest_offset/nest_list CSR of NestedSubhalos (batch-local subhalo indices), may be NULL (no nesting).
 *                       Subhaloes that appear in nobody's list are roots.  Each root is processed
 *                       like Subhalo_t::RecursiveUnbind: children first, each child's unbound tail
 *                       Particles[Nbound..] is appended to its parent's source in list order.
 *  io[nsub]             per-subhalo scalars, updated in place
 *  order_offset[nsub+1] [out] subhalo s's final particle list is
 *                       order_out[order_offset[s] .. order_offset[s]+io[s].nsource)
 *  order_out            [out] indices into the batch's input particle arrays: the new
 *                       Subhalo_t::Particles order (bound by E ascending, then removal batches)
 *  energy_out           [out, optional] binding energy aligned with order_out (valid for the first
 *                       io[s].nbound entries of each subhalo; SAVE_BINDING_ENERGY), else NULL
 */
int hbtu_unbind_batch(hbtu_ctx *ctx, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset,
                      const float *pos_mass, const float *vel, const int64_t *nest_offset,
                      const int32_t *nest_list, hbtu_sub_io *io, int32_t flags, int64_t order_capacity,
                      int64_t *order_offset, int32_t *order_out, float *energy_out);

/* Pinned (page-locked) host memory for the caller's staging arrays.  hbtu_unbind_batch uploads the particle arrays on a copy
 * stream in two waves (everything but the dominant root subhalo first, that root - which the level-synchronous scheduler
 * reaches last - behind the kernels of the deeper levels); from pinned memory these are asynchronous DMA transfers, from any
 * other host memory they are plain staged copies.  A batch of many independent hierarchies without a dominant one (a
 * cosmological box; at least "pipeline_min_particles" particles - hbtu_set_tuning / HBTU_PIPELINE_MIN, default 2^24, 0 = never -
 * laid out hierarchy by hierarchy, parents in front of their nested subhaloes) is run in two parts (1/4 and 3/4 of the particles) instead, the second
 * uploading behind the kernels of the first; the results do not depend on it, bit for bit, but afterwards only the last
 * part is resident (hbtu_profile_executed refuses).  NULL when the allocation fails. */
void *hbtu_host_alloc(size_t bytes);
void hbtu_host_free(void *p);

/* The same call split so that a caller can keep a snapshot resident in HBM:
 *   hbtu_stage   : validate + host->device copies of the batch
 *   hbtu_execute : all kernels (may be called repeatedly on the staged batch; inputs are not modified)
 *   hbtu_fetch   : device->host copies of io / order / energies
 * hbtu_unbind_batch == stage + execute + fetch. */
int hbtu_stage(hbtu_ctx *ctx, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset,
               const float *pos_mass, const float *vel, const int64_t *nest_offset,
               const int32_t *nest_list, const hbtu_sub_io *io, int32_t flags);
int hbtu_execute(hbtu_ctx *ctx);
int hbtu_fetch(hbtu_ctx *ctx, hbtu_sub_io *io, int64_t order_capacity, int64_t *order_offset,
               int32_t *order_out, float *energy_out);

/* ---------------------------------------------------------------------------------------------------
 * Target split of the walk over several GPUs ("the non-natural case": one subhalo >> all others, so
 * that whole hierarchies cannot balance the GPUs - the AqA2 central holds 72 % of the particles).  The reference itself
 * parallelises over PARTICLES there (the OpenMP target loops of src/subhalo_unbind.cpp:319,341, chosen at
 * src/subhalo_tracking.cpp:420).  Here `nranks` cooperating contexts (one per GPU; in one process or in several) are given
 * the SAME batch and execute it in lock step: everything except the walk is replicated (it is ~15 % of a step), every round's
 * walk targets are dealt block-cyclically (16 CTAs = 2048..8192 targets at a time) to the contexts, each writes the new
 * binding energies of its targets into a zero-initialised array in target order, and `allreduce` must sum that array
 * element-wise over all contexts, in place (every element is non-zero on at most one context, so the sum is exact).
 * The results of all contexts are identical, bit for bit, to the unsplit execution.
 *
 *   allreduce(user, device_buf, count, cuda_stream): device_buf holds `count` floats in this context's device memory; the
 *   library's stream is idle while the function runs; when it returns the summed data must be complete in device_buf (or the
 *   work enqueued on `cuda_stream`, a cudaStream_t).  Return 0, anything else aborts the execution with HBTU_ERR_CUDA.
 *   It is called once per round, the same number of times, in the same order, on every context of the group.
 * nranks <= 1 or allreduce == NULL switches the split off. */
typedef int (*hbtu_allreduce_fn)(void *user, float *device_buf, int64_t count, void *cuda_stream);
int hbtu_set_walk_split(hbtu_ctx *ctx, int rank, int nranks, hbtu_allreduce_fn allreduce, void *user);

/* The same for contexts of ONE process (one host thread per context, e.g. the shim's HBT_UNBIND_DEVICES mode): a built-in
 * all-reduce over peer-mapped device memory - after a host barrier every context launches ONE kernel that sums its 1/nranks
 * slice of all members' arrays through NVLink peer loads and stores the result into every member's array (reduce-scatter +
 * all-gather fused; no NCCL).  hbtu_split_group_join sets the walk split of `ctx`; all members must then execute the same
 * batch concurrently.  Destroy the group after its members stopped executing (their split is switched off). */
typedef struct hbtu_split_group hbtu_split_group;
hbtu_split_group *hbtu_split_group_create(int nranks);
int hbtu_split_group_join(hbtu_split_group *group, hbtu_ctx *ctx, int rank);
void hbtu_split_group_destroy(hbtu_split_group *group);

/* GravityTree_t::Build + EvaluatePotential / BindingEnergy for one particle set
 * (src/gravity_tree.cpp:79-175): tree over the nsrc source particles, potential at ntgt targets.
 *  src_pos_mass / tgt_pos / tgt_vel are float4-strided (x,y,z,mass|unused) HOST arrays
 *  tgt_self_mass  NULL, or per-target mass whose self term m/eps is cancelled (a target that is
 *                 itself a tree source), 0 for foreign targets
 *  tgt_vel/ref_*  if tgt_vel != NULL the result is the binding energy 0.5|dv|^2 + pot w.r.t.
 *                 the frame (ref_pos, ref_vel), else the potential
 *  out[ntgt]      double results (the device sums in fp32 tiles + fp64 carries)              */
int hbtu_tree_potential(hbtu_ctx *ctx, const hbtu_epoch *epoch, int64_t nsrc, const float *src_pos_mass,
                        int64_t ntgt, const float *tgt_pos, const float *tgt_self_mass,
                        const float *tgt_vel, const double *ref_pos, const double *ref_vel, double *out);


/* ---------------------------------------------------------------------------------------------------
 * Post-unbinding per-subhalo properties: the step right after the path in
 * SubhaloSnapshot_t::UpdateTracks (src/subhalo_tracking.cpp:901-906),
 *   void Subhalo_t::CalculateProfileProperties(const Snapshot_t&)   src/subhalo.h:124, src/subhalo.cpp:242-332
 *   void Subhalo_t::CalculateShape()                                src/subhalo.h:125, src/subhalo.cpp:334-398
 * with Snapshot_t::SphericalOverdensitySize (src/snapshot.cpp:264-281, virial factor c200 = 200 of
 * HaloVirialFactors, :328-336) and PeriodicDistance (src/config_parser.h:143-156).
 * The eigen-vectors (EigenAxis, HAS_GSL builds only) stay on the host: they are a 3x3 problem per subhalo. */
typedef struct hbtu_profile_io
{
  double mostbound_pos[3];                 /* [in]  ComovingMostBoundPosition (centre of the profile)        */
  int64_t nbound;                          /* [in]  Nbound: the first nbound particles of the list are used  */
  float mbound;                            /* [in]  Mbound (normalises the inertia tensors, :391-392)        */
  float rmax_comoving;                     /* [out] RmaxComoving                                             */
  float vmax_physical;                     /* [out] VmaxPhysical                                             */
  float last_max_vmax_physical;            /* [io]  LastMaxVmaxPhysical                                      */
  int32_t snapshot_index_of_last_max_vmax; /* [io]  SnapshotIndexOfLastMaxVmax                               */
  float r2sigma_comoving;                  /* [out] R2SigmaComoving                                          */
  float rhalf_comoving;                    /* [out] RHalfComoving                                            */
  float bound_r200crit_comoving;           /* [io]  BoundR200CritComoving: untouched when no radius encloses 200 rho_crit */
  float bound_m200crit;                    /* [io]  BoundM200Crit                                            */
  float inertial_tensor[6];                /* [out] InertialTensor {xx,xy,xz,yy,yz,zz}                        */
  float inertial_tensor_weighted[6];       /* [out] InertialTensorWeighted                                   */
  int32_t reserved;
} hbtu_profile_io;

/*  part_offset[nsub+1]  subhalo s's particle list (bound particles first, Particles[0] = most bound) is
 *                       pos_mass[part_offset[s] .. part_offset[s+1]); only the first io[s].nbound are read
 *  pos_mass[4*N]        x,y,z (comoving), mass (HOST memory)                                                  */
int hbtu_profile_batch(hbtu_ctx *ctx, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset,
                       const float *pos_mass, hbtu_profile_io *io);
/* The same for the batch that hbtu_execute left resident in HBM (after hbtu_stage + hbtu_execute, before the next stage):
 * the particle lists are the new orders the unbinding just produced, the epoch is the staged one, io[s] belongs to
 * subhalo s of that batch; only the per-subhalo records cross PCIe. */
int hbtu_profile_executed(hbtu_ctx *ctx, hbtu_profile_io *io);

/* ---------------------------------------------------------------------------------------------------
 * Source preparation: exclusive particle ownership before unbinding,
 *   void SubhaloSnapshot_t::MaskSubhalos()  + SubhaloMasker_t::Mask   src/subhalo_tracking.cpp:793-841
 * Every ROOT of the nest forest (a host's central with the other heads appended to its list, :832-838) owns one
 * exclusion set.  Its hierarchy is visited depth first, children before their parent, in list order; a subhalo keeps, in
 * order, the particles whose Id no earlier visited subhalo (or earlier position of its own list) holds.  Subhaloes with
 * nbound <= 1 (orphans) are skipped: they keep their list and exclude nothing (:806).
 *
 *  particle_id[N]       Particle_t::Id of every list entry (HOST memory; any 64-bit values)
 *  nbound[nsub]         Subhalo_t::Nbound on entry (only "<= 1" matters)
 *  new_count[nsub]      [out] particles subhalo s keeps
 *  keep_index[N]        [out] subhalo s keeps the entries keep_index[part_offset[s] .. part_offset[s]+new_count[s])
 *                       (indices into the batch arrays, ascending)                                              */
int hbtu_mask_batch(hbtu_ctx *ctx, int64_t nsub, const int64_t *part_offset, const int64_t *particle_id,
                    const int64_t *nest_offset, const int32_t *nest_list, const int64_t *nbound, int64_t *new_count,
                    int32_t *keep_index);

/* ---------------------------------------------------------------------------------------------------
 * Particle query, the on-node part: Id -> index of the snapshot's particle, as
 *   MappedIndexTable_t::Fill        src/hash.tpp:18-32    (sort the (Id, index) pairs)
 *   MappedIndexTable_t::GetIndices  src/hash_remote.tpp:9-88  (batch binary search; what ParticleSnapshot_t::GetIndices,
 *                                   :98-105, runs for ParticleExchanger_t::QueryParticles, src/particle_exchanger.h:196-211)
 * The table lives in the context until the next build / clear / hbtu_destroy.  Ids are 64-bit (HBTInt = int builds widen).
 *  index_out[nq]  index into the build array, or -1 (SpecialConst::NullParticleId, src/datatypes.h:87) when absent.
 *  Ids are expected to be unique; of equal Ids the lowest index answers.                                         */
int hbtu_idtable_build(hbtu_ctx *ctx, int64_t n, const int64_t *particle_id);
int hbtu_idtable_query(hbtu_ctx *ctx, int64_t nq, const int64_t *query_id, int64_t *index_out);
int hbtu_idtable_clear(hbtu_ctx *ctx);

/* ---------------------------------------------------------------------------------------------------
 * Merger trap detection: the detection part of SubhaloSnapshot_t::MergeSubhalos,
 *   SubHelper_t::BuildPosition / BuildVelocity   src/subhalo_merge.cpp:29-123   mass-weighted mean and dispersion of the
 *                                                                              <= NumPartCoreMax = 20 most bound particles
 *   SinkDistance                                 src/subhalo_merge.cpp:125-130 d/sigma_R + v/sigma_V to a host's core
 *   FillHostTrackIds + DetectTraps               src/subhalo_merge.cpp:132-172 walk up the host chain, sink if delta < 2
 * The host relation is the nest forest (with the other heads glued to the central's list, GlueHeadNests,
 * src/subhalo_tracking.cpp:843-861).  Merging itself (MergeTo) and the re-unbinding of the hosts flagged in is_merged
 * (src/subhalo_merge.cpp:201-214) stay with the caller: the latter is one more hbtu_unbind_batch. */
typedef struct hbtu_trap_io
{
  double mostbound_pos[3];        /* [in]  ComovingMostBoundPosition                                            */
  double mostbound_vel[3];        /* [in]  PhysicalMostBoundVelocity                                            */
  int64_t nbound;                 /* [in]  Nbound                                                               */
  int64_t sink_track_id;          /* [io]  SinkTrackId (batch-local index; >= 0 on entry: already trapped)      */
  int32_t snapshot_index_of_sink; /* [io]  SnapshotIndexOfSink                                                  */
  int32_t is_merged;              /* [out] SubHelper_t::IsMerged: a real subhalo (Nbound > 1) sank into this one */
} hbtu_trap_io;

/*  part_offset / pos_mass / vel   particle lists in bound order; only the first min(nbound, 20) of each are read, so a list
 *                                 may be truncated to that                                                              */
int hbtu_detect_traps(hbtu_ctx *ctx, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset,
                      const float *pos_mass, const float *vel, const int64_t *nest_offset,
                      const int32_t *nest_list, hbtu_trap_io *io);

/* Counters of the last hbtu_execute / hbtu_tree_potential (for bench.py's roofline and
 * gpu_launches fields). */
typedef struct hbtu_stats
{
  int64_t kernel_launches;     /* kernels launched by this library (incl. CUB passes)  */
  int64_t tree_builds;         /* subhalo trees built                                  */
  int64_t walk_targets;        /* target particles walked / direct-summed              */
  int64_t pair_interactions;   /* accepted target x source interactions evaluated      */
  int64_t nodes_visited;       /* warp-level node visits of the tree walk              */
  int64_t rounds;              /* batched unbinding rounds                             */
  double walk_ms;              /* CUDA-event time of walk / direct-sum kernels         */
  double build_ms;             /* CUDA-event time of tree build kernels                */
  double other_ms;             /* partition / sort / reductions                        */
  double h2d_ms, d2h_ms;       /* copies in hbtu_stage / hbtu_fetch                    */
  int64_t h2d_bytes, d2h_bytes;
  double execute_ms;           /* CUDA-event time of the whole hbtu_execute (host planning gaps included) */
  int64_t tree_sources;        /* source particles over all tree builds (sum over rounds)                */
  int64_t walk_fallbacks;      /* counting on: groups of the masked walk redone per lane (chain stack exhausted) */
  double stage_wall_ms;        /* host wall clock inside hbtu_stage / the staging part of hbtu_unbind_batch          */
  double execute_wall_ms;      /* host wall clock inside hbtu_execute                                                */
  double fetch_wall_ms;        /* host wall clock inside hbtu_fetch                                                  */
  double phase_ms[8];          /* CUDA-event time of a step by phase, summed over its rounds: 0 source gather + bounding boxes,
                                  1 tree build (keys, sorts, cells, moments), 2 walk targets, 3 walk, 4 bound count + iteration
                                  state + partition bookkeeping, 5 energy sort + permutation, 6 frame reductions, 7 kinematics +
                                  finalisation                                                                           */
} hbtu_stats;
int hbtu_get_stats(const hbtu_ctx *ctx, hbtu_stats *out);
/* diagnostics (no reference counterpart): when on, the walk kernels of subsequent calls count accepted
 * interactions and warp node visits into hbtu_stats (costs a few percent; off by default). */
int hbtu_set_counting(hbtu_ctx *ctx, int on);
/* diagnostics (no reference counterpart): process-wide kernel-routing knobs of the walk ("walk_group_min", "walk_masked_pairs",
 * "walk_masked_blocks", "walk_tpl", "walk_big2", "walk_big4", "walk_small_max") and of the upload pipeline
 * ("pipeline_min_particles"; defaults = measured best, also settable
 * through HBTU_WALK_* environment variables read at the first use).  Results do not depend on them beyond fp64 summation
 * order.  hbtu_get_tuning returns -1 for an unknown key. */
/* diagnostics (no reference counterpart, host only): the parts hbtu_unbind_batch would run this batch in - see hbtu_host_alloc
 * above.  part_begin[0 .. parts] receives the first subhalo of every part and nsub; returns the number of parts (1 = in one
 * piece), HBTU_ERR_CAPACITY when max_parts is too small. */
int hbtu_plan_pipeline(int64_t nsub, const int64_t *part_offset, const int64_t *nest_offset, const int32_t *nest_list,
                       int64_t *part_begin, int max_parts);
int hbtu_set_tuning(const char *key, int64_t value);
int64_t hbtu_get_tuning(const char *key);

#ifdef __cplusplus
}
#endif
#endif /* HBT_UNBIND_H */
