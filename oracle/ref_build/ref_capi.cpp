/* ref_capi.cpp - C entry points around the UNMODIFIED reference sources.
 *
 * TEST INFRASTRUCTURE ONLY.  This translation unit is compiled together with
 * /root/reference/src/{subhalo_unbind,gravity_tree,...}.cpp (see ../Makefile)
 * into oracle/_ref/libhbtref_<variant>.so.  It contains no algorithm: it fills
 * the reference's own data structures (HBTConfig, Subhalo_t, Particle_t) from
 * the POD arguments of include/hbt_unbind.h, calls the reference's own
 *   Subhalo_t::RecursiveUnbind / Unbind / TruncateSource   (src/subhalo_unbind.cpp:263-458)
 *   SubhaloSnapshot_t::RefineParticles                      (src/subhalo_unbind.cpp:460-516)
 *   GravityTree_t::Build / EvaluatePotential / BindingEnergy (src/gravity_tree.cpp:79-175)
 * and copies the results back.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the resulting library.
 */
#include <omp.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <stdexcept>

#include <unordered_set>
#include <unordered_map>
#include <string>
#include <sstream>
#include <fstream>
#include <algorithm>
#include <memory>

#include "datatypes.h"
#include "config_parser.h"
#include "snapshot.h"
/* SubhaloSnapshot_t::MaskSubhalos is a private member (src/subhalo.h:207); the harness calls the unmodified
 * member, so the access specifier is lifted for this translation unit only (layout is unaffected). */
#define private public
#include "subhalo.h"
#undef private
#include "gravity_tree.h"
#include "hash.h"
#include "hash_remote.tpp"

#include "hbt_unbind.h"

/* defined by integration/subhalo_unbind_b200.cpp in the drop-in build only */
void HBT_B200_MaskSubhalos(SubhaloSnapshot_t &snap) __attribute__((weak));
void HBT_B200_DetectTraps(SubhaloSnapshot_t &snap, std::vector<char> &is_merged) __attribute__((weak));
void HBT_B200_CalculateProperties(SubhaloList_t &Subhalos, const Snapshot_t &epoch) __attribute__((weak));
void HBT_B200_UnbindMerged(SubhaloSnapshot_t &snap, const std::vector<char> &is_merged) __attribute__((weak));

/* the real body lives in src/io/subhalo_io.cpp, which needs libhdf5 (absent) */
void SubhaloSnapshot_t::BuildHDFDataType()
{
  H5T_SubhaloInMem = 0;
  H5T_SubhaloInDisk = 0;
}

namespace
{
class Epoch_t : public Snapshot_t
{ /* the minimum Unbind needs from `epoch`: Cosmology + snapshot index */
  HBTxyz dummy;

public:
  Epoch_t() : dummy{{0, 0, 0}} {}
  HBTInt size() const { return 0; }
  const HBTxyz &GetComovingPosition(const HBTInt) const { return dummy; }
  const HBTxyz &GetPhysicalVelocity(const HBTInt) const { return dummy; }
  HBTReal GetMass(const HBTInt) const { return 0; }
};

class ParticleView_t : public Snapshot_t
{ /* a Snapshot_t over a vector<Particle_t>, as src/halo.h / subhalo_unbind.cpp's views do */
public:
  const std::vector<Particle_t> &P;
  ParticleView_t(const std::vector<Particle_t> &p, const Snapshot_t &epoch) : P(p) { Cosmology = epoch.Cosmology; }
  HBTInt size() const { return P.size(); }
  const HBTxyz &GetComovingPosition(const HBTInt i) const { return P[i].ComovingPosition; }
  const HBTxyz &GetPhysicalVelocity(const HBTInt i) const { return P[i].PhysicalVelocity; }
  HBTReal GetMass(const HBTInt i) const { return P[i].Mass; }
};

void apply_params(const hbtu_params *p)
{
  HBTConfig.MinNumPartOfSub = p->min_num_part_of_sub;
  HBTConfig.PeriodicBoundaryOn = p->periodic_boundary_on != 0;
  HBTConfig.RefineMostboundParticle = p->refine_mostbound_particle != 0;
  HBTConfig.MaxSampleSizeOfPotentialEstimate = p->max_sample_size;
  HBTConfig.BoundMassPrecision = p->bound_mass_precision;
  HBTConfig.SourceSubRelaxFactor = p->source_sub_relax_factor;
  HBTConfig.BoxSize = p->box_size;
  HBTConfig.BoxHalf = p->box_half;
  HBTConfig.SofteningHalo = p->softening_halo;
  HBTConfig.TreeNodeOpenAngleSquare = p->tree_node_open_angle_square;
  HBTConfig.TreeNodeResolution = p->tree_node_resolution;
  HBTConfig.TreeNodeResolutionHalf = p->tree_node_resolution_half;
  HBTConfig.TreeAllocFactor = p->tree_alloc_factor > 0 ? p->tree_alloc_factor : 0.8;
  HBTConfig.TreeMinNumOfCells = p->tree_min_num_of_cells > 0 ? p->tree_min_num_of_cells : 10;
  HBTConfig.MinSnapshotIndex = 0;
  HBTConfig.MaxSnapshotIndex = 1 << 30;
  PhysicalConst::G = p->G;
  PhysicalConst::H0 = 100.;
}

void set_epoch(Snapshot_t &snap, const hbtu_epoch *e)
{
  snap.Cosmology.OmegaM0 = 0.3;
  snap.Cosmology.OmegaLambda0 = 0.7;
  snap.Cosmology.ScaleFactor = e->scale_factor;
  snap.Cosmology.Hz = e->hz;
  snap.Cosmology.OmegaZ = 0.3;
  snap.SetSnapshotIndex(e->snapshot_index);
}

void fill_subhalo(Subhalo_t &sub, int64_t s, const int64_t *part_offset, const float *pos_mass, const float *vel,
                  const hbtu_sub_io &io)
{
  int64_t b = part_offset[s], n = part_offset[s + 1] - b;
  sub.Particles.resize(n);
  for (int64_t i = 0; i < n; i++)
  {
    Particle_t &p = sub.Particles[i];
    p.Id = (HBTInt)(b + i); /* the input index travels as the particle Id */
    for (int j = 0; j < 3; j++)
    {
      p.ComovingPosition[j] = pos_mass[4 * (b + i) + j];
      p.PhysicalVelocity[j] = vel[4 * (b + i) + j];
    }
    p.Mass = pos_mass[4 * (b + i) + 3];
#ifndef DM_ONLY
    p.Type = TypeDM;
#endif
#ifdef HAS_THERMAL_ENERGY
    p.InternalEnergy = vel[4 * (b + i) + 3];
#endif
  }
  for (int j = 0; j < 3; j++)
  {
    sub.ComovingAveragePosition[j] = io.avg_pos[j];
    sub.PhysicalAverageVelocity[j] = io.avg_vel[j];
    sub.ComovingMostBoundPosition[j] = io.mostbound_pos[j];
    sub.PhysicalMostBoundVelocity[j] = io.mostbound_vel[j];
  }
  sub.Nbound = (HBTInt)io.nbound;
  sub.SinkTrackId = (HBTInt)io.sink_track_id;
  sub.SnapshotIndexOfDeath = io.snapshot_index_of_death;
  sub.SnapshotIndexOfSink = io.snapshot_index_of_sink;
  sub.TrackId = (HBTInt)s;
  /* [out] fields a subhalo keeps when nothing unbinds it (Subhalo_t's constructor leaves them uninitialised) */
  sub.Mbound = io.mbound;
  sub.SpecificSelfPotentialEnergy = io.specific_self_potential_energy;
  sub.SpecificSelfKineticEnergy = io.specific_self_kinetic_energy;
  for (int j = 0; j < 3; j++) sub.SpecificAngularMomentum[j] = io.specific_angular_momentum[j];
}

void read_subhalo(const Subhalo_t &sub, hbtu_sub_io &io)
{
  for (int j = 0; j < 3; j++)
  {
    io.avg_pos[j] = sub.ComovingAveragePosition[j];
    io.avg_vel[j] = sub.PhysicalAverageVelocity[j];
    io.mostbound_pos[j] = sub.ComovingMostBoundPosition[j];
    io.mostbound_vel[j] = sub.PhysicalMostBoundVelocity[j];
    io.specific_angular_momentum[j] = sub.SpecificAngularMomentum[j];
  }
  io.nbound = sub.Nbound;
  io.sink_track_id = sub.SinkTrackId;
  io.snapshot_index_of_death = sub.SnapshotIndexOfDeath;
  io.snapshot_index_of_sink = sub.SnapshotIndexOfSink;
  io.mbound = sub.Mbound;
  io.specific_self_potential_energy = sub.SpecificSelfPotentialEnergy;
  io.specific_self_kinetic_energy = sub.SpecificSelfKineticEnergy;
}

int write_orders(const std::vector<Subhalo_t> &subs, const std::vector<int64_t> &full, hbtu_sub_io *io,
                 int64_t order_capacity, int64_t *order_offset, int32_t *order_out, float *energy_out)
{
  int64_t pos = 0, nsub = subs.size();
  for (int64_t s = 0; s < nsub; s++)
  {
    const Subhalo_t &sub = subs[s];
    int64_t n = sub.Particles.size();
    if (pos + n > order_capacity) return HBTU_ERR_CAPACITY;
    order_offset[s] = pos;
    for (int64_t i = 0; i < n; i++) order_out[pos + i] = (int32_t)sub.Particles[i].Id;
    if (energy_out)
    {
      for (int64_t i = 0; i < n; i++) energy_out[pos + i] = 0.f;
#ifdef SAVE_BINDING_ENERGY
      int64_t ne = sub.Energies.size();
      for (int64_t i = 0; i < n && i < ne; i++) energy_out[pos + i] = sub.Energies[i];
#endif
    }
    io[s].nsource = n;
    io[s].nsource_full = full[s];
    io[s].iterations = 0;
    pos += n;
  }
  order_offset[nsub] = pos;
  return HBTU_OK;
}
} // namespace

/* the physics variants are compile-time in the reference: a library built with -DNO_STRIPPING /
 * -DUNBIND_WITH_THERMAL_ENERGY answers only batches that ask for exactly that variant */
static int variant_flags(void)
{
  int f = 0;
#ifdef NO_STRIPPING
  f |= HBTU_FLAG_NO_STRIPPING;
#endif
#ifdef UNBIND_WITH_THERMAL_ENERGY
  f |= HBTU_FLAG_THERMAL_ENERGY;
#endif
  return f;
}
static double g_last_refine_seconds = 0.0; /* wall time of the last SubhaloSnapshot_t::RefineParticles() call alone */
static bool variant_ok(int32_t flags) { return (flags & (HBTU_FLAG_NO_STRIPPING | HBTU_FLAG_THERMAL_ENERGY)) == variant_flags(); }

extern "C" {
int hbtref_variant_flags(void) { return variant_flags(); }

int hbtref_sizeof_particle(void) { return (int)sizeof(Particle_t); }
int hbtref_sizeof_subhalo(void) { return (int)sizeof(Subhalo_t); }
int hbtref_sizeof_hbtint(void) { return (int)sizeof(HBTInt); }
int hbtref_sizeof_hbtreal(void) { return (int)sizeof(HBTReal); }

void hbtref_set_num_threads(int n)
{
  omp_set_num_threads(n);
  omp_set_max_active_levels(1); /* HBT.cpp:22 */
}
int hbtref_get_max_threads(void) { return omp_get_max_threads(); }
/* seconds the last hbtref_refine_particles spent inside SubhaloSnapshot_t::RefineParticles() itself (in the drop-in build: pack +
 * H2D + kernels + D2H + permutation of vector<Particle_t>), without the harness' construction of the snapshot */
double hbtref_last_refine_seconds(void) { return g_last_refine_seconds; }
void hbtref_seed(unsigned s)
{
  srand(s);
  srand48(s);
}

/* Same contract as hbtu_unbind_batch (include/hbt_unbind.h); roots are driven with the
 * reference's RecursiveUnbind, then TruncateSource when asked - the schedule of
 * RefineParticles (src/subhalo_unbind.cpp:479-513) with one OpenMP loop over roots. */
int hbtref_unbind_batch(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset,
                        const float *pos_mass, const float *vel, const int64_t *nest_offset, const int32_t *nest_list,
                        hbtu_sub_io *io, int32_t flags, int64_t order_capacity, int64_t *order_offset,
                        int32_t *order_out, float *energy_out)
{
  if ((params->real_bytes != 4 && params->real_bytes != (int)sizeof(HBTReal)) || !variant_ok(flags)) return HBTU_ERR_UNSUPPORTED;
  apply_params(params);
  Epoch_t snap;
  set_epoch(snap, epoch);
  omp_set_max_active_levels(1);

  std::vector<Subhalo_t> subs(nsub);
  std::vector<char> is_child(nsub, 0);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t s = 0; s < nsub; s++) fill_subhalo(subs[s], s, part_offset, pos_mass, vel, io[s]);
  if (nest_offset)
    for (int64_t s = 0; s < nsub; s++)
      for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++)
      {
        int32_t c = nest_list[k];
        if (c < 0 || c >= nsub || is_child[c] || c == s) return HBTU_ERR_INVALID;
        is_child[c] = 1;
        subs[s].NestedSubhalos.push_back(c);
      }
  /* largest roots first so that dynamic scheduling balances, as the reference's
   * mass-sorted member lists effectively do */
  for (int64_t s = 0; s < nsub; s++)
    if ((io[s].flags & HBTU_SUB_PLAIN_UNBIND) && (is_child[s] || !subs[s].NestedSubhalos.empty())) return HBTU_ERR_INVALID;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t s = 0; s < nsub; s++)
    if (!is_child[s])
    { /* HBTU_SUB_PLAIN_UNBIND: the reference's plain Unbind call sites (src/subhalo_unbind.cpp:498-510, subhalo_merge.cpp:210) */
      if (io[s].flags & HBTU_SUB_PLAIN_UNBIND) subs[s].Unbind(snap);
      else subs[s].RecursiveUnbind(subs, snap);
    }

  std::vector<int64_t> full(nsub);
  for (int64_t s = 0; s < nsub; s++) full[s] = subs[s].Particles.size();
  if (flags & HBTU_FLAG_TRUNCATE_SOURCE)
  {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t s = 0; s < nsub; s++) subs[s].TruncateSource();
  }
  for (int64_t s = 0; s < nsub; s++) read_subhalo(subs[s], io[s]);
  return write_orders(subs, full, io, order_capacity, order_offset, order_out, energy_out);
}

/* GravityTree_t::Build + EvaluatePotential/BindingEnergy; same contract as hbtu_tree_potential. */
int hbtref_tree_potential(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsrc, const float *src_pos_mass,
                          int64_t ntgt, const float *tgt_pos, const float *tgt_self_mass, const float *tgt_vel,
                          const double *ref_pos, const double *ref_vel, double *out)
{
  if ((params->real_bytes != 4 && params->real_bytes != (int)sizeof(HBTReal))) return HBTU_ERR_UNSUPPORTED;
  apply_params(params);
  Epoch_t snap;
  set_epoch(snap, epoch);
  std::vector<Particle_t> P(nsrc);
  for (int64_t i = 0; i < nsrc; i++)
  {
    P[i].Id = i;
    for (int j = 0; j < 3; j++)
    {
      P[i].ComovingPosition[j] = src_pos_mass[4 * i + j];
      P[i].PhysicalVelocity[j] = 0;
    }
    P[i].Mass = src_pos_mass[4 * i + 3];
  }
  ParticleView_t view(P, snap);
  GravityTree_t tree;
  tree.Reserve(nsrc);
  tree.Build(view);
  HBTxyz rp, rv;
  if (tgt_vel)
    for (int j = 0; j < 3; j++)
    {
      rp[j] = ref_pos[j];
      rv[j] = ref_vel[j];
    }
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < ntgt; i++)
  {
    HBTxyz x{{(HBTReal)tgt_pos[4 * i], (HBTReal)tgt_pos[4 * i + 1], (HBTReal)tgt_pos[4 * i + 2]}};
    HBTReal m = tgt_self_mass ? tgt_self_mass[i] : 0;
    if (tgt_vel)
    {
      HBTxyz v{{(HBTReal)tgt_vel[4 * i], (HBTReal)tgt_vel[4 * i + 1], (HBTReal)tgt_vel[4 * i + 2]}};
      out[i] = tree.BindingEnergy(x, v, rp, rv, m);
    }
    else
      out[i] = tree.EvaluatePotential(x, m);
  }
  return HBTU_OK;
}


/* SubhaloSnapshot_t::RefineParticles() itself (src/subhalo_unbind.cpp:460-516) on an in-memory snapshot:
 * host haloes with a central + heads + nests, field subhaloes (host_halo_id = -1) and new-born subhaloes
 * (index >= n_old, unknown to the MemberTable).  The central of a host is its member with the largest mbound_in.
 * The same translation unit is linked a second time against integration/subhalo_unbind_b200.o instead of the
 * reference's subhalo_unbind.o (libhbtdropin_v32.so): same driver, two backends. */
int hbtref_refine_particles(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset,
                            const float *pos_mass, const float *vel, const int64_t *nest_offset, const int32_t *nest_list,
                            const int32_t *host_halo_id, int64_t n_old, int32_t nhalos, const float *mbound_in, hbtu_sub_io *io,
                            int64_t order_capacity, int64_t *order_offset, int32_t *order_out, float *energy_out)
{
  if ((params->real_bytes != 4 && params->real_bytes != (int)sizeof(HBTReal))) return HBTU_ERR_UNSUPPORTED;
  apply_params(params);
  omp_set_max_active_levels(1);
  SubhaloSnapshot_t snap;
  set_epoch(snap, epoch);
  std::vector<char> is_child(nsub, 0);
  snap.Subhalos.resize(n_old);
  auto fill = [&](Subhalo_t &sub, int64_t s) {
    fill_subhalo(sub, s, part_offset, pos_mass, vel, io[s]);
    sub.HostHaloId = host_halo_id[s];
    sub.Mbound = mbound_in[s];
    sub.Rank = 0;
    if (nest_offset)
      for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++) sub.NestedSubhalos.push_back(nest_list[k]);
  };
  for (int64_t s = 0; s < n_old; s++) fill(snap.Subhalos[s], s);
  if (nest_offset)
    for (int64_t k = 0; k < nest_offset[nsub]; k++) is_child[nest_list[k]] = 1;
#pragma omp parallel
  snap.MemberTable.Build(nhalos, snap.Subhalos, true); /* orphaned worksharing inside: must be called in parallel */
  snap.MemberTable.SubGroupsOfHeads.assign(nhalos, std::vector<HBTInt>());
  for (HBTInt h = 0; h < nhalos; h++)
  {
    auto &grp = snap.MemberTable.SubGroups[h];
    for (HBTInt i = 0; i < grp.size(); i++)
      if (!is_child[grp[i]]) snap.MemberTable.SubGroupsOfHeads[h].push_back(grp[i]); /* mass-sorted: central first */
  }
  for (int64_t s = n_old; s < nsub; s++)
  {
    snap.Subhalos.emplace_back();
    fill(snap.Subhalos.back(), s);
  }
  try
  {
    const double t0 = omp_get_wtime();
    snap.RefineParticles();
    g_last_refine_seconds = omp_get_wtime() - t0;
  }
  catch (const std::exception &ex)
  { /* only the drop-in build can throw (the reference path cannot fail) */
    fprintf(stderr, "RefineParticles threw: %s\n", ex.what());
    return HBTU_ERR_CUDA;
  }
  std::vector<int64_t> full(nsub);
  for (int64_t s = 0; s < nsub; s++)
  {
    full[s] = snap.Subhalos[s].Particles.size();
    read_subhalo(snap.Subhalos[s], io[s]);
  }
  return write_orders(snap.Subhalos, full, io, order_capacity, order_offset, order_out, energy_out);
}

/* Subhalo_t::CalculateProfileProperties + CalculateShape of the reference (src/subhalo.cpp:242-398); same contract as
 * hbtu_profile_batch.  In the drop-in build (libhbtdropin_*.so) the shim's batched replacement of the loop at
 * src/subhalo_tracking.cpp:901-906 is linked in and used instead of the two member functions. */
int hbtref_profile_batch(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset,
                         const float *pos_mass, hbtu_profile_io *io)
{
  if ((params->real_bytes != 4 && params->real_bytes != (int)sizeof(HBTReal))) return HBTU_ERR_UNSUPPORTED;
  apply_params(params);
  Epoch_t snap;
  set_epoch(snap, epoch);
  omp_set_max_active_levels(1);
  std::vector<Subhalo_t> subs(nsub);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t s = 0; s < nsub; s++)
  {
    Subhalo_t &sub = subs[s];
    const int64_t b = part_offset[s], n = part_offset[s + 1] - b;
    sub.Particles.resize(n);
    for (int64_t i = 0; i < n; i++)
    {
      Particle_t &p = sub.Particles[i];
      p.Id = (HBTInt)(b + i);
      for (int j = 0; j < 3; j++)
      {
        p.ComovingPosition[j] = pos_mass[4 * (b + i) + j];
        p.PhysicalVelocity[j] = 0;
      }
      p.Mass = pos_mass[4 * (b + i) + 3];
#ifndef DM_ONLY
      p.Type = TypeDM;
#endif
    }
    hbtu_profile_io &o = io[s];
    sub.Nbound = (HBTInt)o.nbound;
    sub.Mbound = o.mbound;
    for (int j = 0; j < 3; j++) sub.ComovingMostBoundPosition[j] = o.mostbound_pos[j];
    sub.LastMaxVmaxPhysical = o.last_max_vmax_physical;
    sub.SnapshotIndexOfLastMaxVmax = o.snapshot_index_of_last_max_vmax;
    sub.BoundR200CritComoving = o.bound_r200crit_comoving;
    sub.BoundM200Crit = o.bound_m200crit;
  }
  if (HBT_B200_CalculateProperties)
    HBT_B200_CalculateProperties(subs, snap);
  else
  {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t s = 0; s < nsub; s++)
    {
      subs[s].CalculateProfileProperties(snap);
      subs[s].CalculateShape();
    }
  }
  for (int64_t s = 0; s < nsub; s++)
  {
    const Subhalo_t &sub = subs[s];
    hbtu_profile_io &o = io[s];
    o.rmax_comoving = sub.RmaxComoving;
    o.vmax_physical = sub.VmaxPhysical;
    o.last_max_vmax_physical = sub.LastMaxVmaxPhysical;
    o.snapshot_index_of_last_max_vmax = sub.SnapshotIndexOfLastMaxVmax;
    o.r2sigma_comoving = sub.R2SigmaComoving;
    o.rhalf_comoving = sub.RHalfComoving;
    o.bound_r200crit_comoving = sub.BoundR200CritComoving;
    o.bound_m200crit = sub.BoundM200Crit;
    for (int j = 0; j < 6; j++)
    {
      o.inertial_tensor[j] = sub.InertialTensor[j];
      o.inertial_tensor_weighted[j] = sub.InertialTensorWeighted[j];
    }
  }
  return HBTU_OK;
}

/* SubhaloSnapshot_t::MaskSubhalos of the reference (src/subhalo_tracking.cpp:824-841) on an explicit nest forest; same
 * contract as hbtu_mask_batch.  Every root becomes the central (and only head) of its own host halo, so that the
 * reference's loop over MemberTable.SubGroups visits exactly the given hierarchies.  In the drop-in build the shim's
 * HBT_B200_MaskSubhalos (-> hbtu_mask_batch) is used instead of the member. */
int hbtref_mask_batch(const hbtu_params *params, int64_t nsub, const int64_t *part_offset, const int64_t *particle_id,
                      const int64_t *nest_offset, const int32_t *nest_list, const int64_t *nbound, int64_t *new_count,
                      int32_t *keep_index)
{
  apply_params(params);
  omp_set_max_active_levels(1);
  SubhaloSnapshot_t snap;
  snap.Subhalos.resize(nsub);
  std::vector<int32_t> root(nsub, -1);
  std::vector<char> is_child(nsub, 0);
  if (nest_offset)
    for (int64_t k = 0; k < nest_offset[nsub]; k++) is_child[nest_list[k]] = 1;
  int32_t nhalos = 0;
  std::vector<int64_t> stack;
  for (int64_t s = 0; s < nsub; s++)
  {
    if (is_child[s]) continue;
    stack.assign(1, s);
    while (!stack.empty())
    {
      int64_t q = stack.back();
      stack.pop_back();
      root[q] = nhalos;
      if (nest_offset)
        for (int64_t k = nest_offset[q]; k < nest_offset[q + 1]; k++) stack.push_back(nest_list[k]);
    }
    nhalos++;
  }
  for (int64_t s = 0; s < nsub; s++)
  {
    Subhalo_t &sub = snap.Subhalos[s];
    const int64_t b = part_offset[s], n = part_offset[s + 1] - b;
    sub.Particles.resize(n);
    for (int64_t i = 0; i < n; i++) sub.Particles[i].Id = (HBTInt)particle_id[b + i];
    sub.Nbound = (HBTInt)nbound[s];
    sub.Mbound = is_child[s] ? 1.f : 1e30f; /* the root sorts first in its group: it is the central */
    sub.HostHaloId = root[s];
    sub.TrackId = (HBTInt)s;
    sub.Rank = 0;
    if (nest_offset)
      for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++) sub.NestedSubhalos.push_back(nest_list[k]);
  }
#pragma omp parallel
  snap.MemberTable.Build(nhalos, snap.Subhalos, true);
  snap.MemberTable.SubGroupsOfHeads.assign(nhalos, std::vector<HBTInt>());
  for (HBTInt h = 0; h < nhalos; h++) snap.MemberTable.SubGroupsOfHeads[h].push_back(snap.MemberTable.SubGroups[h][0]);
  if (HBT_B200_MaskSubhalos)
    HBT_B200_MaskSubhalos(snap);
  else
    snap.MaskSubhalos();
  for (int64_t s = 0; s < nsub; s++)
  { /* the kept list is a subsequence of the input list: recover the positions */
    const Subhalo_t &sub = snap.Subhalos[s];
    const int64_t b = part_offset[s], e = part_offset[s + 1];
    int64_t i = b, save = b;
    for (const auto &p : sub.Particles)
    {
      while (i < e && (HBTInt)particle_id[i] != p.Id) i++;
      if (i >= e) return HBTU_ERR_INVALID;
      keep_index[save++] = (int32_t)i++;
    }
    new_count[s] = save - b;
  }
  return HBTU_OK;
}

/* MappedIndexTable_t<HBTInt,HBTInt>::Fill + GetIndices of the reference (src/hash.tpp:18-32, src/hash_remote.tpp:9-88),
 * queried the way ParticleExchanger_t::QueryParticles does (src/particle_exchanger.h:196-211): queries sorted by Id, batch
 * binary search, original order restored.  Same contract as hbtu_idtable_build + hbtu_idtable_query. */
namespace
{
struct IdList_t : public KeyList_t<HBTInt, HBTInt>
{
  const int64_t *ids;
  int64_t n;
  IdList_t(const int64_t *p, int64_t m) : ids(p), n(m) {}
  HBTInt GetKey(const HBTInt i) const { return (HBTInt)ids[i]; }
  HBTInt GetIndex(const HBTInt i) const { return i; }
  HBTInt size() const { return (HBTInt)n; }
};
struct QueryItem_t
{
  HBTInt Id;
  int64_t Order;
};
} // namespace

int hbtref_idtable_query(const hbtu_params *params, int64_t n, const int64_t *particle_id, int64_t nq, const int64_t *query_id,
                         int64_t *index_out)
{
  (void)params;
  MappedIndexTable_t<HBTInt, HBTInt> table;
  IdList_t keys(particle_id, n);
  table.Fill(keys, SpecialConst::NullParticleId);
  std::vector<QueryItem_t> q(nq);
  for (int64_t i = 0; i < nq; i++)
  {
    q[i].Id = (HBTInt)query_id[i];
    q[i].Order = i;
  }
  std::sort(q.begin(), q.end(), [](const QueryItem_t &a, const QueryItem_t &b) { return a.Id < b.Id; });
  table.GetIndices(q);
  for (const auto &x : q) index_out[x.Order] = x.Id;
  return HBTU_OK;
}

/* The detection part of SubhaloSnapshot_t::MergeSubhalos of the reference (src/subhalo_merge.cpp:187-199, with
 * HBTConfig.MergeTrappedSubhalos = false so that nothing is merged or re-unbound) on an explicit nest forest; same contract
 * as hbtu_detect_traps.  In the drop-in build the shim's HBT_B200_DetectTraps (-> hbtu_detect_traps) is used instead. */
int hbtref_detect_traps(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass,
                        const float *vel, const int64_t *nest_offset, const int32_t *nest_list, hbtu_trap_io *io)
{
  apply_params(params);
  HBTConfig.MergeTrappedSubhalos = false;
  omp_set_max_active_levels(1);
  SubhaloSnapshot_t snap;
  set_epoch(snap, epoch);
  snap.Subhalos.resize(nsub);
  std::vector<int32_t> root(nsub, -1);
  std::vector<char> is_child(nsub, 0);
  if (nest_offset)
    for (int64_t k = 0; k < nest_offset[nsub]; k++) is_child[nest_list[k]] = 1;
  int32_t nhalos = 0;
  std::vector<int64_t> stack;
  for (int64_t s = 0; s < nsub; s++)
  {
    if (is_child[s]) continue;
    stack.assign(1, s);
    while (!stack.empty())
    {
      int64_t q = stack.back();
      stack.pop_back();
      root[q] = nhalos;
      if (nest_offset)
        for (int64_t k = nest_offset[q]; k < nest_offset[q + 1]; k++) stack.push_back(nest_list[k]);
    }
    nhalos++;
  }
  for (int64_t s = 0; s < nsub; s++)
  {
    Subhalo_t &sub = snap.Subhalos[s];
    const int64_t b = part_offset[s], n = part_offset[s + 1] - b;
    sub.Particles.resize(n);
    for (int64_t i = 0; i < n; i++)
    {
      Particle_t &p = sub.Particles[i];
      p.Id = (HBTInt)(b + i);
      for (int j = 0; j < 3; j++)
      {
        p.ComovingPosition[j] = pos_mass[4 * (b + i) + j];
        p.PhysicalVelocity[j] = vel[4 * (b + i) + j];
      }
      p.Mass = pos_mass[4 * (b + i) + 3];
    }
    sub.Nbound = (HBTInt)io[s].nbound;
    sub.Mbound = is_child[s] ? 1.f : 1e30f;
    sub.HostHaloId = root[s];
    sub.TrackId = (HBTInt)s;
    sub.Rank = 0;
    sub.SinkTrackId = (HBTInt)io[s].sink_track_id;
    sub.SnapshotIndexOfSink = io[s].snapshot_index_of_sink;
    for (int j = 0; j < 3; j++)
    {
      sub.ComovingMostBoundPosition[j] = io[s].mostbound_pos[j];
      sub.PhysicalMostBoundVelocity[j] = io[s].mostbound_vel[j];
    }
    if (nest_offset)
      for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++) sub.NestedSubhalos.push_back(nest_list[k]);
  }
#pragma omp parallel
  snap.MemberTable.Build(nhalos, snap.Subhalos, true);
  snap.MemberTable.SubGroupsOfHeads.assign(nhalos, std::vector<HBTInt>());
  for (HBTInt h = 0; h < nhalos; h++) snap.MemberTable.SubGroupsOfHeads[h].push_back(snap.MemberTable.SubGroups[h][0]);
  std::vector<char> merged(nsub, 0);
  std::vector<int64_t> sink_before(nsub);
  for (int64_t s = 0; s < nsub; s++) sink_before[s] = io[s].sink_track_id;
  if (HBT_B200_DetectTraps)
    HBT_B200_DetectTraps(snap, merged);
  else
  {
    snap.MergeSubhalos();
    /* SubHelper_t::IsMerged is local to MergeSubhalos: it is set exactly for the sinks of real subhaloes trapped now (:152-153) */
    for (int64_t s = 0; s < nsub; s++)
      if (sink_before[s] == -1 && snap.Subhalos[s].SinkTrackId != SpecialConst::NullTrackId && snap.Subhalos[s].Nbound > 1)
        merged[snap.Subhalos[s].SinkTrackId] = 1;
  }
  for (int64_t s = 0; s < nsub; s++)
  {
    io[s].sink_track_id = snap.Subhalos[s].SinkTrackId;
    io[s].snapshot_index_of_sink = snap.Subhalos[s].SnapshotIndexOfSink;
    io[s].is_merged = merged[s];
  }
  return HBTU_OK;
}

/* SubhaloSnapshot_t::MergeSubhalos() of the reference with HBTConfig.MergeTrappedSubhalos ON (src/subhalo_merge.cpp:187-220):
 * trap detection, MergeRecursive / MergeTo, then Subhalo_t::Unbind of every host flagged IsMerged FROM THE OpenMP LOOP at
 * :207-210 and TruncateSource (:211-214).  Every subhalo is a member of one host halo per nest tree (root = central).
 *   mode 0  the unmodified member function.  In libhbtref_* its Unbind is the reference's; in libhbtdropin_* it is the shim's
 *           (integration/subhalo_unbind_b200.cpp), called concurrently by `nthreads` OpenMP threads: the thread-safety test.
 *   mode 1  (drop-in build only) the patched sequence INTEGRATION.md documents: HBT_B200_DetectTraps, the reference's own
 *           MergeRecursive, then ONE batch through HBT_B200_UnbindMerged.
 * Particle Ids are the input indices, so order_out reports the final lists the same way hbtu_unbind_batch does. */
int hbtref_merge_subhalos(const hbtu_params *params, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass,
                          const float *vel, const int64_t *nest_offset, const int32_t *nest_list, hbtu_sub_io *io, int32_t mode,
                          int32_t nthreads, int64_t order_capacity, int64_t *order_offset, int32_t *order_out, int32_t *is_merged_out)
{
  if ((params->real_bytes != 4 && params->real_bytes != (int)sizeof(HBTReal))) return HBTU_ERR_UNSUPPORTED;
  if (mode == 1 && !(HBT_B200_DetectTraps && HBT_B200_UnbindMerged)) return HBTU_ERR_UNSUPPORTED;
  apply_params(params);
  HBTConfig.MergeTrappedSubhalos = true;
  omp_set_max_active_levels(1);
  if (nthreads > 0) omp_set_num_threads(nthreads);
  SubhaloSnapshot_t snap;
  set_epoch(snap, epoch);
  snap.ParallelizeHaloes = true;
  snap.Subhalos.resize(nsub);
  std::vector<int32_t> root(nsub, -1);
  std::vector<char> is_child(nsub, 0);
  if (nest_offset)
    for (int64_t k = 0; k < nest_offset[nsub]; k++) is_child[nest_list[k]] = 1;
  int32_t nhalos = 0;
  std::vector<int64_t> stack;
  for (int64_t s = 0; s < nsub; s++)
  {
    if (is_child[s]) continue;
    stack.assign(1, s);
    while (!stack.empty())
    {
      int64_t q = stack.back();
      stack.pop_back();
      root[q] = nhalos;
      if (nest_offset)
        for (int64_t k = nest_offset[q]; k < nest_offset[q + 1]; k++) stack.push_back(nest_list[k]);
    }
    nhalos++;
  }
  for (int64_t s = 0; s < nsub; s++)
  {
    Subhalo_t &sub = snap.Subhalos[s];
    fill_subhalo(sub, s, part_offset, pos_mass, vel, io[s]);
    sub.Mbound = is_child[s] ? 1.f : 1e30f; /* the root sorts first in its group: it is the central */
    sub.HostHaloId = root[s];
    sub.Rank = 0;
    if (nest_offset)
      for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++) sub.NestedSubhalos.push_back(nest_list[k]);
  }
#pragma omp parallel
  snap.MemberTable.Build(nhalos, snap.Subhalos, true);
  snap.MemberTable.SubGroupsOfHeads.assign(nhalos, std::vector<HBTInt>());
  for (HBTInt h = 0; h < nhalos; h++) snap.MemberTable.SubGroupsOfHeads[h].push_back(snap.MemberTable.SubGroups[h][0]);
  std::vector<int64_t> sink_before(nsub), nbound_before(nsub);
  for (int64_t s = 0; s < nsub; s++)
  {
    sink_before[s] = io[s].sink_track_id;
    nbound_before[s] = io[s].nbound;
  }
  std::vector<char> merged(nsub, 0);
  try
  {
    if (mode == 0)
    {
      snap.MergeSubhalos();
      /* SubHelper_t::IsMerged is local to MergeSubhalos: set exactly for the sinks of real subhaloes trapped now (:152-153) */
      for (int64_t s = 0; s < nsub; s++)
        if (sink_before[s] == -1 && snap.Subhalos[s].SinkTrackId != SpecialConst::NullTrackId && nbound_before[s] > 1)
          merged[snap.Subhalos[s].SinkTrackId] = 1;
    }
    else
    {
      HBT_B200_DetectTraps(snap, merged);
      for (HBTInt grpid = 0; grpid < nhalos; grpid++)
        if (snap.MemberTable.SubGroups[grpid].size()) snap.MergeRecursive(snap.MemberTable.SubGroups[grpid][0]);
      HBT_B200_UnbindMerged(snap, merged);
    }
  }
  catch (const std::exception &ex)
  {
    fprintf(stderr, "MergeSubhalos threw: %s\n", ex.what());
    return HBTU_ERR_CUDA;
  }
  std::vector<int64_t> full(nsub);
  for (int64_t s = 0; s < nsub; s++)
  {
    full[s] = snap.Subhalos[s].Particles.size();
    read_subhalo(snap.Subhalos[s], io[s]);
    if (is_merged_out) is_merged_out[s] = merged[s];
  }
  return write_orders(snap.Subhalos, full, io, order_capacity, order_offset, order_out, nullptr);
}

} // extern "C"
