// subhalo_unbind_b200.cpp - drop-in replacement object for the reference's src/subhalo_unbind.cpp.
//
// It defines exactly the four member functions that file defines
//     void SubhaloSnapshot_t::RefineParticles()                         (src/subhalo.h:245)
//     void Subhalo_t::Unbind(const Snapshot_t &epoch)                   (src/subhalo.h:111)
//     void Subhalo_t::RecursiveUnbind(SubhaloList_t&, const Snapshot_t&) (src/subhalo.h:112)
//     void Subhalo_t::TruncateSource()                                  (src/subhalo.h:114)
// with unchanged signatures, so HBT.cpp:75, subhalo_merge.cpp:210 and every other caller link against it
// unmodified.  Each one packs the reference's own data structures (vector<Particle_t>, Subhalo_t scalars,
// HBTConfig, Cosmology) into the POD arguments of include/hbt_unbind.h, calls the CUDA library and unpacks.
// There is no algorithm and no CPU fallback here: a failing call throws std::runtime_error (the reference's own
// error style, src/config_parser.cpp:70).  Compile with the same -D flags as the rest of HBT+ (V32: -DDM_ONLY).
//
// Build:  g++ -std=c++11 -O2 -fopenmp $(HBT_DEFS) -I$(HBT)/src -I<repo>/include -c subhalo_unbind_b200.cpp
// Link :  replace subhalo_unbind.o by subhalo_unbind_b200.o and add -L<repo>/hbtplus_b200/csrc -lhbtunbind
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#ifdef __linux__
#include <sys/mman.h>
#endif

#include "datatypes.h"
#include "snapshot_number.h"
#include "subhalo.h"

#include "hbt_unbind.h"

// HBTReal = double builds (-DHBT_REAL8) are served too: the pack loops below narrow positions, velocities and masses to the
// float4 arrays of the ABI (the device path computes in fp32 + fp64 sums either way), the records travel as double.

namespace
{
hbtu_params g_params;

void fill_params(hbtu_params &p)
{
  std::memset(&p, 0, sizeof(p));
  p.struct_size = sizeof(hbtu_params);
  p.real_bytes = 4; // width of the particle arrays this shim hands over (see the note on HBT_REAL8 above)
  p.min_num_part_of_sub = HBTConfig.MinNumPartOfSub;
  p.periodic_boundary_on = HBTConfig.PeriodicBoundaryOn;
  p.refine_mostbound_particle = HBTConfig.RefineMostboundParticle;
  const char *dev = getenv("HBT_UNBIND_DEVICE"); // one MPI rank <-> one GPU; default: local rank modulo device count is the launcher's job
  p.device = dev ? atoi(dev) : 0;
  p.max_sample_size = HBTConfig.MaxSampleSizeOfPotentialEstimate;
  p.bound_mass_precision = HBTConfig.BoundMassPrecision;
  p.source_sub_relax_factor = HBTConfig.SourceSubRelaxFactor;
  p.box_size = HBTConfig.BoxSize;
  p.box_half = HBTConfig.BoxHalf;
  p.softening_halo = HBTConfig.SofteningHalo;
  p.tree_node_open_angle_square = HBTConfig.TreeNodeOpenAngleSquare;
  p.tree_node_resolution = HBTConfig.TreeNodeResolution;
  p.tree_node_resolution_half = HBTConfig.TreeNodeResolutionHalf;
  p.tree_alloc_factor = HBTConfig.TreeAllocFactor;
  p.tree_min_num_of_cells = HBTConfig.TreeMinNumOfCells;
  p.G = PhysicalConst::G;
  p.direct_sum_max = 0;
  const char *seed = getenv("HBT_UNBIND_SHUFFLE_SEED");
  p.shuffle_seed = seed ? atoll(seed) : 20240001;
}

// Devices this process drives.  Default: one (HBT_UNBIND_DEVICE, the usual one-MPI-rank-per-GPU launch).  With
// HBT_UNBIND_DEVICES=0,1,2,3 one rank shards its hierarchies over several GPUs of the box (whole
// hierarchies, cost-weighted longest-processing-time-first, no data-path collective; see run_sharded below).
//
// Threading.  An hbtu_ctx is not re-entrant, while the reference calls Subhalo_t::Unbind from an OpenMP
// worksharing loop (src/subhalo_merge.cpp:207-210, `if(ParallelizeHaloes)`).  So: the device table is guarded by
// g_table_mutex, every library call on a context runs under that Device's own mutex (devices of a sharded batch still run
// concurrently), Device objects are never freed while the process lives (a context is re-created in place, under its mutex,
// when HBTConfig changed), and concurrent Unbind callers are COMBINED into one batch per flight (UnbindCombiner below)
// instead of queueing one-subhalo batches behind a lock.
// grow-only pinned host buffer (hbtu_host_alloc): from page-locked memory the library's uploads are asynchronous DMA that
// overlaps its kernels; the page-locking itself is paid once per process, not per snapshot
template <class T>
struct PinnedBuffer
{
  T *p = nullptr;
  size_t cap = 0;
  T *get(size_t n)
  {
    if (n > cap)
    {
      hbtu_host_free(p);
      cap = n + n / 8 + 1024;
      p = static_cast<T *>(hbtu_host_alloc(cap * sizeof(T)));
      if (!p)
      {
        cap = 0;
        throw std::runtime_error("hbtu_host_alloc failed (pinned staging buffer)");
      }
    }
    return p;
  }
};

struct Device
{
  hbtu_ctx *ctx = nullptr;
  std::mutex mu; // guards ctx AND the staging buffers below
  PinnedBuffer<float> pos_mass, vel;
  PinnedBuffer<int32_t> order;
  PinnedBuffer<float> energy;
  std::vector<Particle_t> all; // batch-wide copy of the particle records (grow-only: its pages are faulted in once per process)
};
std::mutex g_table_mutex;
// a freshly allocated block of >= 4 MB that is about to be written once, front to back, by all threads: ask for transparent huge
// pages so that first-touching it takes one page fault per 2 MB instead of per 4 KB (no effect where THP is off)
inline void prefer_huge_pages(void *p, size_t bytes)
{
#ifdef __linux__
  const size_t kHuge = size_t(2) << 20;
  if (bytes < 2 * kHuge) return;
  const uintptr_t a = (reinterpret_cast<uintptr_t>(p) + kHuge - 1) & ~(uintptr_t)(kHuge - 1);
  const uintptr_t b = (reinterpret_cast<uintptr_t>(p) + bytes) & ~(uintptr_t)(kHuge - 1);
  if (b > a) madvise(reinterpret_cast<void *>(a), b - a, MADV_HUGEPAGE);
#else
  (void)p; (void)bytes;
#endif
}
inline double wall_seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
std::mutex g_times_mutex;
double g_last_times[4] = {0, 0, 0, 0}; // seconds of pack / library call / unpack of the last batch, and its particle count
std::vector<Device *> g_devs;    // current device list
std::vector<Device *> g_retired; // devices of an earlier HBT_UNBIND_DEVICES list (contexts destroyed, objects kept)
std::string g_devices;

std::vector<Device *> devices()
{ // one context per device; re-created if the configuration or the device list changed
  std::lock_guard<std::mutex> table_lock(g_table_mutex);
  hbtu_params p;
  fill_params(p);
  const char *list = getenv("HBT_UNBIND_DEVICES");
  const std::string devs = list ? list : std::to_string(p.device);
  if (!g_devs.empty() && devs == g_devices && std::memcmp(&p, &g_params, sizeof(p)) == 0) return g_devs;
  std::vector<int> ids;
  size_t pos = 0;
  while (pos <= devs.size())
  {
    size_t comma = devs.find(',', pos);
    if (comma == std::string::npos) comma = devs.size();
    if (comma > pos) ids.push_back(atoi(devs.substr(pos, comma - pos).c_str()));
    pos = comma + 1;
  }
  if (ids.empty()) throw std::runtime_error("HBT_UNBIND_DEVICES names no device");
  while (g_devs.size() > ids.size())
  { // the list shrank: retire the surplus devices (their contexts go, the objects stay for late holders of the pointer)
    Device *d = g_devs.back();
    g_devs.pop_back();
    std::lock_guard<std::mutex> lk(d->mu);
    if (d->ctx) hbtu_destroy(d->ctx);
    d->ctx = nullptr;
    g_retired.push_back(d);
  }
  while (g_devs.size() < ids.size()) g_devs.push_back(new Device());
  for (size_t i = 0; i < ids.size(); i++)
  {
    Device *d = g_devs[i];
    std::lock_guard<std::mutex> lk(d->mu); // waits for a call in flight on the old context
    if (d->ctx) hbtu_destroy(d->ctx);
    d->ctx = nullptr;
    hbtu_params q = p;
    q.device = ids[i];
    int rc = hbtu_create(&q, &d->ctx);
    if (rc != HBTU_OK)
    {
      g_devices.clear(); // retry from scratch on the next call
      throw std::runtime_error(std::string("hbtu_create failed: ") + hbtu_last_error(nullptr));
    }
  }
  g_params = p;
  g_devices = devs;
  static bool registered = false;
  if (!registered)
  {
    atexit([] {
      for (auto *d : g_devs)
        if (d->ctx) { hbtu_destroy(d->ctx); d->ctx = nullptr; }
    });
    registered = true;
  }
  return g_devs;
}
Device *device0() { return devices()[0]; }

// run f(ctx); a non-zero return code becomes the reference's error style (config_parser.cpp:70).  The _locked form expects the
// caller to hold d->mu already.
template <class F>
void call_library_locked(Device *d, const char *what, F &&f)
{
  if (!d->ctx) throw std::runtime_error(std::string(what) + ": the device was retired (HBT_UNBIND_DEVICES changed during a call)");
  const int rc = f(d->ctx);
  if (rc != HBTU_OK) throw std::runtime_error(std::string(what) + " failed: " + hbtu_last_error(d->ctx));
}
template <class F>
void call_library(Device *d, const char *what, F &&f)
{
  std::lock_guard<std::mutex> lk(d->mu);
  if (!d->ctx) throw std::runtime_error(std::string(what) + ": the device was retired (HBT_UNBIND_DEVICES changed during a call)");
  const int rc = f(d->ctx);
  if (rc != HBTU_OK) throw std::runtime_error(std::string(what) + " failed: " + hbtu_last_error(d->ctx));
}

struct Batch
{ // pack -> call -> unpack of a set of subhaloes given by index into `Subhalos`
  std::vector<Subhalo_t *> subs;
  std::vector<int32_t> sub_flags; // HBTU_SUB_* per subhalo (missing entries = 0): which reference entry point this subhalo goes through
  std::vector<int64_t> part_offset, nest_offset;
  std::vector<int32_t> nest_list;

  void run(const Snapshot_t &epoch, int32_t flags, Device *dev = nullptr)
  {
    const int64_t nsub = subs.size();
    if (nsub == 0) return;
    if (!dev) dev = device0();
    std::lock_guard<std::mutex> device_lock(dev->mu); // the pinned staging buffers belong to the device
    // the caller's compile-time physics variant travels as batch flags
#ifdef NO_STRIPPING
    flags |= HBTU_FLAG_NO_STRIPPING;
#endif
#ifdef UNBIND_WITH_THERMAL_ENERGY
    flags |= HBTU_FLAG_THERMAL_ENERGY;
#endif
    part_offset.assign(nsub + 1, 0);
    for (int64_t s = 0; s < nsub; s++) part_offset[s + 1] = part_offset[s] + (int64_t)subs[s]->Particles.size();
    const int64_t N = part_offset[nsub];
    const double t_begin = wall_seconds();
    // the reference's Unbind also makes one full copy (subhalo_unbind.cpp:409-415); here the copy lives in a grow-only buffer of
    // the device, so that a snapshot does not pay 10 page faults per thousand particles again
    if ((int64_t)dev->all.size() < N)
    {
      std::vector<Particle_t>().swap(dev->all);
      dev->all.resize(N + N / 8);
      prefer_huge_pages(dev->all.data(), dev->all.size() * sizeof(Particle_t));
    }
    Particle_t *all = dev->all.data();
    float *pos_mass = dev->pos_mass.get(4 * (size_t)N), *vel = dev->vel.get(4 * (size_t)N);
    std::vector<hbtu_sub_io> io(nsub);
    // work items of the pack / unpack loops: (subhalo, chunk of <= kChunk particles), so that a dominant subhalo (an AqA2
    // central holds 72 % of the particles) is packed by all threads instead of one
    constexpr int64_t kChunk = 1 << 16;
    std::vector<int64_t> item_sub, item_begin;
    for (int64_t s = 0; s < nsub; s++)
    {
      const int64_t n = part_offset[s + 1] - part_offset[s];
      for (int64_t b0 = 0; b0 < n || b0 == 0; b0 += kChunk)
      {
        item_sub.push_back(s);
        item_begin.push_back(b0);
      }
    }
    const int64_t nitems = (int64_t)item_sub.size();
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t it = 0; it < nitems; it++)
    {
      const Subhalo_t &sub = *subs[item_sub[it]];
      const int64_t b = part_offset[item_sub[it]];
      const int64_t i1 = std::min<int64_t>(item_begin[it] + kChunk, (int64_t)sub.Particles.size());
      for (int64_t i = item_begin[it]; i < i1; i++)
      {
        const Particle_t &p = sub.Particles[i];
        all[b + i] = p;
        float *x = &pos_mass[4 * (b + i)], *v = &vel[4 * (b + i)];
        x[0] = p.ComovingPosition[0]; x[1] = p.ComovingPosition[1]; x[2] = p.ComovingPosition[2]; x[3] = p.Mass;
        v[0] = p.PhysicalVelocity[0]; v[1] = p.PhysicalVelocity[1]; v[2] = p.PhysicalVelocity[2];
#ifdef UNBIND_WITH_THERMAL_ENERGY
        v[3] = p.InternalEnergy; // E += InternalEnergy (src/subhalo_unbind.cpp:351-353), HBTU_FLAG_THERMAL_ENERGY
#else
        v[3] = 0.f;
#endif
      }
    }
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nsub; s++)
    {
      const Subhalo_t &sub = *subs[s];
      hbtu_sub_io &o = io[s];
      std::memset(&o, 0, sizeof(o));
      for (int j = 0; j < 3; j++)
      {
        o.avg_pos[j] = sub.ComovingAveragePosition[j];
        o.avg_vel[j] = sub.PhysicalAverageVelocity[j];
        o.mostbound_pos[j] = sub.ComovingMostBoundPosition[j];
        o.mostbound_vel[j] = sub.PhysicalMostBoundVelocity[j];
        o.specific_angular_momentum[j] = sub.SpecificAngularMomentum[j];
      }
      o.nbound = sub.Nbound;
      o.sink_track_id = sub.SinkTrackId;
      o.snapshot_index_of_death = sub.SnapshotIndexOfDeath;
      o.snapshot_index_of_sink = sub.SnapshotIndexOfSink;
      o.mbound = sub.Mbound;
      o.specific_self_potential_energy = sub.SpecificSelfPotentialEnergy;
      o.specific_self_kinetic_energy = sub.SpecificSelfKineticEnergy;
      o.flags = (size_t)s < sub_flags.size() ? sub_flags[s] : 0;
    }
    hbtu_epoch e;
    e.scale_factor = epoch.Cosmology.ScaleFactor;
    e.hz = epoch.Cosmology.Hz;
    e.snapshot_index = epoch.GetSnapshotIndex();
    e.reserved = 0;
    const int64_t *no = nest_offset.empty() ? nullptr : nest_offset.data();
    const int32_t *nl = nest_offset.empty() ? nullptr : nest_list.data();
    int64_t cap = hbtu_order_capacity(nsub, part_offset.data(), no, nl);
    if (cap < 0) throw std::runtime_error("hbtu_order_capacity: malformed nesting");
    std::vector<int64_t> order_offset(nsub + 1);
    int32_t *order = dev->order.get(cap > 0 ? cap : 1);
#ifdef SAVE_BINDING_ENERGY
    float *energy = dev->energy.get(cap > 0 ? cap : 1);
    float *pe = energy;
#else
    float *pe = nullptr;
#endif
    const double t_packed = wall_seconds();
    call_library_locked(dev, "hbtu_unbind_batch", [&](hbtu_ctx *ctx) {
      return hbtu_unbind_batch(ctx, &e, nsub, part_offset.data(), pos_mass, vel, no, nl, io.data(), flags, cap, order_offset.data(), order, pe);
    });
    const double t_called = wall_seconds();
    // unpack: the new particle lists (a list can hold particles of nested subhaloes: gather from the batch-wide copy), in chunks
    // like the pack; the vectors are resized first (serially per subhalo, cheap) so that the chunks can write independently
    // A list that GROWS (a host receives the particles stripped from its nests) must not go through vector::resize: that copies
    // the old contents - about to be overwritten - on ONE thread and faults the new block in serially (0.4 s for the 3e7-particle
    // central of the bench's drop-in row).  Drop the old block first; the new one is first touched by the parallel gather below.
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t s = 0; s < nsub; s++)
    {
      std::vector<Particle_t> &pl = subs[s]->Particles;
      const bool fresh = (size_t)io[s].nsource > pl.capacity();
      if (fresh) std::vector<Particle_t>().swap(pl);
      pl.resize(io[s].nsource);
      if (fresh) prefer_huge_pages(pl.data(), pl.size() * sizeof(Particle_t));
    }
    std::vector<int64_t> out_sub, out_begin;
    for (int64_t s = 0; s < nsub; s++)
      for (int64_t b0 = 0; b0 < io[s].nsource; b0 += kChunk)
      {
        out_sub.push_back(s);
        out_begin.push_back(b0);
      }
    const int64_t nout = (int64_t)out_sub.size();
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t it = 0; it < nout; it++)
    {
      const int64_t s = out_sub[it];
      Subhalo_t &sub = *subs[s];
      const int32_t *ord = &order[order_offset[s]];
      const int64_t i1 = std::min<int64_t>(out_begin[it] + kChunk, io[s].nsource);
      Particle_t *dst = sub.Particles.data();
      for (int64_t i = out_begin[it]; i < i1; i++)
      { // random 40-byte reads from the batch-wide copy: keep a few cache misses in flight
        if (i + 12 < i1) __builtin_prefetch(&all[ord[i + 12]]);
        dst[i] = all[ord[i]];
      }
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t s = 0; s < nsub; s++)
    {
      Subhalo_t &sub = *subs[s];
      const hbtu_sub_io &o = io[s];
      for (int j = 0; j < 3; j++)
      {
        sub.ComovingAveragePosition[j] = o.avg_pos[j];
        sub.PhysicalAverageVelocity[j] = o.avg_vel[j];
        sub.ComovingMostBoundPosition[j] = o.mostbound_pos[j];
        sub.PhysicalMostBoundVelocity[j] = o.mostbound_vel[j];
        sub.SpecificAngularMomentum[j] = o.specific_angular_momentum[j];
      }
      sub.Nbound = (HBTInt)o.nbound;
      sub.Mbound = o.mbound;
      sub.SinkTrackId = (HBTInt)o.sink_track_id;
      sub.SnapshotIndexOfDeath = o.snapshot_index_of_death;
      sub.SnapshotIndexOfSink = o.snapshot_index_of_sink;
      sub.SpecificSelfPotentialEnergy = o.specific_self_potential_energy;
      sub.SpecificSelfKineticEnergy = o.specific_self_kinetic_energy;
#ifdef SAVE_BINDING_ENERGY
      sub.Energies.assign(&energy[order_offset[s]], &energy[order_offset[s]] + (o.nbound < o.nsource ? o.nbound : o.nsource));
#endif
#ifndef DM_ONLY
      if (sub.Particles.size() >= 2) sub.CountParticleTypes(); else sub.CountParticles(); // host bookkeeping stays on the host
#endif
    }
    const double t_end = wall_seconds();
    {
      std::lock_guard<std::mutex> lk(g_times_mutex);
      g_last_times[0] = t_packed - t_begin;
      g_last_times[1] = t_called - t_packed;
      g_last_times[2] = t_end - t_called;
      g_last_times[3] = (double)N;
    }
    if (getenv("HBT_B200_TRACE"))
      fprintf(stderr, "[hbt_b200] batch of %lld subhaloes / %lld particles: pack %.1f ms, hbtu_unbind_batch %.1f ms, unpack %.1f ms\n",
              (long long)nsub, (long long)N, 1e3 * (t_packed - t_begin), 1e3 * (t_called - t_packed), 1e3 * (t_end - t_called));
  }
};

// a subhalo the reference unbinds with plain Subhalo_t::Unbind (no RecursiveUnbind, hence no orphan rule)
void add_plain(Batch &b, std::vector<std::vector<int32_t>> &lists, Subhalo_t &sub)
{
  b.sub_flags.resize(b.subs.size(), 0);
  b.subs.push_back(&sub);
  b.sub_flags.push_back(HBTU_SUB_PLAIN_UNBIND);
  lists.emplace_back();
}

// append `sub` and, depth first, everything nested in it; returns its batch index
int64_t add_hierarchy(Batch &b, std::vector<std::vector<int32_t>> &lists, SubhaloList_t &Subhalos, Subhalo_t &sub)
{
  int64_t me = b.subs.size();
  b.subs.push_back(&sub);
  lists.emplace_back();
  for (HBTInt i = 0; i < (HBTInt)sub.NestedSubhalos.size(); i++)
  {
    int64_t ch = add_hierarchy(b, lists, Subhalos, Subhalos[sub.NestedSubhalos[i]]);
    lists[me].push_back((int32_t)ch);
  }
  return me;
}
void close_nests(Batch &b, const std::vector<std::vector<int32_t>> &lists);

// Multi-GPU inside one process.  The batch is cut at hierarchy boundaries (add_hierarchy appends a root and
// everything nested in it contiguously, so a unit is the index range up to the next root), the units are dealt to the
// devices longest-processing-time-first with cost sum n*log2(n) per unit, and every device runs its sub-batch on its own
// host thread and context.  Results do not depend on the split: hierarchies never interact.
void run_sharded(Batch &b, const std::vector<std::vector<int32_t>> &lists, const Snapshot_t &epoch, int32_t flags)
{
  const std::vector<Device *> ctxs = devices();
  const int64_t nsub = b.subs.size();
  const int G = (int)ctxs.size();
  if (G == 1 || nsub < 2)
  {
    b.run(epoch, flags, ctxs[0]);
    return;
  }
  std::vector<char> is_child(nsub, 0);
  for (auto &l : lists)
    for (int32_t ch : l) is_child[ch] = 1;
  struct Unit { int64_t first, last; double cost; };
  std::vector<Unit> units;
  for (int64_t s = 0; s < nsub; s++)
  {
    const double n = (double)b.subs[s]->Particles.size();
    const double c = n * std::log2(n + 2.0);
    if (!is_child[s]) units.push_back(Unit{s, s + 1, c});
    else { units.back().last = s + 1; units.back().cost += c; }
  }
  std::vector<size_t> by_cost(units.size());
  for (size_t i = 0; i < units.size(); i++) by_cost[i] = i;
  std::stable_sort(by_cost.begin(), by_cost.end(), [&](size_t x, size_t y) { return units[x].cost > units[y].cost; });
  std::vector<double> load(G, 0.0);
  std::vector<std::vector<size_t>> mine(G);
  for (size_t u : by_cost)
  {
    int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    load[g] += units[u].cost;
    mine[g].push_back(u);
  }
  std::vector<Batch> parts(G);
  std::vector<std::vector<std::vector<int32_t>>> part_lists(G);
  for (int g = 0; g < G; g++)
  {
    std::sort(mine[g].begin(), mine[g].end()); // keep the reference's visiting order inside a device
    std::vector<int32_t> remap(nsub, -1);
    for (size_t u : mine[g])
      for (int64_t s = units[u].first; s < units[u].last; s++)
      {
        remap[s] = (int32_t)parts[g].subs.size();
        parts[g].subs.push_back(b.subs[s]);
        parts[g].sub_flags.push_back((size_t)s < b.sub_flags.size() ? b.sub_flags[s] : 0);
      }
    part_lists[g].resize(parts[g].subs.size());
    for (size_t u : mine[g])
      for (int64_t s = units[u].first; s < units[u].last; s++)
        for (int32_t ch : lists[s]) part_lists[g][remap[s]].push_back(remap[ch]);
    close_nests(parts[g], part_lists[g]);
  }
  std::vector<std::thread> workers;
  std::vector<std::string> errors(G);
  for (int g = 0; g < G; g++)
    workers.emplace_back([&, g] {
      try { parts[g].run(epoch, flags, ctxs[g]); }
      catch (const std::exception &ex) { errors[g] = ex.what(); }
    });
  for (auto &w : workers) w.join();
  for (auto &e : errors)
    if (!e.empty()) throw std::runtime_error(e);
}

void close_nests(Batch &b, const std::vector<std::vector<int32_t>> &lists)
{
  b.nest_offset.assign(lists.size() + 1, 0);
  b.nest_list.clear();
  for (size_t s = 0; s < lists.size(); s++)
  {
    b.nest_list.insert(b.nest_list.end(), lists[s].begin(), lists[s].end());
    b.nest_offset[s + 1] = b.nest_list.size();
  }
}
} // namespace

namespace
{
// Subhalo_t::Unbind is called one subhalo at a time, concurrently from the OpenMP threads of the merge loop
// (src/subhalo_merge.cpp:207-210).  Flat combining: a caller queues its request; whoever finds no batch in flight becomes the
// leader, takes everything queued for the same epoch (its own request included), runs ONE hbtu_unbind_batch for it and wakes
// the owners.  Every call still returns only when its own subhalo is done, so the unmodified caller sees the reference's
// synchronous semantics; with T threads the merged hosts go through in batches of up to T subhaloes instead of one by one.
// (Results do not depend on how subhaloes are grouped into batches, bit for bit: DESIGN.md section 7.)
struct UnbindCombiner
{
  struct Request
  {
    Subhalo_t *sub;
    const Snapshot_t *epoch;
    bool done;
    std::string error;
  };
  std::mutex m;
  std::condition_variable cv;
  std::vector<Request *> queue;
  bool running = false;

  void submit(Subhalo_t *sub, const Snapshot_t &epoch)
  {
    Request r{sub, &epoch, false, std::string()};
    std::unique_lock<std::mutex> lk(m);
    queue.push_back(&r);
    while (!r.done)
    {
      if (running)
      {
        cv.wait(lk);
        continue;
      }
      running = true;
      std::vector<Request *> take, rest;
      const Snapshot_t *ep = queue.front()->epoch;
      for (Request *q : queue) (q->epoch == ep ? take : rest).push_back(q);
      queue.swap(rest);
      lk.unlock();
      std::string error;
      try
      {
        Batch b;
        std::vector<std::vector<int32_t>> lists;
        for (Request *q : take) add_plain(b, lists, *q->sub);
        b.run(*ep, 0);
      }
      catch (const std::exception &ex) { error = ex.what(); if (error.empty()) error = "unbinding failed"; }
      lk.lock();
      for (Request *q : take)
      {
        q->error = error;
        q->done = true;
      }
      running = false;
      cv.notify_all();
    }
    lk.unlock();
    if (!r.error.empty()) throw std::runtime_error(r.error);
  }
} g_unbind_combiner;
} // namespace

// where the wall time of the last batch went (seconds: AoS -> pinned SoA pack, hbtu_unbind_batch, permutation of the
// vector<Particle_t>s; out[3] = its particle count); for the bench's drop-in row and for a maintainer's own timing
extern "C" void HBT_B200_LastBatchTimes(double out[4])
{
  std::lock_guard<std::mutex> lk(g_times_mutex);
  for (int i = 0; i < 4; i++) out[i] = g_last_times[i];
}

void Subhalo_t::Unbind(const Snapshot_t &epoch)
{ // second caller: subhalo_merge.cpp:207-210 (merged hosts), one subhalo per call, from concurrent OpenMP threads
  g_unbind_combiner.submit(this, epoch);
}

// The batched form of the merge path.  src/subhalo_merge.cpp:207-214 of the reference reads
//     #pragma omp parallel for schedule(dynamic,1) if(ParallelizeHaloes)
//     for(subid...) if(Helpers[subid].IsMerged) Subhalos[subid].Unbind(*this);
//     #pragma omp parallel for
//     for(subid...) if(Helpers[subid].IsMerged) Subhalos[subid].TruncateSource();
// and becomes   HBT_B200_UnbindMerged(*this, merged);   with merged[subid] = Helpers[subid].IsMerged:
// ONE batch (sharded over HBT_UNBIND_DEVICES like RefineParticles) with the truncation done by the library.
void HBT_B200_UnbindMerged(SubhaloSnapshot_t &snap, const std::vector<char> &is_merged)
{
  Batch b;
  std::vector<std::vector<int32_t>> lists;
  for (size_t i = 0; i < snap.Subhalos.size() && i < is_merged.size(); i++)
    if (is_merged[i]) add_plain(b, lists, snap.Subhalos[i]);
  if (b.subs.empty()) return;
  close_nests(b, lists);
  run_sharded(b, lists, snap, HBTU_FLAG_TRUNCATE_SOURCE);
}

void Subhalo_t::RecursiveUnbind(SubhaloList_t &Subhalos, const Snapshot_t &snap)
{
  Batch b;
  std::vector<std::vector<int32_t>> lists;
  add_hierarchy(b, lists, Subhalos, *this);
  close_nests(b, lists);
  b.run(snap, 0);
}

void Subhalo_t::TruncateSource()
{ // pure host bookkeeping (8 lines in the reference, src/subhalo_unbind.cpp:449-458); RefineParticles below lets the
  // library truncate, this is for the merge path that calls it separately (subhalo_merge.cpp:211-214)
  HBTInt n = Nbound <= 1 ? Nbound : (HBTInt)(Nbound * HBTConfig.SourceSubRelaxFactor);
  if (n > (HBTInt)Particles.size()) n = Particles.size();
  Particles.resize(n);
}

void SubhaloSnapshot_t::RefineParticles()
{ // ONE batch for the whole rank: every host halo's hierarchy, the field subhaloes and the new-born ones
  Batch b;
  std::vector<std::vector<int32_t>> lists;
  std::vector<char> done(Subhalos.size(), 0);
#ifdef INCLUSIVE_MASS
  for (auto &sub : Subhalos) add_plain(b, lists, sub); // flat loop of plain Unbind calls (subhalo_unbind.cpp:470-476)
#else
  HBTInt NumHalos = MemberTable.SubGroups.size();
  for (HBTInt haloid = 0; haloid < NumHalos; haloid++)
  {
    auto &subgroup = MemberTable.SubGroups[haloid];
    if (subgroup.size() == 0) continue;
    auto &central = Subhalos[subgroup[0]];
    // the other heads of this host feed the central like nested subhaloes (subhalo_unbind.cpp:485-492)
    auto &heads = MemberTable.SubGroupsOfHeads[haloid];
    int64_t me = add_hierarchy(b, lists, Subhalos, central);
    for (size_t i = 1; i < heads.size(); i++)
    {
      int64_t ch = add_hierarchy(b, lists, Subhalos, Subhalos[heads[i]]);
      lists[me].push_back((int32_t)ch);
    }
  }
  for (auto *s : b.subs) done[s - Subhalos.data()] = 1;
  HBTInt NumField = MemberTable.SubGroups[-1].size();
  for (HBTInt i = 0; i < NumField; i++)
  { // field subhaloes: plain Unbind, no recursion (subhalo_unbind.cpp:498-503)
    HBTInt subid = MemberTable.SubGroups[-1][i];
    if (done[subid]) continue;
    add_plain(b, lists, Subhalos[subid]);
    done[subid] = 1;
  }
  for (HBTInt i = MemberTable.AllMembers.size(); i < (HBTInt)Subhalos.size(); i++)
  { // new-born subhaloes (subhalo_unbind.cpp:505-510)
    if (done[i]) continue;
    add_plain(b, lists, Subhalos[i]);
    done[i] = 1;
  }
#endif
  close_nests(b, lists);
  run_sharded(b, lists, *this, HBTU_FLAG_TRUNCATE_SOURCE);
  // subhaloes the reference's loops never unbind (members of a host whose nest is not reachable from a head) are
  // still truncated by its last loop (subhalo_unbind.cpp:511-513)
  for (size_t i = 0; i < Subhalos.size(); i++)
    if (!done[i]) Subhalos[i].TruncateSource();
}

// ---------------------------------------------------------------------------------------------------
// Post-unbinding properties.  The reference computes them one subhalo at a time at
// the end of SubhaloSnapshot_t::UpdateTracks (src/subhalo_tracking.cpp:901-906):
//     for(i...) { Subhalos[i].CalculateProfileProperties(*this); Subhalos[i].CalculateShape(); }
// Those two members live in src/subhalo.cpp, which is NOT replaced; the maintainer swaps that loop for one call of
//     HBT_B200_CalculateProperties(Subhalos, *this);        // declared in integration/hbt_b200.h
// which packs the bound part of every particle list, calls hbtu_profile_batch and writes the same members back.
// The eigen-vectors (EigenAxis, HAS_GSL builds, src/subhalo.cpp:393-396) are a 3x3 problem per subhalo and stay on
// the host, computed from the tensors returned here.
void HBT_B200_CalculateProperties(SubhaloList_t &Subhalos, const Snapshot_t &epoch)
{
  const int64_t nsub = Subhalos.size();
  if (nsub == 0) return;
  std::vector<int64_t> part_offset(nsub + 1, 0);
  for (int64_t s = 0; s < nsub; s++)
  { // only the bound particles are read (Nbound <= 1: nothing is, the outputs are zeroed)
    const int64_t nb = Subhalos[s].Nbound > 1 ? (int64_t)Subhalos[s].Nbound : 0;
    part_offset[s + 1] = part_offset[s] + nb;
  }
  std::vector<float> pos_mass(4 * (size_t)part_offset[nsub]);
  std::vector<hbtu_profile_io> io(nsub);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t s = 0; s < nsub; s++)
  {
    const Subhalo_t &sub = Subhalos[s];
    const int64_t b = part_offset[s], nb = part_offset[s + 1] - b;
    for (int64_t i = 0; i < nb; i++)
    {
      const Particle_t &p = sub.Particles[i];
      float *x = &pos_mass[4 * (b + i)];
      x[0] = p.ComovingPosition[0]; x[1] = p.ComovingPosition[1]; x[2] = p.ComovingPosition[2]; x[3] = p.Mass;
    }
    hbtu_profile_io &o = io[s];
    std::memset(&o, 0, sizeof(o));
    for (int j = 0; j < 3; j++) o.mostbound_pos[j] = sub.ComovingMostBoundPosition[j];
    o.nbound = nb; // 0 or Nbound
    o.mbound = sub.Mbound;
    o.last_max_vmax_physical = sub.LastMaxVmaxPhysical;
    o.snapshot_index_of_last_max_vmax = sub.SnapshotIndexOfLastMaxVmax;
    o.bound_r200crit_comoving = sub.BoundR200CritComoving;
    o.bound_m200crit = sub.BoundM200Crit;
  }
  hbtu_epoch e;
  e.scale_factor = epoch.Cosmology.ScaleFactor;
  e.hz = epoch.Cosmology.Hz;
  e.snapshot_index = epoch.GetSnapshotIndex();
  e.reserved = 0;
  call_library(device0(), "hbtu_profile_batch",
               [&](hbtu_ctx *ctx) { return hbtu_profile_batch(ctx, &e, nsub, part_offset.data(), pos_mass.data(), io.data()); });
  for (int64_t s = 0; s < nsub; s++)
  {
    Subhalo_t &sub = Subhalos[s];
    const hbtu_profile_io &o = io[s];
    sub.RmaxComoving = o.rmax_comoving;
    sub.VmaxPhysical = o.vmax_physical;
    sub.LastMaxVmaxPhysical = o.last_max_vmax_physical;
    sub.SnapshotIndexOfLastMaxVmax = o.snapshot_index_of_last_max_vmax;
    sub.R2SigmaComoving = o.r2sigma_comoving;
    sub.RHalfComoving = o.rhalf_comoving;
    sub.BoundR200CritComoving = o.bound_r200crit_comoving;
    sub.BoundM200Crit = o.bound_m200crit;
    for (int j = 0; j < 6; j++)
    {
      sub.InertialTensor[j] = o.inertial_tensor[j];
      sub.InertialTensorWeighted[j] = o.inertial_tensor_weighted[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Source preparation.  SubhaloSnapshot_t::MaskSubhalos (src/subhalo_tracking.cpp:
// 824-841, called from PrepareCentrals at :531) is a private member defined in a file that is NOT replaced; the maintainer
// swaps that one call for
//     HBT_B200_MaskSubhalos(*this);                          // declared in integration/hbt_b200.h
// which sends every host's hierarchy (central + old nests + the other heads, exactly the temporary append of :832-838) with
// the particle Ids to hbtu_mask_batch and shrinks the particle lists to the entries it keeps.
void HBT_B200_MaskSubhalos(SubhaloSnapshot_t &snap)
{
  SubhaloList_t &Subhalos = snap.Subhalos;
  Batch b;
  std::vector<std::vector<int32_t>> lists;
  for (HBTInt haloid = 0; haloid < (HBTInt)snap.MemberTable.SubGroups.size(); haloid++)
  {
    auto &subgroup = snap.MemberTable.SubGroups[haloid];
    if (subgroup.size() == 0) continue;
    auto &heads = snap.MemberTable.SubGroupsOfHeads[haloid];
    int64_t me = add_hierarchy(b, lists, Subhalos, Subhalos[subgroup[0]]);
    for (size_t i = 1; i < heads.size(); i++) lists[me].push_back((int32_t)add_hierarchy(b, lists, Subhalos, Subhalos[heads[i]]));
  }
  close_nests(b, lists);
  const int64_t nsub = b.subs.size();
  if (nsub == 0) return;
  std::vector<int64_t> part_offset(nsub + 1, 0), nbound(nsub), new_count(nsub);
  for (int64_t s = 0; s < nsub; s++)
  {
    part_offset[s + 1] = part_offset[s] + (int64_t)b.subs[s]->Particles.size();
    nbound[s] = b.subs[s]->Nbound;
  }
  std::vector<int64_t> ids(part_offset[nsub]);
  std::vector<int32_t> keep(part_offset[nsub] > 0 ? part_offset[nsub] : 1);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t s = 0; s < nsub; s++)
    for (size_t i = 0; i < b.subs[s]->Particles.size(); i++) ids[part_offset[s] + i] = b.subs[s]->Particles[i].Id;
  call_library(device0(), "hbtu_mask_batch", [&](hbtu_ctx *ctx) {
    return hbtu_mask_batch(ctx, nsub, part_offset.data(), ids.data(), b.nest_offset.data(), b.nest_list.data(), nbound.data(),
                           new_count.data(), keep.data());
  });
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t s = 0; s < nsub; s++)
  { // keep is ascending: compact in place, like the reference's move loop (:809-820)
    auto &P = b.subs[s]->Particles;
    const int32_t *k = &keep[part_offset[s]];
    for (int64_t i = 0; i < new_count[s]; i++)
    {
      const int64_t src = k[i] - part_offset[s];
      if (src != i) P[i] = std::move(P[src]);
    }
    P.resize(new_count[s]);
  }
}

// ---------------------------------------------------------------------------------------------------
// Merger trap detection.  SubhaloSnapshot_t::MergeSubhalos (src/subhalo_merge.cpp:187-221,
// not replaced) opens with
//     #pragma omp parallel
//     { GlueHeadNests(); FillHelpers(Helpers, Subhalos); DetectTraps(Subhalos, Helpers, isnap); }
// whose helpers are file-local.  The maintainer replaces FillHelpers + DetectTraps by
//     HBT_B200_DetectTraps(*this, merged);   // then: Helpers[i].IsMerged = merged[i]
// (called outside the parallel region; it builds the glued host relation itself from MemberTable, like RefineParticles does).
void HBT_B200_DetectTraps(SubhaloSnapshot_t &snap, std::vector<char> &is_merged)
{
  SubhaloList_t &Subhalos = snap.Subhalos;
  const int64_t nsub = Subhalos.size();
  is_merged.assign(nsub, 0);
  if (nsub == 0) return;
  // host relation by subhalo index: NestedSubhalos + the other heads of every host halo under its central (GlueHeadNests)
  std::vector<std::vector<int32_t>> lists(nsub);
  for (int64_t s = 0; s < nsub; s++)
    for (auto ch : Subhalos[s].NestedSubhalos) lists[s].push_back((int32_t)ch);
  for (HBTInt haloid = 0; haloid < (HBTInt)snap.MemberTable.SubGroups.size(); haloid++)
  {
    auto &subgroup = snap.MemberTable.SubGroups[haloid];
    if (subgroup.size() == 0) continue;
    auto &heads = snap.MemberTable.SubGroupsOfHeads[haloid];
    for (size_t i = 1; i < heads.size(); i++) lists[subgroup[0]].push_back((int32_t)heads[i]);
  }
  std::vector<int64_t> nest_offset(nsub + 1, 0), part_offset(nsub + 1, 0);
  std::vector<int32_t> nest_list;
  for (int64_t s = 0; s < nsub; s++)
  {
    nest_list.insert(nest_list.end(), lists[s].begin(), lists[s].end());
    nest_offset[s + 1] = nest_list.size();
    const int64_t nb = Subhalos[s].Nbound < 0 ? 0 : (int64_t)Subhalos[s].Nbound;
    part_offset[s + 1] = part_offset[s] + (nb < 20 ? nb : 20); // only the <= NumPartCoreMax most bound particles are read
  }
  std::vector<float> pos_mass(4 * (size_t)part_offset[nsub]), vel(4 * (size_t)part_offset[nsub]);
  std::vector<hbtu_trap_io> io(nsub);
  for (int64_t s = 0; s < nsub; s++)
  {
    const Subhalo_t &sub = Subhalos[s];
    for (int64_t i = 0; i < part_offset[s + 1] - part_offset[s]; i++)
    {
      const Particle_t &p = sub.Particles[i];
      float *x = &pos_mass[4 * (part_offset[s] + i)], *v = &vel[4 * (part_offset[s] + i)];
      x[0] = p.ComovingPosition[0]; x[1] = p.ComovingPosition[1]; x[2] = p.ComovingPosition[2]; x[3] = p.Mass;
      v[0] = p.PhysicalVelocity[0]; v[1] = p.PhysicalVelocity[1]; v[2] = p.PhysicalVelocity[2]; v[3] = 0.f;
    }
    hbtu_trap_io &o = io[s];
    for (int j = 0; j < 3; j++)
    {
      o.mostbound_pos[j] = sub.ComovingMostBoundPosition[j];
      o.mostbound_vel[j] = sub.PhysicalMostBoundVelocity[j];
    }
    o.nbound = sub.Nbound < 0 ? 0 : (int64_t)sub.Nbound; // the list handed over holds min(Nbound, 20) particles
    o.sink_track_id = sub.SinkTrackId;
    o.snapshot_index_of_sink = sub.SnapshotIndexOfSink;
    o.is_merged = 0;
  }
  hbtu_epoch e;
  e.scale_factor = snap.Cosmology.ScaleFactor;
  e.hz = snap.Cosmology.Hz;
  e.snapshot_index = snap.GetSnapshotIndex();
  e.reserved = 0;
  call_library(device0(), "hbtu_detect_traps", [&](hbtu_ctx *ctx) {
    return hbtu_detect_traps(ctx, &e, nsub, part_offset.data(), pos_mass.data(), vel.data(), nest_offset.data(), nest_list.data(), io.data());
  });
  for (int64_t s = 0; s < nsub; s++)
  {
    Subhalos[s].SinkTrackId = (HBTInt)io[s].sink_track_id;
    Subhalos[s].SnapshotIndexOfSink = io[s].snapshot_index_of_sink;
    is_merged[s] = (char)io[s].is_merged;
  }
}
