// profile.cu - post-unbinding per-subhalo properties on the device, sm_100a.
//
// Replaces the per-subhalo loop of SubhaloSnapshot_t::UpdateTracks (src/subhalo_tracking.cpp:901-906):
//   Subhalo_t::CalculateProfileProperties  src/subhalo.cpp:242-332   radius sort, cumulative mass, Vmax/Rmax, RHalf,
//                                                                     R2Sigma, 200 rho_crit overdensity size
//   Subhalo_t::CalculateShape              src/subhalo.cpp:334-398   plain and 1/r^2-weighted inertia tensors
//   Snapshot_t::SphericalOverdensitySize   src/snapshot.cpp:264-281  scan from the outermost particle inwards
//   PeriodicDistance                       src/config_parser.h:143-156
// as segmented, batched passes over the concatenated bound particles of all subhaloes (B = sum Nbound):
//
//   radius + shape (K1)   16 B read per particle, 12 B key/index write; fp64 sums of the tensor terms with a fixed
//                         summation tree (seg_reduce.cuh)                                                          HBM
//   CUB radix sort        (subhalo, radius bits) 64-bit key + 32-bit index                                       HBM
//   CUB scan-by-key       cumulative mass in double (the reference's serial double sum; exact for equal masses)  HBM
//   select (K3)           v^2 = M(<r)/max(r,eps): first maximum, last radius with M > 200 rho_crit (4/3 pi) r^3  HBM
//   finalize (K4)         one thread per subhalo
//
// Arithmetic widths follow the reference expressions (HBTReal = float products without FMA contraction, double sums).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <cmath>
#include <vector>

#include "context.cuh"
#include "seg_reduce.cuh"
#include "det_scan.cuh"

namespace hbt
{

static constexpr int kPB = 256;
static inline int pgrid(int64_t n) { return n > 0 ? div_up(n, kPB) : 1; }

struct ProfSub
{ // per subhalo, uploaded once
  int64_t part_off;  // first particle of the list in the batch arrays
  int64_t bound_off; // first element of this subhalo in the B-concatenation
  int nb;            // Nbound if > 1, else 0 (nothing to compute, src/subhalo.cpp:265,336)
  float cx, cy, cz;  // ComovingMostBoundPosition as HBTReal
  float mbound;
};

struct ProfScratch
{ // per subhalo, device
  double sums[12];           // Ixx Ixy Ixz Iyy Iyz Izz, then the weighted six
  unsigned long long argmax; // (bits of v^2) << 32 | ~index : atomicMax gives the first of the largest
  int so_last;               // 1 + last sorted index whose enclosed mass exceeds 200 rho_crit, 0 = none
  int pad;
};

__device__ __forceinline__ int prof_find(const ProfSub *__restrict__ subs, int n, int64_t e)
{ // largest s with bound_off[s] <= e (subhaloes without bound particles share their successor's offset: skip them)
  int lo = 0, hi = n;
  while (hi - lo > 1)
  {
    int mid = (lo + hi) >> 1;
    if (subs[mid].bound_off <= e) lo = mid; else hi = mid;
  }
  return lo;
}

struct ProfDone
{
  ProfScratch *scr;
  __device__ void operator()(int s, const double (&x)[12]) const
  {
#pragma unroll
    for (int j = 0; j < 12; j++) scr[s].sums[j] = x[j];
  }
};

// K1: radius to the most-bound position (sort key) and the inertia-tensor terms.  Block b works on one 256-particle chunk
// of one subhalo's bound list (chunk table, seg_reduce.cuh), so the tensor sums have a summation tree that depends on the
// subhalo alone.
__global__ void __launch_bounds__(kPB) prof_radius_shape_kernel(const ProfSub *__restrict__ subs, int nsub, const int *__restrict__ chunk_off,
                                                                 const float4 *__restrict__ pos, const int *__restrict__ ids, DevConfig cfg,
                                                                 uint64_t *__restrict__ key, int *__restrict__ val, double *__restrict__ partial)
{
  int s, c;
  seg_chunk_of_block(chunk_off, nsub, blockIdx.x, s, c);
  const ProfSub sb = subs[s];
  const int i = c * kPB + threadIdx.x;
  double v[12];
#pragma unroll
  for (int j = 0; j < 12; j++) v[j] = 0.0;
  if (i < sb.nb)
  {
    const int64_t e = sb.bound_off + i;
    const float4 p = ids ? pos[ids[sb.part_off + i]] : pos[sb.part_off + i]; // ids: the Elist of a resident unbinding batch
    float dx = __fsub_rn(p.x, sb.cx), dy = __fsub_rn(p.y, sb.cy), dz = __fsub_rn(p.z, sb.cz);
    // the reference takes cen - pos for the radius and pos - cen for the tensor: the squares and the pair products agree
    if (cfg.periodic)
    {
      dx = nearest_f(dx, cfg.box_size, cfg.box_half);
      dy = nearest_f(dy, cfg.box_size, cfg.box_half);
      dz = nearest_f(dz, cfg.box_size, cfg.box_half);
    }
    const float dx2 = __fmul_rn(dx, dx), dy2 = __fmul_rn(dy, dy), dz2 = __fmul_rn(dz, dz);
    const float r = __fsqrt_rn(__fadd_rn(__fadd_rn(dx2, dy2), dz2)); // PeriodicDistance
    key[e] = ((uint64_t)(uint32_t)s << 32) | __float_as_uint(r);
    val[e] = i;
    if (i >= 1)
    { // src/subhalo.cpp:354-386: HBTReal products, double accumulation
      const float m = p.w;
      v[0] = (double)__fmul_rn(dx2, m);
      v[1] = (double)__fmul_rn(__fmul_rn(dx, dy), m);
      v[2] = (double)__fmul_rn(__fmul_rn(dx, dz), m);
      v[3] = (double)__fmul_rn(dy2, m);
      v[4] = (double)__fmul_rn(__fmul_rn(dy, dz), m);
      v[5] = (double)__fmul_rn(dz2, m);
      float dr2 = __fadd_rn(__fadd_rn(dx2, dy2), dz2);
      dr2 = __fdiv_rn(dr2, m);
      v[6] = (double)__fdiv_rn(dx2, dr2);
      v[7] = (double)__fdiv_rn(__fmul_rn(dx, dy), dr2);
      v[8] = (double)__fdiv_rn(__fmul_rn(dx, dz), dr2);
      v[9] = (double)__fdiv_rn(dy2, dr2);
      v[10] = (double)__fdiv_rn(__fmul_rn(dy, dz), dr2);
      v[11] = (double)__fdiv_rn(dz2, dr2);
    }
  }
  seg_reduce_chunk<12>(v, partial);
}

__global__ void __launch_bounds__(kPB) prof_finish_kernel(int nsub, const int *__restrict__ chunk_off, const double *__restrict__ partial,
                                                           ProfScratch *__restrict__ scr)
{
  const int a = blockIdx.x;
  if (a >= nsub) return;
  seg_reduce_finish_block<12>(a, chunk_off, partial, ProfDone{scr});
}

struct SortedMass
{ // mass of the particle at sorted position k, as double (src/subhalo.cpp:298-299 accumulates in double)
  const uint64_t *key;
  const int *val;
  const ProfSub *subs;
  const float4 *pos;
  const int *ids;
  __device__ double operator()(int64_t k) const
  {
    const ProfSub &sb = subs[(int)(key[k] >> 32)];
    const int64_t e = sb.part_off + val[k];
    return (double)pos[ids ? ids[e] : e].w;
  }
};
struct PlusD
{
  __device__ double operator()(double a, double b) const { return a + b; }
};

// K3: per sorted element: v^2 = M(<r)/max(r, eps) (src/subhalo.cpp:302-306), argmax, overdensity test (src/snapshot.cpp:272-280)
__global__ void __launch_bounds__(kPB) prof_select_kernel(const ProfSub *__restrict__ subs, int64_t B, const uint64_t *__restrict__ key,
                                                           const double *__restrict__ mcum, float softening, float rho_virial,
                                                           ProfScratch *__restrict__ scr)
{
  const int64_t k = (int64_t)blockIdx.x * kPB + threadIdx.x;
  if (k >= B) return;
  const uint64_t kk = key[k];
  const int s = (int)(kk >> 32);
  float r = __uint_as_float((uint32_t)kk);
  if (r < softening) r = softening;
  const float m = (float)mcum[k];
  const float v = __fdiv_rn(m, r);
  const uint32_t i = (uint32_t)(k - subs[s].bound_off);
  const unsigned long long packed = ((unsigned long long)__float_as_uint(v) << 32) | (0xffffffffu - i);
  // warp-aggregate when the warp is inside one subhalo
  const int s0 = __shfl_sync(0xffffffffu, s, 0);
  const bool so = m > __fmul_rn(__fmul_rn(__fmul_rn(rho_virial, r), r), r);
  if (__all_sync(0xffffffffu, s == s0))
  {
    unsigned long long best = packed;
    for (int o = 16; o > 0; o >>= 1)
    {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    int last = so ? (int)i + 1 : 0;
    last = __reduce_max_sync(0xffffffffu, last);
    if ((threadIdx.x & 31) == 0)
    {
      atomicMax(&scr[s].argmax, best);
      if (last > 0) atomicMax(&scr[s].so_last, last);
    }
  }
  else
  {
    atomicMax(&scr[s].argmax, packed);
    if (so) atomicMax(&scr[s].so_last, (int)i + 1);
  }
}

// K4: one thread per subhalo (src/subhalo.cpp:308-326,389-392)
__global__ void prof_finalize_kernel(const ProfSub *__restrict__ subs, int nsub, const uint64_t *__restrict__ key,
                                     const double *__restrict__ mcum, const ProfScratch *__restrict__ scr, float softening,
                                     float velocity_unit, float rho_virial, int snapshot_index, hbtu_profile_io *__restrict__ io)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsub) return;
  hbtu_profile_io &o = io[s];
  const ProfSub sb = subs[s];
  if (sb.nb == 0)
  { // Nbound <= 1 (src/subhalo.cpp:265-286,336-349): the Vmax record is left alone
    o.rmax_comoving = o.vmax_physical = o.r2sigma_comoving = o.rhalf_comoving = 0.f;
    o.bound_r200crit_comoving = o.bound_m200crit = 0.f;
    for (int j = 0; j < 6; j++) o.inertial_tensor[j] = o.inertial_tensor_weighted[j] = 0.f;
    return;
  }
  auto radius = [&](int64_t i) {
    float r = __uint_as_float((uint32_t)key[sb.bound_off + i]);
    return r < softening ? softening : r;
  };
  const ProfScratch sc = scr[s];
  const int imax = (int)(0xffffffffu - (uint32_t)sc.argmax);
  const float vmax2 = __uint_as_float((uint32_t)(sc.argmax >> 32));
  o.rmax_comoving = radius(imax);
  o.vmax_physical = __fsqrt_rn(__fmul_rn(vmax2, velocity_unit));
  o.rhalf_comoving = radius(sb.nb / 2);
  o.r2sigma_comoving = radius((int)((double)sb.nb * 0.955));
  if (sc.so_last > 0)
  {
    const float m = (float)mcum[sb.bound_off + sc.so_last - 1];
    o.bound_m200crit = m;
    o.bound_r200crit_comoving = (float)pow((double)__fdiv_rn(m, rho_virial), 1.0 / 3);
  }
  if (o.vmax_physical >= o.last_max_vmax_physical)
  {
    o.snapshot_index_of_last_max_vmax = snapshot_index;
    o.last_max_vmax_physical = o.vmax_physical;
  }
  const int map[6] = {0, 1, 2, 3, 4, 5};
  for (int j = 0; j < 6; j++)
  {
    o.inertial_tensor[j] = __fdiv_rn((float)sc.sums[map[j]], sb.mbound);
    o.inertial_tensor_weighted[j] = __fdiv_rn((float)sc.sums[6 + map[j]], sb.mbound);
  }
}

// common core.  Particle list of subhalo s = `list_off[s]` + i, read through `d_ids` when given (indices into d_pos),
// else directly.  host_pos != nullptr: the positions are uploaded first (N_host particles).
static void profile_core(Context &c, const hbtu_epoch *epoch, int64_t nsub, const int64_t *list_off, const int64_t *list_len, const float *host_pos,
                         int64_t N_host, const float4 *resident_pos, const int *d_ids, hbtu_profile_io *io)
{
  std::vector<ProfSub> subs(nsub);
  int64_t B = 0;
  for (int64_t s = 0; s < nsub; s++)
  {
    if (list_len[s] < 0 || io[s].nbound < 0 || io[s].nbound > list_len[s])
      throw CudaError{HBTU_ERR_INVALID, "nbound exceeds the particle list of a subhalo"};
    if (io[s].nbound > 0x7fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "subhalo larger than 2^31 particles"};
    ProfSub &sb = subs[s];
    sb.part_off = list_off[s];
    sb.bound_off = B;
    sb.nb = io[s].nbound > 1 ? (int)io[s].nbound : 0;
    sb.cx = (float)io[s].mostbound_pos[0];
    sb.cy = (float)io[s].mostbound_pos[1];
    sb.cz = (float)io[s].mostbound_pos[2];
    sb.mbound = io[s].mbound;
    B += sb.nb;
  }
  DevConfig cfg = c.cfg;
  cfg.scale_factor = (float)epoch->scale_factor;
  cfg.hz = (float)epoch->hz;
  cfg.snapshot_index = epoch->snapshot_index;
  // HBTReal VelocityUnit = G/ScaleFactor (src/subhalo.cpp:287); RhoVirial = 200*Hz*Hz/2.0/G*a*a*a (src/snapshot.cpp:272)
  const float velocity_unit = cfg.G / cfg.scale_factor;
  const float rho_virial = (float)((double)(200.f * cfg.hz * cfg.hz) / 2.0 / (double)cfg.G * (double)cfg.scale_factor * (double)cfg.scale_factor *
                                   (double)cfg.scale_factor);
  cudaStream_t st = c.stream;
  Arena &ar = c.arena;
  ar.reset();
  ar.reserve(N_host * 16 + B * 56 + nsub * (int64_t)(sizeof(ProfSub) + sizeof(ProfScratch) + sizeof(hbtu_profile_io)) + (64 << 20));
  c.ls.launches = 0;
  const float4 *d_pos = resident_pos;
  if (host_pos)
  {
    float4 *up = ar.alloc<float4>(N_host);
    if (N_host > 0) HBT_CUDA(cudaMemcpyAsync(up, host_pos, sizeof(float4) * (size_t)N_host, cudaMemcpyHostToDevice, st));
    d_pos = up;
  }
  ProfSub *d_subs = ar.alloc<ProfSub>(nsub);
  HBT_CUDA(cudaMemcpyAsync(d_subs, subs.data(), sizeof(ProfSub) * (size_t)nsub, cudaMemcpyHostToDevice, st));
  hbtu_profile_io *d_io = ar.alloc<hbtu_profile_io>(nsub);
  HBT_CUDA(cudaMemcpyAsync(d_io, io, sizeof(hbtu_profile_io) * (size_t)nsub, cudaMemcpyHostToDevice, st));
  ProfScratch *d_scr = ar.alloc<ProfScratch>(nsub);
  HBT_CUDA(cudaMemsetAsync(d_scr, 0, sizeof(ProfScratch) * (size_t)nsub, st));
  uint64_t *key_a = ar.alloc<uint64_t>(B), *key_b = ar.alloc<uint64_t>(B);
  int *val_a = ar.alloc<int>(B), *val_b = ar.alloc<int>(B);
  double *mcum = ar.alloc<double>(B);
  if (B > 0x7fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "more than 2^31 bound particles in one batch"};
  const uint64_t *skey = key_a;
  HBT_CUDA(cudaEventRecord(c.ev_exec[0], st)); // kernels only: the H2D copies above are queued before it
  if (B > 0)
  {
    static_assert(kPB == kSegBlock, "seg_reduce.cuh blocks");
    std::vector<int> chunk_off, bound_off32((size_t)nsub + 1), tile_off;
    const int nchunk = seg_chunk_table((int)nsub, [&](int a) { return subs[a].nb; }, chunk_off);
    for (int64_t a = 0; a < nsub; a++) bound_off32[a] = (int)subs[a].bound_off;
    bound_off32[nsub] = (int)B;
    const int ntiles = scan_tile_table(bound_off32.data(), (int)nsub, tile_off);
    int *d_chunk_off = ar.alloc<int>(nsub + 1), *d_bound_off = ar.alloc<int>(nsub + 1), *d_tile_off = ar.alloc<int>(nsub + 1);
    HBT_CUDA(cudaMemcpyAsync(d_chunk_off, chunk_off.data(), sizeof(int) * (size_t)(nsub + 1), cudaMemcpyHostToDevice, st));
    HBT_CUDA(cudaMemcpyAsync(d_bound_off, bound_off32.data(), sizeof(int) * (size_t)(nsub + 1), cudaMemcpyHostToDevice, st));
    HBT_CUDA(cudaMemcpyAsync(d_tile_off, tile_off.data(), sizeof(int) * (size_t)(nsub + 1), cudaMemcpyHostToDevice, st));
    double *partial = ar.alloc<double>((int64_t)nchunk * 12);
    prof_radius_shape_kernel<<<nchunk, kPB, 0, st>>>(d_subs, (int)nsub, d_chunk_off, d_pos, d_ids, cfg, key_a, val_a, partial);
    HBT_CHECK_LAUNCH();
    prof_finish_kernel<<<(unsigned)nsub, kPB, 0, st>>>((int)nsub, d_chunk_off, partial, d_scr);
    HBT_CHECK_LAUNCH();
    int bits = 32;
    while ((1ll << (bits - 32)) < nsub) bits++;
    cub::DoubleBuffer<uint64_t> dk(key_a, key_b);
    cub::DoubleBuffer<int> dv(val_a, val_b);
    size_t tb = 0;
    HBT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, B, 0, bits, st));
    void *tmp = ar.alloc<char>((int64_t)tb);
    HBT_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, B, 0, bits, st));
    skey = dk.Current();
    // cumulative mass in double with a fixed combination tree aligned to each subhalo (det_scan.cuh; exact anyway for equal
    // masses).  The sorted array keeps every subhalo in its own range [bound_off, bound_off + nb).
    det_inclusive_scan_segments<double, PlusD>(ar, st, d_bound_off, d_tile_off, (int)nsub, ntiles, SortedMass{skey, dv.Current(), d_subs, d_pos, d_ids},
                                               0.0, mcum, c.ls.launches);
    HBT_CUDA(cudaStreamSynchronize(st)); // the tables above are host temporaries
    prof_select_kernel<<<pgrid(B), kPB, 0, st>>>(d_subs, B, skey, mcum, cfg.softening, rho_virial, d_scr);
    HBT_CHECK_LAUNCH();
    c.ls.launches += 3 + 1 + (bits + 7) / 8;
  }
  prof_finalize_kernel<<<pgrid(nsub), kPB, 0, st>>>(d_subs, (int)nsub, skey, mcum, d_scr, cfg.softening, velocity_unit, rho_virial,
                                                    cfg.snapshot_index, d_io);
  HBT_CHECK_LAUNCH();
  c.ls.launches++;
  HBT_CUDA(cudaEventRecord(c.ev_exec[1], st));
  HBT_CUDA(cudaMemcpyAsync(io, d_io, sizeof(hbtu_profile_io) * (size_t)nsub, cudaMemcpyDeviceToHost, st));
  HBT_CUDA(cudaStreamSynchronize(st));
  hbtu_stats stats;
  std::memset(&stats, 0, sizeof(stats));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c.ev_exec[0], c.ev_exec[1]);
  stats.execute_ms = ms; // radius/shape + sort + scan + select + finalize
  stats.other_ms = ms;
  stats.walk_targets = B;
  stats.kernel_launches = c.ls.launches;
  stats.h2d_bytes = N_host * 16 + nsub * (int64_t)(sizeof(ProfSub) + sizeof(hbtu_profile_io));
  stats.d2h_bytes = nsub * (int64_t)sizeof(hbtu_profile_io);
  c.stats = stats;
}

void profile_batch(Context &c, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass, hbtu_profile_io *io)
{
  if (!epoch || nsub < 0 || !part_offset || !io) throw CudaError{HBTU_ERR_INVALID, "bad argument"};
  if (nsub == 0) return;
  if (nsub > 0x7fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "too many subhaloes"};
  const int64_t N = part_offset[nsub];
  if (N > 0 && !pos_mass) throw CudaError{HBTU_ERR_INVALID, "null particle array"};
  std::vector<int64_t> len(nsub);
  for (int64_t s = 0; s < nsub; s++) len[s] = part_offset[s + 1] - part_offset[s];
  c.staged = c.executed = false; // the arena is shared with a staged batch's rounds
  profile_core(c, epoch, nsub, part_offset, len.data(), pos_mass, N, nullptr, nullptr, io);
}

// The same on the batch that hbtu_execute left in HBM: the particle lists are the Elists (new order, bound first) of the
// subhaloes, read through d_ids from the resident positions - nothing but the per-subhalo records crosses PCIe.
void profile_executed(Context &c, hbtu_profile_io *io)
{
  if (!io) throw CudaError{HBTU_ERR_INVALID, "bad argument"};
  if (!c.staged || !c.executed) throw CudaError{HBTU_ERR_INVALID, "hbtu_profile_executed needs an executed batch (hbtu_stage + hbtu_execute)"};
  if (c.pipelined)
    throw CudaError{HBTU_ERR_INVALID, "the last hbtu_unbind_batch ran in pipelined parts, only the last of which is resident: use hbtu_profile_batch"};
  const int64_t nsub = c.nsub;
  if (nsub == 0) return;
  std::vector<int64_t> off(nsub), len(nsub);
  for (int64_t s = 0; s < nsub; s++)
  {
    off[s] = c.subs[s].slot_base;
    len[s] = c.subs[s].n_src; // Particles.size() after unbinding, before truncation: the bound part is its head
  }
  hbtu_epoch e;
  e.scale_factor = c.cfg.scale_factor;
  e.hz = c.cfg.hz;
  e.snapshot_index = c.cfg.snapshot_index;
  e.reserved = 0;
  profile_core(c, &e, nsub, off.data(), len.data(), nullptr, 0, c.d_pos, c.d_ids, io);
}

} // namespace hbt
