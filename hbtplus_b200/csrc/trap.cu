// trap.cu - merger trap detection on the device, sm_100a.
//
// The detection part of SubhaloSnapshot_t::MergeSubhalos (src/subhalo_merge.cpp:187-199):
//   SubHelper_t::BuildPosition / BuildVelocity (:29-123)  mass-weighted mean and dispersion of the <= 20 most bound
//                                                          particles of every subhalo, serial double sums (K1: one thread
//                                                          per subhalo, the same operation order, no FMA contraction)
//   FillHostTrackIds + DetectTraps + SinkDistance (:125-172)  every untrapped subhalo walks up its host chain and sinks into
//                                                          the first host with Nbound > 1 and d/sigma_R + v/sigma_V < 2
//                                                          (K2: one thread per subhalo; hosts are only read)
// Negligible work (20 particles per subhalo); it is here so that the merger step can stay on the device between the
// unbinding of a snapshot and the re-unbinding of the hosts it flags (IsMerged).
#include <cstring>
#include <vector>

#include "context.cuh"

namespace hbt
{

static constexpr int kCoreMax = 20; // NumPartCoreMax, src/subhalo_merge.cpp:11
static constexpr int kTB = 128;

struct TrapSub
{
  int64_t nbound;
  int host;          // Helpers[i].HostTrackId
  int ncore;         // particles stored for this subhalo (<= kCoreMax)
  float mb_pos[3], mb_vel[3];
  int64_t sink;
  int sink_snap;
};
struct TrapHelper
{
  float pos[3], vel[3], sigma_r, sigma_v;
};

__global__ void __launch_bounds__(kTB) trap_helpers_kernel(const TrapSub *__restrict__ subs, int nsub, const float4 *__restrict__ core_pm,
                                                            const float4 *__restrict__ core_vel, DevConfig cfg, TrapHelper *__restrict__ help)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsub) return;
  const TrapSub sb = subs[s];
  TrapHelper h;
  h.sigma_r = h.sigma_v = 0.f;
  for (int j = 0; j < 3; j++) h.pos[j] = h.vel[j] = 0.f;
  const float4 *pm = core_pm + (int64_t)s * kCoreMax, *pv = core_vel + (int64_t)s * kCoreMax;
  if (sb.nbound == 1)
  {
    h.pos[0] = pm[0].x; h.pos[1] = pm[0].y; h.pos[2] = pm[0].z;
    h.vel[0] = pv[0].x; h.vel[1] = pv[0].y; h.vel[2] = pv[0].z;
  }
  else if (sb.nbound > 1)
  {
    double sx[3] = {0, 0, 0}, sx2[3] = {0, 0, 0}, sv[3] = {0, 0, 0}, sv2[3] = {0, 0, 0}, origin[3] = {0, 0, 0}, msum = 0.;
    if (cfg.periodic) { origin[0] = pm[0].x; origin[1] = pm[0].y; origin[2] = pm[0].z; }
    for (int i = 0; i < sb.ncore; i++)
    {
      const float4 x = pm[i], u = pv[i];
      const double m = (double)x.w;
      msum = __dadd_rn(msum, m);
      const float xs[3] = {x.x, x.y, x.z}, us[3] = {u.x, u.y, u.z};
      for (int j = 0; j < 3; j++)
      {
        double dx = cfg.periodic ? nearest_d(__dsub_rn((double)xs[j], origin[j]), (double)cfg.box_size, (double)cfg.box_half) : (double)xs[j];
        sx[j] = __dadd_rn(sx[j], __dmul_rn(dx, m));
        sx2[j] = __dadd_rn(sx2[j], __dmul_rn(__dmul_rn(dx, dx), m));
        const double dv = (double)us[j];
        sv[j] = __dadd_rn(sv[j], __dmul_rn(dv, m));
        sv2[j] = __dadd_rn(sv2[j], __dmul_rn(__dmul_rn(dv, dv), m));
      }
    }
    for (int j = 0; j < 3; j++)
    {
      sx[j] = __ddiv_rn(sx[j], msum); sx2[j] = __ddiv_rn(sx2[j], msum);
      float p = (float)sx[j];
      if (cfg.periodic) p = (float)__dadd_rn((double)p, origin[j]); // HBTReal += double (:75)
      h.pos[j] = p;
      sx2[j] = __dsub_rn(sx2[j], __dmul_rn(sx[j], sx[j]));
      sv[j] = __ddiv_rn(sv[j], msum); sv2[j] = __ddiv_rn(sv2[j], msum);
      h.vel[j] = (float)sv[j];
      sv2[j] = __dsub_rn(sv2[j], __dmul_rn(sv[j], sv[j]));
    }
    h.sigma_r = (float)sqrt(__dadd_rn(__dadd_rn(sx2[0], sx2[1]), sx2[2]));
    h.sigma_v = (float)sqrt(__dadd_rn(__dadd_rn(sv2[0], sv2[1]), sv2[2]));
  }
  help[s] = h;
}

__global__ void __launch_bounds__(kTB) trap_detect_kernel(TrapSub *__restrict__ subs, int nsub, const TrapHelper *__restrict__ help, DevConfig cfg,
                                                           int *__restrict__ merged)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsub) return;
  TrapSub &me = subs[i];
  if (me.sink != -1) return; // IsTrapped (:142)
  int host = me.host;
  while (host >= 0)
  {
    if (subs[host].nbound > 1) // avoid orphans or nulls as hosts (:146); nbound is never written here
    {
      const TrapHelper c = help[host];
      float dx[3], dv[3];
      for (int j = 0; j < 3; j++)
      {
        dx[j] = __fsub_rn(c.pos[j], me.mb_pos[j]);
        if (cfg.periodic) dx[j] = nearest_f(dx[j], cfg.box_size, cfg.box_half);
        dv[j] = __fsub_rn(c.vel[j], me.mb_vel[j]);
      }
      const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx[0], dx[0]), __fmul_rn(dx[1], dx[1])), __fmul_rn(dx[2], dx[2])));
      const float v = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dv[0], dv[0]), __fmul_rn(dv[1], dv[1])), __fmul_rn(dv[2], dv[2])));
      const float delta = __fadd_rn(__fdiv_rn(d, c.sigma_r), __fdiv_rn(v, c.sigma_v)); // SinkDistance (:125-130)
      if ((double)delta < 2.)
      {
        me.sink = host;
        me.sink_snap = cfg.snapshot_index;
        if (me.nbound > 1) merged[host] = 1; // IsMerged (:152-153): benign same-value race, as in the reference
        break;
      }
    }
    host = subs[host].host;
  }
}

void detect_traps(Context &c, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass, const float *vel,
                  const int64_t *nest_offset, const int32_t *nest_list, hbtu_trap_io *io)
{
  if (!epoch || nsub < 0 || !part_offset || !io) throw CudaError{HBTU_ERR_INVALID, "bad argument"};
  if (nsub == 0) return;
  if (nsub > 0x7fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "too many subhaloes"};
  std::vector<TrapSub> subs(nsub);
  std::vector<float4> core_pm((size_t)nsub * kCoreMax, make_float4(0, 0, 0, 0)), core_vel((size_t)nsub * kCoreMax, make_float4(0, 0, 0, 0));
  for (int64_t s = 0; s < nsub; s++)
  {
    const int64_t n = part_offset[s + 1] - part_offset[s];
    if (n < 0 || io[s].nbound < 0 || (io[s].nbound > kCoreMax ? kCoreMax : io[s].nbound) > n)
      throw CudaError{HBTU_ERR_INVALID, "a particle list is shorter than min(nbound, 20)"};
    TrapSub &sb = subs[s];
    sb.nbound = io[s].nbound;
    sb.host = -1;
    sb.ncore = (int)(io[s].nbound > kCoreMax ? kCoreMax : io[s].nbound);
    if (sb.ncore > 0 && (!pos_mass || !vel)) throw CudaError{HBTU_ERR_INVALID, "null particle arrays"};
    for (int j = 0; j < 3; j++)
    {
      sb.mb_pos[j] = (float)io[s].mostbound_pos[j];
      sb.mb_vel[j] = (float)io[s].mostbound_vel[j];
    }
    sb.sink = io[s].sink_track_id;
    sb.sink_snap = io[s].snapshot_index_of_sink;
    for (int i = 0; i < sb.ncore; i++)
    {
      const float *x = &pos_mass[4 * (part_offset[s] + i)], *u = &vel[4 * (part_offset[s] + i)];
      core_pm[(size_t)s * kCoreMax + i] = make_float4(x[0], x[1], x[2], x[3]);
      core_vel[(size_t)s * kCoreMax + i] = make_float4(u[0], u[1], u[2], 0.f);
    }
  }
  if (nest_offset)
    for (int64_t s = 0; s < nsub; s++)
      for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++)
      {
        const int32_t ch = nest_list[k];
        if (ch < 0 || ch >= nsub || ch == s) throw CudaError{HBTU_ERR_INVALID, "malformed nest forest"};
        subs[ch].host = (int)s; // FillHostTrackIds (:163-172)
      }
  c.staged = c.executed = false; // the arena is shared with a staged batch's rounds
  DevConfig cfg = c.cfg;
  cfg.scale_factor = (float)epoch->scale_factor;
  cfg.hz = (float)epoch->hz;
  cfg.snapshot_index = epoch->snapshot_index;
  cudaStream_t st = c.stream;
  Arena &ar = c.arena;
  ar.reset();
  ar.reserve(nsub * (int64_t)(sizeof(TrapSub) + sizeof(TrapHelper) + 2 * kCoreMax * sizeof(float4) + 8) + (1 << 20));
  TrapSub *d_subs = ar.alloc<TrapSub>(nsub);
  TrapHelper *d_help = ar.alloc<TrapHelper>(nsub);
  float4 *d_pm = ar.alloc<float4>(nsub * kCoreMax), *d_vel = ar.alloc<float4>(nsub * kCoreMax);
  int *d_merged = ar.alloc<int>(nsub);
  HBT_CUDA(cudaMemcpyAsync(d_subs, subs.data(), sizeof(TrapSub) * (size_t)nsub, cudaMemcpyHostToDevice, st));
  HBT_CUDA(cudaMemcpyAsync(d_pm, core_pm.data(), sizeof(float4) * core_pm.size(), cudaMemcpyHostToDevice, st));
  HBT_CUDA(cudaMemcpyAsync(d_vel, core_vel.data(), sizeof(float4) * core_vel.size(), cudaMemcpyHostToDevice, st));
  HBT_CUDA(cudaMemsetAsync(d_merged, 0, sizeof(int) * (size_t)nsub, st));
  HBT_CUDA(cudaEventRecord(c.ev_exec[0], st));
  trap_helpers_kernel<<<div_up(nsub, kTB), kTB, 0, st>>>(d_subs, (int)nsub, d_pm, d_vel, cfg, d_help);
  HBT_CHECK_LAUNCH();
  trap_detect_kernel<<<div_up(nsub, kTB), kTB, 0, st>>>(d_subs, (int)nsub, d_help, cfg, d_merged);
  HBT_CHECK_LAUNCH();
  HBT_CUDA(cudaEventRecord(c.ev_exec[1], st));
  std::vector<int> merged(nsub);
  HBT_CUDA(cudaMemcpyAsync(subs.data(), d_subs, sizeof(TrapSub) * (size_t)nsub, cudaMemcpyDeviceToHost, st));
  HBT_CUDA(cudaMemcpyAsync(merged.data(), d_merged, sizeof(int) * (size_t)nsub, cudaMemcpyDeviceToHost, st));
  HBT_CUDA(cudaStreamSynchronize(st));
  for (int64_t s = 0; s < nsub; s++)
  {
    io[s].sink_track_id = subs[s].sink;
    io[s].snapshot_index_of_sink = subs[s].sink_snap;
    io[s].is_merged = merged[s];
  }
  std::memset(&c.stats, 0, sizeof(c.stats));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c.ev_exec[0], c.ev_exec[1]);
  c.stats.execute_ms = c.stats.other_ms = ms;
  c.stats.kernel_launches = 2;
}

} // namespace hbt
