// capi.cu - the extern "C" entry points of include/hbt_unbind.h.
//
// Validation, device/stream ownership, host<->device staging and error translation only; the
// algorithm is in unbind_batch.cu / tree_build.cu / walk.cu.  No entry point has a CPU path: if no
// sm_100 device is usable hbtu_create fails with HBTU_ERR_NODEVICE.
#include <cub/cub.cuh>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>

#include "context.cuh"

using namespace hbt;

namespace
{
thread_local std::string g_create_error;

struct WallTimer
{ // host wall clock of an entry point, added to a hbtu_stats field when it goes out of scope
  double &slot;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit WallTimer(double &s) : slot(s) {}
  ~WallTimer() { slot = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

template <class T>
void grow(T *&p, int64_t &cap, int64_t need)
{
  if (need <= cap) return;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  int64_t ncap = need + need / 8 + 1024;
  HBT_CUDA(cudaMalloc(&p, sizeof(T) * (size_t)ncap));
  cap = ncap;
}

template <class F>
int guarded(hbtu_ctx *ctx, F &&f)
{
  if (!ctx) return HBTU_ERR_INVALID;
  try
  {
    cudaError_t e = cudaSetDevice(ctx->c.device);
    if (e != cudaSuccess) throw CudaError{HBTU_ERR_NODEVICE, std::string("cudaSetDevice: ") + cudaGetErrorString(e)};
    f(ctx->c);
    return HBTU_OK;
  }
  catch (const CudaError &e)
  {
    ctx->c.last_error = e.what;
    cudaGetLastError();
    return e.code;
  }
  catch (const std::bad_alloc &)
  {
    ctx->c.last_error = "host allocation failed";
    return HBTU_ERR_NOMEM;
  }
  catch (const std::exception &e)
  {
    ctx->c.last_error = e.what();
    return HBTU_ERR_INVALID;
  }
}

void build_forest(Context &c, int64_t nsub, const int64_t *part_offset, const int64_t *nest_offset, const int32_t *nest_list)
{
  c.subs.assign((size_t)nsub, SubHost());
  for (int64_t s = 0; s < nsub; s++)
  {
    int64_t n = part_offset[s + 1] - part_offset[s];
    if (n < 0 || n > 0x3fffffff) throw CudaError{HBTU_ERR_INVALID, "part_offset must be non-decreasing with < 2^30 particles per subhalo"};
    c.subs[s].part_begin = part_offset[s];
    c.subs[s].n_own = (int)n;
  }
  if (nest_offset)
    for (int64_t s = 0; s < nsub; s++)
      for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++)
      {
        int ch = nest_list[k];
        if (ch < 0 || ch >= nsub || ch == s || c.subs[ch].parent != -1)
          throw CudaError{HBTU_ERR_INVALID, "nest lists must form a forest of batch-local subhalo indices"};
        c.subs[ch].parent = (int)s;
        c.subs[s].children.push_back(ch);
      }
  // depth, with cycle detection (a cycle has no root, so its members never get a depth)
  c.max_depth = 0;
  std::vector<int> order;
  order.reserve((size_t)nsub);
  for (int64_t s = 0; s < nsub; s++)
    if (c.subs[s].parent < 0) order.push_back((int)s);
  for (size_t i = 0; i < order.size(); i++)
  {
    SubHost &h = c.subs[order[i]];
    for (int ch : h.children)
    {
      c.subs[ch].depth = h.depth + 1;
      if (c.subs[ch].depth > c.max_depth) c.max_depth = c.subs[ch].depth;
      order.push_back(ch);
    }
  }
  if ((int64_t)order.size() != nsub) throw CudaError{HBTU_ERR_INVALID, "nest lists contain a cycle"};
  // capacity: own particles plus everything descendants can feed upwards
  for (size_t i = order.size(); i-- > 0;)
  {
    SubHost &h = c.subs[order[i]];
    h.cap += h.n_own;
    if (h.parent >= 0) c.subs[h.parent].cap += h.cap;
  }
  int64_t slot = 0;
  c.levels.assign((size_t)c.max_depth + 1, std::vector<int>());
  for (int64_t s = 0; s < nsub; s++)
  {
    c.subs[s].slot_base = slot;
    slot += c.subs[s].cap;
    c.levels[c.subs[s].depth].push_back((int)s);
  }
  c.total_cap = slot;
  if (slot > 0x7fffffff00ll) throw CudaError{HBTU_ERR_UNSUPPORTED, "batch too large"};
}

// overlap == false: hbtu_stage, returns when everything is in HBM.  overlap == true (hbtu_unbind_batch): the particle arrays
// travel on the copy stream in two waves - first everything except the dominant root subhalo (the source that holds more
// than a quarter of the batch: an AqA2 central), then that root, which the level-synchronous scheduler reaches last - and
// execute_batch waits for each wave only where it is first needed, so the big upload hides behind the deeper levels' rounds.
void stage(Context &c, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass, const float *vel,
           const int64_t *nest_offset, const int32_t *nest_list, const hbtu_sub_io *io, int32_t flags, bool overlap = false,
           bool preloaded = false)
{ // preloaded: the particle arrays of this batch are already in d_pos_next / d_vel_next (a part of a pipelined hbtu_unbind_batch)
  if (!epoch || nsub < 0 || !part_offset || (nsub > 0 && !io)) throw CudaError{HBTU_ERR_INVALID, "null argument"};
  if (nsub > 0x7ffffff0) throw CudaError{HBTU_ERR_UNSUPPORTED, "too many subhaloes"};
  finish_upload(c); // a previous asynchronous staging that was never executed
  c.staged = c.executed = false;
  const int64_t N = nsub > 0 ? part_offset[nsub] - part_offset[0] : 0;
  if (part_offset[0] != 0) throw CudaError{HBTU_ERR_INVALID, "part_offset[0] must be 0"};
  if (N > 0x7ffffff0) throw CudaError{HBTU_ERR_UNSUPPORTED, "more than 2^31 particles in one batch: split the batch"};
  if (N > 0 && (!pos_mass || !vel)) throw CudaError{HBTU_ERR_INVALID, "null particle arrays"};
  build_forest(c, nsub, part_offset, nest_offset, nest_list);
  for (int64_t s = 0; s < nsub; s++)
    if ((io[s].flags & HBTU_SUB_PLAIN_UNBIND) && (c.subs[s].parent >= 0 || !c.subs[s].children.empty()))
      throw CudaError{HBTU_ERR_INVALID, "HBTU_SUB_PLAIN_UNBIND is only valid for a subhalo without parent and without nested subhaloes"};
  c.nsub = nsub;
  c.N = N;
  c.flags = flags;
  c.cfg.no_stripping = (flags & HBTU_FLAG_NO_STRIPPING) != 0;
  c.cfg.thermal_energy = (flags & HBTU_FLAG_THERMAL_ENERGY) != 0;
  c.cfg.scale_factor = (float)epoch->scale_factor;
  c.cfg.hz = (float)epoch->hz;
  c.cfg.snapshot_index = epoch->snapshot_index;
  c.io_in.assign(io, io + nsub);

  cudaStream_t st = c.stream;
  c.pipelined = false;
  c.order_index_base = 0;
  c.sub_index_base = 0;
  if (preloaded)
  {
    if (c.cap_pos_next < N || c.cap_vel_next < N) throw CudaError{HBTU_ERR_INVALID, "internal: preloaded part larger than its buffers"};
    std::swap(c.d_pos, c.d_pos_next);
    std::swap(c.d_vel, c.d_vel_next);
    std::swap(c.cap_particles, c.cap_pos_next);
    std::swap(c.cap_vel, c.cap_vel_next);
  }
  else
  {
    grow(c.d_pos, c.cap_particles, N);
    grow(c.d_vel, c.cap_vel, N);
  }
  {
    int64_t cap0 = c.cap_slots, cap1 = c.cap_slots, cap2 = c.cap_slots;
    grow(c.d_ids, cap0, c.total_cap);
    grow(c.d_ids_orig, cap1, c.total_cap);
    grow(c.d_E, cap2, c.total_cap);
    c.cap_slots = cap0;
    if (c.cfg.periodic) grow(c.d_rho, c.cap_rho, c.total_cap);
  }
  {
    int64_t cap0 = c.cap_subs, cap1 = c.cap_subs, cap2 = c.cap_subs;
    grow(c.d_subs, cap0, nsub + 1);
    grow(c.d_part_offset, cap1, nsub + 1);
    grow(c.d_slot_base, cap2, nsub + 1);
    c.cap_subs = cap0;
  }
  std::vector<int64_t> sb((size_t)nsub + 1, 0);
  for (int64_t s = 0; s < nsub; s++) sb[s] = c.subs[s].slot_base;
  sb[nsub] = c.total_cap;
  HBT_CUDA(cudaMemcpyAsync(c.d_part_offset, part_offset, sizeof(int64_t) * (size_t)(nsub + 1), cudaMemcpyHostToDevice, st));
  HBT_CUDA(cudaMemcpyAsync(c.d_slot_base, sb.data(), sizeof(int64_t) * (size_t)(nsub + 1), cudaMemcpyHostToDevice, st));
  HBT_CUDA(cudaStreamSynchronize(st)); // `sb` is a local; also orders the (re)allocations above before the copy stream's work
  auto copy_range = [&](cudaStream_t cs, int64_t b, int64_t e) {
    if (e <= b) return;
    HBT_CUDA(cudaMemcpyAsync(c.d_pos + b, pos_mass + 4 * b, sizeof(float4) * (size_t)(e - b), cudaMemcpyHostToDevice, cs));
    HBT_CUDA(cudaMemcpyAsync(c.d_vel + b, vel + 4 * b, sizeof(float4) * (size_t)(e - b), cudaMemcpyHostToDevice, cs));
  };
  c.stats.h2d_bytes = N * 32 + (nsub + 1) * 16;
  c.waves_pending = false;
  c.staged_async = overlap && !preloaded;
  if (preloaded)
    c.stats.h2d_ms = 0; // the pipelined wrapper accounts for the upload
  else if (!overlap)
  {
    HBT_CUDA(cudaEventRecord(c.ev[0], st));
    copy_range(st, 0, N);
    HBT_CUDA(cudaEventRecord(c.ev[1], st));
    HBT_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]);
    c.stats.h2d_ms = ms;
  }
  else
  {
    int64_t big = -1; // the dominant root, if any
    for (int64_t s = 0; s < nsub; s++)
      if (c.subs[s].parent < 0 && c.subs[s].n_own > N / 4 && (big < 0 || c.subs[s].n_own > c.subs[big].n_own)) big = s;
    if (c.max_depth == 0 || N < (1 << 20)) big = -1; // nothing to hide the upload behind
    // ranges of wave 1 (everything but the dominant root) and wave 2 (that root)
    std::vector<std::pair<int64_t, int64_t>> waves[2];
    if (big < 0)
      waves[0].push_back({0, N});
    else
    {
      waves[0].push_back({0, part_offset[big]});
      waves[0].push_back({part_offset[big + 1], N});
      waves[1].push_back({part_offset[big], part_offset[big + 1]});
    }
    // The copies are fed to the copy stream in chunks by a helper thread that waits for every chunk: with the whole upload
    // enqueued at once, nothing issued afterwards on ANY stream started before the last byte had crossed PCIe (measured: the
    // first kernel of the step waited the full 104 ms) - the copy engine's reads starve the command fetch; the gaps between
    // chunks let the kernels of the deeper levels through.
    const char *env = getenv("HBTU_UPLOAD_CHUNK_MB");
    const int64_t chunk = std::max<int64_t>(1, env ? atoll(env) : 32) * (int64_t)(1 << 20) / 32; // particles per chunk (32 B each)
    finish_upload(c);
    c.up_wave_done = 0;
    c.up_error.clear();
    c.waves_pending = true;
    const int device = c.device;
    cudaStream_t cs = c.copy_stream;
    float4 *d_pos = c.d_pos, *d_vel = c.d_vel;
    Context *cp = &c;
    c.uploader = std::thread([=]() {
      const auto t0 = std::chrono::steady_clock::now();
      std::string err;
      if (cudaSetDevice(device) != cudaSuccess) err = "cudaSetDevice failed in the upload helper";
      for (int w = 0; w < 2; w++)
      {
        for (const auto &r : waves[w])
          for (int64_t b = r.first; b < r.second && err.empty(); b += chunk)
          {
            const int64_t e = std::min(r.second, b + chunk);
            cudaError_t e1 = cudaMemcpyAsync(d_pos + b, pos_mass + 4 * b, sizeof(float4) * (size_t)(e - b), cudaMemcpyHostToDevice, cs);
            cudaError_t e2 = cudaMemcpyAsync(d_vel + b, vel + 4 * b, sizeof(float4) * (size_t)(e - b), cudaMemcpyHostToDevice, cs);
            cudaError_t e3 = cudaStreamSynchronize(cs);
            if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
              err = std::string("particle upload failed: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3));
          }
        std::lock_guard<std::mutex> lk(cp->up_m);
        if (!err.empty()) cp->up_error = err;
        cp->up_wave_done = w + 1;
        if (w == 1) cp->up_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        cp->up_cv.notify_all();
      }
    });
  }
  c.staged = true;
}

// ---- stand-alone GravityTree_t::Build + EvaluatePotential / BindingEnergy -------------------------
__global__ void fill_seg0_kernel(int *ts_seg, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ts_seg[i] = 0;
}
__global__ void target_keys_kernel(const float4 *tgt, int n, const SegRoot *roots, uint64_t *key, int *val)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = tgt[i];
  key[i] = morton_key(p.x, p.y, p.z, roots[0]);
  val[i] = i;
}
__global__ void target_gather_kernel(const int *perm, const float4 *tgt, const float *self_mass, const float4 *vel, int n,
                                     float4 *tgt_pm, float4 *tgt_vel)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int src = perm[i];
  float4 p = tgt[src];
  p.w = self_mass ? self_mass[src] : 0.f;
  tgt_pm[i] = p;
  if (vel) tgt_vel[i] = vel[src];
}
__global__ void scatter_out_kernel(const int *perm, const double *sorted_out, int n, double *out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[perm[i]] = sorted_out[i];
}

void tree_potential(Context &c, const hbtu_epoch *epoch, int64_t nsrc, const float *src_pos_mass, int64_t ntgt, const float *tgt_pos,
                    const float *tgt_self_mass, const float *tgt_vel, const double *ref_pos, const double *ref_vel, double *out)
{
  if (!epoch || nsrc < 1 || ntgt < 0 || !src_pos_mass || (ntgt > 0 && (!tgt_pos || !out)))
    throw CudaError{HBTU_ERR_INVALID, "bad argument"};
  if (tgt_vel && (!ref_pos || !ref_vel)) throw CudaError{HBTU_ERR_INVALID, "binding energy needs a reference frame"};
  if (nsrc > 0x3fffffff || ntgt > 0x3fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "too many particles"};
  c.staged = c.executed = false; // the arena is shared with a staged batch's rounds
  DevConfig cfg = c.cfg;
  cfg.scale_factor = (float)epoch->scale_factor;
  cfg.hz = (float)epoch->hz;
  cfg.snapshot_index = epoch->snapshot_index;
  cudaStream_t st = c.stream;
  const int S = (int)nsrc, T = (int)ntgt, kB = 256;
  c.ls.launches = 0;
  std::memset(&c.stats, 0, sizeof(c.stats));
  Arena &ar = c.arena;
  ar.reset();
  ar.reserve(tree_arena_bytes(S, 1) + (int64_t)T * 96 + (1 << 20));
  const WalkClass wcl = walk_class(T, S);
  const int tpl = wcl.targets_per_lane;
  std::vector<int> tree_off{0, S}, warp_off{0, (T + wcl.targets_per_warp - 1) / wcl.targets_per_warp};
  Segment sg{};
  sg.mode = tgt_vel ? kWalkBindingEnergy : kWalkPotential;
  sg.tree_n = S;
  sg.tgt_n = T;
  std::vector<Segment> segs{sg};
  Segment *d_segs = ar.alloc<Segment>(1);
  int *d_tree_off = ar.alloc<int>(2), *d_warp_off = ar.alloc<int>(2);
  HBT_CUDA(cudaMemcpyAsync(d_segs, segs.data(), sizeof(Segment), cudaMemcpyHostToDevice, st));
  HBT_CUDA(cudaMemcpyAsync(d_tree_off, tree_off.data(), 2 * sizeof(int), cudaMemcpyHostToDevice, st));
  HBT_CUDA(cudaMemcpyAsync(d_warp_off, warp_off.data(), 2 * sizeof(int), cudaMemcpyHostToDevice, st));
  TreeArrays tr;
  tr.S = S;
  tr.nseg = 1;
  tr.tree_off = d_tree_off;
  tr.h_tree_off = tree_off.data();
  tr.tpos = ar.alloc<float4>(S);
  tr.ts_seg = ar.alloc<int>(S);
  tr.bbox = ar.alloc<uint32_t>(6);
  HBT_CUDA(cudaMemcpyAsync(tr.tpos, src_pos_mass, sizeof(float4) * (size_t)S, cudaMemcpyHostToDevice, st));
  fill_seg0_kernel<<<div_up(S, kB), kB, 0, st>>>(tr.ts_seg, S);
  HBT_CHECK_LAUNCH();
  HBT_CUDA(cudaEventRecord(c.ev[0], st));
  launch_init_bbox(tr.bbox, 1, st, c.ls);
  launch_bbox(tr.tpos, tr.ts_seg, S, tr.bbox, st, c.ls);
  build_trees(tr, ar, cfg, st, c.ls);
  HBT_CUDA(cudaEventRecord(c.ev[1], st));
  if (T > 0)
  {
    float4 *d_tgt = ar.alloc<float4>(T), *d_vel = tgt_vel ? ar.alloc<float4>(T) : nullptr;
    float *d_sm = tgt_self_mass ? ar.alloc<float>(T) : nullptr;
    HBT_CUDA(cudaMemcpyAsync(d_tgt, tgt_pos, sizeof(float4) * (size_t)T, cudaMemcpyHostToDevice, st));
    if (tgt_vel) HBT_CUDA(cudaMemcpyAsync(d_vel, tgt_vel, sizeof(float4) * (size_t)T, cudaMemcpyHostToDevice, st));
    if (tgt_self_mass) HBT_CUDA(cudaMemcpyAsync(d_sm, tgt_self_mass, sizeof(float) * (size_t)T, cudaMemcpyHostToDevice, st));
    // walk the targets in key order (spatially coherent warps), scatter the results back
    uint64_t *ka = ar.alloc<uint64_t>(T), *kb = ar.alloc<uint64_t>(T);
    int *va = ar.alloc<int>(T), *vb = ar.alloc<int>(T);
    target_keys_kernel<<<div_up(T, kB), kB, 0, st>>>(d_tgt, T, tr.roots, ka, va);
    HBT_CHECK_LAUNCH();
    cub::DoubleBuffer<uint64_t> dk(ka, kb);
    cub::DoubleBuffer<int> dv(va, vb);
    size_t tb = 0;
    HBT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, T, 0, 63, st));
    void *tmp = ar.alloc<char>((int64_t)tb);
    HBT_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, T, 0, 63, st));
    float4 *tgt_pm = ar.alloc<float4>(T), *tgt_v = tgt_vel ? ar.alloc<float4>(T) : nullptr;
    target_gather_kernel<<<div_up(T, kB), kB, 0, st>>>(dv.Current(), d_tgt, d_sm, d_vel, T, tgt_pm, tgt_v);
    HBT_CHECK_LAUNCH();
    double *d_sorted = ar.alloc<double>(T), *d_out = ar.alloc<double>(T);
    c.ls.launches += 12;
    HBT_CUDA(cudaMemsetAsync(c.d_counters, 0, kWalkCounters * sizeof(unsigned long long), st));
    WalkArgs wa{};
    wa.node_xm = tr.node_xm;
    wa.node_aux = tr.node_aux;
    wa.cellcount = tr.cellcount;
    wa.tree_off = d_tree_off;
    wa.segs = d_segs;
    wa.warp_off = d_warp_off;
    wa.nseg = 1;
    wa.nwarps = warp_off[1];
    wa.targets_per_lane = tpl;
    wa.tgt_pm = tgt_pm;
    wa.vel = tgt_v;
    wa.out = d_sorted;
    wa.counters = c.count_interactions ? c.d_counters : nullptr;
    if (tgt_vel)
      for (int j = 0; j < 3; j++)
      {
        wa.ref_pos[j] = (float)ref_pos[j];
        wa.ref_vel[j] = (float)ref_vel[j];
      }
    HBT_CUDA(cudaEventRecord(c.ev[2], st));
    launch_walk(wa, cfg, st, c.ls);
    HBT_CUDA(cudaEventRecord(c.ev[3], st));
    scatter_out_kernel<<<div_up(T, kB), kB, 0, st>>>(dv.Current(), d_sorted, T, d_out);
    HBT_CHECK_LAUNCH();
    c.ls.launches++;
    HBT_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)T, cudaMemcpyDeviceToHost, st));
    HBT_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]);
    c.stats.walk_ms = ms;
    if (c.count_interactions)
    {
      unsigned long long cnt[kWalkCounters];
      HBT_CUDA(cudaMemcpy(cnt, c.d_counters, sizeof(cnt), cudaMemcpyDeviceToHost));
      c.stats.pair_interactions = (int64_t)cnt[0];
      c.stats.nodes_visited = (int64_t)cnt[1];
      c.stats.walk_fallbacks = (int64_t)cnt[2];
    }
  }
  HBT_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]);
  c.stats.build_ms = ms;
  c.stats.tree_builds = 1;
  c.stats.walk_targets = T;
  c.stats.rounds = 1;
  c.stats.kernel_launches = c.ls.launches;
}

// ---- pipelined hbtu_unbind_batch ------------------------------------------------------------------
// A batch of MANY independent hierarchies (a cosmological box: 5e5 subhaloes per GPU in the EAGLE-shaped configuration) has no
// dominant root whose upload could hide behind the deeper levels: all of it is needed by the first round, and the whole
// upload (104 ms per 1.7e8 particles alone on the link, 280 ms with eight GPUs sharing the host's memory) sat in front of
// the kernels.  Hierarchies never interact and results do not depend on how subhaloes are grouped into batches, so such a
// batch is cut at a hierarchy boundary into two parts of 1/4 and 3/4 of the particles; the second part is uploaded into a
// second pair of buffers while the first executes.  Only the first quarter of the upload stays exposed.
std::atomic<long long> g_pipeline_min{-1}; // particles; 0 = never.  HBTU_PIPELINE_MIN / hbtu_set_tuning("pipeline_min_particles")
long long pipeline_min()
{
  long long v = g_pipeline_min.load(std::memory_order_relaxed);
  if (v < 0)
  {
    const char *e = getenv("HBTU_PIPELINE_MIN");
    v = e ? atoll(e) : (1ll << 24);
    if (v < 0) v = 0;
    g_pipeline_min.store(v, std::memory_order_relaxed);
  }
  return v;
}

struct BatchPart
{
  int64_t s0, s1; // subhaloes [s0, s1): whole hierarchies
};

// parts of a batch that qualifies for pipelining (empty: run it in one piece).  Needs every hierarchy to be a contiguous index
// range with parents in front of their children - the layout integration/subhalo_unbind_b200.cpp::add_hierarchy produces.
std::vector<BatchPart> plan_parts(const Context &c, int64_t nsub, const int64_t *part_offset, const int64_t *nest_offset, const int32_t *nest_list)
{
  std::vector<BatchPart> none;
  const long long min_particles = pipeline_min();
  if (min_particles <= 0 || nsub < 64 || c.split_n > 1) return none;
  const int64_t N = part_offset[nsub] - part_offset[0];
  if (N < min_particles || part_offset[0] != 0) return none;
  std::vector<int32_t> parent((size_t)nsub, -1);
  if (nest_offset)
    for (int64_t s = 0; s < nsub; s++)
      for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++)
      {
        const int32_t ch = nest_list[k];
        if (ch < 0 || ch >= nsub || parent[ch] != -1) return none; // malformed: let stage() report it
        parent[ch] = (int32_t)s;
      }
  std::vector<int64_t> hstart; // first subhalo of every hierarchy
  for (int64_t s = 0; s < nsub; s++)
  {
    if (parent[s] < 0) hstart.push_back(s);
    else if (hstart.empty() || parent[s] < hstart.back() || parent[s] >= s) return none; // not depth-first contiguous
  }
  if (hstart.size() < 16) return none;
  hstart.push_back(nsub);
  for (size_t h = 0; h + 1 < hstart.size(); h++)
    if (part_offset[hstart[h + 1]] - part_offset[hstart[h]] > N / 4) return none; // a dominant hierarchy: the two-wave upload handles it
  // Two parts: the first quarter's upload is the exposed one (26 ms per 1.7e8 particles), the rest travels behind its kernels.
  // Every part pays its own ~28 rounds of planning and small launches (~1 ms each), which is why three parts (1/8, 3/8, 1/2)
  // measured no better (tools/gpu/c25.sh, c27.sh).
  const int64_t cuts[1] = {N / 4};
  std::vector<BatchPart> parts;
  int64_t begin = 0;
  size_t h = 0;
  for (int k = 0; k < 1; k++)
  {
    while (h + 1 < hstart.size() && part_offset[hstart[h]] < cuts[k]) h++;
    if (hstart[h] > begin && hstart[h] < nsub)
    {
      parts.push_back(BatchPart{begin, hstart[h]});
      begin = hstart[h];
    }
  }
  parts.push_back(BatchPart{begin, nsub});
  if (parts.size() < 2) return none;
  return parts;
}

// upload particles [p0, p1) of the caller's arrays into d_pos_next / d_vel_next on the copy stream, from the helper thread
void start_prefetch(Context &c, const float *pos_mass, const float *vel, int64_t p0, int64_t p1)
{
  finish_upload(c);
  grow(c.d_pos_next, c.cap_pos_next, p1 - p0);
  grow(c.d_vel_next, c.cap_vel_next, p1 - p0);
  const char *env = getenv("HBTU_UPLOAD_CHUNK_MB");
  const int64_t chunk = std::max<int64_t>(1, env ? atoll(env) : 32) * (int64_t)(1 << 20) / 32;
  c.up_wave_done = 0;
  c.up_error.clear();
  const int device = c.device;
  cudaStream_t cs = c.copy_stream;
  float4 *d_pos = c.d_pos_next, *d_vel = c.d_vel_next;
  Context *cp = &c;
  c.uploader = std::thread([=]() {
    const auto t0 = std::chrono::steady_clock::now();
    std::string err;
    if (cudaSetDevice(device) != cudaSuccess) err = "cudaSetDevice failed in the upload helper";
    for (int64_t b = p0; b < p1 && err.empty(); b += chunk)
    {
      const int64_t e = std::min(p1, b + chunk);
      cudaError_t e1 = cudaMemcpyAsync(d_pos + (b - p0), pos_mass + 4 * b, sizeof(float4) * (size_t)(e - b), cudaMemcpyHostToDevice, cs);
      cudaError_t e2 = cudaMemcpyAsync(d_vel + (b - p0), vel + 4 * b, sizeof(float4) * (size_t)(e - b), cudaMemcpyHostToDevice, cs);
      cudaError_t e3 = cudaStreamSynchronize(cs);
      if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        err = std::string("particle upload failed: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3));
    }
    std::lock_guard<std::mutex> lk(cp->up_m);
    if (!err.empty()) cp->up_error = err;
    cp->up_wave_done = 2;
    cp->up_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    cp->up_cv.notify_all();
  });
}

void add_stats(hbtu_stats &a, const hbtu_stats &b)
{
  a.kernel_launches += b.kernel_launches;
  a.tree_builds += b.tree_builds;
  a.walk_targets += b.walk_targets;
  a.pair_interactions += b.pair_interactions;
  a.nodes_visited += b.nodes_visited;
  a.rounds += b.rounds;
  a.walk_ms += b.walk_ms;
  a.build_ms += b.build_ms;
  a.other_ms += b.other_ms;
  a.h2d_ms += b.h2d_ms;
  a.d2h_ms += b.d2h_ms;
  a.h2d_bytes += b.h2d_bytes;
  a.d2h_bytes += b.d2h_bytes;
  a.execute_ms += b.execute_ms;
  a.tree_sources += b.tree_sources;
  a.walk_fallbacks += b.walk_fallbacks;
  a.stage_wall_ms += b.stage_wall_ms;
  a.execute_wall_ms += b.execute_wall_ms;
  a.fetch_wall_ms += b.fetch_wall_ms;
  for (int i = 0; i < 8; i++) a.phase_ms[i] += b.phase_ms[i];
}

void unbind_batch_pipelined(Context &c, const std::vector<BatchPart> &parts, const hbtu_epoch *epoch, const int64_t *part_offset, const float *pos_mass,
                            const float *vel, const int64_t *nest_offset, const int32_t *nest_list, hbtu_sub_io *io, int32_t flags,
                            int64_t order_capacity, int64_t *order_offset, int32_t *order_out, float *energy_out)
{
  if (!epoch || !io || !order_offset || !pos_mass || !vel || (!order_out && order_capacity > 0)) throw CudaError{HBTU_ERR_INVALID, "null argument"};
  hbtu_stats total;
  std::memset(&total, 0, sizeof(total));
  int64_t o0 = 0; // order entries written so far
  start_prefetch(c, pos_mass, vel, part_offset[parts[0].s0], part_offset[parts[0].s1]);
  std::vector<int64_t> po, no;
  std::vector<int32_t> nl;
  for (size_t k = 0; k < parts.size(); k++)
  {
    const int64_t s0 = parts[k].s0, s1 = parts[k].s1, ns = s1 - s0, p0 = part_offset[s0];
    po.resize((size_t)ns + 1);
    for (int64_t i = 0; i <= ns; i++) po[i] = part_offset[s0 + i] - p0;
    const int64_t *nop = nullptr;
    const int32_t *nlp = nullptr;
    if (nest_offset)
    {
      const int64_t n0 = nest_offset[s0];
      no.resize((size_t)ns + 1);
      for (int64_t i = 0; i <= ns; i++) no[i] = nest_offset[s0 + i] - n0;
      nl.resize((size_t)no[ns]);
      for (int64_t j = 0; j < no[ns]; j++) nl[j] = nest_list[n0 + j] - (int32_t)s0;
      nop = no.data();
      nlp = nl.data();
    }
    double t_stage = 0, t_exec = 0, t_fetch = 0, t_up;
    {
      WallTimer wt(t_stage);
      wait_upload_wave(c, 2); // this part's particles are in d_pos_next / d_vel_next
      finish_upload(c);
      t_up = c.up_ms;
      stage(c, epoch, ns, po.data(), pos_mass + 4 * p0, vel + 4 * p0, nop, nlp, io + s0, flags, true, true);
      c.sub_index_base = s0;
      if (k + 1 < parts.size()) start_prefetch(c, pos_mass, vel, part_offset[parts[k + 1].s0], part_offset[parts[k + 1].s1]);
    }
    {
      WallTimer wt(t_exec);
      execute_batch(c);
    }
    {
      WallTimer wt(t_fetch);
      c.order_index_base = p0;
      fetch_batch(c, io + s0, order_capacity - o0, order_offset + s0, order_out ? order_out + o0 : nullptr, energy_out ? energy_out + o0 : nullptr);
      for (int64_t i = 0; i <= ns; i++) order_offset[s0 + i] += o0;
      o0 = order_offset[s1];
    }
    hbtu_stats part = c.stats;
    part.stage_wall_ms = t_stage;
    part.execute_wall_ms = t_exec;
    part.fetch_wall_ms = t_fetch;
    part.h2d_ms = t_up;
    add_stats(total, part);
  }
  c.stats = total;
  c.pipelined = true;
}
} // namespace

extern "C" {

int hbtu_abi_version(void) { return HBTU_ABI_VERSION; }

const char *hbtu_last_error(const hbtu_ctx *ctx) { return ctx ? ctx->c.last_error.c_str() : g_create_error.c_str(); }

int hbtu_create(const hbtu_params *p, hbtu_ctx **out)
{
  if (out) *out = nullptr;
  if (!p || !out || p->struct_size != (int32_t)sizeof(hbtu_params))
  {
    g_create_error = "hbtu_params is null or has the wrong struct_size (ABI mismatch)";
    return HBTU_ERR_INVALID;
  }
  if (p->real_bytes != 4)
  {
    g_create_error = "only the HBTReal=float ABI variant is built";
    return HBTU_ERR_UNSUPPORTED;
  }
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); // no effect if the process has already created its CUDA context
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || p->device < 0 || p->device >= ndev)
  {
    g_create_error = std::string("no usable CUDA device (there is no CPU fallback): ") + (e != cudaSuccess ? cudaGetErrorString(e) : "bad ordinal");
    cudaGetLastError();
    return HBTU_ERR_NODEVICE;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, p->device);
  if (prop.major != 10)
  {
    g_create_error = "device is not sm_100 (this library is built for sm_100a only)";
    return HBTU_ERR_NODEVICE;
  }
  hbtu_ctx *ctx = new (std::nothrow) hbtu_ctx();
  if (!ctx) return HBTU_ERR_NOMEM;
  Context &c = ctx->c;
  c.params = *p;
  c.device = p->device;
  c.cfg.box_size = (float)p->box_size;
  c.cfg.box_half = (float)p->box_half;
  c.cfg.softening = (float)p->softening_halo;
  c.cfg.theta2 = (float)p->tree_node_open_angle_square;
  c.cfg.resolution = (float)p->tree_node_resolution;
  c.cfg.resolution_half = (float)p->tree_node_resolution_half;
  c.cfg.G = (float)p->G;
  c.cfg.bound_mass_precision = (float)p->bound_mass_precision;
  c.cfg.relax_factor = (float)p->source_sub_relax_factor;
  c.cfg.periodic = p->periodic_boundary_on != 0;
  c.cfg.min_num_part = p->min_num_part_of_sub;
  c.cfg.max_sample = p->max_sample_size;
  c.cfg.scale_factor = 1.f;
  int rc = guarded(ctx, [&](Context &cc) {
    // the compute stream and the upload stream must not share a hardware work queue: commands of aliased streams are dispatched in
    // issue order, and the multi-GB particle upload of hbtu_unbind_batch then delays every kernel issued after it (measured: the
    // whole upload, 104 ms per step).  Different priorities map to different queues; CUDA_DEVICE_MAX_CONNECTIONS (read when the
    // CUDA context is created: bench.py and the shim set it before the first CUDA call) widens the pool for equal priorities.
    int prio_least = 0, prio_greatest = 0;
    HBT_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    HBT_CUDA(cudaStreamCreateWithPriority(&cc.stream, cudaStreamNonBlocking, prio_greatest));
    HBT_CUDA(cudaStreamCreateWithPriority(&cc.copy_stream, cudaStreamNonBlocking, prio_least));
    for (auto &ev : cc.ev_wave) HBT_CUDA(cudaEventCreate(&ev));
    HBT_CUDA(cudaEventCreate(&cc.ev_copy0));
    for (auto &ev : cc.ev) HBT_CUDA(cudaEventCreate(&ev));
    for (auto &ev : cc.ev_exec) HBT_CUDA(cudaEventCreate(&ev));
    for (auto &ev : cc.ev_ph) HBT_CUDA(cudaEventCreate(&ev));
    HBT_CUDA(cudaMalloc(&cc.d_counters, kWalkCounters * sizeof(unsigned long long)));
  });
  if (rc != HBTU_OK)
  {
    g_create_error = c.last_error;
    hbtu_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return HBTU_OK;
}

void hbtu_destroy(hbtu_ctx *ctx)
{
  if (!ctx) return;
  Context &c = ctx->c;
  cudaSetDevice(c.device);
  finish_upload(c);
  if (c.copy_stream) cudaStreamSynchronize(c.copy_stream);
  if (c.stream) cudaStreamSynchronize(c.stream);
  c.arena.release();
  cudaFree(c.d_pos);
  cudaFree(c.d_vel);
  cudaFree(c.d_pos_next);
  cudaFree(c.d_vel_next);
  cudaFree(c.d_ids);
  cudaFree(c.d_ids_orig);
  cudaFree(c.d_E);
  cudaFree(c.d_rho);
  idtable_clear(c);
  cudaFree(c.d_subs);
  cudaFree(c.d_part_offset);
  cudaFree(c.d_slot_base);
  cudaFree(c.d_counters);
  if (c.h_ring) cudaFreeHost(c.h_ring);
  if (c.h_back) cudaFreeHost(c.h_back);
  for (auto &ev : c.ev)
    if (ev) cudaEventDestroy(ev);
  for (auto &ev : c.ev_exec)
    if (ev) cudaEventDestroy(ev);
  for (auto &ev : c.ev_ph)
    if (ev) cudaEventDestroy(ev);
  for (auto &ev : c.ev_wave)
    if (ev) cudaEventDestroy(ev);
  if (c.ev_copy0) cudaEventDestroy(c.ev_copy0);
  if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
  if (c.stream) cudaStreamDestroy(c.stream);
  delete ctx;
}

int64_t hbtu_order_capacity(int64_t nsub, const int64_t *part_offset, const int64_t *nest_offset, const int32_t *nest_list)
{
  if (nsub < 0 || !part_offset) return HBTU_ERR_INVALID;
  try
  {
    { // fast path (parents in front of their nested subhaloes): capacity = sum of own particles x (nesting depth + 1), since every
      // subhalo's slot range holds its own particles plus everything its descendants can feed upwards.  Anything unusual falls
      // through to the full forest construction below, which also diagnoses malformed nesting.
      std::vector<int32_t> parent((size_t)nsub, -1);
      bool ok = true;
      if (nest_offset)
        for (int64_t s = 0; s < nsub && ok; s++)
          for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++)
          {
            const int32_t ch = nest_list[k];
            if (ch <= s || ch >= nsub || parent[ch] != -1) { ok = false; break; }
            parent[ch] = (int32_t)s;
          }
      if (ok)
      {
        std::vector<int32_t> depth((size_t)nsub, 0);
        int64_t total = 0;
        for (int64_t s = 0; s < nsub; s++)
        {
          const int64_t n = part_offset[s + 1] - part_offset[s];
          if (n < 0 || n > 0x3fffffff) { ok = false; break; }
          if (parent[s] >= 0) depth[s] = depth[parent[s]] + 1;
          total += n * (depth[s] + 1);
        }
        if (ok && total <= 0x7fffffff00ll) return total;
      }
    }
    Context tmp;
    build_forest(tmp, nsub, part_offset, nest_offset, nest_list);
    return tmp.total_cap;
  }
  catch (...)
  {
    return HBTU_ERR_INVALID;
  }
}

int hbtu_plan_pipeline(int64_t nsub, const int64_t *part_offset, const int64_t *nest_offset, const int32_t *nest_list, int64_t *part_begin,
                       int max_parts)
{ // diagnostics: the parts hbtu_unbind_batch would run this batch in (host only, no device needed)
  if (nsub < 0 || !part_offset || !part_begin || max_parts < 1) return HBTU_ERR_INVALID;
  try
  {
    Context tmp;
    std::vector<BatchPart> parts;
    if (nsub > 0) parts = plan_parts(tmp, nsub, part_offset, nest_offset, nest_list);
    if (parts.empty()) parts.push_back(BatchPart{0, nsub});
    if ((int)parts.size() > max_parts) return HBTU_ERR_CAPACITY;
    for (size_t k = 0; k < parts.size(); k++) part_begin[k] = parts[k].s0;
    part_begin[parts.size()] = nsub;
    return (int)parts.size();
  }
  catch (...)
  {
    return HBTU_ERR_INVALID;
  }
}

int hbtu_stage(hbtu_ctx *ctx, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass,
               const float *vel, const int64_t *nest_offset, const int32_t *nest_list, const hbtu_sub_io *io, int32_t flags)
{
  return guarded(ctx, [&](Context &c) {
    WallTimer wt(c.stats.stage_wall_ms);
    stage(c, epoch, nsub, part_offset, pos_mass, vel, nest_offset, nest_list, io, flags);
  });
}
int hbtu_execute(hbtu_ctx *ctx)
{
  return guarded(ctx, [&](Context &c) {
    WallTimer wt(c.stats.execute_wall_ms);
    execute_batch(c);
  });
}
int hbtu_fetch(hbtu_ctx *ctx, hbtu_sub_io *io, int64_t order_capacity, int64_t *order_offset, int32_t *order_out, float *energy_out)
{
  return guarded(ctx, [&](Context &c) {
    if (!io || !order_offset || (!order_out && order_capacity > 0)) throw CudaError{HBTU_ERR_INVALID, "null output"};
    WallTimer wt(c.stats.fetch_wall_ms);
    fetch_batch(c, io, order_capacity, order_offset, order_out, energy_out);
  });
}
int hbtu_unbind_batch(hbtu_ctx *ctx, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass,
                      const float *vel, const int64_t *nest_offset, const int32_t *nest_list, hbtu_sub_io *io, int32_t flags,
                      int64_t order_capacity, int64_t *order_offset, int32_t *order_out, float *energy_out)
{
  // many independent hierarchies and no dominant one: run the batch in parts, uploading the next part behind the kernels of the
  // current one (unbind_batch_pipelined above); results do not depend on the grouping
  if (ctx && nsub > 0 && part_offset)
  {
    std::vector<BatchPart> parts;
    int rc0 = guarded(ctx, [&](Context &c) { parts = plan_parts(c, nsub, part_offset, nest_offset, nest_list); });
    if (rc0 != HBTU_OK) return rc0;
    if (!parts.empty())
    {
      int rc = guarded(ctx, [&](Context &c) {
        unbind_batch_pipelined(c, parts, epoch, part_offset, pos_mass, vel, nest_offset, nest_list, io, flags, order_capacity, order_offset, order_out, energy_out);
      });
      finish_upload(ctx->c); // never leave copies from the caller's buffers in flight behind the return
      return rc;
    }
  }
  // stage with the uploads on the copy stream (they overlap the kernels of the deeper nesting levels when the source arrays
  // are pinned - hbtu_host_alloc - and are plain staged copies otherwise), then execute + fetch
  int rc = guarded(ctx, [&](Context &c) {
    WallTimer wt(c.stats.stage_wall_ms);
    stage(c, epoch, nsub, part_offset, pos_mass, vel, nest_offset, nest_list, io, flags, true);
  });
  if (rc == HBTU_OK) rc = hbtu_execute(ctx);
  if (ctx) finish_upload(ctx->c); // never leave copies from the caller's buffers in flight behind the return
  if (rc != HBTU_OK) return rc;
  return hbtu_fetch(ctx, io, order_capacity, order_offset, order_out, energy_out);
}

/* Pinned host memory for the caller's staging arrays (pos_mass / vel / order_out): with it the uploads of hbtu_unbind_batch
 * are true asynchronous DMA that overlaps the kernels; any other host memory works too, just slower. */
void *hbtu_host_alloc(size_t bytes)
{
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess)
  {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void hbtu_host_free(void *p)
{
  if (p) cudaFreeHost(p);
}

int hbtu_tree_potential(hbtu_ctx *ctx, const hbtu_epoch *epoch, int64_t nsrc, const float *src_pos_mass, int64_t ntgt,
                        const float *tgt_pos, const float *tgt_self_mass, const float *tgt_vel, const double *ref_pos,
                        const double *ref_vel, double *out)
{
  return guarded(ctx, [&](Context &c) {
    tree_potential(c, epoch, nsrc, src_pos_mass, ntgt, tgt_pos, tgt_self_mass, tgt_vel, ref_pos, ref_vel, out);
  });
}

int hbtu_profile_batch(hbtu_ctx *ctx, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass,
                       hbtu_profile_io *io)
{
  return guarded(ctx, [&](Context &c) { profile_batch(c, epoch, nsub, part_offset, pos_mass, io); });
}

int hbtu_profile_executed(hbtu_ctx *ctx, hbtu_profile_io *io)
{
  return guarded(ctx, [&](Context &c) { profile_executed(c, io); });
}

int hbtu_mask_batch(hbtu_ctx *ctx, int64_t nsub, const int64_t *part_offset, const int64_t *particle_id, const int64_t *nest_offset,
                    const int32_t *nest_list, const int64_t *nbound, int64_t *new_count, int32_t *keep_index)
{
  return guarded(ctx, [&](Context &c) { mask_batch(c, nsub, part_offset, particle_id, nest_offset, nest_list, nbound, new_count, keep_index); });
}

int hbtu_idtable_build(hbtu_ctx *ctx, int64_t n, const int64_t *particle_id)
{
  return guarded(ctx, [&](Context &c) { idtable_build(c, n, particle_id); });
}
int hbtu_idtable_query(hbtu_ctx *ctx, int64_t nq, const int64_t *query_id, int64_t *index_out)
{
  return guarded(ctx, [&](Context &c) { idtable_query(c, nq, query_id, index_out); });
}
int hbtu_idtable_clear(hbtu_ctx *ctx)
{
  return guarded(ctx, [&](Context &c) { idtable_clear(c); });
}

int hbtu_detect_traps(hbtu_ctx *ctx, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass, const float *vel,
                      const int64_t *nest_offset, const int32_t *nest_list, hbtu_trap_io *io)
{
  return guarded(ctx, [&](Context &c) { detect_traps(c, epoch, nsub, part_offset, pos_mass, vel, nest_offset, nest_list, io); });
}

int hbtu_get_stats(const hbtu_ctx *ctx, hbtu_stats *out)
{
  if (!ctx || !out) return HBTU_ERR_INVALID;
  *out = ctx->c.stats;
  return HBTU_OK;
}

/* diagnostics switch (not part of the reference seam): count accepted interactions and warp node visits
 * in the walk kernels of subsequent calls; costs a few percent, so bench.py enables it for one untimed pass */
int hbtu_set_counting(hbtu_ctx *ctx, int on)
{
  if (!ctx) return HBTU_ERR_INVALID;
  ctx->c.count_interactions = on != 0;
  return HBTU_OK;
}

/* diagnostics (no reference counterpart): kernel-routing knobs of the walk, process-wide (device_tree.cuh: WalkTuning).
 * Results do not depend on them up to fp64 summation order; tests force every route, tools/ab_walk.py times them. */
int hbtu_set_tuning(const char *key, int64_t value)
{
  if (!key) return HBTU_ERR_INVALID;
  WalkTuning &t = walk_tuning();
  const std::string k(key);
  if (k == "walk_tpl") t.forced_tpl = (int)value;
  else if (k == "walk_big4") t.big4 = (int)value;
  else if (k == "walk_big2") t.big2 = (int)value;
  else if (k == "walk_group_min") t.group_min = (int)value;
  else if (k == "pipeline_min_particles") g_pipeline_min.store(value < 0 ? 0 : value, std::memory_order_relaxed);
  else if (k == "walk_masked_pairs") t.masked_pairs = value == 1 ? 1 : 2;
  else if (k == "walk_masked_blocks") t.masked_blocks = (int)value;
  else if (k == "walk_small_max") t.small_max = (int)value;
  else return HBTU_ERR_INVALID;
  return HBTU_OK;
}
int64_t hbtu_get_tuning(const char *key)
{
  if (!key) return -1;
  const WalkTuning &t = walk_tuning();
  const std::string k(key);
  if (k == "walk_tpl") return t.forced_tpl;
  if (k == "walk_big4") return t.big4;
  if (k == "walk_big2") return t.big2;
  if (k == "walk_group_min") return t.group_min;
  if (k == "pipeline_min_particles") return pipeline_min();
  if (k == "walk_masked_pairs") return t.masked_pairs;
  if (k == "walk_masked_blocks") return t.masked_blocks;
  if (k == "walk_small_max") return t.small_max;
  return -1;
}

} // extern "C"
