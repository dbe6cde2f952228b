// context.cuh - the hbtu_ctx object: one CUDA device, one stream, one staged batch.
#pragma once
#include <condition_variable>
#include <mutex>
#include <string>
#include <algorithm>
#include <thread>
#include <utility>
#include <vector>

#include "device_tree.cuh"

namespace hbt
{

struct SubHost
{ // host mirror of a subhalo's place in the batch and of its iteration state
  int64_t slot_base = 0, part_begin = 0, cap = 0;
  int n_own = 0, n_src = 0, nbound = 0, nlast = 0, correction = 0, iterations = 0;
  int depth = 0, parent = -1;
  bool done = false, disrupted = false, is_orphan = false;
  std::vector<int> children; // NestedSubhalos, in list order
};

struct Context
{
  hbtu_params params{};
  DevConfig cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;                 // uploads that run behind the kernels of the deeper nesting levels
  cudaEvent_t ev_wave[2] = {nullptr, nullptr};        // [0]: everything but the dominant root is in HBM, [1]: the dominant root too
  cudaEvent_t ev_copy0 = nullptr;                     // start of the uploads on copy_stream (h2d_ms)
  // asynchronous staging (hbtu_unbind_batch): a helper host thread feeds the particle arrays to the copy stream in chunks and
  // reports the two waves done; the scheduler waits on the HOST for a wave right before the first round that needs it
  std::thread uploader;
  std::mutex up_m;
  std::condition_variable up_cv;
  int up_wave_done = 0;   // 0: nothing, 1: everything but the dominant root is in HBM, 2: all of it
  std::string up_error;   // set by the helper if a copy failed
  double up_ms = 0.0;     // wall time of the whole upload
  bool staged_async = false;                          // ... and its h2d_ms is read from the copy stream's events after the execution
  bool waves_pending = false;                         // the staged batch was uploaded asynchronously (hbtu_unbind_batch)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_exec[2] = {nullptr, nullptr};
  cudaEvent_t ev_ph[9] = {}; // phase boundaries of a round (hbtu_stats.phase_ms)
  std::string last_error;

  // staged batch -------------------------------------------------------------------------------
  bool staged = false, executed = false;
  int64_t nsub = 0, N = 0, total_cap = 0;
  int32_t flags = 0;
  int max_depth = 0;
  std::vector<SubHost> subs;
  std::vector<std::vector<int>> levels;
  std::vector<hbtu_sub_io> io_in;
  std::vector<int> refine_list; // converged subhaloes of the current level that need RefineBindingEnergyOrder

  // persistent device buffers (grow-only)
  float4 *d_pos = nullptr, *d_vel = nullptr;
  int64_t cap_particles = 0, cap_vel = 0;
  // pipelined hbtu_unbind_batch (capi.cu): while one part of the batch executes, the upload helper fills this second pair of
  // buffers with the next part's particles; staging that part swaps the pairs
  float4 *d_pos_next = nullptr, *d_vel_next = nullptr;
  int64_t cap_pos_next = 0, cap_vel_next = 0;
  int64_t order_index_base = 0; // added to the particle indices fetch_batch returns (first particle of the part in the caller's arrays)
  int64_t sub_index_base = 0;   // first subhalo of the part in the caller's batch (the sampled mode's permutation is keyed by subhalo index)
  bool pipelined = false;       // the last hbtu_unbind_batch ran in parts: only its last part is resident
  int *d_ids = nullptr, *d_ids_orig = nullptr;
  float *d_E = nullptr;
  int *d_rho = nullptr; // periodic runs: index of every Elist entry in the REFERENCE's Elist order (unbind_batch.cu, rho_*)
  int64_t cap_slots = 0, cap_rho = 0;
  SubState *d_subs = nullptr;
  int64_t *d_part_offset = nullptr, *d_slot_base = nullptr;
  int64_t cap_subs = 0;
  void *d_idt_slots = nullptr; // idtable.cu: open-addressing table (Id, index) of the last hbtu_idtable_build, idt_cap + 1 slots
  int64_t idt_cap = 0;
  int64_t idt_n = 0;
  unsigned long long *d_counters = nullptr;
  bool count_interactions = false;

  // target split of the walk across cooperating contexts (hbtu_set_walk_split / hbtu_split_group_*)
  int split_rank = 0, split_n = 1;
  hbtu_allreduce_fn split_fn = nullptr;
  void *split_user = nullptr;

  // Small host tables (segments, offsets, job lists, the initial SubState records) reach the device through a pinned, MAPPED
  // staging ring and a copy kernel on the compute stream, not through cudaMemcpyAsync: the DMA engine serves host-to-device
  // copies first-in-first-out across streams, so a 16-byte table copy issued while hbtu_unbind_batch's multi-GB particle
  // upload is in flight would wait for all of it (measured: 104 ms per step) - the SMs read the ring over PCIe instead.
  char *h_ring = nullptr, *d_ring = nullptr; // host pointer / device alias of the ring
  size_t ring_cap = 0, ring_used = 0;
  char *h_back = nullptr; // pinned landing buffer of the device-to-host readbacks (round results, final SubState records): a pageable
  size_t back_cap = 0;    // destination costs a staging pass per copy, which shows with 1e5..1e6 subhaloes per batch

  Arena arena;
  LaunchStats ls;
  hbtu_stats stats{};
};

// Host-side loops over the subhaloes / segments of a batch with 1e5..1e6 of them are bound by a few hundred bytes of host
// memory traffic per element and run while the GPU waits for the round they plan.  [0, n) is cut into up to 8 contiguous
// chunks of at least `grain` elements; run_chunks calls fn(chunk, begin, end) for every chunk, each on its own host thread
// (the caller's thread takes chunk 0).  fn must not throw.
inline std::vector<std::pair<int64_t, int64_t>> chunk_ranges(int64_t n, int64_t grain)
{
  const unsigned hw = std::thread::hardware_concurrency();
  int64_t k = std::min<int64_t>(std::min<unsigned>(8u, hw ? hw : 1u), grain > 0 ? n / grain : 1);
  if (k < 1) k = 1;
  const int64_t per = (n + k - 1) / k;
  std::vector<std::pair<int64_t, int64_t>> r;
  for (int64_t t = 0; t < k; t++) r.push_back({std::min(n, t * per), std::min(n, (t + 1) * per)});
  return r;
}
template <class F>
inline void run_chunks(const std::vector<std::pair<int64_t, int64_t>> &r, F &&fn)
{
  std::vector<std::thread> pool;
  for (size_t t = 1; t < r.size(); t++) pool.emplace_back([&fn, &r, t] { fn((int)t, r[t].first, r[t].second); });
  if (!r.empty()) fn(0, r[0].first, r[0].second);
  for (auto &th : pool) th.join();
}
template <class F>
inline void parallel_ranges(int64_t n, int64_t grain, F &&fn)
{
  run_chunks(chunk_ranges(n, grain), [&fn](int, int64_t b, int64_t e) { fn(b, e); });
}

// reserve `bytes` in the pinned mapped staging ring (returns the host pointer; may synchronise the stream to wrap) ...
char *ring_reserve(Context &c, size_t bytes);
// ... and, once the host has filled them, copy them to `dst` (device) with a kernel on c.stream
void ring_commit(Context &c, void *dst, const char *staged, size_t bytes);
void execute_batch(Context &c);
// block the calling host thread until upload wave `wave` (1 or 2) of an asynchronously staged batch has landed
void wait_upload_wave(Context &c, int wave);
// join the upload helper (no-op if none is running)
void finish_upload(Context &c);
// stage `bytes` of host data in the ring and copy them to `dst` (device) with a kernel on c.stream (context.cuh: h_ring)
void upload_bytes(Context &c, void *dst, const void *src, size_t bytes);
// pinned host buffer of at least `bytes` for a readback (grow-only; valid until the next call)
void *readback_buffer(Context &c, size_t bytes);
// profile.cu: Subhalo_t::CalculateProfileProperties + CalculateShape for a batch of particle lists
void profile_batch(Context &c, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass, hbtu_profile_io *io);
void profile_executed(Context &c, hbtu_profile_io *io);
// mask.cu: SubhaloSnapshot_t::MaskSubhalos for a forest of particle-Id lists
void mask_batch(Context &c, int64_t nsub, const int64_t *part_offset, const int64_t *particle_id, const int64_t *nest_offset,
                const int32_t *nest_list, const int64_t *nbound, int64_t *new_count, int32_t *keep_index);
// idtable.cu: MappedIndexTable_t::Fill / GetIndices
void idtable_build(Context &c, int64_t n, const int64_t *particle_id);
void idtable_query(Context &c, int64_t nq, const int64_t *query_id, int64_t *index_out);
void idtable_clear(Context &c);
// trap.cu: detection part of SubhaloSnapshot_t::MergeSubhalos
void detect_traps(Context &c, const hbtu_epoch *epoch, int64_t nsub, const int64_t *part_offset, const float *pos_mass, const float *vel,
                  const int64_t *nest_offset, const int32_t *nest_list, hbtu_trap_io *io);
void fetch_batch(Context &c, hbtu_sub_io *io, int64_t order_capacity, int64_t *order_offset, int32_t *order_out, float *energy_out);

} // namespace hbt

struct hbtu_ctx
{
  hbt::Context c;
};
