// split_group.cu - cooperation of several contexts on ONE batch: the walk's target split (include/hbt_unbind.h,
// hbtu_set_walk_split) and the built-in all-reduce for contexts of one process.
//
// The built-in exchange is one kernel per member and round: after a host barrier (every member's walk has finished and its
// staging array is published) member r sums elements [r*n/R, (r+1)*n/R) of ALL members' arrays through peer-mapped loads and
// writes the sum into ALL members' arrays through peer-mapped stores - reduce-scatter and all-gather fused over NVLink /
// NVSwitch peer memory, no NCCL, no staging through the host.  Slices are disjoint, so members never touch the same element.
// A second host barrier makes every member's stores visible before anybody reads its array.
#include <condition_variable>
#include <mutex>
#include <vector>

#include "context.cuh"

namespace hbt
{
static constexpr int kMaxSplit = 16;
struct PeerBufs
{
  float *p[kMaxSplit];
};

__global__ void __launch_bounds__(256) peer_allreduce_kernel(PeerBufs bufs, int nranks, int64_t begin, int64_t end)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = begin + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < end; i += stride)
  {
    if (i + 4 <= end && (i & 3) == 0)
    { // 16-byte peer loads / stores
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < nranks; r++)
      {
        const float4 v = *reinterpret_cast<const float4 *>(bufs.p[r] + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      for (int r = 0; r < nranks; r++) *reinterpret_cast<float4 *>(bufs.p[r] + i) = s;
    }
    else
      for (int64_t j = i; j < end && j < i + 4; j++)
      {
        float s = 0.f;
        for (int r = 0; r < nranks; r++) s += bufs.p[r][j];
        for (int r = 0; r < nranks; r++) bufs.p[r][j] = s;
      }
  }
}
} // namespace hbt

using namespace hbt;

struct hbtu_split_group
{
  int nranks = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t generation = 0;
  bool failed = false;
  float *buf[kMaxSplit] = {};
  int64_t count[kMaxSplit] = {};
  int device[kMaxSplit] = {};
  hbtu_ctx *member[kMaxSplit] = {};
  struct Slot { hbtu_split_group *g; int rank; } slot[kMaxSplit];

  // returns false if some member reported a failure in this phase
  bool barrier(bool ok)
  {
    std::unique_lock<std::mutex> lk(m);
    if (!ok) failed = true;
    const uint64_t gen = generation;
    if (++arrived == nranks)
    {
      arrived = 0;
      generation++;
      cv.notify_all();
    }
    else
      cv.wait(lk, [&] { return generation != gen; });
    return !failed;
  }
};

static int group_allreduce(void *user, float *device_buf, int64_t count, void *cuda_stream)
{
  auto *slot = static_cast<hbtu_split_group::Slot *>(user);
  hbtu_split_group *g = slot->g;
  const int r = slot->rank, R = g->nranks;
  g->buf[r] = device_buf;
  g->count[r] = count;
  if (!g->barrier(true)) return 1; // everybody's walk is done and its array is published
  bool ok = true;
  for (int q = 0; q < R; q++) ok = ok && g->count[q] == count; // the members run the same batch in lock step
  if (ok && count > 0)
  {
    PeerBufs pb{};
    for (int q = 0; q < R; q++) pb.p[q] = g->buf[q];
    const int64_t chunk = ((count + R - 1) / R + 3) & ~(int64_t)3;
    const int64_t b = std::min<int64_t>(count, chunk * r), e = std::min<int64_t>(count, chunk * (r + 1));
    if (e > b)
    {
      const int grid = (int)std::min<int64_t>(148 * 8, (e - b + 1023) / 1024);
      peer_allreduce_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(pb, R, b, e);
      ok = cudaGetLastError() == cudaSuccess;
    }
    ok = ok && cudaStreamSynchronize((cudaStream_t)cuda_stream) == cudaSuccess;
  }
  return g->barrier(ok) ? 0 : 1; // all members' stores have landed
}

extern "C" {

int hbtu_set_walk_split(hbtu_ctx *ctx, int rank, int nranks, hbtu_allreduce_fn allreduce, void *user)
{
  if (!ctx) return HBTU_ERR_INVALID;
  Context &c = ctx->c;
  if (nranks <= 1 || !allreduce)
  {
    c.split_rank = 0;
    c.split_n = 1;
    c.split_fn = nullptr;
    c.split_user = nullptr;
    return HBTU_OK;
  }
  if (rank < 0 || rank >= nranks)
  {
    c.last_error = "hbtu_set_walk_split: rank out of range";
    return HBTU_ERR_INVALID;
  }
  c.split_rank = rank;
  c.split_n = nranks;
  c.split_fn = allreduce;
  c.split_user = user;
  return HBTU_OK;
}

hbtu_split_group *hbtu_split_group_create(int nranks)
{
  if (nranks < 1 || nranks > kMaxSplit) return nullptr;
  hbtu_split_group *g = new (std::nothrow) hbtu_split_group();
  if (g) g->nranks = nranks;
  return g;
}

int hbtu_split_group_join(hbtu_split_group *group, hbtu_ctx *ctx, int rank)
{
  if (!group || !ctx || rank < 0 || rank >= group->nranks) return HBTU_ERR_INVALID;
  std::lock_guard<std::mutex> lk(group->m);
  group->member[rank] = ctx;
  group->device[rank] = ctx->c.device;
  group->slot[rank] = hbtu_split_group::Slot{group, rank};
  // peer access between the members' devices (a no-op for contexts of the same device)
  for (int q = 0; q < group->nranks; q++)
    if (group->member[q] && group->device[q] != ctx->c.device)
    {
      for (int dir = 0; dir < 2; dir++)
      {
        const int from = dir ? group->device[q] : ctx->c.device, to = dir ? ctx->c.device : group->device[q];
        int can = 0;
        cudaDeviceCanAccessPeer(&can, from, to);
        if (!can)
        {
          ctx->c.last_error = "hbtu_split_group_join: the devices of the group cannot access each other's memory";
          return HBTU_ERR_UNSUPPORTED;
        }
        cudaSetDevice(from);
        cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        {
          ctx->c.last_error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
          cudaGetLastError();
          return HBTU_ERR_CUDA;
        }
        cudaGetLastError();
      }
    }
  return hbtu_set_walk_split(ctx, rank, group->nranks, group->nranks > 1 ? group_allreduce : nullptr, &group->slot[rank]);
}

void hbtu_split_group_destroy(hbtu_split_group *group)
{
  if (!group) return;
  for (int q = 0; q < group->nranks; q++)
    if (group->member[q]) hbtu_set_walk_split(group->member[q], 0, 1, nullptr, nullptr);
  delete group;
}

} // extern "C"
