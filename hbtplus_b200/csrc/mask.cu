// mask.cu - exclusive particle ownership on the device, sm_100a.
//
// Replaces SubhaloSnapshot_t::MaskSubhalos / SubhaloMasker_t::Mask (src/subhalo_tracking.cpp:793-841): per host group
// the reference walks the hierarchy depth first (children before their parent, in list order) with ONE
// unordered_set<HBTInt> and lets every subhalo keep, in order, the particles whose Id was not inserted before.  That is:
// among all list entries of one hierarchy that carry the same Id, the entry with the smallest (visit rank of its
// subhalo, position) survives; orphans (Nbound <= 1) neither exclude nor lose anything (:806).
//
// Data-parallel form (the hash set is the reference's serial bottleneck; here everything is a streaming pass):
//   keys (K1)     8 B Id read + 12 B (Id, entry) write per entry; orphans' entries are kept at once            HBM
//   CUB radix sort by Id (64-bit key, 32-bit entry index)                                                     HBM
//   runs (K2)     one thread per sorted entry: it survives unless an entry of the same Id AND the same root
//                 hierarchy with a smaller (rank, position) exists; runs of equal Id are a few entries long     HBM
//   CUB scan of the keep flags + scatter (K3): ascending survivor indices per subhalo, new counts              HBM
// The visit ranks (post-order of the nest forest) are computed on the host: O(nsub).
#include <cub/cub.cuh>

#include <vector>

#include "context.cuh"

namespace hbt
{

static constexpr int kMB = 256;
static inline int mgrid(int64_t n) { return n > 0 ? div_up(n, kMB) : 1; }

struct MaskSub
{
  int64_t part_off;
  int root;   // hierarchy (exclusion set) this subhalo belongs to
  int rank;   // visit order inside the hierarchy: post-order, children in list order; -1 = orphan (skipped)
};

__device__ __forceinline__ int mask_find(const MaskSub *__restrict__ subs, int n, int64_t e)
{ // largest s with part_off[s] <= e and a non-empty list there (empty lists share the successor's offset)
  int lo = 0, hi = n;
  while (hi - lo > 1)
  {
    int mid = (lo + hi) >> 1;
    if (subs[mid].part_off <= e) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kMB) mask_keys_kernel(const MaskSub *__restrict__ subs, int nsub, int64_t N, const int64_t *__restrict__ ids,
                                                         uint64_t *__restrict__ key, int *__restrict__ val, int *__restrict__ sub_of,
                                                         int *__restrict__ keep)
{
  const int64_t e = (int64_t)blockIdx.x * kMB + threadIdx.x;
  if (e >= N) return;
  const int s = mask_find(subs, nsub, e);
  key[e] = (uint64_t)ids[e];
  val[e] = (int)e;
  sub_of[e] = s;
  keep[e] = subs[s].rank < 0 ? 1 : 0; // orphans keep their whole list
}

// one thread per sorted entry; the run of equal Ids around it is scanned in both directions
__global__ void __launch_bounds__(kMB) mask_runs_kernel(const MaskSub *__restrict__ subs, int64_t N, const uint64_t *__restrict__ skey,
                                                         const int *__restrict__ sval, const int *__restrict__ sub_of, int *__restrict__ keep)
{
  const int64_t k = (int64_t)blockIdx.x * kMB + threadIdx.x;
  if (k >= N) return;
  const int e = sval[k];
  const MaskSub me = subs[sub_of[e]];
  if (me.rank < 0) return;
  const uint64_t id = skey[k];
  bool beaten = false;
  for (int64_t j = k - 1; j >= 0 && skey[j] == id && !beaten; j--)
  {
    const int f = sval[j];
    const MaskSub o = subs[sub_of[f]];
    beaten = o.rank >= 0 && o.root == me.root && (o.rank < me.rank || (o.rank == me.rank && f < e));
  }
  for (int64_t j = k + 1; j < N && skey[j] == id && !beaten; j++)
  {
    const int f = sval[j];
    const MaskSub o = subs[sub_of[f]];
    beaten = o.rank >= 0 && o.root == me.root && (o.rank < me.rank || (o.rank == me.rank && f < e));
  }
  if (!beaten) keep[e] = 1;
}

__global__ void __launch_bounds__(kMB) mask_scatter_kernel(const MaskSub *__restrict__ subs, int64_t N, const int *__restrict__ sub_of,
                                                            const int *__restrict__ keep, const int *__restrict__ scan,
                                                            int *__restrict__ keep_index)
{ // scan = exclusive prefix sum of keep over all entries
  const int64_t e = (int64_t)blockIdx.x * kMB + threadIdx.x;
  if (e >= N || !keep[e]) return;
  const int64_t b = subs[sub_of[e]].part_off;
  keep_index[b + (scan[e] - scan[b])] = (int)e;
}

__global__ void mask_counts_kernel(const MaskSub *__restrict__ subs, int nsub, int64_t N, const int *__restrict__ keep,
                                   const int *__restrict__ scan, int64_t *__restrict__ new_count)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsub) return;
  const int64_t b = subs[s].part_off, e = s + 1 < nsub ? subs[s + 1].part_off : N;
  if (e <= b) { new_count[s] = 0; return; }
  new_count[s] = (int64_t)(scan[e - 1] + keep[e - 1]) - scan[b];
}

void mask_batch(Context &c, int64_t nsub, const int64_t *part_offset, const int64_t *particle_id, const int64_t *nest_offset,
                const int32_t *nest_list, const int64_t *nbound, int64_t *new_count, int32_t *keep_index)
{
  if (nsub < 0 || !part_offset || !nbound || !new_count) throw CudaError{HBTU_ERR_INVALID, "bad argument"};
  if (nsub == 0) return;
  if (nsub > 0x7fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "too many subhaloes"};
  const int64_t N = part_offset[nsub];
  if (N > 0x7fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "batch larger than 2^31 list entries"};
  if (N > 0 && (!particle_id || !keep_index)) throw CudaError{HBTU_ERR_INVALID, "null particle arrays"};
  // forest: roots, visit ranks (post-order, children in list order) - src/subhalo_tracking.cpp:801-805
  std::vector<int> parent(nsub, -1);
  if (nest_offset)
  {
    if (nest_offset[0] != 0 || (nest_offset[nsub] > 0 && !nest_list)) throw CudaError{HBTU_ERR_INVALID, "malformed nest lists"};
    for (int64_t s = 0; s < nsub; s++)
      for (int64_t k = nest_offset[s]; k < nest_offset[s + 1]; k++)
      {
        const int32_t ch = nest_list[k];
        if (ch < 0 || ch >= nsub || ch == s || parent[ch] >= 0) throw CudaError{HBTU_ERR_INVALID, "malformed nest forest"};
        parent[ch] = (int)s;
      }
  }
  std::vector<MaskSub> subs(nsub);
  {
    struct Item { int64_t s, next; };
    std::vector<Item> stack;
    int nroot = 0;
    int64_t visited = 0;
    for (int64_t r = 0; r < nsub; r++)
    {
      if (parent[r] >= 0) continue;
      int rank = 0;
      stack.assign(1, Item{r, nest_offset ? nest_offset[r] : 0});
      while (!stack.empty())
      {
        Item &it = stack.back();
        const int64_t end = nest_offset ? nest_offset[it.s + 1] : 0;
        if (nest_offset && it.next < end)
        {
          const int64_t ch = nest_list[it.next++];
          stack.push_back(Item{ch, nest_offset[ch]});
          continue;
        }
        subs[it.s].root = nroot;
        subs[it.s].rank = nbound[it.s] <= 1 ? -1 : rank;
        rank++;
        visited++;
        stack.pop_back();
      }
      nroot++;
    }
    if (visited != nsub) throw CudaError{HBTU_ERR_INVALID, "nest lists contain a cycle"};
  }
  for (int64_t s = 0; s < nsub; s++)
  {
    if (part_offset[s + 1] < part_offset[s]) throw CudaError{HBTU_ERR_INVALID, "part_offset not monotone"};
    subs[s].part_off = part_offset[s];
  }
  if (N == 0)
  {
    for (int64_t s = 0; s < nsub; s++) new_count[s] = 0;
    return;
  }
  c.staged = c.executed = false; // the arena is shared with a staged batch's rounds
  cudaStream_t st = c.stream;
  Arena &ar = c.arena;
  ar.reset();
  ar.reserve(N * 60 + nsub * (int64_t)(sizeof(MaskSub) + 8) + (64 << 20));
  c.ls.launches = 0;
  int64_t *d_ids = ar.alloc<int64_t>(N);
  HBT_CUDA(cudaMemcpyAsync(d_ids, particle_id, sizeof(int64_t) * (size_t)N, cudaMemcpyHostToDevice, st));
  MaskSub *d_subs = ar.alloc<MaskSub>(nsub);
  HBT_CUDA(cudaMemcpyAsync(d_subs, subs.data(), sizeof(MaskSub) * (size_t)nsub, cudaMemcpyHostToDevice, st));
  uint64_t *key_a = ar.alloc<uint64_t>(N), *key_b = ar.alloc<uint64_t>(N);
  int *val_a = ar.alloc<int>(N), *val_b = ar.alloc<int>(N);
  int *sub_of = ar.alloc<int>(N), *keep = ar.alloc<int>(N), *scan = ar.alloc<int>(N);
  int *d_index = ar.alloc<int>(N);
  int64_t *d_count = ar.alloc<int64_t>(nsub);
  HBT_CUDA(cudaEventRecord(c.ev_exec[0], st)); // kernels only
  mask_keys_kernel<<<mgrid(N), kMB, 0, st>>>(d_subs, (int)nsub, N, d_ids, key_a, val_a, sub_of, keep);
  HBT_CHECK_LAUNCH();
  cub::DoubleBuffer<uint64_t> dk(key_a, key_b);
  cub::DoubleBuffer<int> dv(val_a, val_b);
  size_t tb = 0;
  HBT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, N, 0, 64, st));
  void *tmp = ar.alloc<char>((int64_t)tb);
  HBT_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, N, 0, 64, st));
  mask_runs_kernel<<<mgrid(N), kMB, 0, st>>>(d_subs, N, dk.Current(), dv.Current(), sub_of, keep);
  HBT_CHECK_LAUNCH();
  size_t sb = 0;
  HBT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, keep, scan, N, st));
  void *tmp2 = ar.alloc<char>((int64_t)sb);
  HBT_CUDA(cub::DeviceScan::ExclusiveSum(tmp2, sb, keep, scan, N, st));
  mask_scatter_kernel<<<mgrid(N), kMB, 0, st>>>(d_subs, N, sub_of, keep, scan, d_index);
  HBT_CHECK_LAUNCH();
  mask_counts_kernel<<<mgrid(nsub), kMB, 0, st>>>(d_subs, (int)nsub, N, keep, scan, d_count);
  HBT_CHECK_LAUNCH();
  c.ls.launches += 4 + 9 + 2;
  HBT_CUDA(cudaEventRecord(c.ev_exec[1], st));
  HBT_CUDA(cudaMemcpyAsync(keep_index, d_index, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, st));
  HBT_CUDA(cudaMemcpyAsync(new_count, d_count, sizeof(int64_t) * (size_t)nsub, cudaMemcpyDeviceToHost, st));
  HBT_CUDA(cudaStreamSynchronize(st));
  std::memset(&c.stats, 0, sizeof(c.stats));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c.ev_exec[0], c.ev_exec[1]);
  c.stats.execute_ms = ms;
  c.stats.other_ms = ms;
  c.stats.kernel_launches = c.ls.launches;
  c.stats.walk_targets = N;
  c.stats.h2d_bytes = N * 8 + nsub * (int64_t)sizeof(MaskSub);
  c.stats.d2h_bytes = N * 4 + nsub * 8;
}

} // namespace hbt
