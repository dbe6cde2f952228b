// idtable.cu - particle Id -> index lookup on the device, sm_100a.
//
// Replaces the on-node part of ParticleExchanger_t::QueryParticles (src/particle_exchanger.h:196-211):
//   MappedIndexTable_t::Fill        src/hash.tpp:18-32         sort the (Id, index) pairs of the snapshot
//   MappedIndexTable_t::GetIndices  src/hash_remote.tpp:9-88   batch binary search of sorted query Ids
// build : one CUB radix sort of (order-preserving uint64 of the signed Id, index); the sorted table stays in the context
// query : one thread per query, lower_bound over the sorted keys (the top ~20 levels of the search stay in the 126 MB L2;
//         the queries need no sorting and no order restoration), -1 (SpecialConst::NullParticleId) when absent.
// HBM-bound integer work: 8 B read + 12 B written per table entry (one algorithmic sort pass), 8 B read + 8 B written +
// one 8-B probe per query.
#include <cub/cub.cuh>

#include "context.cuh"

namespace hbt
{

static constexpr int kIB = 256;
static inline int igrid(int64_t n) { return n > 0 ? div_up(n, kIB) : 1; }

__device__ __forceinline__ uint64_t id_key(int64_t id) { return (uint64_t)id ^ 0x8000000000000000ull; } // signed order

__global__ void __launch_bounds__(kIB) idtable_keys_kernel(const int64_t *__restrict__ ids, int64_t n, uint64_t *__restrict__ key, int *__restrict__ val)
{
  const int64_t i = (int64_t)blockIdx.x * kIB + threadIdx.x;
  if (i >= n) return;
  key[i] = id_key(ids[i]);
  val[i] = (int)i;
}

__global__ void __launch_bounds__(kIB) idtable_query_kernel(const uint64_t *__restrict__ skey, const int *__restrict__ sval, int64_t n,
                                                             const int64_t *__restrict__ query, int64_t nq, int64_t *__restrict__ out)
{
  const int64_t q = (int64_t)blockIdx.x * kIB + threadIdx.x;
  if (q >= nq) return;
  const uint64_t key = id_key(query[q]);
  int64_t lo = 0, hi = n;
  while (lo < hi)
  { // lower_bound (src/hash_remote.tpp:76)
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(&skey[mid]) < key) lo = mid + 1; else hi = mid;
  }
  out[q] = (lo < n && skey[lo] == key) ? (int64_t)sval[lo] : -1; // :77-83
}

void idtable_clear(Context &c)
{
  cudaFree(c.d_idt_key);
  cudaFree(c.d_idt_val);
  c.d_idt_key = nullptr;
  c.d_idt_val = nullptr;
  c.idt_n = 0;
}

void idtable_build(Context &c, int64_t n, const int64_t *particle_id)
{
  if (n < 0 || (n > 0 && !particle_id)) throw CudaError{HBTU_ERR_INVALID, "bad argument"};
  if (n > 0x7fffffff) throw CudaError{HBTU_ERR_UNSUPPORTED, "table larger than 2^31 entries"};
  idtable_clear(c);
  if (n == 0) return;
  c.staged = c.executed = false; // the arena is shared with a staged batch's rounds
  cudaStream_t st = c.stream;
  Arena &ar = c.arena;
  ar.reset();
  ar.reserve(n * 36 + (64 << 20));
  HBT_CUDA(cudaMalloc(&c.d_idt_key, sizeof(uint64_t) * (size_t)n));
  HBT_CUDA(cudaMalloc(&c.d_idt_val, sizeof(int) * (size_t)n));
  int64_t *d_ids = ar.alloc<int64_t>(n);
  HBT_CUDA(cudaMemcpyAsync(d_ids, particle_id, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, st));
  uint64_t *key_a = ar.alloc<uint64_t>(n);
  int *val_a = ar.alloc<int>(n);
  HBT_CUDA(cudaEventRecord(c.ev_exec[0], st));
  idtable_keys_kernel<<<igrid(n), kIB, 0, st>>>(d_ids, n, key_a, val_a);
  HBT_CHECK_LAUNCH();
  size_t tb = 0;
  HBT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key_a, c.d_idt_key, val_a, c.d_idt_val, n, 0, 64, st));
  void *tmp = ar.alloc<char>((int64_t)tb);
  HBT_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, key_a, c.d_idt_key, val_a, c.d_idt_val, n, 0, 64, st));
  HBT_CUDA(cudaEventRecord(c.ev_exec[1], st));
  HBT_CUDA(cudaStreamSynchronize(st));
  c.idt_n = n;
  std::memset(&c.stats, 0, sizeof(c.stats));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c.ev_exec[0], c.ev_exec[1]);
  c.stats.execute_ms = c.stats.other_ms = ms;
  c.stats.kernel_launches = 1 + 9;
  c.stats.h2d_bytes = n * 8;
}

void idtable_query(Context &c, int64_t nq, const int64_t *query_id, int64_t *index_out)
{
  if (nq < 0 || (nq > 0 && (!query_id || !index_out))) throw CudaError{HBTU_ERR_INVALID, "bad argument"};
  if (nq == 0) return;
  c.staged = c.executed = false;
  cudaStream_t st = c.stream;
  Arena &ar = c.arena;
  ar.reset();
  ar.reserve(nq * 16 + (1 << 20));
  int64_t *d_q = ar.alloc<int64_t>(nq), *d_out = ar.alloc<int64_t>(nq);
  HBT_CUDA(cudaMemcpyAsync(d_q, query_id, sizeof(int64_t) * (size_t)nq, cudaMemcpyHostToDevice, st));
  HBT_CUDA(cudaEventRecord(c.ev_exec[0], st));
  idtable_query_kernel<<<igrid(nq), kIB, 0, st>>>(c.d_idt_key, c.d_idt_val, c.idt_n, d_q, nq, d_out);
  HBT_CHECK_LAUNCH();
  HBT_CUDA(cudaEventRecord(c.ev_exec[1], st));
  HBT_CUDA(cudaMemcpyAsync(index_out, d_out, sizeof(int64_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
  HBT_CUDA(cudaStreamSynchronize(st));
  std::memset(&c.stats, 0, sizeof(c.stats));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c.ev_exec[0], c.ev_exec[1]);
  c.stats.execute_ms = c.stats.other_ms = ms;
  c.stats.kernel_launches = 1;
  c.stats.h2d_bytes = nq * 8;
  c.stats.d2h_bytes = nq * 8;
}

} // namespace hbt
