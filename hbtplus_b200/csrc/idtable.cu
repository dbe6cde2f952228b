// idtable.cu - particle Id -> index lookup on the device, sm_100a.
//
// Replaces the on-node part of ParticleExchanger_t::QueryParticles (src/particle_exchanger.h:196-211):
//   MappedIndexTable_t::Fill        src/hash.tpp:18-32         sort the (Id, index) pairs of the snapshot
//   MappedIndexTable_t::GetIndices  src/hash_remote.tpp:9-88   sort the queries, batch binary search
// The reference's answer is a pure function of the (Id, index) set - "the index of this Id, or NullParticleId" - so the
// sorted table is an implementation detail.  Here: an open-addressing hash table in HBM, 16-byte slots
// {Id, index, pad}, load factor 1/2 (capacity 2n, linear probing: 1.5 expected probes for a present Id, 2.5 for an absent
// one, and two slots share a 32-byte DRAM sector).
//   build : memset + ONE kernel; an entry claims its slot with a 64-bit atomicCAS on the Id and publishes its index with
//           atomicMin (of equal Ids the lowest index answers, as documented in the ABI: deterministic whatever the race)
//   query : ONE kernel, one thread per query, one 16-byte load per probe; the queries need no sort and no order restoration
// HBM-bound integer work dominated by random 32-byte sectors: 8 B read + one sector read-modify-write per entry,
// 8 B read + 8 B written + ~1.5 sectors per query (round 1: a 28-level binary search per query, 35 ms per 1.8e8 queries).
// The empty-slot marker is the all-ones Id = SpecialConst::NullParticleId (-1, src/datatypes.h:87); a table entry that
// really carries Id -1 is kept in a side slot so that even that query answers like the reference.
#include <algorithm>
#include <cstring>

#include "context.cuh"

namespace hbt
{

static constexpr int kIB = 256;
static inline unsigned igrid(int64_t n) { return (unsigned)(n > 0 ? div_up(n, kIB) : 1); }

struct __align__(16) IdSlot
{
  unsigned long long id; // raw bits of the signed Id; kEmptyId = free
  unsigned index;        // 0xffffffff until published
  unsigned pad;
};
static constexpr unsigned long long kEmptyId = ~0ull;

__device__ __forceinline__ uint64_t id_slot(uint64_t id, uint64_t cap)
{ // splitmix64 finaliser, then multiply-shift onto [0, cap)
  uint64_t x = id;
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return __umul64hi(x, cap);
}

__global__ void __launch_bounds__(kIB) idtable_insert_kernel(const int64_t *__restrict__ ids, int64_t n, IdSlot *__restrict__ slots, uint64_t cap)
{
  const int64_t i = (int64_t)blockIdx.x * kIB + threadIdx.x;
  if (i >= n) return;
  const unsigned long long id = (unsigned long long)ids[i];
  if (id == kEmptyId)
  { // Id -1: the side slot behind the table
    atomicMin(&slots[cap].index, (unsigned)i);
    return;
  }
  uint64_t s = id_slot(id, cap);
  for (;;)
  {
    const unsigned long long old = atomicCAS(&slots[s].id, kEmptyId, id);
    if (old == kEmptyId || old == id)
    {
      atomicMin(&slots[s].index, (unsigned)i);
      return;
    }
    if (++s == cap) s = 0;
  }
}

__global__ void __launch_bounds__(kIB) idtable_query_kernel(const IdSlot *__restrict__ slots, uint64_t cap, const int64_t *__restrict__ query, int64_t nq,
                                                             int64_t *__restrict__ out)
{
  const int64_t q = (int64_t)blockIdx.x * kIB + threadIdx.x;
  if (q >= nq) return;
  const unsigned long long id = (unsigned long long)query[q];
  unsigned found = 0xffffffffu;
  if (id == kEmptyId)
    found = slots[cap].index;
  else
  {
    uint64_t s = id_slot(id, cap);
    for (;;)
    {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(&slots[s]));
      const unsigned long long k = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
      if (k == id)
      {
        found = v.z;
        break;
      }
      if (k == kEmptyId) break; // absent (src/hash_remote.tpp:77-83: NullParticleId)
      if (++s == cap) s = 0;
    }
  }
  out[q] = found == 0xffffffffu ? -1 : (int64_t)found;
}

void idtable_clear(Context &c)
{
  cudaFree(c.d_idt_slots);
  c.d_idt_slots = nullptr;
  c.idt_cap = 0;
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
  ar.reserve(n * 8 + (1 << 20));
  const uint64_t cap = (uint64_t)std::max<int64_t>(2 * n, 64);
  HBT_CUDA(cudaMalloc(&c.d_idt_slots, sizeof(IdSlot) * (size_t)(cap + 1)));
  IdSlot *slots = static_cast<IdSlot *>(c.d_idt_slots);
  int64_t *d_ids = ar.alloc<int64_t>(n);
  HBT_CUDA(cudaMemcpyAsync(d_ids, particle_id, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, st));
  HBT_CUDA(cudaEventRecord(c.ev_exec[0], st));
  HBT_CUDA(cudaMemsetAsync(slots, 0xff, sizeof(IdSlot) * (size_t)(cap + 1), st));
  idtable_insert_kernel<<<igrid(n), kIB, 0, st>>>(d_ids, n, slots, cap);
  HBT_CHECK_LAUNCH();
  HBT_CUDA(cudaEventRecord(c.ev_exec[1], st));
  HBT_CUDA(cudaStreamSynchronize(st));
  c.idt_n = n;
  c.idt_cap = (int64_t)cap;
  std::memset(&c.stats, 0, sizeof(c.stats));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c.ev_exec[0], c.ev_exec[1]);
  c.stats.execute_ms = c.stats.other_ms = ms;
  c.stats.kernel_launches = 1;
  c.stats.h2d_bytes = n * 8;
}

void idtable_query(Context &c, int64_t nq, const int64_t *query_id, int64_t *index_out)
{
  if (nq < 0 || (nq > 0 && (!query_id || !index_out))) throw CudaError{HBTU_ERR_INVALID, "bad argument"};
  if (nq == 0) return;
  if (c.idt_n == 0)
  { // empty table: nothing is found (and there is no device table to probe)
    for (int64_t q = 0; q < nq; q++) index_out[q] = -1;
    std::memset(&c.stats, 0, sizeof(c.stats));
    return;
  }
  c.staged = c.executed = false;
  cudaStream_t st = c.stream;
  Arena &ar = c.arena;
  ar.reset();
  ar.reserve(nq * 16 + (1 << 20));
  int64_t *d_q = ar.alloc<int64_t>(nq), *d_out = ar.alloc<int64_t>(nq);
  HBT_CUDA(cudaMemcpyAsync(d_q, query_id, sizeof(int64_t) * (size_t)nq, cudaMemcpyHostToDevice, st));
  HBT_CUDA(cudaEventRecord(c.ev_exec[0], st));
  idtable_query_kernel<<<igrid(nq), kIB, 0, st>>>(static_cast<const IdSlot *>(c.d_idt_slots), (uint64_t)c.idt_cap, d_q, nq, d_out);
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
