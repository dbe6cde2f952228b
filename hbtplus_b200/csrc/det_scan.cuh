// det_scan.cuh - deterministic, composition-invariant inclusive prefix sums inside segments (sm_100a).
//
// cub::DeviceScan::*ByKey is a single-pass scan with decoupled look-back: how the partial sums of earlier tiles are
// grouped depends on the order in which tiles happen to publish them, so for fp64 addition (not associative) the last
// bit of a prefix can change from run to run.  The node moments of the octree and the cumulative masses of the density
// profiles must not - and they must not depend on WHERE in a batch a subhalo sits either (another sharding over GPUs has
// to give the same catalogue).  So the combination tree is fixed and aligned to each segment:
//   K1  tile sums        every segment is cut into 1024-element tiles counted from ITS first element; cub::BlockReduce
//   K2  one block per segment scans that segment's tile sums in order (running carry) -> carry-in of every tile
//   K3  every tile scans its elements (cub::BlockScan) starting from its carry-in and writes the prefixes.
// The tile table (tile_off[a] = number of tiles of the segments before a) comes from the host: O(nseg).
// Reads the input twice (values come from a functor, typically 20-24 B per element) and writes it once.
#pragma once
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>

#include <vector>

#include "common.cuh"

namespace hbt
{

constexpr int kScanThreads = 256, kScanItems = 4, kScanTile = kScanThreads * kScanItems;

// largest a in [0,n) with off[a] <= k (off non-decreasing; empty segments share their successor's offset and are skipped)
__device__ __forceinline__ int scan_find(const int *__restrict__ off, int n, int k)
{
  int lo = 0, hi = n;
  while (hi - lo > 1)
  {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= k) lo = mid; else hi = mid;
  }
  return lo;
}

template <class V, class ValFn>
__device__ __forceinline__ void scan_load(int64_t first, int64_t seg_end, ValFn val, V zero, V (&x)[kScanItems])
{
  const int64_t i0 = first + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int j = 0; j < kScanItems; j++) x[j] = (i0 + j < seg_end) ? val(i0 + j) : zero;
}

template <class V, class Plus, class ValFn>
__global__ void __launch_bounds__(kScanThreads) det_scan_sums_kernel(const int *__restrict__ seg_off, const int *__restrict__ tile_off, int nseg, ValFn val,
                                                                      V zero, V *__restrict__ tile_sum)
{
  using BlockReduce = cub::BlockReduce<V, kScanThreads>;
  __shared__ typename BlockReduce::TempStorage tmp;
  const int a = scan_find(tile_off, nseg, (int)blockIdx.x);
  const int64_t first = (int64_t)seg_off[a] + (int64_t)((int)blockIdx.x - tile_off[a]) * kScanTile;
  V x[kScanItems];
  scan_load<V>(first, seg_off[a + 1], val, zero, x);
  const V s = BlockReduce(tmp).Reduce(x, Plus());
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = s;
}

// one block per segment: carry[t] = sum of the segment's tile sums before tile t, combined in order
template <class V, class Plus>
__global__ void __launch_bounds__(kScanThreads) det_scan_carry_kernel(const int *__restrict__ tile_off, V zero, const V *__restrict__ tile_sum,
                                                                       V *__restrict__ carry)
{
  using BlockScan = cub::BlockScan<V, kScanThreads>;
  __shared__ typename BlockScan::TempStorage tmp;
  __shared__ V running;
  const int t0 = tile_off[blockIdx.x], t1 = tile_off[blockIdx.x + 1];
  if (t1 - t0 <= 1)
  {
    if (t1 - t0 == 1 && threadIdx.x == 0) carry[t0] = zero;
    return;
  }
  if (threadIdx.x == 0) running = zero;
  __syncthreads();
  Plus plus{};
  for (int base = t0; base < t1; base += kScanTile)
  {
    V x[kScanItems];
    const int i0 = base + (int)threadIdx.x * kScanItems;
#pragma unroll
    for (int j = 0; j < kScanItems; j++) x[j] = (i0 + j < t1) ? tile_sum[i0 + j] : zero;
    V agg;
    BlockScan(tmp).ExclusiveScan(x, x, zero, plus, agg);
    const V run = running;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kScanItems; j++)
      if (i0 + j < t1) carry[i0 + j] = plus(run, x[j]);
    if (threadIdx.x == 0) running = plus(run, agg);
    __syncthreads();
  }
}

template <class V, class Plus, class ValFn>
__global__ void __launch_bounds__(kScanThreads) det_scan_apply_kernel(const int *__restrict__ seg_off, const int *__restrict__ tile_off, int nseg, ValFn val,
                                                                       V zero, const V *__restrict__ carry, V *__restrict__ out)
{
  using BlockScan = cub::BlockScan<V, kScanThreads>;
  __shared__ typename BlockScan::TempStorage tmp;
  const int a = scan_find(tile_off, nseg, (int)blockIdx.x);
  const int64_t first = (int64_t)seg_off[a] + (int64_t)((int)blockIdx.x - tile_off[a]) * kScanTile;
  const int64_t seg_end = seg_off[a + 1];
  V x[kScanItems];
  scan_load<V>(first, seg_end, val, zero, x);
  Plus plus{};
  BlockScan(tmp).InclusiveScan(x, x, plus);
  const V c = carry[blockIdx.x];
  const int64_t i0 = first + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int j = 0; j < kScanItems; j++)
    if (i0 + j < seg_end) out[i0 + j] = plus(c, x[j]);
}

// host: tile table of segments given by element offsets seg_off[0..nseg]
inline int scan_tile_table(const int *seg_off, int nseg, std::vector<int> &tile_off)
{
  tile_off.resize((size_t)nseg + 1);
  int t = 0;
  for (int a = 0; a < nseg; a++)
  {
    tile_off[a] = t;
    t += (seg_off[a + 1] - seg_off[a] + kScanTile - 1) / kScanTile;
  }
  tile_off[nseg] = t;
  return t;
}

// out[i] = val(first element of i's segment) + ... + val(i), combined in a fixed order that depends only on the segment
//   d_seg_off / d_tile_off: device copies of the element offsets and of scan_tile_table(); ntiles = tile_off[nseg]
template <class V, class Plus, class ValFn>
inline void det_inclusive_scan_segments(Arena &arena, cudaStream_t stream, const int *d_seg_off, const int *d_tile_off, int nseg, int ntiles, ValFn val,
                                        V zero, V *out, int64_t &launches)
{
  if (ntiles <= 0 || nseg <= 0) return;
  V *tile_sum = arena.alloc<V>(ntiles), *carry = arena.alloc<V>(ntiles);
  det_scan_sums_kernel<V, Plus><<<(unsigned)ntiles, kScanThreads, 0, stream>>>(d_seg_off, d_tile_off, nseg, val, zero, tile_sum);
  HBT_CHECK_LAUNCH();
  det_scan_carry_kernel<V, Plus><<<(unsigned)nseg, kScanThreads, 0, stream>>>(d_tile_off, zero, tile_sum, carry);
  HBT_CHECK_LAUNCH();
  det_scan_apply_kernel<V, Plus><<<(unsigned)ntiles, kScanThreads, 0, stream>>>(d_seg_off, d_tile_off, nseg, val, zero, carry, out);
  HBT_CHECK_LAUNCH();
  launches += 3;
}

} // namespace hbt
