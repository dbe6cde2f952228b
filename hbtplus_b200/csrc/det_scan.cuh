// det_scan.cuh - run-to-run deterministic inclusive scan-by-key for floating-point values (sm_100a).
//
// cub::DeviceScan::*ByKey is a single-pass scan with decoupled look-back: how the partial sums of earlier tiles are
// grouped depends on the order in which tiles happen to publish them, so for fp64 addition (not associative) the last
// bit of a prefix can change from run to run.  The node moments of the octree and the cumulative masses of the density
// profiles must not: this is the classic three-kernel reduce-then-scan with a fixed combination tree,
//   K1  tile summaries   (segmented aggregate of every 1024-element tile, cub::BlockScan: fixed order)
//   K2  one block scans the tile summaries in order (running carry) -> carry-in of every tile
//   K3  every tile scans its elements starting from its carry-in and writes the prefixes.
// Reads the input twice (values come from a functor, typically 20-24 B per element) and writes it once.
#pragma once
#include <cub/block/block_scan.cuh>

#include "common.cuh"

namespace hbt
{

constexpr int kScanThreads = 256, kScanItems = 4, kScanTile = kScanThreads * kScanItems;

template <class V>
struct SegVal
{ // segmented-scan element: `head` = a segment starts at (or inside) this partial result
  V v;
  int head;
};
template <class V, class Plus>
struct SegOp
{
  Plus plus;
  __device__ __forceinline__ SegVal<V> operator()(const SegVal<V> &a, const SegVal<V> &b) const
  {
    SegVal<V> r;
    r.head = a.head | b.head;
    r.v = b.head ? b.v : plus(a.v, b.v);
    return r;
  }
};

template <class V, class ValFn, class KeyFn>
__device__ __forceinline__ void scan_load(int64_t base, int64_t n, ValFn val, KeyFn key, V zero, SegVal<V> (&x)[kScanItems])
{
  const int64_t i0 = base + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int j = 0; j < kScanItems; j++)
  {
    const int64_t i = i0 + j;
    if (i < n)
    {
      x[j].v = val(i);
      x[j].head = (i == 0 || key(i) != key(i - 1)) ? 1 : 0;
    }
    else
    { // padding behaves like a new, empty segment
      x[j].v = zero;
      x[j].head = 1;
    }
  }
}

template <class V, class Plus, class ValFn, class KeyFn>
__global__ void __launch_bounds__(kScanThreads) det_scan_summary_kernel(int64_t n, ValFn val, KeyFn key, V zero, SegVal<V> *__restrict__ summary)
{
  using BlockScan = cub::BlockScan<SegVal<V>, kScanThreads>;
  __shared__ typename BlockScan::TempStorage tmp;
  SegVal<V> x[kScanItems];
  scan_load<V>((int64_t)blockIdx.x * kScanTile, n, val, key, zero, x);
  SegVal<V> agg;
  BlockScan(tmp).InclusiveScan(x, x, SegOp<V, Plus>(), agg);
  if (threadIdx.x == 0) summary[blockIdx.x] = agg;
}

// one block: carry[t] = summaries 0..t-1 combined in order
template <class V, class Plus>
__global__ void __launch_bounds__(kScanThreads) det_scan_carry_kernel(int64_t ntiles, const SegVal<V> *__restrict__ summary, V zero,
                                                                       SegVal<V> *__restrict__ carry)
{
  using BlockScan = cub::BlockScan<SegVal<V>, kScanThreads>;
  __shared__ typename BlockScan::TempStorage tmp;
  __shared__ SegVal<V> running;
  if (threadIdx.x == 0)
  {
    running.v = zero;
    running.head = 1;
  }
  __syncthreads();
  SegOp<V, Plus> op{};
  for (int64_t base = 0; base < ntiles; base += kScanTile)
  {
    SegVal<V> x[kScanItems];
    const int64_t i0 = base + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
    for (int j = 0; j < kScanItems; j++)
    {
      if (i0 + j < ntiles) x[j] = summary[i0 + j];
      else { x[j].v = zero; x[j].head = 1; }
    }
    SegVal<V> agg;
    BlockScan(tmp).ExclusiveScan(x, x, op, agg); // x[j] = combination of this chunk's summaries before i0+j (undefined for the first)
    const SegVal<V> run = running;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kScanItems; j++)
      if (i0 + j < ntiles) carry[i0 + j] = (threadIdx.x == 0 && j == 0) ? run : op(run, x[j]);
    if (threadIdx.x == 0) running = op(run, agg);
    __syncthreads();
  }
}

template <class V, class Plus, class ValFn, class KeyFn>
__global__ void __launch_bounds__(kScanThreads) det_scan_apply_kernel(int64_t n, ValFn val, KeyFn key, V zero, const SegVal<V> *__restrict__ carry,
                                                                       V *__restrict__ out)
{
  using BlockScan = cub::BlockScan<SegVal<V>, kScanThreads>;
  __shared__ typename BlockScan::TempStorage tmp;
  SegVal<V> x[kScanItems];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  scan_load<V>(base, n, val, key, zero, x);
  SegOp<V, Plus> op{};
  BlockScan(tmp).InclusiveScan(x, x, op);
  const SegVal<V> c = carry[blockIdx.x];
  const int64_t i0 = base + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int j = 0; j < kScanItems; j++)
    if (i0 + j < n) out[i0 + j] = op(c, x[j]).v;
}

// out[i] = plus-combination of val(j) over the j <= i with key(j) == key(i) (keys are grouped), in a fixed order
template <class V, class Plus, class ValFn, class KeyFn>
inline void det_inclusive_scan_by_key(Arena &arena, cudaStream_t stream, int64_t n, ValFn val, KeyFn key, V zero, V *out, int64_t &launches)
{
  if (n <= 0) return;
  const int64_t ntiles = (n + kScanTile - 1) / kScanTile;
  SegVal<V> *summary = arena.alloc<SegVal<V>>(ntiles), *carry = arena.alloc<SegVal<V>>(ntiles);
  det_scan_summary_kernel<V, Plus><<<(unsigned)ntiles, kScanThreads, 0, stream>>>(n, val, key, zero, summary);
  HBT_CHECK_LAUNCH();
  det_scan_carry_kernel<V, Plus><<<1, kScanThreads, 0, stream>>>(ntiles, summary, zero, carry);
  HBT_CHECK_LAUNCH();
  det_scan_apply_kernel<V, Plus><<<(unsigned)ntiles, kScanThreads, 0, stream>>>(n, val, key, zero, carry, out);
  HBT_CHECK_LAUNCH();
  launches += 3;
}

} // namespace hbt
