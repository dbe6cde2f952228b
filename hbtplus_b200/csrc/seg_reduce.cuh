// seg_reduce.cuh - deterministic, composition-invariant segmented sums of NV doubles per element (sm_100a).
//
// Elements are grouped into contiguous segments (subhaloes).  fp64 addition is not associative, so sums combined by
// atomics depend on the order in which blocks happen to finish, and sums over blocks aligned to the batch depend on where
// in the batch a subhalo sits.  The reductions of the unbinding path (frames, kinematics, inertia tensors) use a FIXED
// summation tree aligned to each segment, so the same subhalo gives the same bits on every run and in every batch:
//
//   pass A (inside the producing kernel): every segment is cut into 256-element chunks counted from ITS first element;
//           block b works on chunk (a, c) found through the chunk table, reduces its NV sums with a fixed shuffle /
//           shared-memory tree and stores them in partial[b].
//   pass B (seg_reduce_finish_block, one block per segment): thread j adds the segment's chunks j, j+256, .. in order,
//           then the same fixed block tree.
// The chunk table (chunk_off[a] = number of chunks of the segments before a) comes from the host: O(nseg).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace hbt
{

constexpr int kSegBlock = 256; // threads (= elements) per chunk

// host: chunk table for segment lengths len[0..nseg)
template <class LenFn>
inline int seg_chunk_table(int nseg, LenFn len, std::vector<int> &chunk_off)
{
  chunk_off.resize((size_t)nseg + 1);
  int64_t t = 0;
  for (int a = 0; a < nseg; a++)
  {
    chunk_off[a] = (int)t;
    t += ((int64_t)len(a) + kSegBlock - 1) / kSegBlock;
  }
  chunk_off[nseg] = (int)t;
  return (int)t;
}

// device: segment and chunk of block b (largest a with chunk_off[a] <= b; empty segments are skipped)
__device__ __forceinline__ void seg_chunk_of_block(const int *__restrict__ chunk_off, int nseg, int b, int &a, int &c)
{
  int lo = 0, hi = nseg;
  while (hi - lo > 1)
  {
    int mid = (lo + hi) >> 1;
    if (chunk_off[mid] <= b) lo = mid; else hi = mid;
  }
  a = lo;
  c = b - chunk_off[lo];
}

__device__ __forceinline__ double seg_warp_sum(double v)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fixed block tree: xor-shuffle inside warps, then warp sums added in warp order by thread 0.  Result valid in thread 0.
template <int NV>
__device__ __forceinline__ void seg_block_sum(double (&s)[NV])
{
  __shared__ double s_red[NV][kSegBlock / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++)
  {
    const double x = seg_warp_sum(s[i]);
    if (lane == 0) s_red[i][w] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int i = 0; i < NV; i++)
    {
      double x = 0.0;
      for (int k = 0; k < kSegBlock / 32; k++) x += s_red[i][k];
      s[i] = x;
    }
  }
}

// Pass A: every thread of the block calls it with its contribution (zeros when it has no element).
template <int NV>
__device__ __forceinline__ void seg_reduce_chunk(double (&v)[NV], double *__restrict__ partial)
{
  seg_block_sum<NV>(v);
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int i = 0; i < NV; i++) partial[(int64_t)blockIdx.x * NV + i] = v[i];
  }
}

// Pass B: one 256-thread block per segment a; done(a, s) is called by thread 0 with the NV sums.
template <int NV, class DoneFn>
__device__ __forceinline__ void seg_reduce_finish_block(int a, const int *__restrict__ chunk_off, const double *__restrict__ partial, DoneFn done)
{
  const int c0 = chunk_off[a], c1 = chunk_off[a + 1];
  double s[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) s[i] = 0.0;
  for (int c = c0 + (int)threadIdx.x; c < c1; c += kSegBlock)
#pragma unroll
    for (int i = 0; i < NV; i++) s[i] += partial[(int64_t)c * NV + i];
  seg_block_sum<NV>(s);
  if (threadIdx.x == 0) done(a, s);
}

} // namespace hbt
