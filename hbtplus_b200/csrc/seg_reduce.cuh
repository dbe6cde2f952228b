// seg_reduce.cuh - deterministic segmented sums of NV doubles per element (sm_100a).
//
// Elements 0..N-1 are grouped into contiguous segments (subhaloes).  fp64 addition is not associative, so sums that are
// combined by atomics depend on the order in which blocks happen to finish; the reductions of the unbinding path
// (frames, kinematics, inertia tensors) instead use a FIXED summation tree, so that the same input gives the same bits:
//
//   pass A (at the end of the producing kernel, one call per 256-thread block): the block's elements are cut into
//           pieces by segment.  A block that lies inside one segment reduces with a fixed shuffle/shared-memory tree;
//           a block that straddles segments lets NV threads walk its staged values in element order.  A piece that is a
//           whole segment is final and handed to `done`; the piece of a segment that began in an earlier block goes to
//           head[block], the piece of a segment that continues in a later block to tail[block].
//   pass B (seg_reduce_finish_block, one block per segment that spans several blocks):
//           total = tail[first block] + (head[] of the blocks in between, strided over the threads in a fixed order,
//           fixed block tree) + head[last block].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hbt
{

constexpr int kSegBlock = 256; // threads (= elements) per block of pass A

template <int NV>
struct SegPartials
{
  double *head; // [nblocks][NV] piece of the segment that was already open when the block started
  double *tail; // [nblocks][NV] piece of the segment that continues after the block (and did not start before it)
};

__device__ __forceinline__ double seg_warp_sum(double v)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Pass A.  Every thread of the block calls it (also threads past the end, with valid = false).
//   v                the thread's contribution (ignored when !valid)
//   seg              segment of the thread's element; seg_begin/seg_end: element range of that segment
//   done(seg, s)     called by ONE thread per finished segment with the NV sums
template <int NV, class DoneFn>
__device__ __forceinline__ void seg_reduce_block(const double (&v)[NV], bool valid, int seg, int64_t seg_begin, int64_t seg_end, int64_t n_total,
                                                 const SegPartials<NV> &part, DoneFn done)
{
  __shared__ double s_val[kSegBlock][NV];
  __shared__ int s_seg[kSegBlock];
  __shared__ int64_t s_rng[kSegBlock][2];
  __shared__ double s_red[NV][kSegBlock / 32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t block_begin = (int64_t)blockIdx.x * kSegBlock;
  const int64_t block_end = block_begin + kSegBlock < n_total ? block_begin + kSegBlock : n_total;
  const int nvalid = (int)(block_end - block_begin);
  s_seg[tid] = valid ? seg : -1;
  __syncthreads();
  const bool uniform = nvalid > 0 && s_seg[0] == s_seg[nvalid - 1]; // segments are contiguous
  if (uniform)
  {
#pragma unroll
    for (int i = 0; i < NV; i++)
    {
      const double s = seg_warp_sum(valid ? v[i] : 0.0);
      if (lane == 0) s_red[i][w] = s;
    }
    __syncthreads();
    if (tid == 0)
    {
      double s[NV];
#pragma unroll
      for (int i = 0; i < NV; i++)
      {
        double x = 0.0;
        for (int k = 0; k < kSegBlock / 32; k++) x += s_red[i][k];
        s[i] = x;
      }
      const bool before = seg_begin < block_begin, after = seg_end > block_end;
      if (!before && !after) done(seg, s);
      else
      {
        double *dst = (before ? part.head : part.tail) + (int64_t)blockIdx.x * NV;
#pragma unroll
        for (int i = 0; i < NV; i++) dst[i] = s[i];
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < NV; i++) s_val[tid][i] = valid ? v[i] : 0.0;
  s_rng[tid][0] = seg_begin;
  s_rng[tid][1] = seg_end;
  __syncthreads();
  if (tid < NV)
  { // component tid of every piece, in element order
    int i = 0;
    while (i < nvalid)
    {
      const int a = s_seg[i];
      double x = 0.0;
      int j = i;
      for (; j < nvalid && s_seg[j] == a; j++) x += s_val[j][tid];
      const bool before = s_rng[i][0] < block_begin, after = s_rng[i][1] > block_end;
      if (before) part.head[(int64_t)blockIdx.x * NV + tid] = x;
      else if (after) part.tail[(int64_t)blockIdx.x * NV + tid] = x;
      else s_red[tid][0] = x; // whole segment inside the block: collected below
      if (!before && !after)
      {
        // all NV components of this piece must reach `done` together: thread 0 gathers them
        __syncwarp((1u << NV) - 1u);
        if (tid == 0)
        {
          double s[NV];
#pragma unroll
          for (int c = 0; c < NV; c++) s[c] = s_red[c][0];
          done(a, s);
        }
        __syncwarp((1u << NV) - 1u);
      }
      i = j;
    }
  }
}

// Pass B: one 256-thread block per segment (blockIdx.x = segment).  seg_range(a, begin, end) gives the element range;
// done(a, s) as in pass A.  Thread j adds the middle blocks j, j+256, .. in order, then the fixed block tree.
template <int NV, class RangeFn, class DoneFn>
__device__ __forceinline__ void seg_reduce_finish_block(int a, int64_t n_total, const SegPartials<NV> &part, RangeFn seg_range, DoneFn done)
{
  __shared__ double s_red[NV][kSegBlock / 32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int64_t begin, end;
  seg_range(a, begin, end);
  if (end <= begin)
  {
    if (tid == 0)
    {
      double s[NV];
#pragma unroll
      for (int i = 0; i < NV; i++) s[i] = 0.0;
      done(a, s);
    }
    return;
  }
  const int64_t b0 = begin / kSegBlock, b1 = (end - 1) / kSegBlock;
  if (b0 == b1) return; // finished in pass A
  double s[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) s[i] = 0.0;
  for (int64_t b = b0 + 1 + tid; b < b1; b += kSegBlock)
#pragma unroll
    for (int i = 0; i < NV; i++) s[i] += part.head[b * NV + i];
#pragma unroll
  for (int i = 0; i < NV; i++)
  {
    const double x = seg_warp_sum(s[i]);
    if (lane == 0) s_red[i][w] = x;
  }
  __syncthreads();
  if (tid == 0)
  {
#pragma unroll
    for (int i = 0; i < NV; i++)
    {
      double x = 0.0;
      for (int k = 0; k < kSegBlock / 32; k++) x += s_red[i][k];
      s[i] = (part.tail[b0 * NV + i] + x) + part.head[b1 * NV + i];
    }
    done(a, s);
  }
}

} // namespace hbt
