// tree_build.cu - batched octree construction for all active subhaloes of a round (sm_100a).
//
// Replaces OctTree_t::Build (src/oct_tree.tpp:17-144, sequential insertion, serial on the CPU) and
// GravityTree_t::UpdateInternalNodes (src/gravity_tree.cpp:48-77, recursive post-order moments) with
// data-parallel passes over the concatenated source particles of every active subhalo:
//
//   bbox (K1)      segmented min/max by warp/block aggregation + ordered-uint atomics         HBM
//   keys (K1)      63-bit octal key by the reference's own double-precision descent            HBM
//   sort (K2)      CUB radix sort by key, then stable by segment                               HBM
//   cells (K3)     one thread per adjacent pair: gallop+bisect for the cell range              latency/L2
//   scan           cell counts (int, CUB)                                                       HBM
//   emit (K3/K4)   pre-order node array: particles, cell links (len^2/theta^2, end), then cell mass and centre of mass bottom-up,
//                  one launch per tree depth, with the reference's own (hierarchically rounded) summation   HBM
//
// Every pass is a streaming, coalesced read of 4..32 B per source particle; grids are sized in
// multiples of the SM count by the launch helpers.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "device_tree.cuh"

namespace hbt
{

static constexpr int kBlock = 256;

static inline int grid_for(int64_t n, int per_block = kBlock) { return n > 0 ? div_up(n, per_block) : 1; }

// ---------------------------------------------------------------------------------------------------
// bbox
// ---------------------------------------------------------------------------------------------------
__global__ void init_bbox_kernel(uint32_t *bbox, int nseg)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nseg * 6) bbox[i] = (i % 6 < 3) ? 0xffffffffu : 0u; // min slots / max slots
}

__device__ __forceinline__ float warp_min(float v)
{
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v)
{
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 4 elements per thread, 1024 per block.  ts_seg is non-decreasing, so a block whose first and last
// element share a segment is uniform: reduce in the block and issue 6 atomics; otherwise fall back to
// warp-uniform or per-lane atomics (only blocks that straddle segment boundaries).
__global__ void __launch_bounds__(kBlock) bbox_kernel(const float4 *__restrict__ tpos, const int *__restrict__ ts_seg, int S,
                                                       uint32_t *__restrict__ bbox)
{
  __shared__ float red[6][kBlock / 32];
  const int base = blockIdx.x * (kBlock * 4);
  const int last = min(base + kBlock * 4, S) - 1;
  const int seg_first = ts_seg[base], seg_last = ts_seg[last];
  const bool block_uniform = seg_first == seg_last;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int it = 0; it < 4; it++)
  {
    int k = base + it * kBlock + threadIdx.x;
    bool valid = k < S;
    float4 p = valid ? tpos[k] : make_float4(0, 0, 0, 0);
    if (block_uniform)
    {
      if (valid)
      {
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
      }
    }
    else
    {
      int seg = valid ? ts_seg[k] : -1;
      int seg0 = __shfl_sync(0xffffffffu, seg, 0);
      bool uni = __all_sync(0xffffffffu, seg == seg0) && seg0 >= 0;
      if (uni)
      {
        float a0 = warp_min(p.x), a1 = warp_min(p.y), a2 = warp_min(p.z);
        float b0 = warp_max(p.x), b1 = warp_max(p.y), b2 = warp_max(p.z);
        if ((threadIdx.x & 31) == 0)
        {
          uint32_t *bb = bbox + 6 * (int64_t)seg0;
          atomic_min_float(bb + 0, a0); atomic_min_float(bb + 1, a1); atomic_min_float(bb + 2, a2);
          atomic_max_float(bb + 3, b0); atomic_max_float(bb + 4, b1); atomic_max_float(bb + 5, b2);
        }
      }
      else if (valid)
      {
        uint32_t *bb = bbox + 6 * (int64_t)seg;
        atomic_min_float(bb + 0, p.x); atomic_min_float(bb + 1, p.y); atomic_min_float(bb + 2, p.z);
        atomic_max_float(bb + 3, p.x); atomic_max_float(bb + 4, p.y); atomic_max_float(bb + 5, p.z);
      }
    }
  }
  if (block_uniform)
  {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
      float a = warp_min(mn[j]), b = warp_max(mx[j]);
      if (l == 0) { red[j][w] = a; red[3 + j][w] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 6)
    {
      float v = red[threadIdx.x][0];
      for (int i = 1; i < kBlock / 32; i++) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][i]) : fmaxf(v, red[threadIdx.x][i]);
      uint32_t *bb = bbox + 6 * (int64_t)seg_first;
      if (threadIdx.x < 3) atomic_min_float(bb + threadIdx.x, v); else atomic_max_float(bb + threadIdx.x, v);
    }
  }
}

void launch_init_bbox(uint32_t *bbox, int nseg, cudaStream_t stream, LaunchStats &ls)
{
  init_bbox_kernel<<<grid_for((int64_t)nseg * 6), kBlock, 0, stream>>>(bbox, nseg);
  HBT_CHECK_LAUNCH();
  ls.launches++;
}
void launch_bbox(const float4 *tpos, const int *ts_seg, int S, uint32_t *bbox, cudaStream_t stream, LaunchStats &ls)
{
  if (S <= 0) return;
  bbox_kernel<<<grid_for(S, kBlock * 4), kBlock, 0, stream>>>(tpos, ts_seg, S, bbox);
  HBT_CHECK_LAUNCH();
  ls.launches++;
}

// root cube per segment: Len = max extent, Center = mid-range, in double (src/oct_tree.tpp:44-51)
__global__ void roots_kernel(const uint32_t *__restrict__ bbox, int nseg, double resolution, SegRoot *__restrict__ roots)
{
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nseg) return;
  const uint32_t *bb = bbox + 6 * (int64_t)a;
  double mn[3], mx[3];
  for (int j = 0; j < 3; j++) { mn[j] = (double)ordered_to_float(bb[j]); mx[j] = (double)ordered_to_float(bb[3 + j]); }
  double len = mx[0] - mn[0];
  for (int j = 1; j < 3; j++) if ((mx[j] - mn[j]) > len) len = mx[j] - mn[j];
  SegRoot r;
  r.cx = 0.5 * (mx[0] + mn[0]); r.cy = 0.5 * (mx[1] + mn[1]); r.cz = 0.5 * (mx[2] + mn[2]);
  r.len = len;
  r.halvings = count_halvings(len, resolution);
  r.pad = 0;
  roots[a] = r;
}

__global__ void __launch_bounds__(kBlock) keys_kernel(const float4 *__restrict__ tpos, const int *__restrict__ ts_seg, int S,
                                                       const SegRoot *__restrict__ roots, uint64_t *__restrict__ key, int *__restrict__ val)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  float4 p = tpos[k];
  key[k] = morton_key(p.x, p.y, p.z, roots[ts_seg[k]]);
  val[k] = k;
}

__global__ void __launch_bounds__(kBlock) gather_seg_kernel(const int *__restrict__ perm, const int *__restrict__ ts_seg, int S,
                                                             int *__restrict__ segkey, int *__restrict__ iota)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  segkey[k] = ts_seg[perm[k]];
  iota[k] = k;
}

// final sorted arrays from the (optional) second sort's index `order2` into the first sort's outputs
__global__ void __launch_bounds__(kBlock) sorted_gather_kernel(const uint64_t *__restrict__ key1, const int *__restrict__ perm1,
                                                                const int *__restrict__ order2, const float4 *__restrict__ tpos, int S,
                                                                uint64_t *__restrict__ skey, int *__restrict__ sperm, float4 *__restrict__ spos)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  int j = order2 ? order2[k] : k;
  int src = perm1[j];
  skey[k] = key1[j];
  sperm[k] = src;
  spos[k] = tpos[src];
}

// ---------------------------------------------------------------------------------------------------
// cells
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) pairs_kernel(const uint64_t *__restrict__ skey, const int *__restrict__ ts_seg,
                                                        const int *__restrict__ tree_off, int S, int2 *__restrict__ cell_lr,
                                                        int8_t *__restrict__ cell_depth, uint32_t *__restrict__ depthmask)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  int8_t depth = -1;
  if (k + 1 < S)
  {
    int a = ts_seg[k];
    if (ts_seg[k + 1] == a)
    {
      CellRange c = cell_of_pair(skey, k, tree_off[a], tree_off[a + 1]);
      if (c.is_rep)
      {
        depth = (int8_t)c.depth;
        cell_lr[k] = make_int2(c.l, c.r);
        atomicOr(&depthmask[c.l], 1u << c.depth);
      }
    }
  }
  cell_depth[k] = depth;
}

struct PopcOp
{
  __host__ __device__ int operator()(uint32_t m) const
  {
#if defined(__CUDA_ARCH__)
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
  }
};

__global__ void __launch_bounds__(kBlock) emit_particles_kernel(const float4 *__restrict__ spos, const int *__restrict__ cellcount, int S,
                                                                 float4 *__restrict__ node_xm, float2 *__restrict__ node_aux)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  int64_t pos = particle_node_pos(k, cellcount);
  node_xm[pos] = spos[k];
  node_aux[pos] = make_float2(0.f, __int_as_float((int)(pos + 1)));
}

// cells, pass 1: geometry and links of every cell (len^2/theta^2 and `end`), and where it sits in the pre-order array
__global__ void __launch_bounds__(kBlock) emit_cell_links_kernel(const int2 *__restrict__ cell_lr, const int8_t *__restrict__ cell_depth,
                                                                  const uint32_t *__restrict__ depthmask, const int *__restrict__ cellcount,
                                                                  const int *__restrict__ ts_seg, const SegRoot *__restrict__ roots, float theta2, int S,
                                                                  float2 *__restrict__ node_aux, int *__restrict__ cell_pos)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  int depth = cell_depth[k];
  if (depth < 0) return;
  CellRange c;
  int2 lr = cell_lr[k];
  c.l = lr.x; c.r = lr.y; c.depth = depth; c.is_rep = 1;
  float lenf = cell_len(roots[ts_seg[k]], depth);
  // reference MAC: (float)(len*len) > r2*theta2 (src/gravity_tree.cpp:135); stored as len^2/theta^2
  float lenq = __fdiv_rn(__fmul_rn(lenf, lenf), theta2);
  int64_t pos = cell_node_pos(c, cellcount, depthmask);
  node_aux[pos] = make_float2(lenq, __int_as_float((int)cell_node_end(c, cellcount)));
  cell_pos[k] = (int)pos;
}

// cells, pass 2 (one launch per depth, deepest first): mass and centre of mass exactly as the reference computes them
// (GravityTree_t::UpdateInternalNodes / ProcessNode / FillNodeCenter, src/gravity_tree.cpp:18-77): a cell sums its CHILDREN's
// stored HBTReal = float mass and centre (a particle's position, a child cell's already rounded centre) in son order, in double,
// and rounds the results to float once more.  Rounding therefore accumulates level by level; summing the particles directly
// (one rounding) gives centres that differ from the reference's by up to an ulp of the COORDINATE - 4e-6 at x ~ 50, which is
// 1e-3 of the node distances inside an AqA2-sized subhalo and showed up as a 1e-5 mean potential difference (and with it a
// few E ~ 0 membership flips per thousand subhaloes).  With the reference's own summation the potentials agree to 1e-8.
// Children have more shared key digits than their parent, so depth order is dependency order; cells of one depth are independent.
__global__ void __launch_bounds__(kBlock) cell_moments_level_kernel(const int8_t *__restrict__ cell_depth, const int *__restrict__ cell_pos, int S,
                                                                     int depth, const float2 *__restrict__ node_aux, float4 *__restrict__ node_xm)
{
  const int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (k0 >= S) return;
  // 16 depth bytes per thread (S is padded to a multiple of 16 by the caller's allocation)
  const uint4 d16 = *reinterpret_cast<const uint4 *>(cell_depth + k0);
  const unsigned w[4] = {d16.x, d16.y, d16.z, d16.w};
#pragma unroll
  for (int q = 0; q < 16; q++)
  {
    const int d = (int)(int8_t)((w[q >> 2] >> (8 * (q & 3))) & 0xffu);
    if (d != depth || k0 + q >= S) continue;
    const int pos = cell_pos[k0 + q];
    const int end = __float_as_int(node_aux[pos].y);
    double M = 0., cx = 0., cy = 0., cz = 0.;
    for (int c = pos + 1; c < end;)
    {
      const float4 xm = node_xm[c];
      const int nxt = __float_as_int(node_aux[c].y);
      const double m = (double)xm.w;
      M = __dadd_rn(M, m);
      cx = __dadd_rn(cx, __dmul_rn((double)xm.x, m)); // VectorAdd(CoM, pos, thismass): the product is exact in double
      cy = __dadd_rn(cy, __dmul_rn((double)xm.y, m));
      cz = __dadd_rn(cz, __dmul_rn((double)xm.z, m));
      c = nxt;
    }
    node_xm[pos] = make_float4((float)__ddiv_rn(cx, M), (float)__ddiv_rn(cy, M), (float)__ddiv_rn(cz, M), (float)M);
  }
}

// ---------------------------------------------------------------------------------------------------
int64_t tree_arena_bytes(int64_t S, int64_t nseg)
{
  // keys/vals double buffers, sorted copies, cell arrays, scans, nodes (2S), CUB temp (~S*8 worst)
  int64_t per = 8 * 2 + 4 * 2 + 4 * 3 + 8 + 4 + 16 + 8 + 1 + 4 + 4 + 32 + 2 * 16 + 2 * 8 + 16;
  return per * (S + 1024) + nseg * (int64_t)(sizeof(SegRoot) + 64) + (64 << 20);
}

void build_trees(TreeArrays &t, Arena &arena, const DevConfig &cfg, cudaStream_t stream, LaunchStats &ls)
{
  const int S = t.S, nseg = t.nseg;
  if (S <= 0) return;
  t.roots = arena.alloc<SegRoot>(nseg);
  roots_kernel<<<grid_for(nseg), kBlock, 0, stream>>>(t.bbox, nseg, (double)cfg.resolution, t.roots);
  HBT_CHECK_LAUNCH();
  ls.launches++;

  // keys + sort ---------------------------------------------------------------------------------------
  uint64_t *key_a = arena.alloc<uint64_t>(S), *key_b = arena.alloc<uint64_t>(S);
  int *val_a = arena.alloc<int>(S), *val_b = arena.alloc<int>(S);
  keys_kernel<<<grid_for(S), kBlock, 0, stream>>>(t.tpos, t.ts_seg, S, t.roots, key_a, val_a);
  HBT_CHECK_LAUNCH();
  ls.launches++;
  cub::DoubleBuffer<uint64_t> dkeys(key_a, key_b);
  cub::DoubleBuffer<int> dvals(val_a, val_b);
  size_t tmp_bytes = 0;
  HBT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dkeys, dvals, S, 0, 63, stream));
  void *tmp = arena.alloc<char>((int64_t)tmp_bytes);
  HBT_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dkeys, dvals, S, 0, 63, stream));
  ls.launches += 9; // onesweep: histogram + 8 digit passes
  const uint64_t *key1 = dkeys.Current();
  const int *perm1 = dvals.Current();
  const int *order2 = nullptr;
  if (nseg > 1)
  {
    int *seg_a = arena.alloc<int>(S), *seg_b = arena.alloc<int>(S);
    int *ord_a = arena.alloc<int>(S), *ord_b = arena.alloc<int>(S);
    gather_seg_kernel<<<grid_for(S), kBlock, 0, stream>>>(perm1, t.ts_seg, S, seg_a, ord_a);
    HBT_CHECK_LAUNCH();
    ls.launches++;
    int bits = 1;
    while ((1ll << bits) < nseg) bits++;
    cub::DoubleBuffer<int> dseg(seg_a, seg_b), dord(ord_a, ord_b);
    size_t tmp2 = 0;
    HBT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, dseg, dord, S, 0, bits, stream));
    void *tmpb = arena.alloc<char>((int64_t)tmp2);
    HBT_CUDA(cub::DeviceRadixSort::SortPairs(tmpb, tmp2, dseg, dord, S, 0, bits, stream));
    ls.launches += 1 + (bits + 7) / 8;
    order2 = dord.Current();
  }
  t.skey = arena.alloc<uint64_t>(S);
  t.sperm = arena.alloc<int>(S);
  t.spos = arena.alloc<float4>(S);
  sorted_gather_kernel<<<grid_for(S), kBlock, 0, stream>>>(key1, perm1, order2, t.tpos, S, t.skey, t.sperm, t.spos);
  HBT_CHECK_LAUNCH();
  ls.launches++;

  // cells ---------------------------------------------------------------------------------------------
  t.cell_lr = arena.alloc<int2>(S);
  t.cell_depth = arena.alloc<int8_t>(S + 16); // read 16 bytes at a time by cell_moments_level_kernel
  t.depthmask = arena.alloc<uint32_t>(S);
  t.cellcount = arena.alloc<int>(S);
  HBT_CUDA(cudaMemsetAsync(t.depthmask, 0, sizeof(uint32_t) * (size_t)S, stream));
  pairs_kernel<<<grid_for(S), kBlock, 0, stream>>>(t.skey, t.ts_seg, t.tree_off, S, t.cell_lr, t.cell_depth, t.depthmask);
  HBT_CHECK_LAUNCH();
  ls.launches++;
  {
    auto in = thrust::make_transform_iterator(t.depthmask, PopcOp());
    size_t b = 0;
    HBT_CUDA(cub::DeviceScan::InclusiveSum(nullptr, b, in, t.cellcount, S, stream));
    void *tm = arena.alloc<char>((int64_t)b);
    HBT_CUDA(cub::DeviceScan::InclusiveSum(tm, b, in, t.cellcount, S, stream));
    ls.launches += 2;
  }
  // nodes ---------------------------------------------------------------------------------------------
  t.node_xm = arena.alloc<float4>(2 * (int64_t)S + 64); // +pad: the walk stages 32 nodes without a bounds check
  t.node_aux = arena.alloc<float2>(2 * (int64_t)S + 64);
  int *cell_pos = arena.alloc<int>(S);
  emit_particles_kernel<<<grid_for(S), kBlock, 0, stream>>>(t.spos, t.cellcount, S, t.node_xm, t.node_aux);
  HBT_CHECK_LAUNCH();
  emit_cell_links_kernel<<<grid_for(S), kBlock, 0, stream>>>(t.cell_lr, t.cell_depth, t.depthmask, t.cellcount, t.ts_seg, t.roots, cfg.theta2, S,
                                                             t.node_aux, cell_pos);
  HBT_CHECK_LAUNCH();
  ls.launches += 2;
  for (int depth = kMaxDepth; depth >= 0; depth--)
  { // the reference's bottom-up moments, level by level (see cell_moments_level_kernel)
    cell_moments_level_kernel<<<grid_for(div_up(S, 16)), kBlock, 0, stream>>>(t.cell_depth, cell_pos, S, depth, t.node_aux, t.node_xm);
    HBT_CHECK_LAUNCH();
  }
  ls.launches += kMaxDepth + 1;
}

} // namespace hbt
