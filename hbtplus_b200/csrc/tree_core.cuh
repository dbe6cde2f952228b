// tree_core.cuh - per-element logic of the device octree, written as host+device inline functions
// so that the same code is (a) called from the CUDA kernels in tree_build.cu / walk.cu and
// (b) compiled with plain g++ by tests/host_emul (a CPU emulation used ONLY to unit-test the
// index arithmetic without a GPU; it is never linked into libhbtunbind.so).
//
// The tree reproduces the geometry of the reference's sequential-insertion octree
// (src/oct_tree.tpp:17-144, src/gravity_tree.cpp:18-77) without inserting sequentially:
//
//   * root cube  : bbox of raw coordinates, Len = max extent, Center = mid-range (oct_tree.tpp:30-51)
//   * octant rule: digit bit set iff pos > center, centre moved by +-lenhalf in double (oct_tree.tpp:69-95)
//                  -> 21 octal digits = a 63-bit key; sorting by key is the reference's son order 0..7
//   * cells      : a reference cell exists for every key prefix shared by >= 2 particles.  A chain of
//                  single-child cells has identical mass/CoM and decreasing len, and the walk accepts the
//                  first chain member that passes len^2 <= r^2 theta^2, so the chain is equivalent to its
//                  DEEPEST member: cells here are the prefixes at which >= 2 particles actually branch.
//   * len        : root len = Len; a child halves it only while the parent's len >= TreeNodeResolution
//                  (gravity_tree.cpp:37-40), stored as HBTReal=float
//   * layout     : one array of nodes (particles AND cells) in depth-first pre-order; `end` is the index
//                  of the first node after the subtree, i.e. the reference's `sibling` link, and
//                  index+1 is its `nextnode` link (gravity_tree.h / oct_tree.h:26-42).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HBT_HD __host__ __device__ __forceinline__
#else
#define HBT_HD inline
#endif

namespace hbt
{

constexpr int kMaxDepth = 21; // octal digits in a 63-bit key

struct SegRoot
{ // per-tree root cube (one per subhalo tree of a round)
  double cx, cy, cz; // Center
  double len;        // Len (max bbox extent)
  int halvings;      // K: number of halvings until len < TreeNodeResolution
  int pad;
};

// 63-bit key by the reference's own descent (oct_tree.tpp:60-95), in double like the reference.
HBT_HD uint64_t morton_key(float x, float y, float z, const SegRoot &root)
{
  double c0 = root.cx, c1 = root.cy, c2 = root.cz, lh = 0.5 * root.len;
  uint64_t key = 0;
#pragma unroll 1
  for (int lev = 0; lev < kMaxDepth; lev++)
  {
    lh *= 0.5;
    unsigned sub = 0;
    if (x > c0) { c0 += lh; sub |= 1u; } else c0 -= lh;
    if (y > c1) { c1 += lh; sub |= 2u; } else c1 -= lh;
    if (z > c2) { c2 += lh; sub |= 4u; } else c2 -= lh;
    key = (key << 3) | sub;
  }
  return key;
}

HBT_HD int clz64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}
HBT_HD int popc32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}

// number of leading octal digits two keys share (0..21)
HBT_HD int common_digits(uint64_t a, uint64_t b) { return (clz64(a ^ b) - 1) / 3; }

// len of a cell whose particles share `depth` digits (gravity_tree.cpp:37-40,50-51), as HBTReal
HBT_HD float cell_len(const SegRoot &root, int depth)
{
  int k = depth < root.halvings ? depth : root.halvings;
  double len = root.len;
  for (int i = 0; i < k; i++) len *= 0.5; // exact: power-of-two scaling, same values as len/2. chains
  return (float)len;
}

// K = number of halvings applied before len drops below the resolution (children of a cell with
// len < TreeNodeResolution keep its len)
HBT_HD int count_halvings(double len, double resolution)
{
  int k = 0;
  while (len >= resolution && k < 1100) { len *= 0.5; k++; }
  return k;
}

struct CellRange
{
  int l, r;   // inclusive particle range (sorted order, global indices)
  int depth;  // shared digits
  int is_rep; // this adjacent pair is the representative (leftmost pair of that depth) of its cell
};

// Cell spanned by the adjacent sorted pair (i, i+1), both inside the segment [seg_lo, seg_hi).
// keys[] sorted ascending within the segment.  O(log range) key loads (gallop + bisect).
HBT_HD CellRange cell_of_pair(const uint64_t *__restrict__ keys, int i, int seg_lo, int seg_hi)
{
  CellRange c;
  const uint64_t ki = keys[i];
  const int D = common_digits(ki, keys[i + 1]);
  c.depth = D;
  // left bound: smallest l with common_digits(keys[l], ki) >= D  (monotone in l)
  int lo = i, step = 1;
  while (lo - step >= seg_lo && common_digits(keys[lo - step], ki) >= D) { lo -= step; step <<= 1; }
  // answer in [max(seg_lo, lo-step+1), lo]; bisect the remaining gap
  int bad = lo - step; if (bad < seg_lo - 1) bad = seg_lo - 1; // keys[bad] known (or boundary) to fail
  while (lo - bad > 1)
  {
    int mid = bad + ((lo - bad) >> 1);
    if (common_digits(keys[mid], ki) >= D) lo = mid; else bad = mid;
  }
  c.l = lo;
  // right bound: largest r with common_digits(ki, keys[r]) >= D
  int hi = i + 1; step = 1;
  while (hi + step < seg_hi && common_digits(ki, keys[hi + step]) >= D) { hi += step; step <<= 1; }
  int badr = hi + step; if (badr > seg_hi) badr = seg_hi;
  while (badr - hi > 1)
  {
    int mid = hi + ((badr - hi) >> 1);
    if (common_digits(ki, keys[mid]) >= D) hi = mid; else badr = mid;
  }
  c.r = hi;
  c.is_rep = (c.l == i) || (common_digits(keys[c.l], ki) > D);
  return c;
}

// Pre-order positions.  cellcount_incl[j] = number of cells whose left end is <= j (inclusive scan of
// popc(depthmask)), taken over the whole concatenated array so that positions are global.
HBT_HD int64_t particle_node_pos(int j, const int *__restrict__ cellcount_incl) { return (int64_t)j + cellcount_incl[j]; }
HBT_HD int64_t cell_node_pos(const CellRange &c, const int *__restrict__ cellcount_incl, const uint32_t *__restrict__ depthmask)
{
  int before = c.l > 0 ? cellcount_incl[c.l - 1] : 0;
  int shallower = popc32(depthmask[c.l] & ((1u << c.depth) - 1u));
  return (int64_t)c.l + before + shallower;
}
HBT_HD int64_t cell_node_end(const CellRange &c, const int *__restrict__ cellcount_incl) { return (int64_t)c.r + 1 + cellcount_incl[c.r]; }

// Counter-based sampling permutation (sampled mode): positions j of subhalo `sub` are ordered by this 40-bit key
// (ties by j).  Replaces std::random_shuffle on libc rand() (src/subhalo_unbind.cpp:302); oracle/hbt_oracle.c
// restates the same formula for its shuffle mode 1.
HBT_HD uint64_t shuffle_key(uint64_t seed, uint64_t sub, uint64_t j)
{
  uint64_t z = seed ^ (0x9E3779B97F4A7C15ULL * (sub + 1)) ^ (j * 0xBF58476D1CE4E5B9ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return z >> 24;
}

// order-preserving float <-> uint32 maps (for atomic min/max and radix sorting by energy)
HBT_HD uint32_t float_to_ordered(float f)
{
#if defined(__CUDA_ARCH__)
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } cv; cv.f = f; uint32_t b = cv.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
HBT_HD float ordered_to_float(uint32_t u)
{
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } cv; cv.u = b; return cv.f;
#endif
}

} // namespace hbt
