// emul_tree.h - TEST-ONLY: builds the device tree's pre-order node array on the CPU from tree_core.cuh (the same
// per-element logic the CUDA kernels call).  Shared by emul.cpp and masked_emul.cpp.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "tree_core.cuh"
#include "hbt_unbind.h"

struct Node { float x, y, z, m, lenq; int end; };

// returns 0, or a negative code when a pre-order slot is written twice / never; sp = sources in key order
static int emul_build_nodes(const hbtu_params *p, int64_t n, const float *src, std::vector<Node> &nodes, std::vector<float> &sp, int64_t *ncells_out)
{
  using namespace hbt;
  // bbox -> root (oct_tree.tpp:30-51)
  float mn[3], mx[3];
  for (int j = 0; j < 3; j++) mn[j] = mx[j] = src[j];
  for (int64_t i = 1; i < n; i++)
    for (int j = 0; j < 3; j++) { mn[j] = std::min(mn[j], src[4 * i + j]); mx[j] = std::max(mx[j], src[4 * i + j]); }
  SegRoot root;
  double len = (double)mx[0] - mn[0];
  for (int j = 1; j < 3; j++) len = std::max(len, (double)mx[j] - mn[j]);
  root.cx = 0.5 * ((double)mx[0] + mn[0]); root.cy = 0.5 * ((double)mx[1] + mn[1]); root.cz = 0.5 * ((double)mx[2] + mn[2]);
  root.len = len;
  root.halvings = count_halvings(len, p->tree_node_resolution);
  std::vector<uint64_t> key(n);
  std::vector<int> perm(n);
  for (int64_t i = 0; i < n; i++) { key[i] = morton_key(src[4 * i], src[4 * i + 1], src[4 * i + 2], root); perm[i] = i; }
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
  std::vector<uint64_t> skey(n);
  sp.assign(4 * n, 0.f);
  for (int64_t i = 0; i < n; i++) { skey[i] = key[perm[i]]; memcpy(&sp[4 * i], &src[4 * perm[i]], 16); }
  // cells
  std::vector<CellRange> cells(n);
  std::vector<uint32_t> mask(n, 0);
  for (int64_t i = 0; i + 1 < n; i++) { cells[i] = cell_of_pair(skey.data(), (int)i, 0, (int)n); if (cells[i].is_rep) mask[cells[i].l] |= 1u << cells[i].depth; }
  std::vector<int> cinc(n);
  int run = 0;
  for (int64_t i = 0; i < n; i++) { run += popc32(mask[i]); cinc[i] = run; }
  int64_t nn = n + run;
  if (ncells_out) *ncells_out = run;
  std::vector<double> S(4 * (n + 1), 0.0); // prefix sums of m, m*(x-c)
  for (int64_t i = 0; i < n; i++)
  {
    double m = sp[4 * i + 3];
    S[4 * (i + 1)] = S[4 * i] + m;
    S[4 * (i + 1) + 1] = S[4 * i + 1] + m * ((double)sp[4 * i] - root.cx);
    S[4 * (i + 1) + 2] = S[4 * i + 2] + m * ((double)sp[4 * i + 1] - root.cy);
    S[4 * (i + 1) + 3] = S[4 * i + 3] + m * ((double)sp[4 * i + 2] - root.cz);
  }
  nodes.assign(nn, Node());
  std::vector<char> written(nn, 0);
  float theta2 = (float)p->tree_node_open_angle_square;
  for (int64_t i = 0; i < n; i++)
  {
    int64_t pos = particle_node_pos((int)i, cinc.data());
    if (written[pos]) return -100;
    written[pos] = 1;
    nodes[pos] = Node{sp[4 * i], sp[4 * i + 1], sp[4 * i + 2], sp[4 * i + 3], 0.f, (int)(pos + 1)};
  }
  for (int64_t i = 0; i + 1 < n; i++)
    if (cells[i].is_rep)
    {
      const CellRange &c = cells[i];
      int64_t pos = cell_node_pos(c, cinc.data(), mask.data());
      if (pos < 0 || pos >= nn || written[pos]) return -101;
      written[pos] = 1;
      double M = S[4 * (c.r + 1)] - S[4 * c.l];
      float lenf = cell_len(root, c.depth);
      float lenq = (lenf * lenf) / theta2;
      nodes[pos] = Node{(float)(root.cx + (S[4 * (c.r + 1) + 1] - S[4 * c.l + 1]) / M), (float)(root.cy + (S[4 * (c.r + 1) + 2] - S[4 * c.l + 2]) / M),
                        (float)(root.cz + (S[4 * (c.r + 1) + 3] - S[4 * c.l + 3]) / M), (float)M, lenq, (int)cell_node_end(c, cinc.data())};
    }
  for (int64_t i = 0; i < nn; i++) if (!written[i]) return -102;
  // Node masses and centres as the REFERENCE computes them (src/gravity_tree.cpp:18-77): a cell sums its CHILDREN's stored
  // HBTReal = float mass and centre (not the particles), in son order, in double, and rounds to float once more - so rounding
  // accumulates level by level.  Children have more shared digits than their parent: deepest cells first.
  if (!getenv("EMUL_FLAT_MOMENTS"))
    for (int d = kMaxDepth; d >= 0; d--)
      for (int64_t i = 0; i + 1 < n; i++)
        if (cells[i].is_rep && cells[i].depth == d)
        {
          const int64_t pos = cell_node_pos(cells[i], cinc.data(), mask.data());
          const int end = nodes[pos].end;
          double M = 0., cx = 0., cy = 0., cz = 0.;
          for (int64_t c = pos + 1; c < end; c = nodes[c].end)
          {
            const double m = (double)nodes[c].m;
            M += m;
            cx += (double)nodes[c].x * m;
            cy += (double)nodes[c].y * m;
            cz += (double)nodes[c].z * m;
          }
          nodes[pos].x = (float)(cx / M);
          nodes[pos].y = (float)(cy / M);
          nodes[pos].z = (float)(cz / M);
          nodes[pos].m = (float)M;
        }
  return 0;
}
