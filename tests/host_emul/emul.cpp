// emul.cpp - CPU emulation of the device octree index arithmetic (tree_core.cuh) + a scalar fp32
// walk over the pre-order node array.  TEST-ONLY: built by tests/test_tree_core.py with g++ to
// unit-test hbtplus_b200/csrc/tree_core.cuh without a GPU.  Never linked into the product library.
#include "emul_tree.h"
using namespace hbt;

extern "C" int emul_tree_potential(const hbtu_params *p, const hbtu_epoch *e, int64_t n, const float *src, int64_t ntgt,
                                   const float *tgt, const float *self_mass, double *out, int64_t *accepted, int64_t *visited,
                                   int64_t *ncells_out)
{
  std::vector<Node> nodes;
  std::vector<float> sp;
  if (int rc = emul_build_nodes(p, n, src, nodes, sp, ncells_out)) return rc;
  const int64_t nn = (int64_t)nodes.size();
  // scalar fp32 walk
  const float eps = (float)p->softening_halo, h = 2.8f * eps, h2 = h * h, hinv = 1.f / h;
  const bool periodic = p->periodic_boundary_on;
  const float box = (float)p->box_size, half = (float)p->box_half;
  for (int64_t t = 0; t < ntgt; t++)
  {
    float px = tgt[4 * t], py = tgt[4 * t + 1], pz = tgt[4 * t + 2];
    double pot = 0;
    int64_t acc = 0, vis = 0;
    int64_t no = 0;
    while (no < nn)
    {
      const Node &nd = nodes[no];
      float dx = nd.x - px, dy = nd.y - py, dz = nd.z - pz;
      if (periodic)
      {
        dx = dx > half ? dx - box : (dx < -half ? dx + box : dx);
        dy = dy > half ? dy - box : (dy < -half ? dy + box : dy);
        dz = dz > half ? dz - box : (dz < -half ? dz + box : dz);
      }
      float r2 = dx * dx + dy * dy + dz * dz;
      vis++;
      if (nd.lenq > r2) { no++; continue; }
      no = nd.end;
      acc++;
      if (r2 >= h2) pot -= nd.m / std::sqrt(r2);
      else
      {
        float u = std::sqrt(r2) * hinv, wp;
        if (u < 0.5f) wp = -2.8f + u * u * (5.333333333333f + u * u * (6.4f * u - 9.6f));
        else wp = -3.2f + 0.066666666667f / u + u * u * (10.666666666667f + u * (-16.0f + u * (9.6f - 2.133333333333f * u)));
        pot += nd.m * hinv * wp;
      }
    }
    double self = self_mass ? (double)self_mass[t] / (double)eps : 0.0;
    out[t] = (pot + self) * p->G / e->scale_factor;
    if (accepted) accepted[t] = acc;
    if (visited) visited[t] = vis;
  }
  return 0;
}
