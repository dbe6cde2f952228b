// masked_emul.cpp - TEST-ONLY: runs hbtplus_b200/csrc/walk_masked.cuh (the core of the masked group walk) on the CPU,
// one warp = 32 fibers (warp_emul.h), over the pre-order node array built from tree_core.cuh, and returns per-target
// sums and accepted-interaction counts next to those of a scalar per-target walk with the same fp32 arithmetic.
// Built by tests/test_masked_walk_emul.py with g++; never linked into the product library.
#include "emul_tree.h"
#include "warp_emul.h"
using hbt::float_to_ordered;
using hbt::ordered_to_float;
static int g_max_ncs = 0;
static int64_t g_hist[16];
#define HBT_MASKED_STAT_ON 1
#define HBT_MASKED_STAT(what, n) do { if (wemu::g_cur == 0) g_hist[what] += (n); } while (0)
#define HBT_MASKED_TRACK(ncs) do { if ((ncs) > g_max_ncs) g_max_ncs = (ncs); } while (0)
#include "walk_masked.cuh"

#ifndef EMUL_STACK
#define EMUL_STACK 184
#endif
#ifndef EMUL_NP
#define EMUL_NP 2 // slice pairs per warp: 1 = 64-target groups, 2 = 128-target groups
#endif
typedef hbt::MaskedSmemT<EMUL_STACK, EMUL_NP> SmemT;
static constexpr int kT = 2 * EMUL_NP, kGroup = 32 * kT;
extern "C" int emul_group_size(void) { return kGroup; }

extern "C" int emul_masked_walk(const hbtu_params *p, int64_t n, const float *src, int64_t ntgt_in, int64_t group_stride, double *sum_masked, double *sum_scalar,
                                int64_t *acc_masked, int64_t *acc_scalar, int64_t *stats /*[4]: overflows, iterations, collectives, groups*/)
{ // targets = the sources in key order (a full evaluation), groups of kGroup (64 or 128) consecutive targets; only the first ntgt_in
  // targets, and of those only every group_stride-th group, are walked (the others keep their zeros)
  std::vector<Node> nodes;
  std::vector<float> sp;
  if (int rc = emul_build_nodes(p, n, src, nodes, sp, nullptr)) return rc;
  const int64_t nn = (int64_t)nodes.size();
  std::vector<float4> node_xm(nn + 64);
  std::vector<float2> node_aux(nn + 64);
  for (int64_t i = 0; i < nn; i++)
  {
    node_xm[i] = make_float4(nodes[i].x, nodes[i].y, nodes[i].z, nodes[i].m);
    node_aux[i] = make_float2(nodes[i].lenq, __int_as_float(nodes[i].end));
  }
  const float eps = (float)p->softening_halo, h = 2.8f * eps, h2 = h * h;
  const double hinv_d = 1.0 / (2.8 * (double)eps);
  const bool periodic = p->periodic_boundary_on;
  const float box = (float)p->box_size, half = (float)p->box_half;
  const int64_t ntgt = std::min<int64_t>(ntgt_in, n);
  if (const char *e = getenv("EMUL_LANE_ORDER_SEED")) wemu::g_shuffle = strtoull(e, nullptr, 10); // random lane order between rendez-vous
  int64_t overflows = 0, iters = 0, groups = 0;
  struct Guarded { uint64_t c0[8]; SmemT sm; uint64_t c1[8]; };
  static Guarded g;
  if (group_stride < 1) group_stride = 1;
  for (int64_t g0 = 0; g0 < ntgt; g0 += kGroup * group_stride)
  {
    const int n0 = (int)std::min<int64_t>(kGroup, ntgt - g0);
    for (int q = 0; q < 8; q++) g.c0[q] = g.c1[q] = 0x5a5a5a5a5a5a5a5aull;
    memset(&g.sm, 0xff, sizeof(g.sm));
    bool ok_all = true;
    unsigned vis_lane0 = 0;
    groups++;
    wemu::run_warp([&](int lane) {
      float px[kT], py[kT], pz[kT];
      bool valid[kT];
      const float *r = &sp[4 * g0];
      for (int k = 0; k < kT; k++)
      {
        const int j = lane + 32 * k;
        valid[k] = j < n0;
        const float *t = &sp[4 * (g0 + (valid[k] ? j : 0))];
        px[k] = t[0]; py[k] = t[1]; pz[k] = t[2];
        if (periodic)
        { // un-wrap towards the group's first target (walk_masked.cu does the same)
          const float ax = t[0] - r[0], ay = t[1] - r[1], az = t[2] - r[2];
          if (ax > half) px[k] = t[0] - box; else if (ax < -half) px[k] = t[0] + box;
          if (ay > half) py[k] = t[1] - box; else if (ay < -half) py[k] = t[1] + box;
          if (az > half) pz[k] = t[2] - box; else if (az < -half) pz[k] = t[2] + box;
        }
      }
      double accd[kT] = {};
      unsigned long long nacc = 0;
      unsigned n_acc = 0, n_vis = 0;
      bool ok;
      if (periodic)
        ok = hbt::masked_group_walk<true, true>(g.sm, lane, node_xm.data(), node_aux.data(), 0, (int)nn, px, py, pz, valid, n0, box, half, eps, accd, nacc, n_acc, n_vis);
      else
        ok = hbt::masked_group_walk<false, true>(g.sm, lane, node_xm.data(), node_aux.data(), 0, (int)nn, px, py, pz, valid, n0, box, half, eps, accd, nacc, n_acc, n_vis);
      if (!ok) ok_all = false;
      if (lane == 0) vis_lane0 = n_vis;
      for (int k = 0; k < kT; k++)
        if (valid[k])
        {
          sum_masked[g0 + lane + 32 * k] = accd[k];
          // the dense FAR part of the count is warp-uniform (nacc / n0 per target), the masked part per lane (not per slice):
          // report per-lane totals on slice 0 and the uniform part on every target
          acc_masked[g0 + lane + 32 * k] = (int64_t)(nacc / (unsigned long long)n0) + (k == 0 ? (int64_t)n_acc : 0);
        }
    });
    for (int q = 0; q < 8; q++)
      if (g.c0[q] != 0x5a5a5a5a5a5a5a5aull || g.c1[q] != 0x5a5a5a5a5a5a5a5aull) return -200;
    if (!ok_all) overflows++;
    if (getenv("EMUL_VERBOSE2")) fprintf(stderr, "group %ld: max stack %d iters %u ok %d\n", (long)g0, g_max_ncs, vis_lane0, (int)ok_all);
    g_max_ncs = 0;
    iters += vis_lane0;
  }
  // scalar per-target walk, same fp32 arithmetic (FMUL, FFMA, FFMA; criterion on the fp32 r^2)
  for (int64_t t = 0; t < ntgt; t++)
  {
    if ((t / kGroup) % group_stride != 0) continue;
    const float px = sp[4 * t], py = sp[4 * t + 1], pz = sp[4 * t + 2];
    double pot = 0;
    int64_t acc = 0, no = 0;
    float accf = 0.f;
    while (no < nn)
    {
      const Node &nd = nodes[no];
      float dx = nd.x - px, dy = nd.y - py, dz = nd.z - pz;
      if (periodic) { dx = hbt::nearest_f(dx, box, half); dy = hbt::nearest_f(dy, box, half); dz = hbt::nearest_f(dz, box, half); }
      const float r2 = std::fma(dz, dz, std::fma(dy, dy, dx * dx));
      if (nd.lenq > r2) { no++; continue; }
      no = nd.end;
      acc++;
      if (r2 >= h2) pot -= (double)(nd.m / std::sqrt(r2));
      else pot += (double)nd.m * hinv_d * hbt::spline_wp(r2, hinv_d);
    }
    (void)accf;
    sum_scalar[t] = pot;
    acc_scalar[t] = acc;
  }
  if (getenv("EMUL_VERBOSE")) fprintf(stderr, "per group: dense-ring nodes %.0f accept-all elements %.0f deciding elements (bare) %.0f pending chains %.0f open cells %.0f\n", (double)g_hist[0] / groups, (double)g_hist[1] / groups, (double)g_hist[2] / groups, (double)g_hist[3] / groups, (double)g_hist[4] / groups);
  if (getenv("EMUL_VERBOSE")) fprintf(stderr, "  single-slice pair-elements: accept-all %.0f deciding %.0f; targets inside the masks per pair-element: accept-all %.1f deciding %.1f (of 64)\n", (double)g_hist[5] / groups, (double)g_hist[6] / groups, (double)g_hist[7] / std::max<int64_t>(1, g_hist[1]), (double)g_hist[8] / std::max<int64_t>(1, g_hist[2]));
  for (int q = 0; q < 16; q++) g_hist[q] = 0;
  if (stats) { stats[0] = overflows; stats[1] = iters; stats[2] = (int64_t)wemu::g_ncollectives; stats[3] = groups; }
  return 0;
}
