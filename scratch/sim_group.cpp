// Scratch: CPU simulation of walk strategies on the device tree layout, to estimate instruction budgets before
// writing CUDA.  Not part of the product or tests.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "tree_core.cuh"
using namespace hbt;
struct Node { float x, y, z, m, lenq; int end; };
static std::vector<Node> nodes;
static std::vector<float> sp; // sorted particles
static int64_t nn;

static void build(const std::vector<float> &src, int64_t n, double resolution, float theta2)
{
  float mn[3], mx[3];
  for (int j = 0; j < 3; j++) mn[j] = mx[j] = src[j];
  for (int64_t i = 1; i < n; i++) for (int j = 0; j < 3; j++) { mn[j] = std::min(mn[j], src[4 * i + j]); mx[j] = std::max(mx[j], src[4 * i + j]); }
  SegRoot root; double len = (double)mx[0] - mn[0];
  for (int j = 1; j < 3; j++) len = std::max(len, (double)mx[j] - mn[j]);
  root.cx = 0.5 * ((double)mx[0] + mn[0]); root.cy = 0.5 * ((double)mx[1] + mn[1]); root.cz = 0.5 * ((double)mx[2] + mn[2]);
  root.len = len; root.halvings = count_halvings(len, resolution);
  std::vector<uint64_t> key(n); std::vector<int> perm(n);
  for (int64_t i = 0; i < n; i++) { key[i] = morton_key(src[4 * i], src[4 * i + 1], src[4 * i + 2], root); perm[i] = i; }
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
  std::vector<uint64_t> skey(n); sp.resize(4 * n);
  for (int64_t i = 0; i < n; i++) { skey[i] = key[perm[i]]; memcpy(&sp[4 * i], &src[4 * perm[i]], 16); }
  std::vector<CellRange> cells(n); std::vector<uint32_t> mask(n, 0);
  for (int64_t i = 0; i + 1 < n; i++) { cells[i] = cell_of_pair(skey.data(), (int)i, 0, (int)n); if (cells[i].is_rep) mask[cells[i].l] |= 1u << cells[i].depth; }
  std::vector<int> cinc(n); int run = 0;
  for (int64_t i = 0; i < n; i++) { run += popc32(mask[i]); cinc[i] = run; }
  nn = n + run;
  std::vector<double> S(4 * (n + 1), 0.0);
  for (int64_t i = 0; i < n; i++) { double m = sp[4 * i + 3]; S[4 * (i + 1)] = S[4 * i] + m; for (int j = 0; j < 3; j++) S[4 * (i + 1) + 1 + j] = S[4 * i + 1 + j] + m * ((double)sp[4 * i + j] - (&root.cx)[j]); }
  nodes.assign(nn, Node());
  for (int64_t i = 0; i < n; i++) { int64_t pos = particle_node_pos((int)i, cinc.data()); nodes[pos] = Node{sp[4 * i], sp[4 * i + 1], sp[4 * i + 2], sp[4 * i + 3], 0.f, (int)(pos + 1)}; }
  for (int64_t i = 0; i + 1 < n; i++) if (cells[i].is_rep) {
    const CellRange &c = cells[i]; int64_t pos = cell_node_pos(c, cinc.data(), mask.data());
    double M = S[4 * (c.r + 1)] - S[4 * c.l]; float lenf = cell_len(root, c.depth); float lenq = (lenf * lenf) / theta2;
    nodes[pos] = Node{(float)(root.cx + (S[4 * (c.r + 1) + 1] - S[4 * c.l + 1]) / M), (float)(root.cy + (S[4 * (c.r + 1) + 2] - S[4 * c.l + 2]) / M), (float)(root.cz + (S[4 * (c.r + 1) + 3] - S[4 * c.l + 3]) / M), (float)M, lenq, (int)cell_node_end(c, cinc.data())};
  }
}
struct Stats { double steps = 0, lane_acc = 0, tiles = 0, nA = 0, nC = 0, nO = 0, nM = 0, mSteps = 0, mAcc = 0, mTiles = 0, kids = 0, mStepsUniA = 0, mStepsUniO = 0; };

// current algorithm on range [b,e): per-lane DFS with skip; returns warp steps
static void lane_dfs(int b, int e, const float *tg, int G, Stats &st, bool inM)
{
  std::vector<int> skip(G, b);
  int no = b; int tile_base = -1000;
  while (no < e)
  {
    if (no >= tile_base + 32 || no < tile_base) { tile_base = no; (inM ? st.mTiles : st.tiles)++; }
    const Node &nd = nodes[no];
    bool any_open = false; int nact = 0, nacc = 0, nopen = 0;
    for (int k = 0; k < G; k++)
    {
      if (no < skip[k]) continue;
      nact++;
      float dx = nd.x - tg[4 * k], dy = nd.y - tg[4 * k + 1], dz = nd.z - tg[4 * k + 2];
      float r2 = dx * dx + dy * dy + dz * dz;
      if (nd.lenq > r2) { any_open = true; nopen++; }
      else { skip[k] = nd.end; nacc++; }
    }
    if (inM) { st.mSteps++; st.mAcc += nacc; if (nacc == nact) st.mStepsUniA++; if (nopen == nact) st.mStepsUniO++; }
    else { st.steps++; st.lane_acc += nacc; }
    no = any_open ? no + 1 : nd.end;
  }
}
static void group_walk(const float *tg, int G, float h2, Stats &st)
{
  float lo[3], hi[3];
  for (int j = 0; j < 3; j++) { lo[j] = hi[j] = tg[j]; }
  for (int k = 1; k < G; k++) for (int j = 0; j < 3; j++) { lo[j] = std::min(lo[j], tg[4 * k + j]); hi[j] = std::max(hi[j], tg[4 * k + j]); }
  float c[3], hw[3];
  for (int j = 0; j < 3; j++) { c[j] = 0.5f * (lo[j] + hi[j]); hw[j] = 0.5f * (hi[j] - lo[j]) * 1.00001f + 1e-30f; }
  std::vector<std::pair<int, int>> stack; // ranges of children to classify: (first child, parent end)
  // root
  auto classify = [&](int no) {
    const Node &nd = nodes[no];
    float r2min = 0, r2max = 0;
    const float p[3] = {nd.x, nd.y, nd.z};
    for (int j = 0; j < 3; j++) { float d = std::fabs(p[j] - c[j]); float dmin = std::max(0.f, d - hw[j]); float dmax = d + hw[j]; r2min += dmin * dmin; r2max += dmax * dmax; }
    if (nd.lenq == 0.f) { if (r2min < h2) st.nC++; else st.nA++; return; }
    if (nd.lenq > r2max * 1.00001f) { st.nO++; stack.push_back({no + 1, nd.end}); }
    else if (!(nd.lenq > r2min * 0.99999f)) { if (r2min < h2) st.nC++; else st.nA++; }
    else { st.nM++; lane_dfs(no, nd.end, tg, G, st, true); }
  };
  classify(0);
  while (!stack.empty())
  {
    auto pr = stack.back(); stack.pop_back();
    int ch = pr.first;
    while (ch < pr.second) { st.kids++; int nx = nodes[ch].end; classify(ch); ch = nx; }
  }
}
int main(int argc, char **argv)
{
  int64_t n = argc > 1 ? atoll(argv[1]) : 2000000;
  float eps = argc > 2 ? atof(argv[2]) : 4.8e-5f;
  double a = argc > 3 ? atof(argv[3]) : 0.03;
  std::mt19937_64 rng(12345); std::uniform_real_distribution<double> U(0, 1); std::normal_distribution<double> Nn(0, 1);
  std::vector<float> src(4 * n);
  for (int64_t i = 0; i < n; i++) {
    double u = U(rng) * 0.97, s = std::sqrt(u), r = a * s / (1 - s);
    double x = Nn(rng), y = Nn(rng), z = Nn(rng), q = r / std::sqrt(x * x + y * y + z * z);
    src[4 * i] = 50 + x * q; src[4 * i + 1] = 50 + y * q; src[4 * i + 2] = 50 + z * q; src[4 * i + 3] = 1e-6f;
  }
  float theta2 = 0.45f * 0.45f;
  build(src, n, 0.1 * eps, theta2);
  printf("n=%ld nodes=%ld\n", (long)n, (long)nn);
  float h = 2.8f * eps, h2 = h * h;
  for (int G : {32, 64, 128})
  {
    Stats cur, neu;
    int ngroups = 300; double tot_tg = 0;
    for (int g = 0; g < ngroups; g++)
    {
      int64_t start = (int64_t)((double)g / ngroups * (n - G)); start -= start % G;
      lane_dfs(0, (int)nn, &sp[4 * start], G, cur, false);
      group_walk(&sp[4 * start], G, h2, neu);
      tot_tg += G;
    }
    int T = G / 32;
    double acc_per_t = cur.lane_acc / tot_tg;
    double cost_cur = cur.steps * (13.0 * T + 15) + cur.tiles * 40;
    double cost_new = neu.nA * (8.0 * T + 2) + neu.nC * (25.0 * T + 4) + neu.mSteps * (13.0 * T + 15) + neu.mTiles * 40 + neu.kids * 6.0 + ngroups * 200;
    printf("G=%d: accepted/target %.0f | current: steps/warp %.0f accept-frac %.3f tiles/step %.2f slots/interaction %.1f\n", G, acc_per_t, cur.steps / ngroups,
           cur.lane_acc / (cur.steps * G), cur.tiles / cur.steps, cost_cur / (cur.lane_acc / 32));
    printf("      new: per warp A %.0f C %.0f O %.0f M %.0f kids %.0f | Msteps %.0f (uniA %.2f uniO %.2f) Macc-frac %.3f | interactions: A %.2f C %.2f M %.2f | slots/interaction %.1f\n", neu.nA / ngroups,
           neu.nC / ngroups, neu.nO / ngroups, neu.nM / ngroups, neu.kids / ngroups, neu.mSteps / ngroups, neu.mStepsUniA / neu.mSteps, neu.mStepsUniO / neu.mSteps, neu.mAcc / (neu.mSteps * G),
           neu.nA * G / cur.lane_acc, neu.nC * G / cur.lane_acc, neu.mAcc / cur.lane_acc, cost_new / (cur.lane_acc / 32));
  }
  return 0;
}
