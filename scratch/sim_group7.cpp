// Scratch: groups aligned to tree cells (maximal cells with <= GMAX particles) vs fixed runs of G sorted particles.
#define main main_unused
#include "sim_group.cpp"
#undef main
struct Cost { double A = 0, Aint = 0, iters = 0, nM = 0, dfs = 0, dfsint = 0, dfssteps = 0, tiles = 0, act = 0, opn = 0, tg = 0, warps = 0; };
static float g_h2;
static void sub_dfs(int b, int e, const float *tg, int n, Cost &c)
{
  int Tp = (n + 31) / 32; std::vector<int> skip(n, b); int no = b; int tile_base = -1000;
  while (no < e)
  {
    if (no >= tile_base + 32) { tile_base = no; c.tiles++; }
    const Node &nd = nodes[no]; bool any_open = false;
    for (int q = 0; q < n; q++) { if (no < skip[q]) continue; c.act++;
      float dx = nd.x - tg[4 * q], dy = nd.y - tg[4 * q + 1], dz = nd.z - tg[4 * q + 2]; float r2 = dx * dx + dy * dy + dz * dz;
      if (nd.lenq > r2) { any_open = true; c.opn++; } else { skip[q] = nd.end; c.dfsint++; } }
    c.dfs += 13.0 * Tp + 15; c.dfssteps++;
    no = any_open ? no + 1 : nd.end;
  }
}
static void walk_group(const float *tg, int G, Cost &c)
{
  int T = (G + 31) / 32;
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (int k = 0; k < G; k++) for (int j = 0; j < 3; j++) { lo[j] = std::min(lo[j], tg[4 * k + j]); hi[j] = std::max(hi[j], tg[4 * k + j]); }
  float cc[3], hw[3]; for (int j = 0; j < 3; j++) { cc[j] = 0.5f * (lo[j] + hi[j]); hw[j] = 0.5f * (hi[j] - lo[j]) * 1.00001f + 1e-30f; }
  std::vector<std::pair<int, int>> stack{{0, (int)nn}}; std::vector<int> mlist;
  while (!stack.empty())
  {
    int take = std::min<size_t>(32, stack.size());
    std::vector<std::pair<int, int>> batch(stack.end() - take, stack.end()); stack.resize(stack.size() - take);
    int maxlen = 0;
    for (auto pr : batch)
    {
      int ch = pr.first, len = 0;
      while (ch < pr.second)
      {
        len++; const Node &nd = nodes[ch]; int nx = nd.end;
        float r2min = 0, r2max = 0; const float p[3] = {nd.x, nd.y, nd.z};
        for (int j = 0; j < 3; j++) { float d = std::fabs(p[j] - cc[j]); float dmin = std::max(0.f, d - hw[j]); float dmax = d + hw[j]; r2min += dmin * dmin; r2max += dmax * dmax; }
        bool isA = false;
        if (nd.lenq == 0.f) isA = true;
        else if (nd.lenq > r2max * 1.00002f) stack.push_back({ch + 1, nd.end});
        else if (!(nd.lenq > r2min * 0.99998f)) isA = true;
        else mlist.push_back(ch);
        if (isA) { c.Aint += G; c.A += (r2min < g_h2 ? 25.0 * T + 4 : 8.0 * T + 2); }
        ch = nx;
      }
      maxlen = std::max(maxlen, len);
    }
    c.iters += maxlen;
  }
  c.nM += mlist.size();
  for (int no : mlist) { c.dfs += 10; sub_dfs(no, nodes[no].end, tg, G, c); }
  c.tg += G; c.warps++;
}
int main(int argc, char **argv)
{
  int64_t n = argc > 1 ? atoll(argv[1]) : 2000000; float eps = argc > 2 ? atof(argv[2]) : 4.8e-5f; double a = argc > 3 ? atof(argv[3]) : 0.03;
  std::mt19937_64 rng(12345); std::uniform_real_distribution<double> U(0, 1); std::normal_distribution<double> Nn(0, 1);
  std::vector<float> src(4 * n);
  for (int64_t i = 0; i < n; i++) { double u = U(rng) * 0.97, s = std::sqrt(u), r = a * s / (1 - s); double x = Nn(rng), y = Nn(rng), z = Nn(rng), q = r / std::sqrt(x * x + y * y + z * z);
    src[4 * i] = 50 + x * q; src[4 * i + 1] = 50 + y * q; src[4 * i + 2] = 50 + z * q; src[4 * i + 3] = 1e-6f; }
  build(src, n, 0.1 * eps, 0.45f * 0.45f);
  float h = 2.8f * eps; g_h2 = h * h;
  auto report = [&](const char *name, Cost &c) {
    double inter = c.Aint + c.dfsint; double cost = c.A + c.iters * 105 + c.dfs + c.tiles * 40;
    printf("%-28s targets/warp %.1f | inter share A %.2f | cost share A %.2f iters %.2f dfs %.2f tiles %.2f | dfs lane-slots: accept %.2f open %.2f idle %.2f | slots per 32 inter %.1f\n", name, c.tg / c.warps,
           c.Aint / inter, c.A / cost, c.iters * 105 / cost, c.dfs / cost, c.tiles * 40 / cost, c.dfsint / (c.dfssteps * (c.tg / c.warps)), c.opn / (c.dfssteps * (c.tg / c.warps)),
           1 - c.act / (c.dfssteps * (c.tg / c.warps)), cost / (inter / 32));
  };
  // fixed runs
  for (int G : {64, 128}) { Cost c; int ng = 200; for (int g = 0; g < ng; g++) { int64_t start = (int64_t)((double)g / ng * (n - G)); start -= start % G; walk_group(&sp[4 * start], G, c); } char nm[64]; snprintf(nm, 64, "fixed runs of %d", G); report(nm, c); }
  // cell-aligned: maximal cells with <= GMAX particles. particle index range of cell at node i: count leaves in [i, end)
  std::vector<int> leaf_prefix(nn + 1, 0);
  for (int64_t i = 0; i < nn; i++) leaf_prefix[i + 1] = leaf_prefix[i] + (nodes[i].lenq == 0.f);
  for (int GMAX : {64, 128, 256})
  {
    std::vector<std::pair<int, int>> groups; // (first sorted particle, count)
    int64_t i = 0;
    while (i < nn)
    {
      int cnt = leaf_prefix[nodes[i].end] - leaf_prefix[i];
      if (cnt <= GMAX) { groups.push_back({leaf_prefix[i], cnt}); i = nodes[i].end; } else i++;
    }
    // merge consecutive small groups (siblings) while total <= GMAX, as a packing heuristic
    std::vector<std::pair<int, int>> merged;
    for (auto g : groups) { if (!merged.empty() && merged.back().second + g.second <= GMAX && merged.back().first + merged.back().second == g.first) merged.back().second += g.second; else merged.push_back(g); }
    for (int pass = 0; pass < 2; pass++)
    {
      auto &gs = pass ? merged : groups;
      Cost c; int ng = 300;
      for (int g = 0; g < ng; g++) { auto gr = gs[(size_t)((double)g / ng * gs.size())]; if (gr.second == 0) continue; walk_group(&sp[4 * (size_t)gr.first], gr.second, c); }
      char nm[64]; snprintf(nm, 64, "cells<=%d%s (%zu groups)", GMAX, pass ? " merged" : "", gs.size()); report(nm, c);
    }
  }
  return 0;
}
