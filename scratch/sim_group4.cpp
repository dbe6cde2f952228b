// Scratch: single-frame masked design: chains carry opener masks; classification against the top-level bbox.
#include <deque>
#define main main_unused
#include "sim_group.cpp"
#undef main
struct Cost { double A = 0, Am = 0, Aint = 0, M = 0, Mint = 0, iters = 0, nM = 0, dfs = 0, dfsint = 0, dfssteps = 0, ndfs = 0, Amnodes = 0, Amuse = 0; };
static float g_h2; static double C_ITER = 120, C_DFSSETUP = 150;
static void sub_dfs(int b, int e, const std::vector<int> &idx, const float *tg, Cost &c)
{
  int n = idx.size(); int Tp = (n + 31) / 32; if (Tp == 3) Tp = 4;
  std::vector<int> skip(n, b); int no = b; c.ndfs++;
  while (no < e)
  {
    const Node &nd = nodes[no]; bool any_open = false;
    for (int q = 0; q < n; q++) { if (no < skip[q]) continue; int k = idx[q];
      float dx = nd.x - tg[4 * k], dy = nd.y - tg[4 * k + 1], dz = nd.z - tg[4 * k + 2]; float r2 = dx * dx + dy * dy + dz * dz;
      if (nd.lenq > r2) any_open = true; else { skip[q] = nd.end; c.dfsint++; } }
    c.dfs += 13.0 * Tp + 15; c.dfssteps++;
    no = any_open ? no + 1 : nd.end;
  }
}
struct Chain { int first, pend; int mask; }; // mask = index into masks (-1 = full)
static void walk(const float *tg, int G, Cost &c, int minmask)
{
  int T = G / 32;
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (int k = 0; k < G; k++) for (int j = 0; j < 3; j++) { lo[j] = std::min(lo[j], tg[4 * k + j]); hi[j] = std::max(hi[j], tg[4 * k + j]); }
  float cc[3], hw[3]; for (int j = 0; j < 3; j++) { cc[j] = 0.5f * (lo[j] + hi[j]); hw[j] = 0.5f * (hi[j] - lo[j]) * 1.00001f + 1e-30f; }
  std::deque<std::vector<int>> masks; std::vector<int> full(G); for (int k = 0; k < G; k++) full[k] = k;
  std::vector<Chain> stack; stack.push_back({0, (int)nn, -1});
  while (!stack.empty())
  {
    int take = std::min<size_t>(32, stack.size());
    std::vector<Chain> batch(stack.end() - take, stack.end()); stack.resize(stack.size() - take);
    int maxlen = 0;
    for (auto ch0 : batch)
    {
      const std::vector<int> &S = ch0.mask < 0 ? full : masks[ch0.mask];
      int ch = ch0.first, len = 0;
      while (ch < ch0.pend)
      {
        len++; const Node &nd = nodes[ch]; int nx = nd.end;
        float r2min = 0, r2max = 0; const float p[3] = {nd.x, nd.y, nd.z};
        for (int j = 0; j < 3; j++) { float d = std::fabs(p[j] - cc[j]); float dmin = std::max(0.f, d - hw[j]); float dmax = d + hw[j]; r2min += dmin * dmin; r2max += dmax * dmax; }
        bool isA = false;
        if (nd.lenq == 0.f) isA = true;
        else if (nd.lenq > r2max * 1.00002f) stack.push_back({ch + 1, nd.end, ch0.mask});
        else if (!(nd.lenq > r2min * 0.99998f)) isA = true;
        else
        {
          c.nM++; c.M += 15.0 * T + 30;
          std::vector<int> op;
          for (int k : S) { float dx = nd.x - tg[4 * k], dy = nd.y - tg[4 * k + 1], dz = nd.z - tg[4 * k + 2]; float r2 = dx * dx + dy * dy + dz * dz; if (nd.lenq > r2) op.push_back(k); else c.Mint++; }
          if (!op.empty())
          {
            if ((int)op.size() == G) stack.push_back({ch + 1, nd.end, -1});
            else if ((int)op.size() > minmask) { masks.push_back(op); stack.push_back({ch + 1, nd.end, (int)masks.size() - 1}); }
            else { c.dfs += C_DFSSETUP; sub_dfs(ch + 1, nd.end, op, tg, c); }
          }
        }
        if (isA) { c.Aint += S.size(); if (ch0.mask < 0) c.A += (r2min < g_h2 ? 25.0 * T + 4 : 8.0 * T + 2); else { c.Am += (r2min < g_h2 ? 27.0 * T + 4 : 10.0 * T + 4); c.Amnodes++; c.Amuse += (double)S.size() / G; } }
        ch = nx;
      }
      maxlen = std::max(maxlen, len);
    }
    c.iters += maxlen;
  }
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
  for (int G : {64, 128, 256})
    for (int minmask : {0, 4, 8, 16, 32, 64})
    {
      if (minmask >= G) continue;
      Cost c; int ngroups = 200; double tot = 0;
      for (int g = 0; g < ngroups; g++) { int64_t start = (int64_t)((double)g / ngroups * (n - G)); start -= start % G; walk(&sp[4 * start], G, c, minmask); tot += G; }
      double inter = c.Aint + c.Mint + c.dfsint;
      double cost = c.A + c.Am + c.M + c.iters * C_ITER + c.dfs;
      printf("G=%3d minmask=%3d: inter/target %.0f | cost share A %.2f Amasked %.2f (%.0f nodes, util %.2f) M %.2f iters %.2f dfs %.2f | per warp: iters %.0f nM %.0f ndfs %.0f dfssteps %.0f | slots per 32 inter %.1f\n", G, minmask, inter / tot,
             c.A / cost, c.Am / cost, c.Amnodes / ngroups, c.Amuse / c.Amnodes, c.M / cost, c.iters * C_ITER / cost, c.dfs / cost, c.iters / ngroups, c.nM / ngroups, c.ndfs / ngroups, c.dfssteps / ngroups, cost / (inter / 32));
    }
  return 0;
}
