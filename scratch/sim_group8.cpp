// Scratch: like sim_group6 (single level, in-place dfs) but MIXED nodes are first evaluated per target: all open -> OPEN (children
// re-classified), all accept -> counted as FAR-like (cost of one per-target evaluation), else dfs over the subtree for the openers.
#include <array>
#define main main_unused
#include "sim_group.cpp"
#undef main
struct Cost { double A = 0, Aint = 0, iters = 0, nM = 0, nMallopen = 0, nMallacc = 0, Meval = 0, Mint = 0, dfs = 0, dfsint = 0, dfssteps = 0, tiles = 0; };
static float g_h2;
static void sub_dfs(int b, int e, const float *tg, int n, std::vector<int> &skip, Cost &c)
{
  int Tp = (n + 31) / 32; int no = b; int tile_base = -1000;
  while (no < e)
  {
    if (no >= tile_base + 32) { tile_base = no; c.tiles++; }
    const Node &nd = nodes[no]; bool any_open = false;
    for (int q = 0; q < n; q++) { if (no < skip[q]) continue;
      float dx = nd.x - tg[4 * q], dy = nd.y - tg[4 * q + 1], dz = nd.z - tg[4 * q + 2]; float r2 = dx * dx + dy * dy + dz * dz;
      if (nd.lenq > r2) any_open = true; else { skip[q] = nd.end; c.dfsint++; } }
    c.dfs += 10.0 * Tp + 15; c.dfssteps++;
    no = any_open ? no + 1 : nd.end;
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
  for (int mode : {0, 1})
  for (int G : {128})
  {
    Cost c; int ngroups = 200; double tot = 0; int T = G / 32;
    for (int g = 0; g < ngroups; g++)
    {
      int64_t start = (int64_t)((double)g / ngroups * (n - G)); start -= start % G; const float *tg = &sp[4 * start];
      float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
      for (int k = 0; k < G; k++) for (int j = 0; j < 3; j++) { lo[j] = std::min(lo[j], tg[4 * k + j]); hi[j] = std::max(hi[j], tg[4 * k + j]); }
      float cc[3], hw[3]; for (int j = 0; j < 3; j++) { cc[j] = 0.5f * (lo[j] + hi[j]); hw[j] = 0.5f * (hi[j] - lo[j]) * 1.00001f + 1e-30f; }
      std::vector<std::pair<int, int>> stack{{0, (int)nn}};
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
            else
            {
              c.nM++;
              std::vector<int> skip(G);
              int nopen = 0;
              for (int q = 0; q < G; q++) { float dx = nd.x - tg[4 * q], dy = nd.y - tg[4 * q + 1], dz = nd.z - tg[4 * q + 2]; float r2 = dx * dx + dy * dy + dz * dz;
                if (nd.lenq > r2) { nopen++; skip[q] = ch + 1; } else skip[q] = 0x7fffffff; }
              if (mode == 1)
              {
                c.Meval += 10.0 * T + 12;
                if (nopen == G) { c.nMallopen++; stack.push_back({ch + 1, nd.end}); }
                else { c.Mint += G - nopen; if (nopen == 0) c.nMallacc++; else { c.dfs += 10; sub_dfs(ch + 1, nd.end, tg, G, skip, c); } }
              }
              else
              {
                std::vector<int> sk(G, ch);
                c.dfs += 10; sub_dfs(ch, nd.end, tg, G, sk, c);
              }
            }
            if (isA) { c.Aint += G; c.A += (r2min < g_h2 ? 25.0 * T + 4 : 5.0 * T + 2); }
            ch = nx;
          }
          maxlen = std::max(maxlen, len);
        }
        c.iters += maxlen;
      }
      tot += G;
    }
    double inter = c.Aint + c.Mint + c.dfsint; double cost = c.A + c.Meval + c.iters * 105 + c.dfs + c.tiles * 40;
    printf("mode %d G=%d: inter/target %.0f | share A %.2f Mnode %.2f dfs %.2f | per warp nM %.0f (all-open %.0f all-accept %.0f) iters %.0f dfssteps %.0f | cost share A %.2f Meval %.2f iters %.2f dfs %.2f tiles %.2f | slots per 32 inter %.1f\n",
           mode, G, inter / tot, c.Aint / inter, c.Mint / inter, c.dfsint / inter, c.nM / ngroups, c.nMallopen / ngroups, c.nMallacc / ngroups, c.iters / ngroups, c.dfssteps / ngroups,
           c.A / cost, c.Meval / cost, c.iters * 105 / cost, c.dfs / cost, c.tiles * 40 / cost, cost / (inter / 32));
  }
  return 0;
}
