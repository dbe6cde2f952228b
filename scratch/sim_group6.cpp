// Scratch: single-level classification with NB sub-boxes (slices) + in-place dfs with all G targets on MIXED subtrees.
#include <array>
#define main main_unused
#include "sim_group.cpp"
#undef main
struct Cost { double slices = 0, slices_sticky = 0; double A = 0, Aint = 0, iters = 0, nM = 0, dfs = 0, dfsint = 0, dfssteps = 0, tiles = 0; };
static float g_h2;
static void sub_dfs(int b, int e, const float *tg, int n, Cost &c)
{
  int Tp = (n + 31) / 32; std::vector<int> skip(n, b); int no = b; int tile_base = -1000;
  while (no < e)
  {
    if (no >= tile_base + 32) { tile_base = no; c.tiles++; }
    const Node &nd = nodes[no]; bool any_open = false; unsigned sl = 0;
    for (int q = 0; q < n; q++) { if (no < skip[q]) continue; sl |= 1u << (q / 32);
      float dx = nd.x - tg[4 * q], dy = nd.y - tg[4 * q + 1], dz = nd.z - tg[4 * q + 2]; float r2 = dx * dx + dy * dy + dz * dz;
      if (nd.lenq > r2) any_open = true; else { skip[q] = nd.end; c.dfsint++; } }
    c.dfs += 13.0 * Tp + 15; c.dfssteps++; c.slices += __builtin_popcount(sl);
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
  for (int G : {64, 128, 256})
    for (int NB : {1, 2, 4, 8})
    {
      if (G / NB < 16) continue;
      Cost c; int ngroups = 200; double tot = 0; int T = G / 32;
      for (int g = 0; g < ngroups; g++)
      {
        int64_t start = (int64_t)((double)g / ngroups * (n - G)); start -= start % G; const float *tg = &sp[4 * start];
        std::vector<std::array<float, 6>> bx(NB);
        for (int b = 0; b < NB; b++) { float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
          for (int k = b * (G / NB); k < (b + 1) * (G / NB); k++) for (int j = 0; j < 3; j++) { lo[j] = std::min(lo[j], tg[4 * k + j]); hi[j] = std::max(hi[j], tg[4 * k + j]); }
          for (int j = 0; j < 3; j++) { bx[b][j] = 0.5f * (lo[j] + hi[j]); bx[b][3 + j] = 0.5f * (hi[j] - lo[j]) * 1.00001f + 1e-30f; } }
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
              float r2min = 1e30f, r2max = 0; const float p[3] = {nd.x, nd.y, nd.z};
              for (int b = 0; b < NB; b++) { float mn = 0, mx = 0;
                for (int j = 0; j < 3; j++) { float d = std::fabs(p[j] - bx[b][j]); float dmin = std::max(0.f, d - bx[b][3 + j]); float dmax = d + bx[b][3 + j]; mn += dmin * dmin; mx += dmax * dmax; }
                r2min = std::min(r2min, mn); r2max = std::max(r2max, mx); }
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
        for (int no : mlist) { c.dfs += 30; sub_dfs(no, nodes[no].end, tg, G, c); }
        tot += G;
      }
      double inter = c.Aint + c.dfsint; double CIT = 80 + 25 * NB;
      double cost = c.A + c.iters * CIT + c.dfs + c.tiles * 40;
      printf("[active slices/step %.2f of %d] G=%3d boxes=%d: inter/target %.0f | inter share A %.2f dfs %.2f | cost share A %.2f iters %.2f dfs %.2f tiles %.2f | per warp: iters %.0f nM %.0f dfssteps %.0f tiles %.0f dfs lane-util %.2f | slots per 32 inter %.1f\n", c.slices / c.dfssteps, T, G, NB, inter / tot,
             c.Aint / inter, c.dfsint / inter, c.A / cost, c.iters * CIT / cost, c.dfs / cost, c.tiles * 40 / cost, c.iters / ngroups, c.nM / ngroups, c.dfssteps / ngroups, c.tiles / ngroups, c.dfsint / (c.dfssteps * G), cost / (inter / 32));
    }
  return 0;
}
