// Scratch: two-level static scheme. Level 1: group G (T slices of 32). Level 2: each 32-slice re-classifies level-1 MIXED nodes against its own bbox.
#define main main_unused
#include "sim_group.cpp"
#undef main
struct Cost { double A = 0, Aint = 0, A2 = 0, A2int = 0, iters = 0, iters2 = 0, nM = 0, nM2 = 0, dfs = 0, dfsint = 0, dfssteps = 0; };
static float g_h2; static double C_ITER = 100, C_DFSSETUP = 40;
static void sub_dfs(int b, int e, const float *tg, int n, Cost &c)
{ // all n targets start active at node b (which is the MIXED node itself)
  int Tp = (n + 31) / 32; std::vector<int> skip(n, b); int no = b;
  while (no < e)
  {
    const Node &nd = nodes[no]; bool any_open = false;
    for (int q = 0; q < n; q++) { if (no < skip[q]) continue;
      float dx = nd.x - tg[4 * q], dy = nd.y - tg[4 * q + 1], dz = nd.z - tg[4 * q + 2]; float r2 = dx * dx + dy * dy + dz * dz;
      if (nd.lenq > r2) any_open = true; else { skip[q] = nd.end; c.dfsint++; } }
    c.dfs += 13.0 * Tp + 15; c.dfssteps++;
    no = any_open ? no + 1 : nd.end;
  }
}
// classification walk of `n` targets over chains; MIXED nodes returned in mlist
static void level(const float *tg, int n, std::vector<std::pair<int, int>> stack, std::vector<int> &mlist, double &Acost, double &Aint, double &iters)
{
  int Tp = (n + 31) / 32;
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (int k = 0; k < n; k++) for (int j = 0; j < 3; j++) { lo[j] = std::min(lo[j], tg[4 * k + j]); hi[j] = std::max(hi[j], tg[4 * k + j]); }
  float cc[3], hw[3]; for (int j = 0; j < 3; j++) { cc[j] = 0.5f * (lo[j] + hi[j]); hw[j] = 0.5f * (hi[j] - lo[j]) * 1.00001f + 1e-30f; }
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
        if (isA) { Aint += n; Acost += (r2min < g_h2 ? 25.0 * Tp + 4 : 8.0 * Tp + 2); }
        ch = nx;
      }
      maxlen = std::max(maxlen, len);
    }
    iters += maxlen;
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
  for (int G : {64, 128, 256, 512})
    for (int S2 : {16, 32, 64})
    {
      if (S2 >= G) continue;
      Cost c; int ngroups = 200; double tot = 0;
      for (int g = 0; g < ngroups; g++)
      {
        int64_t start = (int64_t)((double)g / ngroups * (n - G)); start -= start % G; const float *tg = &sp[4 * start];
        std::vector<int> m1;
        level(tg, G, {{0, (int)nn}}, m1, c.A, c.Aint, c.iters); c.nM += m1.size();
        for (int s = 0; s < G / S2; s++)
        {
          std::vector<std::pair<int, int>> st; for (int no : m1) st.push_back({no, nodes[no].end});
          // note: chain (no, end) would iterate siblings; emulate single-node chains by listing kids after classification of the node itself:
          std::vector<int> m2; double it2 = 0;
          // classify each MIXED node itself against the slice: do it by a fake chain that contains only the node
          std::vector<std::pair<int, int>> single; for (int no : m1) single.push_back({no, no + 1 > nodes[no].end ? nodes[no].end : no + 1});
          // a chain [no, no+1) visits node `no` only (its end > no+1 ends the chain)
          level(tg + 4 * s * S2, S2, single, m2, c.A2, c.A2int, it2); c.iters2 += it2; c.nM2 += m2.size();
          for (int no : m2) { c.dfs += C_DFSSETUP; sub_dfs(no, nodes[no].end, tg + 4 * s * S2, S2, c); }
        }
        tot += G;
      }
      double inter = c.Aint + c.A2int + c.dfsint;
      double cost = c.A + c.A2 + (c.iters + c.iters2) * C_ITER + c.dfs;
      printf("G=%3d slice=%2d: inter/target %.0f | inter share A1 %.2f A2 %.2f dfs %.2f | cost share A1 %.2f A2 %.2f iters %.2f dfs %.2f | per warp: it1 %.0f it2 %.0f nM1 %.0f nM2 %.0f dfssteps %.0f | slots per 32 inter %.1f\n", G, S2, inter / tot,
             c.Aint / inter, c.A2int / inter, c.dfsint / inter, c.A / cost, c.A2 / cost, (c.iters + c.iters2) * C_ITER / cost, c.dfs / cost, c.iters / ngroups, c.iters2 / ngroups, c.nM / ngroups, c.nM2 / ngroups, c.dfssteps / ngroups, cost / (inter / 32));
    }
  return 0;
}
