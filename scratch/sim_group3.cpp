// Scratch: like sim_group2 but with realistic overheads: expansion loop iterations (batch of <=32 chains, cost per
// iteration), frame fixed cost, serial kid chains.
#define main main_unused
#include "sim_group.cpp"
#undef main
struct Cost { double A = 0, Aint = 0, M = 0, Mint = 0, iters = 0, frames = 0, nM = 0, dfs = 0, dfsint = 0, C = 0, dfssteps = 0; };
static float g_h2; static double C_ITER = 100, C_FRAME = 400, C_DFS_STEP_BASE = 15, C_DFS_T = 13;
static void sub_dfs(int b, int e, const std::vector<int> &idx, const float *tg, Cost &c)
{
  int n = idx.size(); int Tp = (n + 31) / 32; if (Tp == 3) Tp = 4;
  std::vector<int> skip(n, b); int no = b;
  while (no < e)
  {
    const Node &nd = nodes[no]; bool any_open = false;
    for (int q = 0; q < n; q++) { if (no < skip[q]) continue; int k = idx[q];
      float dx = nd.x - tg[4 * k], dy = nd.y - tg[4 * k + 1], dz = nd.z - tg[4 * k + 2]; float r2 = dx * dx + dy * dy + dz * dz;
      if (nd.lenq > r2) any_open = true; else { skip[q] = nd.end; c.dfsint++; } }
    c.dfs += C_DFS_T * Tp + C_DFS_STEP_BASE; c.dfssteps++;
    no = any_open ? no + 1 : nd.end;
  }
}
static void rec_walk(int first, int pend, const std::vector<int> &idx, const float *tg, Cost &c, int depth, int minrec)
{
  int n = idx.size(); int Tp = (n + 31) / 32; if (Tp == 3) Tp = 4;
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (int k : idx) for (int j = 0; j < 3; j++) { lo[j] = std::min(lo[j], tg[4 * k + j]); hi[j] = std::max(hi[j], tg[4 * k + j]); }
  float cc[3], hw[3]; for (int j = 0; j < 3; j++) { cc[j] = 0.5f * (lo[j] + hi[j]); hw[j] = 0.5f * (hi[j] - lo[j]) * 1.00001f + 1e-30f; }
  c.frames += 1;
  std::vector<std::pair<int, int>> stack; stack.push_back({first, pend});
  std::vector<int> mlist;
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
        if (isA) { if (r2min < g_h2) { c.C += 25.0 * Tp + 4; c.Aint += n; } else { c.A += 8.0 * Tp + 2; c.Aint += n; } }
        ch = nx;
      }
      maxlen = std::max(maxlen, len);
    }
    c.iters += maxlen;
  }
  for (int ch : mlist)
  {
    const Node &nd = nodes[ch];
    c.nM++; c.M += 15.0 * Tp + 20;
    std::vector<int> op;
    for (int k : idx) { float dx = nd.x - tg[4 * k], dy = nd.y - tg[4 * k + 1], dz = nd.z - tg[4 * k + 2]; float r2 = dx * dx + dy * dy + dz * dz; if (nd.lenq > r2) op.push_back(k); else c.Mint++; }
    if (!op.empty())
    {
      if ((int)op.size() >= minrec && depth < 11) rec_walk(ch + 1, nd.end, op, tg, c, depth + 1, minrec);
      else { c.frames += 0.3; sub_dfs(ch + 1, nd.end, op, tg, c); }
    }
  }
}
int main(int argc, char **argv)
{
  int64_t n = argc > 1 ? atoll(argv[1]) : 2000000; float eps = argc > 2 ? atof(argv[2]) : 4.8e-5f; double a = argc > 3 ? atof(argv[3]) : 0.03;
  if (argc > 4) C_ITER = atof(argv[4]); if (argc > 5) C_FRAME = atof(argv[5]);
  std::mt19937_64 rng(12345); std::uniform_real_distribution<double> U(0, 1); std::normal_distribution<double> Nn(0, 1);
  std::vector<float> src(4 * n);
  for (int64_t i = 0; i < n; i++) { double u = U(rng) * 0.97, s = std::sqrt(u), r = a * s / (1 - s); double x = Nn(rng), y = Nn(rng), z = Nn(rng), q = r / std::sqrt(x * x + y * y + z * z);
    src[4 * i] = 50 + x * q; src[4 * i + 1] = 50 + y * q; src[4 * i + 2] = 50 + z * q; src[4 * i + 3] = 1e-6f; }
  build(src, n, 0.1 * eps, 0.45f * 0.45f);
  float h = 2.8f * eps; g_h2 = h * h;
  for (int G : {64, 128, 256})
    for (int minrec : {1, 7, 16, 33, 65, 1000})
    {
      Cost c; int ngroups = 200; double tot = 0;
      for (int g = 0; g < ngroups; g++) { int64_t start = (int64_t)((double)g / ngroups * (n - G)); start -= start % G;
        std::vector<int> idx(G); for (int k = 0; k < G; k++) idx[k] = k;
        rec_walk(0, (int)nn, idx, &sp[4 * start], c, 0, minrec); tot += G; }
      double inter = c.Aint + c.Mint + c.dfsint;
      double ovh = c.iters * C_ITER + c.frames * C_FRAME;
      double cost = c.A + c.C + c.M + ovh + c.dfs;
      printf("G=%3d minrec=%4d: share A %.2f M %.2f dfs %.2f | cost share A %.2f C %.2f M %.2f iters %.2f frames %.2f dfs %.2f | per warp: frames %.0f iters %.0f nM %.0f dfssteps %.0f | slots per 32 inter %.1f\n", G, minrec,
             c.Aint / inter, c.Mint / inter, c.dfsint / inter, c.A / cost, c.C / cost, c.M / cost, c.iters * C_ITER / cost, c.frames * C_FRAME / cost, c.dfs / cost, c.frames / ngroups, c.iters / ngroups, c.nM / ngroups, c.dfssteps / ngroups, cost / (inter / 32));
    }
  return 0;
}
