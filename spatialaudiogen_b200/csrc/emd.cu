// Earth mover's distance between two histograms over a ground-distance matrix: the EMD-hat of Pele & Werman that the
// reference reaches through pyemd.emd(first, second, distance_matrix) (reference pyutils/ambisonics/distance.py:100-127,
// pyemd==0.5.1 in requirements.txt; third-party arithmetic restated from its published definition):
//
//   EMD^(P, Q) = min_F  sum_ij f_ij d_ij  +  |sum P - sum Q| * penalty,     penalty = max_ij d_ij  (extra_mass_penalty = -1)
//   subject to  f >= 0,  sum_j f_ij <= P_i,  sum_i f_ij <= Q_j,  sum_ij f_ij = min(sum P, sum Q).
//
// Host code, like the reference (it runs once per evaluated window on the 84-node energy maps the GPU produced).  The
// unbalanced problem becomes a balanced transportation problem by giving the lighter side a dummy bin that absorbs the
// excess at `penalty` per unit; that is solved exactly by successive shortest augmenting paths with node potentials
// (dense Dijkstra on the bipartite residual graph, doubles throughout).
#include "common.cuh"
#include <cmath>
#include <thread>
#include <algorithm>

namespace sag {

static double emd_hat_one(const double* P, const double* Q, int n, const double* D, double penalty) {
  double sp = 0.0, sq = 0.0, dmax = 0.0;
  for (int i = 0; i < n; ++i) { sp += P[i]; sq += Q[i]; }
  for (int i = 0; i < n * n; ++i) dmax = std::max(dmax, D[i]);
  if (penalty < 0.0) penalty = dmax;
  const int N = n + 1;                                   // + the dummy bin
  std::vector<double> s(N, 0.0), t(N, 0.0);
  for (int i = 0; i < n; ++i) { s[i] = P[i]; t[i] = Q[i]; }
  if (sp > sq) t[n] = sp - sq; else s[n] = sq - sp;
  auto cost = [&](int i, int j) -> double {
    if (i == n && j == n) return 0.0;
    if (i == n || j == n) return penalty;
    return D[i * n + j];
  };
  const double total = std::max(sp, sq);
  if (total <= 0.0) return 0.0;
  const double eps = total * 1e-13;
  std::vector<double> flow((size_t)N * N, 0.0), pu(N, 0.0), pv(N, 0.0);   // potentials of the left / right nodes
  std::vector<double> dl(N), dr(N);
  std::vector<int> prev_l(N), prev_r(N);
  std::vector<char> done_l(N), done_r(N);
  double remaining = 0.0;
  for (int i = 0; i < N; ++i) remaining += s[i];
  double result = 0.0;
  const double INF = 1e300;
  int guard = 0;
  while (remaining > eps && guard++ < 64 * N) {
    // Dijkstra from every left node with supply left; arcs l->r always open (reduced cost c + pu - pv >= 0),
    // arcs r->l open where flow > 0 (reduced cost 0 by complementary slackness)
    for (int i = 0; i < N; ++i) { dl[i] = s[i] > eps ? 0.0 : INF; dr[i] = INF; prev_l[i] = -1; prev_r[i] = -1; done_l[i] = 0; done_r[i] = 0; }
    int target = -1;
    for (;;) {
      int best = -1;
      bool left = true;
      double bd = INF;
      for (int i = 0; i < N; ++i) {
        if (!done_l[i] && dl[i] < bd) { bd = dl[i]; best = i; left = true; }
        if (!done_r[i] && dr[i] < bd) { bd = dr[i]; best = i; left = false; }
      }
      if (best < 0) break;
      if (left) {
        done_l[best] = 1;
        for (int j = 0; j < N; ++j) {
          if (done_r[j]) continue;
          const double rc = std::max(0.0, cost(best, j) + pu[best] - pv[j]);
          if (bd + rc < dr[j]) { dr[j] = bd + rc; prev_r[j] = best; }
        }
      } else {
        done_r[best] = 1;
        if (t[best] > eps) { target = best; break; }
        for (int i = 0; i < N; ++i) {
          if (done_l[i] || flow[(size_t)i * N + best] <= eps) continue;
          if (bd < dl[i]) { dl[i] = bd; prev_l[i] = best; }
        }
      }
    }
    if (target < 0) break;                               // (cannot happen for a balanced problem)
    const double dt = dr[target];
    // bottleneck along the path
    double amt = t[target];
    int j = target;
    for (;;) {
      const int i = prev_r[j];
      if (prev_l[i] < 0) { amt = std::min(amt, s[i]); break; }
      amt = std::min(amt, flow[(size_t)i * N + prev_l[i]]);
      j = prev_l[i];
    }
    j = target;
    for (;;) {
      const int i = prev_r[j];
      flow[(size_t)i * N + j] += amt;
      result += amt * cost(i, j);
      if (prev_l[i] < 0) { s[i] -= amt; break; }
      flow[(size_t)i * N + prev_l[i]] -= amt;
      result -= amt * cost(i, prev_l[i]);
      j = prev_l[i];
    }
    t[target] -= amt;
    remaining -= amt;
    for (int i = 0; i < N; ++i) { pu[i] += std::min(dl[i], dt); pv[i] += std::min(dr[i], dt); }
  }
  // all mass must have been shipped: a tripped iteration guard or an unreachable sink would leave a partial (too small) cost
  // behind -- report it as NaN rather than as a plausible distance
  if (remaining > total * 1e-9) return std::nan("");
  return result;
}

}  // namespace sag

// first, second: [count][n] histograms (non-negative); dist: [n][n] ground distances shared by all pairs;
// extra_mass_penalty < 0 -> max(dist) (pyemd's default -1).  out[count].
int sag_emd_hat(const double* first, const double* second, int n, const double* dist, double extra_mass_penalty, int count,
                double* out) {
  using namespace sag;
  SAG_REQUIRE(first != nullptr && second != nullptr && dist != nullptr && out != nullptr, SAG_EINVAL, "sag_emd_hat: NULL argument");
  SAG_REQUIRE(n > 0 && n <= 4096 && count >= 0, SAG_EINVAL, "sag_emd_hat: bad sizes (n %d, count %d)", n, count);
  for (int64_t i = 0; i < (int64_t)count * n; ++i)
    SAG_REQUIRE(first[i] >= 0.0 && second[i] >= 0.0 && std::isfinite(first[i]) && std::isfinite(second[i]), SAG_EINVAL,
                "sag_emd_hat: histograms must be finite and non-negative");
  unsigned hw = std::thread::hardware_concurrency();
  const int nthreads = (int)std::max(1u, std::min(hw ? hw : 1u, (unsigned)count));
  std::vector<std::thread> pool;
  for (int w = 0; w < nthreads; ++w)
    pool.emplace_back([=]() {
      for (int c = w; c < count; c += nthreads) out[c] = emd_hat_one(first + (size_t)c * n, second + (size_t)c * n, n, dist, extra_mass_penalty);
    });
  for (auto& th : pool) th.join();
  return SAG_OK;
}
