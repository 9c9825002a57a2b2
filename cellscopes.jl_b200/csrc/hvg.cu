// find_variable_genes (src/scrna/processing.jl:133-155), the part after the per-gene moments:
// loess of log10(variance) on log10(mean), variance_expected, variance_standardized and the
// stable descending rank.  G <= a few 10^4 points and ~33 local fits: host code, a few ms.
// The local regression restates Loess.jl 0.6.3 (kd-tree vertices + cubic Hermite blending);
// the operation order follows oracle/cellscopes_oracle.py::loess_fit so both agree to rounding.
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>

#include "common.cuh"

namespace {

// idx[] = the permutation of a stable ascending sort by key (ties, -0.0 == +0.0 included, keep the lower index): what
// std::sort over (key, index) pairs gives.  The order-preserving integer image of the doubles is LSD-radix-sorted on
// its upper 32 bits only -- 8-byte records (key prefix, index), three passes of 11 / 11 / 10 bits, a pass whose digit
// is the same for all keys is skipped --, which leaves groups of keys that agree in sign, exponent and 20 mantissa
// bits (a handful among 33 k genes, unless values repeat) in index order; those groups are then ordered by the full
// key.  (Six passes over all 64 bits with separate key and index arrays cost 2 x 1.2 ms of the 4.6 ms that 33k genes
// took; the pair sort before that 2 x 1.7 ms.)  No NaN keys.
void stable_argsort(const std::vector<double>& key, std::vector<int64_t>& idx) {
  const size_t n = key.size();
  idx.resize(n);
  auto image = [](double d) {
    d += 0.0;                                             // -0.0 -> +0.0: they compare equal
    uint64_t u;
    memcpy(&u, &d, 8);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
  };
  if (n < 4096 || n > 0xffffffffull) {
    std::vector<std::pair<uint64_t, int64_t>> pk(n);
    for (size_t i = 0; i < n; ++i) pk[i] = std::make_pair(image(key[i]), (int64_t)i);
    std::sort(pk.begin(), pk.end());
    for (size_t i = 0; i < n; ++i) idx[i] = pk[i].second;
    return;
  }
  struct Rec {
    uint32_t k, i;
  };
  std::vector<Rec> r0(n), r1(n);
  for (size_t i = 0; i < n; ++i) {
    r0[i].k = (uint32_t)(image(key[i]) >> 32);
    r0[i].i = (uint32_t)i;
  }
  Rec* ra = r0.data();
  Rec* rb = r1.data();
  const int shifts[3] = {0, 11, 22}, bits[3] = {11, 11, 10};
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = shifts[pass];
    const uint32_t mask = (1u << bits[pass]) - 1u;
    size_t cnt[2049] = {0};
    for (size_t i = 0; i < n; ++i) ++cnt[((ra[i].k >> shift) & mask) + 1];
    bool single = false;
    for (int bkt = 0; bkt < 2048; ++bkt) {
      single |= cnt[bkt + 1] == n;
      cnt[bkt + 1] += cnt[bkt];
    }
    if (single) continue;
    for (size_t i = 0; i < n; ++i) rb[cnt[(ra[i].k >> shift) & mask]++] = ra[i];
    std::swap(ra, rb);
  }
  // groups of equal prefix: by (full key, index)
  std::vector<std::pair<uint64_t, uint32_t>> grp;
  for (size_t g0 = 0; g0 < n;) {
    size_t g1 = g0 + 1;
    while (g1 < n && ra[g1].k == ra[g0].k) ++g1;
    if (g1 - g0 > 1) {
      grp.clear();
      for (size_t i = g0; i < g1; ++i) grp.emplace_back(image(key[ra[i].i]), ra[i].i);
      std::sort(grp.begin(), grp.end());
      for (size_t i = g0; i < g1; ++i) ra[i].i = grp[i - g0].second;
    }
    g0 = g1;
  }
  for (size_t i = 0; i < n; ++i) idx[i] = (int64_t)ra[i].i;
}

// fn(begin, end) over [0, n) on a few host threads (the per-gene loops of 33k transcendental calls); the pieces are
// disjoint, so the results do not depend on the split.  Falls back to the calling thread if a thread cannot start.
template <typename F>
void parallel_for(int64_t n, F fn) {
  const int64_t n_thr = n >= 8192 ? std::min<int64_t>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
  if (n_thr <= 1) {
    fn((int64_t)0, n);
    return;
  }
  const int64_t chunk = (n + n_thr - 1) / n_thr;
  std::vector<std::thread> pool;
  int64_t next = chunk;
  try {
    for (; next < n; next += chunk) {
      const int64_t lo = next, hi = std::min(n, next + chunk);
      pool.emplace_back([=] { fn(lo, hi); });
    }
  } catch (...) {
  }
  fn((int64_t)0, std::min(chunk, n));
  for (auto& th : pool) th.join();
  if (next < n) fn(next, n);                              // the pieces whose thread did not start
}

// Householder QR with column pivoting on a q x 3 system (Julia: qr(us, ColumnNorm()) \ vs).  The columns are swapped
// by pointer; every step makes two fused passes over the rows (all dot products with the reflector, then all updates);
// the sums run in four interleaved accumulators (a single fp64 accumulator retires one add per 4 cycles: the eleven
// dependent sums over ~10 k rows were most of the 0.4 ms a vertex cost).  `vbuf`: scratch of q doubles.
void pivoted_qr_solve3(double* a /* q x 3 column-major, destroyed */, double* b /* q, destroyed */, double* vbuf, int64_t q,
                       double coef[3]) {
  const int ncol = 3;
  int perm[3] = {0, 1, 2};
  double* col[3] = {a, a + q, a + 2 * q};
  auto sum4 = [](const double (&s)[4]) { return (s[0] + s[1]) + (s[2] + s[3]); };
  for (int k = 0; k < ncol; ++k) {
    // squared norms of the remaining columns below row k, one pass
    double nrm[3][4] = {{0}};
    {
      int64_t i = k;
      for (; i + 4 <= q; i += 4)
        for (int j = k; j < ncol; ++j)
          for (int u = 0; u < 4; ++u) nrm[j][u] += col[j][i + u] * col[j][i + u];
      for (; i < q; ++i)
        for (int j = k; j < ncol; ++j) nrm[j][0] += col[j][i] * col[j][i];
    }
    int p = k;
    double best = -1.0;
    for (int j = k; j < ncol; ++j) {
      const double t = sum4(nrm[j]);
      if (t > best) {
        best = t;
        p = j;
      }
    }
    if (p != k) {
      std::swap(col[k], col[p]);
      std::swap(perm[k], perm[p]);
    }
    double* x = col[k];
    double alpha = std::sqrt(best);
    if (alpha == 0.0) continue;
    if (x[k] > 0) alpha = -alpha;
    // v = x - alpha e1
    double* v = vbuf;
    for (int64_t i = k; i < q; ++i) v[i] = x[i];
    v[k] -= alpha;
    // pass 1: v.v, v.col_j (j >= k), v.b
    double dv[4] = {0}, dc[3][4] = {{0}}, db[4] = {0};
    {
      int64_t i = k;
      for (; i + 4 <= q; i += 4)
        for (int u = 0; u < 4; ++u) {
          const double t = v[i + u];
          dv[u] += t * t;
          for (int j = k; j < ncol; ++j) dc[j][u] += t * col[j][i + u];
          db[u] += t * b[i + u];
        }
      for (; i < q; ++i) {
        const double t = v[i];
        dv[0] += t * t;
        for (int j = k; j < ncol; ++j) dc[j][0] += t * col[j][i];
        db[0] += t * b[i];
      }
    }
    const double vnorm2 = sum4(dv);
    if (vnorm2 == 0.0) continue;
    double sc[3] = {0, 0, 0};
    for (int j = k; j < ncol; ++j) sc[j] = 2.0 * sum4(dc[j]) / vnorm2;
    const double sb = 2.0 * sum4(db) / vnorm2;
    // pass 2: the reflections
    for (int64_t i = k; i < q; ++i) {
      const double t = v[i];
      for (int j = k; j < ncol; ++j) col[j][i] -= sc[j] * t;
      b[i] -= sb * t;
    }
  }
  double cp[3] = {0, 0, 0};
  for (int k = ncol - 1; k >= 0; --k) {
    double t = b[k];
    for (int j = k + 1; j < ncol; ++j) t -= col[j][k] * cp[j];
    const double d = col[k][k];
    cp[k] = d != 0.0 ? t / d : 0.0;
  }
  for (int k = 0; k < ncol; ++k) coef[perm[k]] = cp[k];
}

void kd_vertices(const std::vector<double>& xs_sorted, int64_t cutoff, std::vector<double>& verts) {
  verts.clear();
  verts.push_back(xs_sorted.front());
  verts.push_back(xs_sorted.back());
  std::vector<std::pair<int64_t, int64_t>> stack;
  stack.emplace_back(0, (int64_t)xs_sorted.size());
  while (!stack.empty()) {
    auto [lo, hi] = stack.back();
    stack.pop_back();
    int64_t n = hi - lo;
    if (n <= cutoff || xs_sorted[hi - 1] - xs_sorted[lo] <= 0.0) continue;
    int64_t mid = n / 2;
    double med = (n % 2 == 1) ? xs_sorted[lo + mid] : (xs_sorted[lo + mid - 1] + xs_sorted[lo + mid]) / 2;
    verts.push_back(med);
    stack.emplace_back(lo, lo + mid);
    stack.emplace_back(lo + mid, hi);
  }
  std::sort(verts.begin(), verts.end());
  verts.erase(std::unique(verts.begin(), verts.end()), verts.end());
}

}  // namespace

// x, y: the points (genes with variance > 0).  out[i] = finish(i, loess prediction at x[i]).
// With a communicator on the context (one process per GPU) the vertex fits -- the expensive part, ~15 ms of one host
// thread for 33 k genes -- are dealt out over the ranks (vertex i to rank i mod world) and their (value, slope) pairs are
// all-reduced: every rank adds zeros for the vertices it did not fit, so all ranks end with bit-identical tables, and
// eight ranks no longer run eight copies of the same 8-thread job on the same host cores.
csc_status comm_allreduce_f64(csc_ctx* ctx, double* host_inout, size_t n);   // comm.cu

// finish(i, prediction at x[i]) -> out[i]
template <typename Finish>
static csc_status loess_fit_predict(csc_ctx* ctx, bool collective, const std::vector<double>& x, const std::vector<double>& y,
                                    double span, std::vector<double>& out, Finish finish) {
  const int64_t n = (int64_t)x.size();
  const double cell = 0.2;
  const int64_t q = std::max<int64_t>(1, (int64_t)std::floor(span * (double)n));
  const int64_t cutoff = (int64_t)std::ceil(cell * span * (double)n);
  // the points sorted by (x, index): the q nearest to a vertex are a window around it, so they come out of a
  // two-pointer walk in O(q) instead of a selection over all n points per vertex
  std::vector<int64_t> sidx;
  stable_argsort(x, sidx);
  if (ctx) CSC_TRACE(ctx, "hvg: argsort of log10 mean");                               // by (x, index) == a stable sort by x
  std::vector<double> xs((size_t)n);
  for (int64_t i = 0; i < n; ++i) xs[i] = x[sidx[i]];
  std::vector<double> verts;
  kd_vertices(xs, cutoff, verts);
  const size_t nv = verts.size();
  std::vector<double> val(nv), slope(nv);
  std::vector<double> ys((size_t)n);
  for (int64_t i = 0; i < n; ++i) ys[i] = y[sidx[i]];
  const double inf = std::numeric_limits<double>::infinity();
  // The local fits are independent: host threads take the vertices round-robin (33 k genes give 33 fits over 9.9 k
  // points each).  The q points nearest to a vertex -- ties of the q-th distance resolved towards the lower gene index,
  // like the stable argsort of |x - v| in the oracle -- are a window of the sorted points plus, at most, a few tied
  // points at its rim: the window comes from a binary search over "how many from the left", the rows of the weighted
  // system are then built from contiguous slices of the sorted arrays.  (The least-squares solution does not depend on
  // the order of the rows; the round-1 code walked outwards from the vertex point by point to visit them nearest-first.)
  auto fit_range = [&](size_t v_begin, size_t v_step) {
    std::vector<double> a((size_t)q * 3), b((size_t)q), vbuf((size_t)q);
    std::vector<std::pair<int64_t, int64_t>> tied;                     // (gene index, sorted position)
    for (size_t vi = v_begin; vi < nv; vi += v_step) {
      const double v = verts[vi];
      const int64_t R0 = (int64_t)(std::lower_bound(xs.begin(), xs.end(), v) - xs.begin());
      const int64_t nL = R0, nR = n - R0;
      auto dl = [&](int64_t i) { return i >= 1 && i <= nL ? std::fabs(xs[R0 - i] - v) : (i < 1 ? -inf : inf); };       // i-th to the left
      auto dr = [&](int64_t j) { return j >= 1 && j <= nR ? std::fabs(xs[R0 + j - 1] - v) : (j < 1 ? -inf : inf); };   // j-th to the right
      // smallest count from the left such that the next left point is not strictly closer than the last right one taken
      int64_t lo_a = std::max<int64_t>(0, q - nR), hi_a = std::min(q, nL);
      while (lo_a < hi_a) {
        const int64_t m = (lo_a + hi_a) / 2;
        if (dl(m + 1) < dr(q - m)) lo_a = m + 1;
        else hi_a = m;
      }
      const double dq = std::max(dl(lo_a), dr(q - lo_a));               // the q-th smallest distance
      // points strictly closer than dq (taken for sure) and points at exactly dq (taken by ascending gene index)
      auto count_below = [&](auto dist, int64_t cnt, bool strict) {
        int64_t lo = 0, hi = cnt;                                       // number of points with dist < dq (or <= dq)
        while (lo < hi) {
          const int64_t m = (lo + hi) / 2;
          const double d = dist(m + 1);
          if (strict ? d < dq : d <= dq) lo = m + 1;
          else hi = m;
        }
        return lo;
      };
      const int64_t aL = count_below(dl, nL, true), aLe = count_below(dl, nL, false);
      const int64_t bR = count_below(dr, nR, true), bRe = count_below(dr, nR, false);
      const int64_t need = q - aL - bR;                                 // >= 1 points of the tied rim
      tied.clear();
      for (int64_t i = aL + 1; i <= aLe; ++i) tied.emplace_back(sidx[R0 - i], R0 - i);
      for (int64_t j = bR + 1; j <= bRe; ++j) tied.emplace_back(sidx[R0 + j - 1], R0 + j - 1);
      if ((int64_t)tied.size() > need) {
        std::nth_element(tied.begin(), tied.begin() + need, tied.end());
        tied.resize((size_t)need);
      }
      const double dmax = dq == 0.0 ? 1.0 : dq;
      double* a0 = a.data();
      double* a1 = a0 + q;
      double* a2 = a1 + q;
      int64_t r = 0;
      auto row = [&](int64_t pos) {
        const double xc = xs[pos] - v;
        const double u = std::fabs(xc) / dmax;
        const double u3 = u * u * u;
        const double t = 1.0 - u3;
        const double w = std::sqrt(t * t * t);
        a0[r] = w;
        a1[r] = w * xc;
        a2[r] = w * xc * xc;
        b[r] = ys[pos] * w;
        ++r;
      };
      for (int64_t pos = R0 - aL; pos < R0 + bR; ++pos) row(pos);
      for (const auto& tp : tied) row(tp.second);
      double coef[3];
      pivoted_qr_solve3(a.data(), b.data(), vbuf.data(), q, coef);
      val[vi] = coef[0];
      slope[vi] = coef[1];
    }
  };
  const size_t world = (collective && ctx && ctx->comm && ctx->comm_world > 1 && nv >= 8) ? (size_t)ctx->comm_world : 1;
  const size_t rank = world > 1 ? (size_t)ctx->comm_rank : 0;
  if (world > 1) {
    std::fill(val.begin(), val.end(), 0.0);
    std::fill(slope.begin(), slope.end(), 0.0);
  }
  const size_t my_nv = (nv + world - 1 - rank) / world;                 // vertices rank, rank + world, ...
  const size_t n_thr = (n >= 2000 && my_nv >= 2) ? std::min<size_t>(std::min<size_t>(16, my_nv), std::max(1u, std::thread::hardware_concurrency())) : 1;
  if (n_thr <= 1) {
    fit_range(rank, world);
  } else {
    // no exception may leave a thread (it would terminate the process): record it and rethrow after the join
    std::vector<int> failed(n_thr, 0);
    auto guarded = [&](size_t tix) {
      try {
        fit_range(rank + world * tix, world * n_thr);
      } catch (...) {
        failed[tix] = 1;
      }
    };
    std::vector<std::thread> pool;
    size_t started = 1;
    try {
      for (; started < n_thr; ++started) pool.emplace_back(guarded, started);
    } catch (...) {
      // could not start every thread: the vertices of the missing ones are done here
    }
    guarded(0);
    for (auto& th : pool) th.join();
    for (size_t tix = started; tix < n_thr; ++tix) guarded(tix);
    for (size_t tix = 0; tix < n_thr; ++tix)
      if (failed[tix]) throw std::bad_alloc();
  }
  if (ctx) CSC_TRACE(ctx, "hvg: vertex fits");
  if (world > 1) {
    std::vector<double> both(2 * nv);
    std::copy(val.begin(), val.end(), both.begin());
    std::copy(slope.begin(), slope.end(), both.begin() + nv);
    CSC_TRY(comm_allreduce_f64(ctx, both.data(), both.size()));
    std::copy(both.begin(), both.begin() + nv, val.begin());
    std::copy(both.begin() + nv, both.end(), slope.begin());
  }
  out.resize((size_t)n);
  parallel_for(n, [&](int64_t i_lo, int64_t i_hi) {
    for (int64_t i = i_lo; i < i_hi; ++i) {
      const double z = x[i];
      size_t hi = (size_t)(std::lower_bound(verts.begin(), verts.end(), z) - verts.begin());
      if (hi < 1) hi = 1;
      if (hi > nv - 1) hi = nv - 1;
      double pred;
      if (nv == 1) {
        pred = val[0];
      } else {
        const size_t lo = hi - 1;
        const double v1 = verts[lo], v2 = verts[hi];
        if (z == v1) {
          pred = val[lo];
        } else if (z == v2) {
          pred = val[hi];
        } else {
          const double h = v2 - v1;
          const double t = (z - v1) / h;
          const double omt = 1.0 - t;
          const double h00 = (1 + 2 * t) * omt * omt;
          const double h10 = t * omt * omt;
          const double h01 = t * t * (3 - 2 * t);
          const double h11 = t * t * (t - 1);
          pred = h00 * val[lo] + h10 * h * slope[lo] + h01 * val[hi] + h11 * h * slope[hi];
        }
      }
      out[i] = finish(i, pred);
    }
  });
  return CSC_OK;
}

csc_status launch_hvg_host(csc_ctx* ctx, const double* s1s2, int64_t n_genes, int64_t n_cells_total,
                           int64_t n_features, double span, double* mean, double* variance, double* var_exp,
                           double* var_std, int64_t* order, bool collective) {
  (void)n_features;  // the caller takes order[0:n_features]; clamping is :134-137
  const double N = (double)n_cells_total;
  std::vector<double> mu((size_t)n_genes), var((size_t)n_genes), vs((size_t)n_genes, 0.0);
  std::vector<double> lx((size_t)n_genes), ly((size_t)n_genes);
  // one pass per gene: moments and their logarithms (a gene without variance gets NaN there and is reported below)
  parallel_for(n_genes, [&](int64_t lo, int64_t hi) {
    for (int64_t g = lo; g < hi; ++g) {
      const double s1 = s1s2[g], s2 = s1s2[n_genes + g];
      mu[g] = s1 / N;                          // exact-moment rule, SURVEY A.3-7
      var[g] = (s2 - s1 * s1 / N) / (N - 1.0);
      const bool keep = var[g] > 0.0;          // filter(:variance => >(0.0)) processing.jl:143
      lx[g] = keep ? std::log10(mu[g]) : std::numeric_limits<double>::quiet_NaN();
      ly[g] = keep ? std::log10(var[g]) : std::numeric_limits<double>::quiet_NaN();
    }
  });
  int64_t n_kept = 0, first_bad = -1;
  for (int64_t g = 0; g < n_genes; ++g) {
    if (var[g] > 0.0) ++n_kept;
    else if (first_bad < 0) first_bad = g;
  }
  CSC_TRACE(ctx, "hvg: moments, log10");
  // The reference requires every gene to have variance > 0: it filters the data frame (:143) and then assigns the
  // gene names of ALL genes to it (:151), which throws a DimensionMismatch when a row was dropped.  Same here (and in
  // the oracle): selecting such a gene would put 0/0 = NaN rows into the scaled matrix and the Gram matrix.
  if (n_kept != n_genes)
    return csc_fail(ctx, CSC_ERR_ARG,
                    "find_variable_genes: %lld of %lld genes have variance <= 0 (first: gene %lld); the reference requires "
                    "every gene to vary (processing.jl:143,151) -- its object constructor drops all-zero genes first "
                    "(subset_matrix, utils.jl:313-320)",
                    (long long)(n_genes - n_kept), (long long)n_genes, (long long)first_bad);
  std::vector<double> ve;
  if (n_genes > 0) {
    // the Loess prediction, 10^pred and the ratio in one pass over the genes
    CSC_TRY(loess_fit_predict(ctx, collective, lx, ly, span, ve, [&](int64_t g, double pred) {
      const double e = std::pow(10.0, pred);
      vs[g] = var[g] / e;
      return e;
    }));
  }
  CSC_TRACE(ctx, "hvg: interpolation, pow, ratio");
  if (mean) std::copy(mu.begin(), mu.end(), mean);
  if (variance) std::copy(var.begin(), var.end(), variance);
  if (var_exp) std::copy(ve.begin(), ve.end(), var_exp);
  if (var_std) std::copy(vs.begin(), vs.end(), var_std);
  if (order) {
    // descending variance_standardized, ties by ascending gene index (= the stable sort of processing.jl:152)
    std::vector<double> neg((size_t)n_genes);
    for (int64_t g = 0; g < n_genes; ++g) neg[g] = -vs[g];
    std::vector<int64_t> ord;
    stable_argsort(neg, ord);
    std::copy(ord.begin(), ord.end(), order);
  }
  CSC_TRACE(ctx, "hvg: final order");
  return CSC_OK;
}
