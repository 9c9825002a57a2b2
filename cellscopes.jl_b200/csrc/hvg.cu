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
// std::sort over (key, index) pairs gives, as an LSD radix sort over the order-preserving integer image of the
// doubles (six 11-bit passes; the pair sort cost 2 x 1.7 ms of the 8 ms that 33k genes take).  No NaN keys.
void stable_argsort(const std::vector<double>& key, std::vector<int64_t>& idx) {
  const size_t n = key.size();
  idx.resize(n);
  if (n < 4096) {
    std::vector<std::pair<double, int64_t>> pk(n);
    for (size_t i = 0; i < n; ++i) pk[i] = std::make_pair(key[i], (int64_t)i);
    std::sort(pk.begin(), pk.end());
    for (size_t i = 0; i < n; ++i) idx[i] = pk[i].second;
    return;
  }
  std::vector<uint64_t> k0(n), k1(n);
  std::vector<int64_t> i1(n);
  for (size_t i = 0; i < n; ++i) {
    const double d = key[i] + 0.0;                        // -0.0 -> +0.0: they compare equal
    uint64_t u;
    memcpy(&u, &d, 8);
    k0[i] = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    idx[i] = (int64_t)i;
  }
  uint64_t* ka = k0.data();
  uint64_t* kb = k1.data();
  int64_t* ia = idx.data();
  int64_t* ib = i1.data();
  for (int pass = 0; pass < 6; ++pass) {
    const int shift = 11 * pass;
    size_t cnt[2049] = {0};
    for (size_t i = 0; i < n; ++i) ++cnt[((ka[i] >> shift) & 2047u) + 1];
    for (int bkt = 0; bkt < 2048; ++bkt) cnt[bkt + 1] += cnt[bkt];
    for (size_t i = 0; i < n; ++i) {
      const size_t o = cnt[(ka[i] >> shift) & 2047u]++;
      kb[o] = ka[i];
      ib[o] = ia[i];
    }
    std::swap(ka, kb);
    std::swap(ia, ib);
  }
  // six passes: the result is back in the first pair of buffers (idx)
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

// Householder QR with column pivoting on a q x 3 system (Julia: qr(us, ColumnNorm()) \ vs)
void pivoted_qr_solve3(std::vector<double>& a /* q x 3 column-major */, std::vector<double>& b, int64_t q,
                       double coef[3]) {
  const int ncol = 3;
  int perm[3] = {0, 1, 2};
  auto col = [&](int j) { return a.data() + (size_t)j * q; };
  for (int k = 0; k < ncol; ++k) {
    int p = k;
    double best = -1.0;
    for (int j = k; j < ncol; ++j) {
      double s = 0.0;
      const double* c = col(j);
      for (int64_t i = k; i < q; ++i) s += c[i] * c[i];
      if (s > best) {
        best = s;
        p = j;
      }
    }
    if (p != k) {
      double* ck = col(k);
      double* cp = col(p);
      for (int64_t i = 0; i < q; ++i) std::swap(ck[i], cp[i]);
      std::swap(perm[k], perm[p]);
    }
    double* x = col(k);
    double nrm2 = 0.0;
    for (int64_t i = k; i < q; ++i) nrm2 += x[i] * x[i];
    double alpha = std::sqrt(nrm2);
    if (alpha == 0.0) continue;
    if (x[k] > 0) alpha = -alpha;
    // v = x - alpha e1 (stored in place of x below the diagonal after use)
    std::vector<double> v((size_t)(q - k));
    for (int64_t i = k; i < q; ++i) v[i - k] = x[i];
    v[0] -= alpha;
    double vnorm2 = 0.0;
    for (double t : v) vnorm2 += t * t;
    if (vnorm2 == 0.0) continue;
    for (int j = k; j < ncol; ++j) {
      double* c = col(j);
      double s = 0.0;
      for (int64_t i = k; i < q; ++i) s += v[i - k] * c[i];
      s = 2.0 * s / vnorm2;
      for (int64_t i = k; i < q; ++i) c[i] -= s * v[i - k];
    }
    double s = 0.0;
    for (int64_t i = k; i < q; ++i) s += v[i - k] * b[i];
    s = 2.0 * s / vnorm2;
    for (int64_t i = k; i < q; ++i) b[i] -= s * v[i - k];
  }
  double cp[3] = {0, 0, 0};
  for (int k = ncol - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < ncol; ++j) s -= col(j)[k] * cp[j];
    double d = col(k)[k];
    cp[k] = d != 0.0 ? s / d : 0.0;
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

// x, y: the points (genes with variance > 0).  out[i] = loess prediction at x[i].
// With a communicator on the context (one process per GPU) the vertex fits -- the expensive part, ~15 ms of one host
// thread for 33 k genes -- are dealt out over the ranks (vertex i to rank i mod world) and their (value, slope) pairs are
// all-reduced: every rank adds zeros for the vertices it did not fit, so all ranks end with bit-identical tables, and
// eight ranks no longer run eight copies of the same 8-thread job on the same host cores.
csc_status comm_allreduce_f64(csc_ctx* ctx, double* host_inout, size_t n);   // comm.cu

static csc_status loess_fit_predict(csc_ctx* ctx, bool collective, const std::vector<double>& x, const std::vector<double>& y,
                                    double span, std::vector<double>& out) {
  const int64_t n = (int64_t)x.size();
  const double cell = 0.2;
  const int64_t q = std::max<int64_t>(1, (int64_t)std::floor(span * (double)n));
  const int64_t cutoff = (int64_t)std::ceil(cell * span * (double)n);
  // the points sorted by (x, index): the q nearest to a vertex are a window around it, so they come out of a
  // two-pointer walk in O(q) instead of a selection over all n points per vertex
  std::vector<int64_t> sidx;
  stable_argsort(x, sidx);                               // by (x, index) == a stable sort by x
  std::vector<double> xs((size_t)n);
  for (int64_t i = 0; i < n; ++i) xs[i] = x[sidx[i]];
  std::vector<double> verts;
  kd_vertices(xs, cutoff, verts);
  const size_t nv = verts.size();
  std::vector<double> val(nv), slope(nv);
  const double inf = std::numeric_limits<double>::infinity();
  // the local fits are independent: a few host threads take the vertices round-robin (33 k genes give 33 fits over
  // 10 k points each; one thread needs ~15 ms for them)
  auto fit_range = [&](size_t v_begin, size_t v_step) {
    std::vector<int64_t> order;
    std::vector<double> d;
    order.reserve((size_t)q + 64);
    d.reserve((size_t)q + 64);
    std::vector<double> a((size_t)q * 3), b((size_t)q);
    for (size_t vi = v_begin; vi < nv; vi += v_step) {
      const double v = verts[vi];
      // nearest first, ties by lower index: exactly the order of a stable sort of |x - v| (the oracle's argsort)
      int64_t R = (int64_t)(std::lower_bound(xs.begin(), xs.end(), v) - xs.begin()), L = R - 1;
      order.clear();
      d.clear();
      while (true) {
        const double dl = L >= 0 ? std::fabs(xs[L] - v) : inf;
        const double dr = R < n ? std::fabs(xs[R] - v) : inf;
        const double dn = std::min(dl, dr);
        if (dn == inf) break;
        if ((int64_t)order.size() >= q && dn > d.back()) break;     // keep every tie of the q-th distance for now
        if (dl <= dr) {
          order.push_back(sidx[L--]);
          d.push_back(dl);
        } else {
          order.push_back(sidx[R++]);
          d.push_back(dr);
        }
      }
      // equal distances: ascending index
      for (size_t i0 = 0; i0 < order.size();) {
        size_t i1 = i0 + 1;
        while (i1 < order.size() && d[i1] == d[i0]) ++i1;
        if (i1 - i0 > 1) std::sort(order.begin() + i0, order.begin() + i1);
        i0 = i1;
      }
      double dmax = 0.0;
      for (int64_t i = 0; i < q; ++i) dmax = std::max(dmax, d[i]);
      if (dmax == 0.0) dmax = 1.0;
      for (int64_t i = 0; i < q; ++i) {
        const int64_t p = order[i];
        const double u = d[i] / dmax;
        const double u3 = u * u * u;
        const double t = 1.0 - u3;
        const double w = std::sqrt(t * t * t);
        const double xc = x[p] - v;
        a[i] = w;
        a[(size_t)q + i] = w * xc;
        a[(size_t)2 * q + i] = w * xc * xc;
        b[i] = y[p] * w;
      }
      double coef[3];
      pivoted_qr_solve3(a, b, q, coef);
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
  const size_t n_thr = (n >= 2000 && my_nv >= 2) ? std::min<size_t>(std::min<size_t>(8, my_nv), std::max(1u, std::thread::hardware_concurrency())) : 1;
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
    if (nv == 1) {
      out[i] = val[0];
      continue;
    }
    const size_t lo = hi - 1;
    const double v1 = verts[lo], v2 = verts[hi];
    if (z == v1) {
      out[i] = val[lo];
      continue;
    }
    if (z == v2) {
      out[i] = val[hi];
      continue;
    }
    const double h = v2 - v1;
    const double t = (z - v1) / h;
    const double omt = 1.0 - t;
    const double h00 = (1 + 2 * t) * omt * omt;
    const double h10 = t * omt * omt;
    const double h01 = t * t * (3 - 2 * t);
    const double h11 = t * t * (t - 1);
    out[i] = h00 * val[lo] + h10 * h * slope[lo] + h01 * val[hi] + h11 * h * slope[hi];
  }
  });
  return CSC_OK;
}

csc_status launch_hvg_host(csc_ctx* ctx, const double* s1s2, int64_t n_genes, int64_t n_cells_total,
                           int64_t n_features, double span, double* mean, double* variance, double* var_exp,
                           double* var_std, int64_t* order, bool collective) {
  (void)n_features;  // the caller takes order[0:n_features]; clamping is :134-137
  const double N = (double)n_cells_total;
  std::vector<double> mu((size_t)n_genes), var((size_t)n_genes), ve((size_t)n_genes, 0.0), vs((size_t)n_genes, 0.0);
  std::vector<double> lx, ly;
  std::vector<int64_t> kept;
  for (int64_t g = 0; g < n_genes; ++g) {
    const double s1 = s1s2[g], s2 = s1s2[n_genes + g];
    mu[g] = s1 / N;                            // exact-moment rule, SURVEY A.3-7
    var[g] = (s2 - s1 * s1 / N) / (N - 1.0);
    if (var[g] > 0.0) kept.push_back(g);       // filter(:variance => >(0.0)) processing.jl:143
  }
  lx.resize(kept.size());
  ly.resize(kept.size());
  parallel_for((int64_t)kept.size(), [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      lx[i] = std::log10(mu[kept[i]]);
      ly[i] = std::log10(var[kept[i]]);
    }
  });
  // The reference requires every gene to have variance > 0: it filters the data frame (:143) and then assigns the
  // gene names of ALL genes to it (:151), which throws a DimensionMismatch when a row was dropped.  Same here (and in
  // the oracle): selecting such a gene would put 0/0 = NaN rows into the scaled matrix and the Gram matrix.
  if ((int64_t)kept.size() != n_genes) {
    int64_t first_bad = 0;
    while (first_bad < n_genes && var[first_bad] > 0.0) ++first_bad;
    return csc_fail(ctx, CSC_ERR_ARG,
                    "find_variable_genes: %lld of %lld genes have variance <= 0 (first: gene %lld); the reference requires "
                    "every gene to vary (processing.jl:143,151) -- its object constructor drops all-zero genes first "
                    "(subset_matrix, utils.jl:313-320)",
                    (long long)(n_genes - (int64_t)kept.size()), (long long)n_genes, (long long)first_bad);
  }
  if (!kept.empty()) {
    std::vector<double> pred;
    CSC_TRY(loess_fit_predict(ctx, collective, lx, ly, span, pred));
    parallel_for((int64_t)kept.size(), [&](int64_t lo, int64_t hi) {
      for (int64_t i = lo; i < hi; ++i) {
        const int64_t g = kept[i];
        ve[g] = std::pow(10.0, pred[i]);
        vs[g] = var[g] / ve[g];
      }
    });
  }
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
  (void)ctx;
  return CSC_OK;
}
