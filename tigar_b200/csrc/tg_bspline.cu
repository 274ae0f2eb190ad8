// B-spline evaluation and 1-D per-element tables.
// Reference semantics: BSplines.py:285-351 (span, nodes, basisFuncs) and the
// embedded C++ basisFuncsInner, BSplines.py:73-120 (Piegl-Tiller A2.2).
#include "tg_common.cuh"

#define TG_MAXP 8

// Cox-de Boor recurrence with the reference's operation order.  Explicit
// _rn intrinsics keep nvcc from contracting a*b+c into an FMA, so results are
// bit-identical to the CPU recurrence.
__device__ void tg_basis_funcs(const double* __restrict__ g, int nG, double u, int p,
                               int i, double* ders) {
  double ndu[(TG_MAXP + 1) * (TG_MAXP + 1)];
  double left[TG_MAXP + 1], right[TG_MAXP + 1];
  const int N = p + 1;
  ndu[0] = 1.0;
  for (int j = 1; j <= p; j++) {
    left[j] = __dsub_rn(u, g[i - j + nG]);
    right[j] = __dsub_rn(g[i + j - 1 + nG], u);
    double saved = 0.0;
    for (int r = 0; r < j; r++) {
      double d = __dadd_rn(right[r + 1], left[j - r]);
      ndu[j * N + r] = d;
      double temp = __ddiv_rn(ndu[r * N + (j - 1)], d);
      ndu[r * N + j] = __dadd_rn(saved, __dmul_rn(right[r + 1], temp));
      saved = __dmul_rn(left[j - r], temp);
    }
    ndu[j * N + j] = saved;
  }
  for (int j = 0; j <= p; j++) ders[j] = ndu[j * N + p];
}

// numpy.searchsorted(knots,u,'left')-1, clamped as BSplines.py:300-308
__device__ int tg_knot_span(const double* __restrict__ knots, int nk, double u, int mult0,
                            int multLast) {
  int lo = 0, hi = nk;  // count of knots < u
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (knots[mid] < u) lo = mid + 1; else hi = mid;
  }
  int span = lo - 1;
  int nspans = nk - 1;
  int smin = mult0 - 1;
  int smax = nspans - (multLast - 1) - 1;
  if (span < smin) span = smin;
  if (span > smax) span = smax;
  return span;
}

__global__ void k_bspline_eval_batch(const double* __restrict__ knots, int nk,
                                     const double* __restrict__ ghost, int nG, int p, int ncp,
                                     int mult0, int multLast, const double* __restrict__ u,
                                     int64_t n, int32_t* __restrict__ span,
                                     int32_t* __restrict__ nodes, double* __restrict__ vals) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  double uu = u[t];
  int s = tg_knot_span(knots, nk, uu, mult0, multLast);
  double d[TG_MAXP + 1];
  tg_basis_funcs(ghost, nG, uu, p, s + 1, d);
  if (span) span[t] = s;
  for (int i = 0; i <= p; i++) {
    if (nodes) {
      int c = (s - p + i) % ncp;
      if (c < 0) c += ncp;
      nodes[t * (p + 1) + i] = c;
    }
    vals[t * (p + 1) + i] = d[i];
  }
}

extern "C" int tg_bspline_eval_batch(const double* knots, int32_t nk, const double* ghostKnots,
                                     int32_t nGhost, int32_t p, int32_t ncp, int32_t mult0,
                                     int32_t multLast, const double* u, int64_t n,
                                     int32_t* span, int32_t* nodes, double* vals,
                                     void* stream) {
  TG_REQUIRE(p >= 1 && p <= TG_MAXP, "degree out of range");
  if (n == 0) return 0;
  int bs = 128;
  k_bspline_eval_batch<<<(unsigned)tg_cdiv(n, bs), bs, 0, tg_stream(stream)>>>(
      knots, nk, ghostKnots, nGhost, p, ncp, mult0, multLast, u, n, span, nodes, vals);
  TG_LAUNCH_CHECK();
  return 0;
}

// basisFuncsInner itself (BSplines.py:73-120): caller-supplied index i (= span + 1) per point
__global__ void k_basis_funcs_inner(const double* __restrict__ ghost, int nG, int p,
                                    const double* __restrict__ u, const int32_t* __restrict__ i,
                                    int64_t n, double* __restrict__ ders) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  double d[TG_MAXP + 1];
  tg_basis_funcs(ghost, nG, u[t], p, i[t], d);
  for (int j = 0; j <= p; j++) ders[t * (p + 1) + j] = d[j];
}

extern "C" int tg_basis_funcs_inner(const double* ghostKnots, int32_t nGhost, int32_t p,
                                    const double* u, const int32_t* i, int64_t n, double* ders,
                                    void* stream) {
  TG_REQUIRE(p >= 1 && p <= TG_MAXP, "degree out of range");
  if (n == 0) return 0;
  k_basis_funcs_inner<<<(unsigned)tg_cdiv(n, 128), 128, 0, tg_stream(stream)>>>(
      ghostKnots, nGhost, p, u, i, n, ders);
  TG_LAUNCH_CHECK();
  return 0;
}

// node e*pf+a ; a=0/pf exactly on the unique knots; interior uk[e]+(a*h)/pf
__device__ inline double tg_fe_node(const double* __restrict__ uk, int e, int a, int pf) {
  if (a == 0) return uk[e];
  if (a == pf) return uk[e + 1];
  double h = __dsub_rn(uk[e + 1], uk[e]);
  return __dadd_rn(uk[e], __ddiv_rn(__dmul_rn((double)a, h), (double)pf));
}

__global__ void k_fe_nodes_1d(const double* __restrict__ uk, int nel, int pf,
                              double* __restrict__ x) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  int n = nel * pf + 1;
  if (g >= n) return;
  int e = g / pf, a = g % pf;
  if (e == nel) { e = nel - 1; a = pf; }
  x[g] = tg_fe_node(uk, e, a, pf);
}

extern "C" int tg_fe_nodes_1d(const double* uniqueKnots, int32_t nel, int32_t pf, double* x,
                              void* stream) {
  int n = nel * pf + 1;
  k_fe_nodes_1d<<<(unsigned)tg_cdiv(n, 128), 128, 0, tg_stream(stream)>>>(uniqueKnots, nel, pf, x);
  TG_LAUNCH_CHECK();
  return 0;
}

__global__ void k_tabulate_1d(const double* __restrict__ ghost, int nG, int p, int ncp,
                              const double* __restrict__ uk, const int32_t* __restrict__ espan,
                              int nel, int pf, int nq, int nder,
                              const double* __restrict__ lag, const double* __restrict__ tq,
                              const double* __restrict__ gw, double* __restrict__ Me,
                              double* __restrict__ tabN, int32_t* __restrict__ idxN,
                              double* __restrict__ tabL, int32_t* __restrict__ idxL,
                              double* __restrict__ wq, double* __restrict__ xq) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nel) return;
  const int np1 = p + 1, nf = pf + 1, nd = nder + 1;
  double* me = Me + (int64_t)e * nf * np1;
  int s = espan[e];
  for (int a = 0; a < nf; a++) {
    double d[TG_MAXP + 1];
    tg_basis_funcs(ghost, nG, tg_fe_node(uk, e, a, pf), p, s + 1, d);
    for (int i = 0; i < np1; i++) me[a * np1 + i] = d[i];
  }
  double h = uk[e + 1] - uk[e];
  double invh = 1.0 / h;
  for (int q = 0; q < nq; q++) {
    double sc = 1.0;
    for (int k = 0; k < nd; k++) {
      for (int i = 0; i < np1; i++) {
        double acc = 0.0;
        for (int a = 0; a < nf; a++) acc += me[a * np1 + i] * lag[(q * nf + a) * nd + k];
        tabN[(((int64_t)e * nq + q) * np1 + i) * nd + k] = acc * sc;
      }
      for (int a = 0; a < nf; a++)
        tabL[(((int64_t)e * nq + q) * nf + a) * nd + k] = lag[(q * nf + a) * nd + k] * sc;
      sc *= invh;
    }
    wq[e * nq + q] = gw[q] * h;
    xq[e * nq + q] = uk[e] + tq[q] * h;
  }
  for (int i = 0; i < np1; i++) {
    int c = (s - p + i) % ncp;
    if (c < 0) c += ncp;
    idxN[e * np1 + i] = c;
  }
  for (int a = 0; a < nf; a++) idxL[e * nf + a] = e * pf + a;
}

extern "C" int tg_tabulate_1d(const double* ghostKnots, int32_t nGhost, int32_t p, int32_t ncp,
                              const double* uniqueKnots, const int32_t* espan, int32_t nel,
                              int32_t pf, int32_t nq, int32_t nder, const double* lag,
                              const double* tq, const double* gw, double* Me, double* tabN,
                              int32_t* idxN, double* tabL, int32_t* idxL, double* wq,
                              double* xq, void* stream) {
  TG_REQUIRE(p >= 1 && p <= TG_MAXP && pf >= p && pf <= TG_MAXP, "degree out of range");
  k_tabulate_1d<<<(unsigned)tg_cdiv(nel, 64), 64, 0, tg_stream(stream)>>>(
      ghostKnots, nGhost, p, ncp, uniqueKnots, espan, nel, pf, nq, nder, lag, tq, gw, Me, tabN,
      idxN, tabL, idxL, wq, xq);
  TG_LAUNCH_CHECK();
  return 0;
}
