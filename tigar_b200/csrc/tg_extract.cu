// Windowed-CSR pattern helpers, extraction operator M, M*U and M^T b.
// Reference: common.py:1460-1578 (generateM*), :97-109 (multTranspose),
// :367-380 (cpFuncs = M_control * P), :1259 (u = M * U).
#include "tg_common.cuh"

__global__ void k_win_rowlen(TgWin w, int64_t nrows, int64_t* __restrict__ rowlen) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  int rc[3];
  tg_decode(r, w.nr, w.dim, rc);
  int64_t len = 1;
  for (int d = 0; d < w.dim; d++) len *= (w.hi[d][rc[d]] - w.lo[d][rc[d]] + 1);
  rowlen[r] = len;
}

extern "C" int tg_win_rowlen(const tg_win* h_w, int64_t* rowlen, void* stream) {
  int64_t nrows = tg_win_nrows(h_w);
  k_win_rowlen<<<(unsigned)tg_cdiv(nrows, 256), 256, 0, tg_stream(stream)>>>(tg_win_dev(h_w),
                                                                             nrows, rowlen);
  TG_LAUNCH_CHECK();
  return 0;
}

struct TgPrefix {
  const int64_t* S[3];
  int64_t T[3];   // totals
};

// rowptr(r0,r1,r2) = S0[r0] len1 len2 + T0 (S1[r1] len2 + T1 S2[r2])
__global__ void k_win_rowptr(TgWin w, TgPrefix P, int64_t nrows, int64_t* __restrict__ rowptr) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > nrows) return;
  if (r == nrows) {
    rowptr[r] = P.T[0] * P.T[1] * P.T[2];
    return;
  }
  int rc[3];
  tg_decode(r, w.nr, w.dim, rc);
  int64_t len1 = 1, len2 = 1, s1 = 0, s2 = 0;
  if (w.dim > 1) {
    len1 = w.hi[1][rc[1]] - w.lo[1][rc[1]] + 1;
    s1 = P.S[1][rc[1]];
  }
  if (w.dim > 2) {
    len2 = w.hi[2][rc[2]] - w.lo[2][rc[2]] + 1;
    s2 = P.S[2][rc[2]];
  }
  rowptr[r] = P.S[0][rc[0]] * len1 * len2 + P.T[0] * (s1 * len2 + P.T[1] * s2);
}

extern "C" int tg_win_rowptr(const tg_win* h_w, const int64_t* const* h_S, int64_t* rowptr,
                             void* stream) {
  int64_t nrows = tg_win_nrows(h_w);
  TgPrefix P;
  cudaStream_t st = tg_stream(stream);
  for (int d = 0; d < 3; d++) {
    P.S[d] = nullptr;
    P.T[d] = 1;
    if (d < h_w->dim) {
      P.S[d] = h_S[d];
      TG_CHECK(cudaMemcpyAsync(&P.T[d], h_S[d] + h_w->nr[d], sizeof(int64_t),
                               cudaMemcpyDeviceToHost, st));
    }
  }
  TG_CHECK(cudaStreamSynchronize(st));
  k_win_rowptr<<<(unsigned)tg_cdiv(nrows + 1, 256), 256, 0, st>>>(tg_win_dev(h_w), P, nrows,
                                                                  rowptr);
  TG_LAUNCH_CHECK();
  return 0;
}

// one warp per row
__global__ void k_win_fill_cols(TgWin w, int64_t nrows, int32_t* __restrict__ cols) {
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  int rc[3];
  tg_decode(r, w.nr, w.dim, rc);
  TgRowWin rw = tg_row_window(w, rc);
  int64_t base = w.rowptr[r];
  int tot = rw.len[0] * rw.len[1] * rw.len[2];
  for (int pos = lane; pos < tot; pos += 32) {
    int c0 = pos % rw.len[0];
    int t = pos / rw.len[0];
    int c1 = t % rw.len[1];
    int c2 = t / rw.len[1];
    int64_t col = (rw.lo[0] + c0) +
                  (int64_t)w.nc[0] * ((rw.lo[1] + c1) + (int64_t)w.nc[1] * (rw.lo[2] + c2));
    cols[base + pos] = (int32_t)col;
  }
}

extern "C" int tg_win_fill_cols(const tg_win* h_w, int32_t* cols, void* stream) {
  int64_t nrows = tg_win_nrows(h_w);
  k_win_fill_cols<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
      tg_win_dev(h_w), nrows, cols);
  TG_LAUNCH_CHECK();
  return 0;
}

struct TgM1d {
  const int32_t* first[3];
  const double* vals[3];
  int np1[3];
};

// M[I, j] = prod_d m1d_d[I_d][j_d - first_d[I_d]]   (BSplines.py:496-502)
__global__ void k_m_fill(TgWin w, TgM1d m, int64_t nrows, double* __restrict__ vals) {
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  int rc[3];
  tg_decode(r, w.nr, w.dim, rc);
  TgRowWin rw = tg_row_window(w, rc);
  int64_t base = w.rowptr[r];
  int tot = rw.len[0] * rw.len[1] * rw.len[2];
  for (int pos = lane; pos < tot; pos += 32) {
    int c[3];
    c[0] = pos % rw.len[0];
    int t = pos / rw.len[0];
    c[1] = t % rw.len[1];
    c[2] = t / rw.len[1];
    double v = 1.0;
    for (int d = 0; d < w.dim; d++) {
      int j = rw.lo[d] + c[d];
      double f = m.vals[d][rc[d] * m.np1[d] + (j - m.first[d][rc[d]])];
      v = (d == 0) ? f : __dmul_rn(v, f);
    }
    vals[base + pos] = v;
  }
}

extern "C" int tg_m_fill(const tg_win* h_wM, const int32_t* const* h_mfirst,
                         const double* const* h_mvals, const int32_t* h_p, double* vals,
                         void* stream) {
  int64_t nrows = tg_win_nrows(h_wM);
  TgM1d m;
  for (int d = 0; d < 3; d++) {
    bool in = d < h_wM->dim;
    m.first[d] = in ? h_mfirst[d] : nullptr;
    m.vals[d] = in ? h_mvals[d] : nullptr;
    m.np1[d] = in ? h_p[d] + 1 : 1;
  }
  k_m_fill<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(tg_win_dev(h_wM), m,
                                                                              nrows, vals);
  TG_LAUNCH_CHECK();
  return 0;
}

// general CSR y = A x, one warp per row, coalesced loads of vals/cols
__global__ void k_spmv(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
                       const double* __restrict__ vals, const double* __restrict__ x,
                       double* __restrict__ y, int64_t nrows) {
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  int64_t b = rowptr[r], e = rowptr[r + 1];
  double acc = 0.0;
  for (int64_t k = b + lane; k < e; k += 32) acc += vals[k] * x[cols[k]];
  acc = tg_warp_sum(acc);
  if (lane == 0) y[r] = acc;
}

extern "C" int tg_spmv(const int64_t* rowptr, const int32_t* cols, const double* vals,
                       const double* x, double* y, int64_t nrows, void* stream) {
  if (nrows == 0) return 0;
  k_spmv<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(rowptr, cols, vals, x,
                                                                            y, nrows);
  TG_LAUNCH_CHECK();
  return 0;
}

// out[j] = sum_{I in support box of j} M[I,j] b[I]; one warp per IGA function
__global__ void k_mt_vec(TgWin wM, TgWin wT, const double* __restrict__ Mvals,
                         const double* __restrict__ b, double* __restrict__ out, int64_t ncols) {
  int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (j >= ncols) return;
  int jc[3];
  tg_decode(j, wT.nr, wT.dim, jc);
  TgRowWin bx = tg_row_window(wT, jc);  // box of FE nodes
  int tot = bx.len[0] * bx.len[1] * bx.len[2];
  double acc = 0.0;
  for (int pos = lane; pos < tot; pos += 32) {
    int I[3];
    I[0] = bx.lo[0] + pos % bx.len[0];
    int t = pos / bx.len[0];
    I[1] = bx.lo[1] + t % bx.len[1];
    I[2] = bx.lo[2] + t / bx.len[1];
    int64_t row = I[0] + (int64_t)wM.nr[0] * (I[1] + (int64_t)wM.nr[1] * I[2]);
    TgRowWin rw = tg_row_window(wM, I);
    acc += Mvals[wM.rowptr[row] + tg_win_pos(rw, jc)] * b[row];
  }
  acc = tg_warp_sum(acc);
  if (lane == 0) out[j] = acc;
}

extern "C" int tg_mt_vec(const tg_win* h_wM, const tg_win* h_wT, const double* Mvals,
                         const double* b, double* out, void* stream) {
  int64_t ncols = tg_win_nrows(h_wT);
  k_mt_vec<<<(unsigned)tg_cdiv(ncols * 32, 256), 256, 0, tg_stream(stream)>>>(
      tg_win_dev(h_wM), tg_win_dev(h_wT), Mvals, b, out, ncols);
  TG_LAUNCH_CHECK();
  return 0;
}

// out[i] = g[i_d] (g == NULL: out[i] = cval): one column of the homogeneous control net of an
// ExplicitBSplineControlMesh on the device -- control point = Greville abscissa of direction d,
// weight 1 (getHomogeneousCoordinate, BSplines.py:935-960, looped per control point by
// common.py:373-375; here one coalesced pass per column).
__global__ void k_tensor_column(double* __restrict__ out, const double* __restrict__ g, int n0,
                                int n1, int n2, int d, double cval) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!g) { out[i] = cval; continue; }
    const int64_t r = i / n0;
    const int id = (d == 0) ? (int)(i - r * n0) : (d == 1) ? (int)(r % n1) : (int)(r / n1);
    out[i] = g[id];
  }
}

extern "C" int tg_tensor_column(double* out, const double* g, int32_t n0, int32_t n1, int32_t n2,
                                int32_t d, double cval, void* stream) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  if (n == 0) return 0;
  int64_t grid = tg_cdiv(n, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  k_tensor_column<<<(unsigned)grid, 256, 0, tg_stream(stream)>>>(out, g, n0, n1, n2, d, cval);
  TG_LAUNCH_CHECK();
  return 0;
}
