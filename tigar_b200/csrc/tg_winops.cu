// Operations on windowed-CSR matrices (rows hold a dense tensor-product box of
// columns, so no column-index array is stored: 8 B per non-zero instead of 12):
// y = C x fused with the CG dot product, symmetric Dirichlet elimination and
// Jacobi diagonal.  These replace, for tensor-product B-spline patches,
// PETSc's MatMult inside KSP (common.py:1255-1258), MatZeroRowsColumns
// (common.py:1199-1200) and MatGetDiagonal of PCJACOBI.
//
// Work decomposition of the SpMV: an item is 32 consecutive rows of one
// (r1,r2) line; a persistent grid walks the items.  Inside a row the lanes of
// one warp cover the (c0,c1) tile of the window and march along c2 with pure
// pointer increments, so the inner loop is 2 loads + 1 DFMA, no index math.
// HBM-bound: the value array is streamed exactly once, x is served by L1/L2.
#include "tg_common.cuh"

#define TG_WS_BLOCK 256
#define TG_WS_ROWS 32   // rows of one line per item

static int g_ws_grid = 0;
int tg_ws_grid_size() {
  if (!g_ws_grid) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_ws_grid = sms * 4;   // 4 CTAs x 256 threads resident per SM, one wave
  }
  return g_ws_grid;
}

__device__ inline double tg_block_sum_ws(double v, double* sh) {
  v = tg_warp_sum(v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    r = tg_warp_sum(r);
  }
  return r;
}

// DOT: also accumulate sum_r x[xoff + r] * y[r] into part[blockIdx.x]
//
// A CTA walks items (32 consecutive rows of one line).  For the line's window
// shape (w0max x len1 x len2) it keeps in shared memory the x-offset of every
// window position, so the fast path of a row is: lanes stride the row's
// contiguous value array (fully coalesced 256 B loads), look the x offset up in
// shared memory and gather x from L1/L2.  No integer divisions in the loop.
// Rows whose first-direction window is clipped by the patch boundary take the
// generic path.
#define TG_WS_MAXTAB 1024
template <bool DOT>
__global__ void __launch_bounds__(TG_WS_BLOCK, 4)
k_win_spmv(TgWin w, const double* __restrict__ vals, const double* __restrict__ x,
           int64_t xoff, double* __restrict__ y, int nchunk, int nitems, int w0max,
           double* __restrict__ part) {
  __shared__ double sh[32];
  __shared__ int xtab[TG_WS_MAXTAB];
  __shared__ int tab_len1, tab_len2;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nr0 = w.nr[0], nr1 = w.nr[1];
  const int nc0 = w.nc[0];
  const int64_t pl = (int64_t)nc0 * w.nc[1];          // x plane stride
  if (threadIdx.x == 0) {
    tab_len1 = -1;
    tab_len2 = -1;
  }
  __syncthreads();
  double dot = 0.0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int line = item / nchunk;
    const int ch = item - line * nchunk;
    const int r2 = line / nr1;
    const int r1 = line - r2 * nr1;
    int lo1 = 0, len1 = 1, lo2 = 0, len2 = 1;
    if (w.dim > 1) {
      lo1 = __ldg(w.lo[1] + r1);
      len1 = __ldg(w.hi[1] + r1) - lo1 + 1;
    }
    if (w.dim > 2) {
      lo2 = __ldg(w.lo[2] + r2);
      len2 = __ldg(w.hi[2] + r2) - lo2 + 1;
    }
    const int tile = len1 * len2;
    const int ntab = w0max * tile;
    const bool usetab = ntab <= TG_WS_MAXTAB;
    if (usetab && (tab_len1 != len1 || tab_len2 != len2)) {      // uniform branch
      __syncthreads();
      for (int p = threadIdx.x; p < ntab; p += TG_WS_BLOCK) {
        int c0 = p % w0max, t = p / w0max, c1 = t % len1, c2 = t / len1;
        xtab[p] = c0 + nc0 * c1 + (int)pl * c2;
      }
      if (threadIdx.x == 0) {
        tab_len1 = len1;
        tab_len2 = len2;
      }
      __syncthreads();
    }
    const int rend = min(nr0, ch * TG_WS_ROWS + TG_WS_ROWS);
    const double* xb = x + (int64_t)nc0 * lo1 + pl * lo2;
    // rowptr in closed form: S0[r0]*len1*len2 + T0*(S1[r1]*len2 + T1*S2[r2])
    int64_t linebase = 0;
    if (w.dim > 1) {
      const int64_t T0 = __ldg(w.S[0] + nr0);
      int64_t inner = __ldg(w.S[1] + r1) * len2;
      if (w.dim > 2) inner += __ldg(w.S[1] + nr1) * __ldg(w.S[2] + r2);
      linebase = T0 * inner;
    }
    for (int r0 = ch * TG_WS_ROWS + wid; r0 < rend; r0 += TG_WS_BLOCK / 32) {
      const int64_t row = r0 + (int64_t)nr0 * line;
      const int lo0 = __ldg(w.lo[0] + r0);
      const int len0 = __ldg(w.hi[0] + r0) - lo0 + 1;
      const double* __restrict__ av = vals + linebase + __ldg(w.S[0] + r0) * tile;
      const double* xr = xb + lo0;
      const int n = len0 * tile;
      double acc = 0.0;
      if (usetab && len0 == w0max) {
        double acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        int p = lane;
        for (; p + 96 < n; p += 128) {
          double a0 = __ldcs(av + p), a1 = __ldcs(av + p + 32), a2 = __ldcs(av + p + 64),
                 a3 = __ldcs(av + p + 96);
          double x0 = xr[xtab[p]], x1 = xr[xtab[p + 32]], x2 = xr[xtab[p + 64]],
                 x3 = xr[xtab[p + 96]];
          acc += a0 * x0;
          acc1 += a1 * x1;
          acc2 += a2 * x2;
          acc3 += a3 * x3;
        }
        for (; p < n; p += 32) acc += __ldcs(av + p) * xr[xtab[p]];
        acc += (acc1 + acc2) + acc3;
      } else {
        for (int p = lane; p < n; p += 32) {
          int c0 = p % len0, t = p / len0, c1 = t % len1, c2 = t / len1;
          acc += av[p] * xr[c0 + nc0 * c1 + pl * c2];
        }
      }
      acc = tg_warp_sum(acc);
      if (lane == 0) {
        y[row] = acc;
        if (DOT) dot += x[xoff + row] * acc;
      }
    }
  }
  if (DOT) {
    dot = tg_block_sum_ws(dot, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = dot;
  }
}

static inline int64_t tg_win_lines(const tg_win* w) {
  int64_t n = 1;
  for (int d = 1; d < w->dim; d++) n *= w->nr[d];
  return n;
}

int tg_win_spmv_launch(const tg_win* h_w, const double* vals, const double* x, int64_t xoff,
                       double* y, double* part, cudaStream_t st) {
  TG_REQUIRE(h_w->dim >= 1 && h_w->dim <= 3, "dim");
  int nchunk = (int)tg_cdiv(h_w->nr[0], TG_WS_ROWS);
  int64_t nitems = tg_win_lines(h_w) * nchunk;
  TG_REQUIRE(nitems < (int64_t)2147483647, "too many row items");
  if (nitems == 0) return 0;
  int g = tg_ws_grid_size();
  TgWin w = tg_win_dev(h_w);
  TG_REQUIRE(h_w->w0max >= 1, "window descriptor lacks w0max");
  if (part)
    k_win_spmv<true><<<g, TG_WS_BLOCK, 0, st>>>(w, vals, x, xoff, y, nchunk, (int)nitems,
                                                h_w->w0max, part);
  else
    k_win_spmv<false><<<g, TG_WS_BLOCK, 0, st>>>(w, vals, x, xoff, y, nchunk, (int)nitems,
                                                 h_w->w0max, nullptr);
  TG_LAUNCH_CHECK();
  return 0;
}

extern "C" int tg_win_spmv(const tg_win* h_w, const double* vals, const double* x, double* y,
                           void* stream) {
  return tg_win_spmv_launch(h_w, vals, x, 0, y, nullptr, tg_stream(stream));
}

// zeroRowsColumns on the window pattern: one warp per row.
// col_shift: added to the row's last-direction coordinate to get its own column
// coordinate (0 on one GPU; halo offset for a slab-local block).
__global__ void k_win_zero_rows_cols(TgWin w, double* __restrict__ vals, int64_t nrows,
                                     const uint8_t* __restrict__ rowmask,
                                     const uint8_t* __restrict__ colmask, double diag,
                                     int col_shift) {
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  int rc[3];
  tg_decode(r, w.nr, w.dim, rc);
  TgRowWin rw = tg_row_window(w, rc);
  int cc[3] = {rc[0], rc[1], rc[2]};
  cc[w.dim - 1] += col_shift;
  const int64_t mycol = cc[0] + (int64_t)w.nc[0] * (cc[1] + (int64_t)w.nc[1] * cc[2]);
  const bool mr = rowmask[r] != 0;
  const int64_t base = w.rowptr[r];
  const int tot = rw.len[0] * rw.len[1] * rw.len[2];
  for (int pos = lane; pos < tot; pos += 32) {
    int c0 = pos % rw.len[0];
    int t = pos / rw.len[0];
    int c1 = t % rw.len[1];
    int c2 = t / rw.len[1];
    int64_t col = (rw.lo[0] + c0) +
                  (int64_t)w.nc[0] * ((rw.lo[1] + c1) + (int64_t)w.nc[1] * (rw.lo[2] + c2));
    if (mr || colmask[col]) vals[base + pos] = (mr && col == mycol) ? diag : 0.0;
  }
}

extern "C" int tg_win_zero_rows_cols(const tg_win* h_w, double* vals, const uint8_t* rowmask,
                                     const uint8_t* colmask, double diag, int32_t col_shift,
                                     void* stream) {
  int64_t nrows = tg_win_nrows(h_w);
  if (nrows == 0) return 0;
  k_win_zero_rows_cols<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
      tg_win_dev(h_w), vals, nrows, rowmask, colmask, diag, col_shift);
  TG_LAUNCH_CHECK();
  return 0;
}

__global__ void k_win_diag_inv(TgWin w, const double* __restrict__ vals, int64_t nrows,
                               int col_shift, double* __restrict__ dinv) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  int rc[3];
  tg_decode(r, w.nr, w.dim, rc);
  TgRowWin rw = tg_row_window(w, rc);
  int cc[3] = {rc[0], rc[1], rc[2]};
  cc[w.dim - 1] += col_shift;
  double d = vals[w.rowptr[r] + tg_win_pos(rw, cc)];
  dinv[r] = (d != 0.0) ? 1.0 / d : 1.0;
}

extern "C" int tg_win_diag_inv(const tg_win* h_w, const double* vals, int32_t col_shift,
                               double* dinv, void* stream) {
  int64_t nrows = tg_win_nrows(h_w);
  if (nrows == 0) return 0;
  k_win_diag_inv<<<(unsigned)tg_cdiv(nrows, 256), 256, 0, tg_stream(stream)>>>(
      tg_win_dev(h_w), vals, nrows, col_shift, dinv);
  TG_LAUNCH_CHECK();
  return 0;
}
