// Direct solve of the extracted system: band Cholesky on the device.
// The reference's default solve() is a sparse direct LU through PETSc (common.py:1255-1256);
// for 2-D patches (and small 3-D ones) the IGA matrix in the reference's own DoF numbering is
// banded with half-bandwidth  bw = sum_d p_d * stride_d, so a blocked band factorisation
// reaches the LU answer where Krylov iterations stall (biharmonic, cond ~ h^-4).
//
// Storage: LAPACK lower band, AB[(i-j) + j*ldab] = A(i,j) for j <= i <= j+bw, ldab >= bw + NB:
// the panel below a diagonal block is then a full rectangle (entries outside the band are zero
// and stay zero).  Element (i,j) is also AB[i + j*(ldab-1)]: a dense column-major view with
// leading dimension ldab-1, so the trailing update is a plain lower-triangular DGEMM
// (tg_dense.cu) on the view.
//
// Right-looking, block size NB = 32, two launches per block column:
//   k_band_panel : every CTA factors the NB x NB diagonal block in the registers of one warp
//                  (lane = row, shuffles; redundantly -- cheaper than a separate launch), CTA 0
//                  writes it back; then each thread solves one panel row  x L11^T = a.
//   k_dgemm<N,T> : trailing (m x m, lower) -= panel * panel^T.
// The triangular solves advance block by block (one launch each, column-oriented updates so no
// cross-CTA reduction is needed).
#include "tg_common.cuh"
#include <math.h>

#define BD_NB 32

int tg_dgemm_lower_nt(int M, int K, double alpha, const double* A, int lda, double beta,
                      double* C, int ldc, cudaStream_t st);

// ---- windowed CSR -> lower band -----------------------------------------------------------
// Block (fr, fc) of an nf x nf equal-order multi-field system goes to the node-major
// interleaved numbering  row' = nf*row + fr, col' = nf*col + fc  (band = nf*bw + nf - 1).
__global__ void k_band_from_win(TgWin w, const double* __restrict__ vals, int bw, int ldab,
                                double* __restrict__ AB, int* __restrict__ info, int nf, int fr,
                                int fc) {
  const int64_t nrows = (int64_t)w.nr[0] * w.nr[1] * w.nr[2];
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; row < nrows;
       row += nw) {
    int rc[3];
    tg_decode(row, w.nr, w.dim, rc);
    const TgRowWin rw = tg_row_window(w, rc);
    const TgRowAddr ra = tg_row_addr(w, rc, rw);
    const int cnt = rw.len[0] * rw.len[1] * rw.len[2];
    for (int pos = lane; pos < cnt; pos += 32) {
      const int d0 = pos % rw.len[0];
      const int t = pos / rw.len[0];
      const int d1 = t % rw.len[1], d2 = t / rw.len[1];
      const int64_t col = (rw.lo[0] + d0) +
                          (int64_t)w.nc[0] * ((rw.lo[1] + d1) + (int64_t)w.nc[1] * (rw.lo[2] + d2));
      const int64_t brow = (int64_t)nf * row + fr, bcol = (int64_t)nf * col + fc;
      if (bcol > brow) continue;
      const int64_t off = brow - bcol;
      const double v = vals[ra.base + ((long long)(d2 * rw.len[1] + d1) * ra.len0 +
                                       (rw.lo[0] + d0 - ra.lo0)) * ra.stride];
      if (off > bw) {
        if (v != 0.0) atomicExch(info, -1);
        continue;
      }
      AB[off + bcol * (int64_t)ldab] = v;
    }
  }
}

// AB must be zero-initialised by the caller.  info (device int, zero-initialised): -1 if an
// entry outside the band was non-zero.
extern "C" int tg_band_from_win(const tg_win* h_w, const double* vals, int32_t bw, int32_t ldab,
                                double* AB, int32_t* info, int32_t nf, int32_t fr, int32_t fc,
                                void* stream) {
  TG_REQUIRE(nf >= 1 && fr >= 0 && fr < nf && fc >= 0 && fc < nf, "field indices");
  TG_REQUIRE(h_w->layout == 0, "band conversion needs the row-major window layout");
  const int64_t n = tg_win_nrows(h_w);
  if (n == 0) return 0;
  int64_t g = tg_cdiv(n * 32, 256);
  if (g > 148 * 32) g = 148 * 32;
  k_band_from_win<<<(unsigned)g, 256, 0, tg_stream(stream)>>>(tg_win_dev(h_w), vals, bw, ldab, AB,
                                                             info, nf, fr, fc);
  TG_LAUNCH_CHECK();
  return 0;
}

// max |A - A^T| and max |A| over the stored entries (symmetry test before Cholesky / CG):
// out2[0] = max_ij |a_ij - a_ji| , out2[1] = max |a_ij|.  An entry whose transpose lies outside
// the pattern counts with a_ji = 0.
__global__ void k_win_asym(TgWin w, const double* __restrict__ vals,
                           unsigned long long* __restrict__ out2) {
  const int64_t nrows = (int64_t)w.nr[0] * w.nr[1] * w.nr[2];
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double dmax = 0.0, amax = 0.0;
  for (int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; row < nrows;
       row += nw) {
    int rc[3];
    tg_decode(row, w.nr, w.dim, rc);
    const TgRowWin rw = tg_row_window(w, rc);
    const TgRowAddr ra = tg_row_addr(w, rc, rw);
    const int cnt = rw.len[0] * rw.len[1] * rw.len[2];
    for (int pos = lane; pos < cnt; pos += 32) {
      const int d0 = pos % rw.len[0];
      const int t = pos / rw.len[0];
      const int d1 = t % rw.len[1], d2 = t / rw.len[1];
      int cc[3] = {rw.lo[0] + d0, rw.lo[1] + d1, rw.lo[2] + d2};
      const double v = vals[ra.base + ((long long)(d2 * rw.len[1] + d1) * ra.len0 +
                                       (cc[0] - ra.lo0)) * ra.stride];
      // transpose entry: row cc, column rc
      const TgRowWin tw = tg_row_window(w, cc);
      double vt = 0.0;
      bool in = true;
#pragma unroll
      for (int d = 0; d < 3; d++) in = in && rc[d] >= tw.lo[d] && rc[d] < tw.lo[d] + tw.len[d];
      if (in) {
        const TgRowAddr ta = tg_row_addr(w, cc, tw);
        vt = vals[ta.base + ((long long)((rc[2] - tw.lo[2]) * tw.len[1] + (rc[1] - tw.lo[1])) *
                                 ta.len0 + (rc[0] - ta.lo0)) * ta.stride];
      }
      dmax = fmax(dmax, fabs(v - vt));
      amax = fmax(amax, fabs(v));
    }
  }
  // non-negative doubles order like their bit patterns
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmax = fmax(dmax, __shfl_down_sync(0xffffffffu, dmax, o));
    amax = fmax(amax, __shfl_down_sync(0xffffffffu, amax, o));
  }
  if (lane == 0) {
    atomicMax(out2, (unsigned long long)__double_as_longlong(dmax));
    atomicMax(out2 + 1, (unsigned long long)__double_as_longlong(amax));
  }
}

// out2: two device doubles, zero-initialised by the caller (square windows only)
extern "C" int tg_win_asym(const tg_win* h_w, const double* vals, double* out2, void* stream) {
  TG_REQUIRE(h_w->layout == 0, "symmetry test needs the row-major window layout");
  for (int d = 0; d < h_w->dim; d++)
    TG_REQUIRE(h_w->nr[d] == h_w->nc[d] && h_w->row0[d] == 0 && h_w->col0[d] == 0,
               "symmetry test needs a whole square matrix");
  const int64_t n = tg_win_nrows(h_w);
  if (n == 0) return 0;
  int64_t g = tg_cdiv(n * 32, 256);
  if (g > 148 * 32) g = 148 * 32;
  k_win_asym<<<(unsigned)g, 256, 0, tg_stream(stream)>>>(tg_win_dev(h_w), vals,
                                                        (unsigned long long*)out2);
  TG_LAUNCH_CHECK();
  return 0;
}

// ---- factorisation ------------------------------------------------------------------------
// Block column k (nb = min(NB, n-k) columns), m panel rows below the diagonal block.
__global__ void __launch_bounds__(128)
k_band_panel(double* __restrict__ AB, int ldab, int k, int nb, int m, int* __restrict__ info) {
  __shared__ double L[BD_NB][BD_NB + 1];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 32) {
    // lane = row i of the diagonal block; missing rows/columns (last block) act as identity
    double a[BD_NB];
#pragma unroll
    for (int j = 0; j < BD_NB; j++) {
      double v = (lane == j) ? 1.0 : 0.0;
      if (lane < nb && j < nb && j <= lane) v = AB[(lane - j) + (int64_t)(k + j) * ldab];
      a[j] = v;
    }
    bool bad = false;
#pragma unroll
    for (int j = 0; j < BD_NB; j++) {
      const double ajj = __shfl_sync(0xffffffffu, a[j], j);
      if (!(ajj > 0.0)) bad = true;
      const double d = sqrt(ajj);
      const double lij = a[j] / d;
      if (lane >= j) a[j] = lij;
#pragma unroll
      for (int q = j + 1; q < BD_NB; q++) {
        const double lqj = __shfl_sync(0xffffffffu, a[j], q);
        if (lane >= q) a[q] = fma(-lij, lqj, a[q]);
      }
    }
#pragma unroll
    for (int j = 0; j < BD_NB; j++) L[lane][j] = (j <= lane) ? a[j] : 0.0;
    if (blockIdx.x == 0) {
      if (bad && lane == 0) atomicCAS(info, 0, k + 1);
#pragma unroll
      for (int j = 0; j < BD_NB; j++)
        if (lane < nb && j < nb && j <= lane) AB[(lane - j) + (int64_t)(k + j) * ldab] = a[j];
    }
  }
  __syncthreads();
  const int r = blockIdx.x * 128 + tid;
  if (r >= m) return;
  // panel row r: A(k+nb+r, k+j) = AB[(nb + r - j) + (k+j)*ldab]
  double x[BD_NB];
#pragma unroll
  for (int j = 0; j < BD_NB; j++)
    x[j] = (j < nb) ? AB[(nb + r - j) + (int64_t)(k + j) * ldab] : 0.0;
#pragma unroll
  for (int j = 0; j < BD_NB; j++) {
    double s = x[j];
#pragma unroll
    for (int t = 0; t < j; t++) s = fma(-x[t], L[j][t], s);
    x[j] = s / L[j][j];
  }
#pragma unroll
  for (int j = 0; j < BD_NB; j++)
    if (j < nb) AB[(nb + r - j) + (int64_t)(k + j) * ldab] = x[j];
}

// In-place Cholesky A = L L^T of a lower band matrix.  info (device int, zero-initialised by the
// caller): k+1 if the pivot block starting at column k is not positive definite.
extern "C" int tg_band_cholesky(int64_t n, int32_t bw, int32_t ldab, double* AB, int32_t* info,
                                void* stream) {
  TG_REQUIRE(ldab >= bw + BD_NB, "ldab must be >= bw + 32");
  TG_REQUIRE(n < (1LL << 31), "band solver: n must fit int32");
  cudaStream_t st = tg_stream(stream);
  for (int64_t k = 0; k < n; k += BD_NB) {
    const int nb = (int)((n - k < BD_NB) ? (n - k) : BD_NB);
    int64_t m64 = n - k - nb;
    if (m64 > bw) m64 = bw;
    const int m = (int)m64;
    const int grid = (m > 0) ? (int)tg_cdiv(m, 128) : 1;
    k_band_panel<<<grid, 128, 0, st>>>(AB, ldab, (int)k, nb, m, info);
    TG_LAUNCH_CHECK();
    if (m > 0) {
      // trailing block at (k+nb, k+nb): dense view with leading dimension ldab-1
      double* Cb = AB + (k + nb) * (int64_t)ldab;
      const double* P = AB + nb + k * (int64_t)ldab;      // A(k+nb, k) = AB[nb + k*ldab]
      int rc = tg_dgemm_lower_nt(m, nb, -1.0, P, ldab - 1, 1.0, Cb, ldab - 1, st);
      if (rc) return rc;
    }
  }
  return 0;
}

// ---- triangular solves ----------------------------------------------------------------------
// forward step k:  y_k = L11^-1 b_k ;  b[k+nb+r] -= sum_j L(k+nb+r, k+j) y_j
__global__ void __launch_bounds__(256)
k_band_fwd(const double* __restrict__ AB, int ldab, int k, int nb, int m, double* __restrict__ b,
           double* __restrict__ y) {
  __shared__ double L[BD_NB][BD_NB + 1];
  __shared__ double ys[BD_NB];
  const int tid = threadIdx.x;
  for (int e = tid; e < BD_NB * BD_NB; e += 256) {
    const int i = e % BD_NB, j = e / BD_NB;     // column j contiguous in i
    double v = (i == j) ? 1.0 : 0.0;
    if (i < nb && j < nb && j <= i) v = AB[(i - j) + (int64_t)(k + j) * ldab];
    L[i][j] = v;
  }
  __syncthreads();
  if (tid < 32) {
    double bi = (tid < nb) ? b[k + tid] : 0.0;
    double yi = 0.0;
#pragma unroll
    for (int j = 0; j < BD_NB; j++) {
      const double cand = bi / L[tid][tid];
      const double yj = __shfl_sync(0xffffffffu, cand, j);
      if (tid == j) yi = yj;
      if (tid > j) bi = fma(-L[tid][j], yj, bi);
    }
    ys[tid] = yi;
    if (blockIdx.x == 0 && tid < nb) y[k + tid] = yi;
  }
  __syncthreads();
  const int r = blockIdx.x * 256 + tid;
  if (r >= m) return;
  double acc = 0.0;
#pragma unroll 8
  for (int j = 0; j < nb; j++) acc = fma(AB[(nb + r - j) + (int64_t)(k + j) * ldab], ys[j], acc);
  b[k + nb + r] -= acc;
}

// backward step k:  x_k = L11^-T y_k ;  y[j] -= sum_t L(k+t, j) x_t  for k-bw <= j < k
__global__ void __launch_bounds__(256)
k_band_bwd(const double* __restrict__ AB, int ldab, int k, int nb, int m, double* __restrict__ y,
           double* __restrict__ x) {
  __shared__ double L[BD_NB][BD_NB + 1];
  __shared__ double xs[BD_NB];
  const int tid = threadIdx.x;
  for (int e = tid; e < BD_NB * BD_NB; e += 256) {
    const int i = e % BD_NB, j = e / BD_NB;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < nb && j < nb && j <= i) v = AB[(i - j) + (int64_t)(k + j) * ldab];
    L[i][j] = v;
  }
  __syncthreads();
  if (tid < 32) {
    double yi = (tid < nb) ? y[k + tid] : 0.0;
    double xi = 0.0;
#pragma unroll
    for (int j = BD_NB - 1; j >= 0; j--) {
      const double cand = yi / L[tid][tid];
      const double xj = __shfl_sync(0xffffffffu, cand, j);
      if (tid == j) xi = xj;
      if (tid < j) yi = fma(-L[j][tid], xj, yi);     // L^T[tid][j] = L[j][tid]
    }
    xs[tid] = xi;
    if (blockIdx.x == 0 && tid < nb) x[k + tid] = xi;
  }
  __syncthreads();
  const int r = blockIdx.x * 256 + tid;       // column j = k - 1 - r
  if (r >= m) return;
  const int64_t j = (int64_t)k - 1 - r;
  double acc = 0.0;
#pragma unroll 8
  for (int t = 0; t < nb; t++) {
    const int64_t off = k + t - j;            // row - col
    acc = fma(AB[off + j * (int64_t)ldab], xs[t], acc);
  }
  y[j] -= acc;
}

// Solve L L^T x = b with the factor from tg_band_cholesky.  b is overwritten (work), x receives
// the solution; b and x must not alias.  work: n doubles.
extern "C" int tg_band_solve(int64_t n, int32_t bw, int32_t ldab, const double* AB, double* b,
                             double* work, double* x, void* stream) {
  cudaStream_t st = tg_stream(stream);
  double* y = work;
  for (int64_t k = 0; k < n; k += BD_NB) {
    const int nb = (int)((n - k < BD_NB) ? (n - k) : BD_NB);
    int64_t m64 = n - k - nb;
    if (m64 > bw) m64 = bw;
    const int m = (int)m64;
    k_band_fwd<<<(m > 0) ? (int)tg_cdiv(m, 256) : 1, 256, 0, st>>>(AB, ldab, (int)k, nb, m, b, y);
    TG_LAUNCH_CHECK();
  }
  const int64_t klast = ((n - 1) / BD_NB) * BD_NB;
  for (int64_t k = klast; k >= 0; k -= BD_NB) {
    const int nb = (int)((n - k < BD_NB) ? (n - k) : BD_NB);
    // columns j in [k - bw, k) can couple to rows of this block (entries between the band and
    // ldab are stored zeros)
    int64_t m64 = bw;                         // offset k+t-j = t+1+r <= nb+bw-1 <= ldab-1
    if (m64 > k) m64 = k;
    const int m = (int)m64;
    k_band_bwd<<<(m > 0) ? (int)tg_cdiv(m, 256) : 1, 256, 0, st>>>(AB, ldab, (int)k, nb, m, y, x);
    TG_LAUNCH_CHECK();
  }
  return 0;
}
