// Homogeneous Dirichlet BCs and Jacobi-preconditioned CG on CSR.
// Reference: common.py:1199-1200 (zeroRowsColumns), :1154-1158 (vector BCs),
// :1236-1263 (solveLinearSystem; DOLFIN/PETSc solve()).
// All reductions are two-stage with a fixed grid -> deterministic, no atomics.
#include "tg_common.cuh"
#include <stdlib.h>

#define TG_CG_BLOCK 256
static int g_cg_grid = 0;

static int tg_cg_grid_size() {
  if (!g_cg_grid) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_cg_grid = sms * 8;   // 8 CTAs of 256 threads per SM: full occupancy, one wave
  }
  return g_cg_grid;
}

extern "C" int tg_cg_scratch_len(void) { return 2 * tg_cg_grid_size() + 8; }   // >= 2*ws grid

__device__ inline double tg_block_sum(double v, double* sh) {
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
  return r;  // valid in thread 0
}

// out[k] = sum_b part[k*nb + b], k < nk ; one block
__global__ void k_final_reduce(const double* __restrict__ part, int nb, int nk,
                               double* __restrict__ out) {
  __shared__ double sh[32];
  for (int k = 0; k < nk; k++) {
    double v = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) v += part[k * nb + b];
    v = tg_block_sum(v, sh);
    if (threadIdx.x == 0) out[k] = v;
  }
}

__global__ void k_zero_rows_cols(const int64_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ cols, double* __restrict__ vals,
                                 int64_t nrows, const uint8_t* __restrict__ mask, double diag) {
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  bool mr = mask[r] != 0;
  for (int64_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) {
    int c = cols[k];
    if (mr || mask[c]) vals[k] = (mr && c == r) ? diag : 0.0;
  }
}

extern "C" int tg_zero_rows_cols(const int64_t* rowptr, const int32_t* cols, double* vals,
                                 int64_t nrows, const uint8_t* mask, double diag, void* stream) {
  if (nrows == 0) return 0;
  k_zero_rows_cols<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
      rowptr, cols, vals, nrows, mask, diag);
  TG_LAUNCH_CHECK();
  return 0;
}

__global__ void k_zero_entries(double* __restrict__ b, const uint8_t* __restrict__ mask,
                               int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && mask[i]) b[i] = 0.0;
}

extern "C" int tg_zero_entries(double* b, const uint8_t* mask, int64_t n, void* stream) {
  if (n == 0) return 0;
  k_zero_entries<<<(unsigned)tg_cdiv(n, 256), 256, 0, tg_stream(stream)>>>(b, mask, n);
  TG_LAUNCH_CHECK();
  return 0;
}

// zeroDofs list -> 0/1 mask over the IGA DoFs (the list is what the reference
// hands to zeroRowsColumns / setValues, common.py:1154-1158,1199-1200; duplicates
// are harmless)
__global__ void k_mask_set(uint8_t* __restrict__ mask, const int64_t* __restrict__ idx,
                           int64_t nidx, int64_t n) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < nidx) {
    const int64_t i = idx[t];
    if (i >= 0 && i < n) mask[i] = 1;
  }
}

extern "C" int tg_mask_set(uint8_t* mask, const int64_t* idx, int64_t nidx, int64_t n,
                           void* stream) {
  if (nidx == 0) return 0;
  k_mask_set<<<(unsigned)tg_cdiv(nidx, 256), 256, 0, tg_stream(stream)>>>(mask, idx, nidx, n);
  TG_LAUNCH_CHECK();
  return 0;
}

// local row r is global row row0+r (column indices are global/extended)
__global__ void k_diag_inv(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
                           const double* __restrict__ vals, int64_t nrows, int64_t row0,
                           double* __restrict__ dinv) {
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  double d = 0.0;
  for (int64_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32)
    if (cols[k] == row0 + r) d = vals[k];
  d = tg_warp_sum(d);
  if (lane == 0) dinv[r] = (d != 0.0) ? 1.0 / d : 1.0;
}

extern "C" int tg_diag_inv(const int64_t* rowptr, const int32_t* cols, const double* vals,
                           int64_t nrows, int64_t row0, double* dinv, void* stream) {
  if (nrows == 0) return 0;
  k_diag_inv<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
      rowptr, cols, vals, nrows, row0, dinv);
  TG_LAUNCH_CHECK();
  return 0;
}

// y = A x (warp per row, grid-stride); part[b] = sum_{rows of block b} x[xoff+r]*y[r]
__global__ void __launch_bounds__(TG_CG_BLOCK)
k_cg_spmv_dot(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
              const double* __restrict__ vals, const double* __restrict__ x, int64_t xoff,
              double* __restrict__ y, int64_t nrows, double* __restrict__ part) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31;
  const int64_t wpb = TG_CG_BLOCK / 32;
  const int64_t nw = (int64_t)gridDim.x * wpb;
  double dot = 0.0;
  for (int64_t r = blockIdx.x * wpb + (threadIdx.x >> 5); r < nrows; r += nw) {
    int64_t b = rowptr[r], e = rowptr[r + 1];
    double acc = 0.0;
    for (int64_t k = b + lane; k < e; k += 32) acc += vals[k] * x[cols[k]];
    acc = tg_warp_sum(acc);
    if (lane == 0) {
      y[r] = acc;
      dot += x[xoff + r] * acc;
    }
  }
  dot = tg_block_sum(dot, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = dot;
}

extern "C" int tg_cg_spmv_dot(const int64_t* rowptr, const int32_t* cols, const double* vals,
                              const double* x, int64_t xoff, double* y, int64_t nrows,
                              double* scratch, double* out1, void* stream) {
  int g = tg_cg_grid_size();
  k_cg_spmv_dot<<<g, TG_CG_BLOCK, 0, tg_stream(stream)>>>(rowptr, cols, vals, x, xoff, y, nrows,
                                                          scratch);
  TG_LAUNCH_CHECK();
  k_final_reduce<<<1, 256, 0, tg_stream(stream)>>>(scratch, g, 1, out1);
  TG_LAUNCH_CHECK();
  return 0;
}

// a = num/den ; x += a p ; r -= a q ; part0 = sum r*dinv*r ; part1 = sum r*r
__global__ void __launch_bounds__(TG_CG_BLOCK)
k_cg_axpy_dot(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
              const double* __restrict__ q, const double* __restrict__ dinv, int64_t n,
              const double* __restrict__ num, const double* __restrict__ den,
              double* __restrict__ part) {
  __shared__ double sh[32];
  double dn = *den;
  double a = (dn != 0.0) ? (*num) / dn : 0.0;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double pi = p[i];
    x[i] += a * pi;
    double ri = r[i] - a * q[i];
    r[i] = ri;
    s0 += ri * dinv[i] * ri;
    s1 += ri * ri;
  }
  s0 = tg_block_sum(s0, sh);
  s1 = tg_block_sum(s1, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = s0;
    part[gridDim.x + blockIdx.x] = s1;
  }
}

extern "C" int tg_cg_axpy_dot(double* x, double* r, const double* p, const double* q,
                              const double* dinv, int64_t n, const double* num,
                              const double* den, double* scratch, double* out2, void* stream) {
  int g = tg_cg_grid_size();
  k_cg_axpy_dot<<<g, TG_CG_BLOCK, 0, tg_stream(stream)>>>(x, r, p, q, dinv, n, num, den, scratch);
  TG_LAUNCH_CHECK();
  k_final_reduce<<<1, 256, 0, tg_stream(stream)>>>(scratch, g, 2, out2);
  TG_LAUNCH_CHECK();
  return 0;
}

// p = dinv*r + (num/den) p
__global__ void k_cg_xpby(double* __restrict__ p, const double* __restrict__ r,
                          const double* __restrict__ dinv, int64_t n,
                          const double* __restrict__ num, const double* __restrict__ den) {
  double dn = *den;
  double b = (dn != 0.0) ? (*num) / dn : 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = dinv[i] * r[i] + b * p[i];
}

extern "C" int tg_cg_xpby(double* p, const double* r, const double* dinv, int64_t n,
                          const double* num, const double* den, void* stream) {
  int g = tg_cg_grid_size();
  k_cg_xpby<<<g, TG_CG_BLOCK, 0, tg_stream(stream)>>>(p, r, dinv, n, num, den);
  TG_LAUNCH_CHECK();
  return 0;
}

// y += a x  (block-row accumulation y_i = sum_j C_ij x_j of a multi-field system)
__global__ void k_axpy(double* __restrict__ y, double a, const double* __restrict__ x,
                       int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __fma_rn(a, x[i], y[i]);
}

extern "C" int tg_axpy(double* y, double a, const double* x, int64_t n, void* stream) {
  if (n <= 0) return 0;
  int g = tg_cg_grid_size();
  k_axpy<<<g, TG_CG_BLOCK, 0, tg_stream(stream)>>>(y, a, x, n);
  TG_LAUNCH_CHECK();
  return 0;
}

// r = b - y ; p = dinv*r ; part0 = r*dinv*r ; part1 = r*r ; part2.. not used
__global__ void __launch_bounds__(TG_CG_BLOCK)
k_cg_init(const double* __restrict__ b, const double* __restrict__ y,
          const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ p,
          int64_t n, double* __restrict__ part) {
  __shared__ double sh[32];
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double ri = b[i] - y[i];
    r[i] = ri;
    double zi = dinv[i] * ri;
    p[i] = zi;
    s0 += ri * zi;
    s1 += ri * ri;
  }
  s0 = tg_block_sum(s0, sh);
  s1 = tg_block_sum(s1, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = s0;
    part[gridDim.x + blockIdx.x] = s1;
  }
}

extern "C" int tg_cg_init(const double* b, const double* y, const double* dinv, double* r,
                          double* p, int64_t n, double* scratch, double* out2, void* stream) {
  int g = tg_cg_grid_size();
  k_cg_init<<<g, TG_CG_BLOCK, 0, tg_stream(stream)>>>(b, y, dinv, r, p, n, scratch);
  TG_LAUNCH_CHECK();
  k_final_reduce<<<1, 256, 0, tg_stream(stream)>>>(scratch, g, 2, out2);
  TG_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(TG_CG_BLOCK)
k_dot(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
      double* __restrict__ part) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    s += a[i] * b[i];
  s = tg_block_sum(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

extern "C" int tg_dot(const double* a, const double* b, int64_t n, double* scratch,
                      double* out1, void* stream) {
  int g = tg_cg_grid_size();
  k_dot<<<g, TG_CG_BLOCK, 0, tg_stream(stream)>>>(a, b, n, scratch);
  TG_LAUNCH_CHECK();
  k_final_reduce<<<1, 256, 0, tg_stream(stream)>>>(scratch, g, 1, out1);
  TG_LAUNCH_CHECK();
  return 0;
}

// ---- optional live timing of the SpMV launches (bench.py roofline) -----------
static int g_prof_on = 0;
static double g_prof_spmv_ms = 0.0;
static int64_t g_prof_spmv_n = 0;
static cudaEvent_t* g_prof_ev = nullptr;
static int g_prof_cap = 0;

extern "C" void tg_prof_enable(int on) {
  g_prof_on = on;
  g_prof_spmv_ms = 0.0;
  g_prof_spmv_n = 0;
}
extern "C" void tg_prof_get(double* spmv_ms, int64_t* spmv_launches) {
  if (spmv_ms) *spmv_ms = g_prof_spmv_ms;
  if (spmv_launches) *spmv_launches = g_prof_spmv_n;
}
static int tg_prof_reserve(int n) {
  if (n <= g_prof_cap) return 0;
  cudaEvent_t* ev = (cudaEvent_t*)realloc(g_prof_ev, sizeof(cudaEvent_t) * n);
  if (!ev) return 1;
  g_prof_ev = ev;
  for (int i = g_prof_cap; i < n; i++) TG_CHECK(cudaEventCreate(&g_prof_ev[i]));
  g_prof_cap = n;
  return 0;
}
static void tg_prof_collect(int npairs) {
  for (int k = 0; k < npairs; k++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof_ev[2 * k], g_prof_ev[2 * k + 1]) == cudaSuccess) {
      g_prof_spmv_ms += ms;
      g_prof_spmv_n++;
    }
  }
}

// Single-GPU driver shared by the general-CSR and the windowed-CSR solvers.
// work: r[n] p[n] q[n] dinv[n] scratch[tg_cg_scratch_len()] s[8]
// s: 0 rz_a, 1 rr_a, 2 pAp, 3 rz_b, 4 rr_b, 5 bb
// SPMV(x, y, dot_out): y = A x, dot_out[0] = x.y ; DIAG(dinv): dinv = 1/diag(A)
template <class SPMV, class DIAG>
static int tg_cg_driver(SPMV spmv, DIAG diag, const double* b, double* x, int64_t n,
                        double rtol, double atol, int32_t maxit, int32_t check_every,
                        double* work, int32_t* h_iters, double* h_relres, void* stream) {
  cudaStream_t st = tg_stream(stream);
  double* r = work;
  double* p = work + n;
  double* q = work + 2 * n;
  double* dinv = work + 3 * n;
  double* scratch = work + 4 * n;
  double* s = scratch + tg_cg_scratch_len();
  if (check_every < 1) check_every = 1;
  int rc;
  const int prof = g_prof_on;
  if (prof && tg_prof_reserve(2 * check_every)) return 1;
  if ((rc = diag(dinv))) return rc;
  if ((rc = tg_dot(b, b, n, scratch, s + 5, stream))) return rc;
  // q = A x0 ; r = b - q ; p = dinv r
  if ((rc = spmv(x, q, s + 2))) return rc;
  if ((rc = tg_cg_init(b, q, dinv, r, p, n, scratch, s + 0, stream))) return rc;
  double hs[6];
  TG_CHECK(cudaMemcpyAsync(hs, s, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TG_CHECK(cudaStreamSynchronize(st));
  double bb = hs[5];
  double tol2 = rtol * rtol * bb;
  if (atol * atol > tol2) tol2 = atol * atol;
  double rr = hs[1];
  int it = 0;
  while (rr > tol2 && it < maxit) {
    int nstep = check_every;
    if (it + nstep > maxit) nstep = maxit - it;
    for (int k = 0; k < nstep; k++, it++) {
      double* cur = (it & 1) ? s + 3 : s + 0;   // rz, rr of current residual
      double* nxt = (it & 1) ? s + 0 : s + 3;
      if (prof) cudaEventRecord(g_prof_ev[2 * k], st);
      if ((rc = spmv(p, q, s + 2))) return rc;
      if (prof) cudaEventRecord(g_prof_ev[2 * k + 1], st);
      if ((rc = tg_cg_axpy_dot(x, r, p, q, dinv, n, cur, s + 2, scratch, nxt, stream))) return rc;
      if ((rc = tg_cg_xpby(p, r, dinv, n, nxt, cur, stream))) return rc;
    }
    double* last = (it & 1) ? s + 3 : s + 0;
    double hpap = 0.0;
    TG_CHECK(cudaMemcpyAsync(hs, last, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    TG_CHECK(cudaMemcpyAsync(&hpap, s + 2, sizeof(double), cudaMemcpyDeviceToHost, st));
    TG_CHECK(cudaStreamSynchronize(st));
    if (prof) tg_prof_collect(nstep);
    rr = hs[1];
    if (!(rr == rr)) {
      tg_set_error("CG produced NaN at iteration %d", it);
      return 3;
    }
    if (rr > tol2 && !(hpap > 0.0)) {
      tg_set_error("CG breakdown at iteration %d: p.Ap = %g (matrix not symmetric positive "
                   "definite)", it, hpap);
      return 4;
    }
  }
  if (h_iters) *h_iters = it;
  if (h_relres) *h_relres = (bb > 0.0) ? sqrt(rr / bb) : 0.0;
  return 0;
}

extern "C" int tg_solve_cg(const int64_t* rowptr, const int32_t* cols, const double* vals,
                           const double* b, double* x, int64_t n, double rtol, double atol,
                           int32_t maxit, int32_t check_every, double* work, int32_t* h_iters,
                           double* h_relres, void* stream) {
  double* scratch = work + 4 * n;
  return tg_cg_driver(
      [&](const double* xx, double* yy, double* dot) {
        return tg_cg_spmv_dot(rowptr, cols, vals, xx, 0, yy, n, scratch, dot, stream);
      },
      [&](double* dinv) { return tg_diag_inv(rowptr, cols, vals, n, 0, dinv, stream); }, b, x, n,
      rtol, atol, maxit, check_every, work, h_iters, h_relres, stream);
}

// windowed-CSR variant (tg_winops.cu)
int tg_win_spmv_launch(const tg_win* h_w, const double* vals, const double* x, int64_t xoff,
                       double* y, double* part, cudaStream_t st);
int tg_ws_grid_size();

extern "C" int tg_win_spmv_dot(const tg_win* h_w, const double* vals, const double* x,
                               int64_t xoff, double* y, double* scratch, double* out1,
                               void* stream) {
  int rc = tg_win_spmv_launch(h_w, vals, x, xoff, y, scratch, tg_stream(stream));
  if (rc) return rc;
  k_final_reduce<<<1, 256, 0, tg_stream(stream)>>>(
      scratch, (h_w->layout == 1) ? 2 * tg_ws_grid_size() : tg_ws_grid_size(), 1, out1);
  TG_LAUNCH_CHECK();
  return 0;
}

extern "C" int tg_win_solve_cg(const tg_win* h_w, const double* vals, const double* b,
                               double* x, double rtol, double atol, int32_t maxit,
                               int32_t check_every, double* work, int32_t* h_iters,
                               double* h_relres, void* stream) {
  int64_t n = tg_win_nrows(h_w);
  double* scratch = work + 4 * n;
  return tg_cg_driver(
      [&](const double* xx, double* yy, double* dot) {
        return tg_win_spmv_dot(h_w, vals, xx, 0, yy, scratch, dot, stream);
      },
      [&](double* dinv) { return tg_win_diag_inv(h_w, vals, 0, dinv, stream); }, b, x, n, rtol,
      atol, maxit, check_every, work, h_iters, h_relres, stream);
}
