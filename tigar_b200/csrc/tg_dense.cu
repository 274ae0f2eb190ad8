// Dense FP64 building blocks of the solve stage (solve() of common.py:1255-1258):
//   * tg_dgemm_batched  -- column-major strided-batched DGEMM on the FP64 pipe (register-tiled,
//                          double-buffered shared memory); the mode products of the
//                          fast-diagonalisation preconditioner and the trailing update of the
//                          band Cholesky factorisation run on it,
//   * tg_fp64_peak      -- measured DFMA peak of the device (roofline denominator of the
//                          FP64-bound kernels; MEASURED_PEAKS.json only holds HBM and bf16),
//   * tg_fd_*           -- fast diagonalisation: scaling in the generalised eigenbasis,
//   * tg_pcg_*          -- vector kernels of the preconditioned CG driver (tigar_b200/solvers.py).
// There is no FP64 tcgen05 path; FP64 DMMA has no throughput advantage over DFMA on B200.
#include "tg_common.cuh"
#include <math.h>

#define GM_BM 64
#define GM_BN 64
#define GM_BK 16
#define GM_LD (GM_BM + 2)

// C = alpha * op(A) * op(B) + beta * C ; op(A) is M x K, op(B) is K x N.
// lower != 0: only entries with (row >= col) are written (trailing update of a band
// factorisation, where the upper triangle of the dense view aliases other band entries).
template <bool TA, bool TB>
__global__ void __launch_bounds__(256, 2)
k_dgemm(int M, int N, int K, double alpha, const double* __restrict__ A, int lda, long long sA,
        const double* __restrict__ B, int ldb, long long sB, double beta, double* __restrict__ C,
        int ldc, long long sC, int lower) {
  __shared__ __align__(16) double As[2][GM_BK][GM_LD];
  __shared__ __align__(16) double Bs[2][GM_BK][GM_LD];
  const int tid = threadIdx.x;
  const int bm = blockIdx.x * GM_BM, bn = blockIdx.y * GM_BN;
  if (lower && bm + GM_BM - 1 < bn) return;
  A += sA * blockIdx.z;
  B += sB * blockIdx.z;
  C += sC * blockIdx.z;
  const int tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
  double ra[4], rb[4];

  auto gload = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int m, kk;
      if (!TA) { m = tid & 63; kk = (tid >> 6) + 4 * j; }
      else     { kk = tid & 15; m = (tid >> 4) + 16 * j; }
      const int gm = bm + m, gk = k0 + kk;
      ra[j] = (gm < M && gk < K)
                  ? (TA ? A[gk + (long long)gm * lda] : A[gm + (long long)gk * lda]) : 0.0;
      int n, kb;
      if (!TB) { kb = tid & 15; n = (tid >> 4) + 16 * j; }
      else     { n = tid & 63; kb = (tid >> 6) + 4 * j; }
      const int gn = bn + n, gkb = k0 + kb;
      rb[j] = (gn < N && gkb < K)
                  ? (TB ? B[gn + (long long)gkb * ldb] : B[gkb + (long long)gn * ldb]) : 0.0;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int m, kk;
      if (!TA) { m = tid & 63; kk = (tid >> 6) + 4 * j; }
      else     { kk = tid & 15; m = (tid >> 4) + 16 * j; }
      As[buf][kk][m] = ra[j];
      int n, kb;
      if (!TB) { kb = tid & 15; n = (tid >> 4) + 16 * j; }
      else     { n = tid & 63; kb = (tid >> 6) + 4 * j; }
      Bs[buf][kb][n] = rb[j];
    }
  };

  const int nk = (K + GM_BK - 1) / GM_BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int t = 0; t < nk; t++) {
    const int cur = t & 1;
    if (t + 1 < nk) gload((t + 1) * GM_BK);
#pragma unroll
    for (int k = 0; k < GM_BK; k++) {
      const double2 a01 = *reinterpret_cast<const double2*>(&As[cur][k][tx * 4]);
      const double2 a23 = *reinterpret_cast<const double2*>(&As[cur][k][tx * 4 + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&Bs[cur][k][ty * 4]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bs[cur][k][ty * 4 + 2]);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y};
      const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < nk) {
      sstore(cur ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int gn = bn + ty * 4 + j;
    if (gn >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int gm = bm + tx * 4 + i;
      if (gm >= M || (lower && gm < gn)) continue;
      double* c = C + gm + (long long)gn * ldc;
      *c = (beta == 0.0) ? alpha * acc[i][j] : fma(alpha, acc[i][j], beta * (*c));
    }
  }
}

static int tg_dgemm_launch(int ta, int tb, int M, int N, int K, double alpha, const double* A,
                           int lda, long long sA, const double* B, int ldb, long long sB,
                           double beta, double* C, int ldc, long long sC, int batch, int lower,
                           cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  dim3 grid((unsigned)tg_cdiv(M, GM_BM), (unsigned)tg_cdiv(N, GM_BN), (unsigned)batch);
  TG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "dgemm grid too large in y/z");
  if (!ta && !tb)
    k_dgemm<false, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, lower);
  else if (ta && !tb)
    k_dgemm<true, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, lower);
  else if (!ta && tb)
    k_dgemm<false, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, lower);
  else
    k_dgemm<true, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, lower);
  TG_LAUNCH_CHECK();
  return 0;
}

// trailing update of the band factorisation: C (M x M, lower) = alpha * A A^T + beta * C
int tg_dgemm_lower_nt(int M, int K, double alpha, const double* A, int lda, double beta,
                      double* C, int ldc, cudaStream_t st) {
  return tg_dgemm_launch(0, 1, M, M, K, alpha, A, lda, 0, A, lda, 0, beta, C, ldc, 0, 1, 1, st);
}

extern "C" int tg_dgemm_batched(int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K,
                                double alpha, const double* A, int32_t lda, int64_t strideA,
                                const double* B, int32_t ldb, int64_t strideB, double beta,
                                double* C, int32_t ldc, int64_t strideC, int32_t batch,
                                void* stream) {
  // grid.y is limited to 65535 tiles: split very wide N (mode-0 product of a 3-D tensor)
  const int64_t NMAX = 65535LL * GM_BN;
  for (int64_t n0 = 0; n0 < N; n0 += NMAX) {
    const int nn = (int)((N - n0 < NMAX) ? (N - n0) : NMAX);
    const double* Bp = transB ? B + n0 : B + n0 * (int64_t)ldb;
    int rc = tg_dgemm_launch(transA, transB, M, nn, K, alpha, A, lda, strideA, Bp, ldb, strideB,
                             beta, C + n0 * (int64_t)ldc, ldc, strideC, batch, 0,
                             tg_stream(stream));
    if (rc) return rc;
  }
  return 0;
}

// ---- measured DFMA peak ----------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) out[0] = s;      // never true: keeps the chains alive
}

// h_tflops[0] = best-of-5 DFMA rate (2 flops per FMA) of a kernel that does nothing else
extern "C" int tg_fp64_peak(double* scratch1, double* h_tflops, void* stream) {
  cudaStream_t st = tg_stream(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int iters = 1 << 15, grid = sms * 8;
  cudaEvent_t e0, e1;
  TG_CHECK(cudaEventCreate(&e0));
  TG_CHECK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    TG_CHECK(cudaEventRecord(e0, st));
    k_fp64_peak<<<grid, 256, 0, st>>>(scratch1, iters, 1.0 + rep);
    TG_LAUNCH_CHECK();
    TG_CHECK(cudaEventRecord(e1, st));
    TG_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    TG_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 8.0 * iters * 256.0 * grid / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (h_tflops) *h_tflops = best;
  return 0;
}

// ---- fast diagonalisation ---------------------------------------------------------------
// t[pl, i2] /= (sigma + l0[i0] + l1[i1] + l2[i2])^pw for the chunk of mq plane entries starting
// at global plane index q0 (pl = q0 + local, i0 = pl % n0, i1 = pl / n0); the whole tensor is
// q0 = 0, mq = n0*n1.  Non-finite or non-positive sums (constrained hyperplanes carry +inf)
// give 0.
__global__ void k_fd_scale(double* __restrict__ t, const double* __restrict__ l0,
                           const double* __restrict__ l1, const double* __restrict__ l2, int n0,
                           int n2, int64_t q0, int64_t mq, double sigma, int pw) {
  const int64_t n = mq * n2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i2 = i / mq;
    const int64_t pl = q0 + (i - i2 * mq);
    const int64_t i1 = pl / n0;
    const int i0 = (int)(pl - i1 * n0);
    double s = sigma + l0[i0] + (l1 ? l1[i1] : 0.0) + (l2 ? l2[i2] : 0.0);
    if (pw == 2) s = s * s;
    t[i] = (isfinite(s) && s > 0.0) ? t[i] / s : 0.0;
  }
}

extern "C" int tg_fd_scale(double* t, const double* l0, const double* l1, const double* l2,
                           int32_t n0, int32_t n1, int32_t n2, int64_t q0, int64_t mq,
                           double sigma, int32_t pw, void* stream) {
  (void)n1;
  const int64_t n = mq * n2;
  if (n <= 0) return 0;
  int g = (int)((n + 255) / 256);
  if (g > 148 * 16) g = 148 * 16;
  k_fd_scale<<<g, 256, 0, tg_stream(stream)>>>(t, l0, l1, l2, n0, n2, q0, mq, sigma, pw);
  TG_LAUNCH_CHECK();
  return 0;
}

// dst = mask ? 0 : src   (input of the preconditioner: constrained entries do not couple)
__global__ void k_masked_copy(double* __restrict__ dst, const double* __restrict__ src,
                              const uint8_t* __restrict__ mask, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (mask && mask[i]) ? 0.0 : src[i];
}
// z = mask ? r * cinv : z   (constrained rows are diag * identity)
__global__ void k_masked_fix(double* __restrict__ z, const double* __restrict__ r,
                             const uint8_t* __restrict__ mask, double cinv, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    if (mask[i]) z[i] = r[i] * cinv;
}

static int tg_vec_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int tg_masked_copy(double* dst, const double* src, const uint8_t* mask, int64_t n,
                              void* stream) {
  if (n <= 0) return 0;
  k_masked_copy<<<tg_vec_grid(n), 256, 0, tg_stream(stream)>>>(dst, src, mask, n);
  TG_LAUNCH_CHECK();
  return 0;
}
extern "C" int tg_masked_fix(double* z, const double* r, const uint8_t* mask, double cinv,
                             int64_t n, void* stream) {
  if (n <= 0 || !mask) return 0;
  k_masked_fix<<<tg_vec_grid(n), 256, 0, tg_stream(stream)>>>(z, r, mask, cinv, n);
  TG_LAUNCH_CHECK();
  return 0;
}

// out4[a] = sum_{i free} d[i] * b_a[i],  b_a = Kronecker products of the 1-D diagonals:
// a = 0,1,2: k_a (x) m (x) m ;  a = 3: m (x) m (x) m.  Least-squares fit of the
// preconditioner's direction weights to diag(C).  One block (the sums are tiny next to the
// solve); deterministic order.
__global__ void __launch_bounds__(1024)
k_fd_fit(const double* __restrict__ d, const uint8_t* __restrict__ mask,
         const double* __restrict__ kd0, const double* __restrict__ kd1,
         const double* __restrict__ kd2, const double* __restrict__ md0,
         const double* __restrict__ md1, const double* __restrict__ md2, int n0, int n1, int n2,
         double* __restrict__ part) {
  __shared__ double sh[4][32];
  const int64_t n = (int64_t)n0 * n1 * n2;
  double s[4] = {0, 0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (mask && mask[i]) continue;
    const int i0 = (int)(i % n0);
    const int64_t r = i / n0;
    const int i1 = (int)(r % n1), i2 = (int)(r / n1);
    const double m0 = md0[i0], m1 = md1 ? md1[i1] : 1.0, m2 = md2 ? md2[i2] : 1.0;
    const double di = d[i];
    s[0] += di * kd0[i0] * m1 * m2;
    s[1] += md1 ? di * m0 * kd1[i1] * m2 : 0.0;
    s[2] += md2 ? di * m0 * m1 * kd2[i2] : 0.0;
    s[3] += di * m0 * m1 * m2;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    double v = tg_warp_sum(s[a]);
    if (lane == 0) sh[a][wid] = v;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int a = 0; a < 4; a++) {
      double v = (lane < (blockDim.x >> 5)) ? sh[a][lane] : 0.0;
      v = tg_warp_sum(v);
      if (lane == 0) part[a * gridDim.x + blockIdx.x] = v;
    }
  }
}

__global__ void k_fd_fit_final(const double* __restrict__ part, int nb, double* __restrict__ out) {
  const int a = threadIdx.x;
  if (a < 4) {
    double v = 0.0;
    for (int b = 0; b < nb; b++) v += part[a * nb + b];
    out[a] = v;
  }
}

// scratch: 4*64 doubles ; out4: 4 doubles (device)
extern "C" int tg_fd_fit(const double* diagC, const uint8_t* mask, const double* kd0,
                         const double* kd1, const double* kd2, const double* md0,
                         const double* md1, const double* md2, int32_t n0, int32_t n1, int32_t n2,
                         double* scratch, double* out4, void* stream) {
  const int nb = 64;
  k_fd_fit<<<nb, 1024, 0, tg_stream(stream)>>>(diagC, mask, kd0, kd1, kd2, md0, md1, md2, n0, n1,
                                              n2, scratch);
  TG_LAUNCH_CHECK();
  k_fd_fit_final<<<1, 32, 0, tg_stream(stream)>>>(scratch, nb, out4);
  TG_LAUNCH_CHECK();
  return 0;
}

// Relative-error version of the fit: minimise sum_i (sum_a x_a col_a(i) / d_i - 1)^2 over the
// unconstrained DoFs, col_a = Kronecker diagonals (a = 0..2: stiffness in direction a, a = 3:
// mass).  The normal equations are not separable (d_i is not): 10 Gram sums + 4 right-hand
// sides + the count, one pass over diag(C).  part: [15][gridDim.x].
__global__ void __launch_bounds__(1024)
k_fd_fit_rel(const double* __restrict__ d, const uint8_t* __restrict__ mask,
             const double* __restrict__ kd0, const double* __restrict__ kd1,
             const double* __restrict__ kd2, const double* __restrict__ md0,
             const double* __restrict__ md1, const double* __restrict__ md2, int n0, int n1,
             int n2, double* __restrict__ part) {
  __shared__ double sh[15][32];
  const int64_t n = (int64_t)n0 * n1 * n2;
  double s[15];
#pragma unroll
  for (int a = 0; a < 15; a++) s[a] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (mask && mask[i]) continue;
    const double di = d[i];
    if (!(di > 0.0)) continue;
    const int i0 = (int)(i % n0);
    const int64_t r = i / n0;
    const int i1 = (int)(r % n1), i2 = (int)(r / n1);
    const double m0 = md0[i0], m1 = md1 ? md1[i1] : 1.0, m2 = md2 ? md2[i2] : 1.0;
    const double inv = 1.0 / di;
    double u[4];
    u[0] = kd0[i0] * m1 * m2 * inv;
    u[1] = md1 ? m0 * kd1[i1] * m2 * inv : 0.0;
    u[2] = md2 ? m0 * m1 * kd2[i2] * inv : 0.0;
    u[3] = m0 * m1 * m2 * inv;
    int k = 0;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = a; b < 4; b++) s[k++] += u[a] * u[b];
#pragma unroll
    for (int a = 0; a < 4; a++) s[10 + a] += u[a];
    s[14] += 1.0;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 15; a++) {
    double v = tg_warp_sum(s[a]);
    if (lane == 0) sh[a][wid] = v;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int a = 0; a < 15; a++) {
      double v = (lane < (blockDim.x >> 5)) ? sh[a][lane] : 0.0;
      v = tg_warp_sum(v);
      if (lane == 0) part[a * gridDim.x + blockIdx.x] = v;
    }
  }
}

__global__ void k_fd_fit_rel_final(const double* __restrict__ part, int nb,
                                   double* __restrict__ out) {
  const int a = threadIdx.x;
  if (a < 15) {
    double v = 0.0;
    for (int b = 0; b < nb; b++) v += part[a * nb + b];
    out[a] = v;
  }
}

// scratch: 15*64 doubles ; out15 (device): G00 G01 G02 G03 G11 G12 G13 G22 G23 G33, r0..r3, count
extern "C" int tg_fd_fit_rel(const double* diagC, const uint8_t* mask, const double* kd0,
                             const double* kd1, const double* kd2, const double* md0,
                             const double* md1, const double* md2, int32_t n0, int32_t n1,
                             int32_t n2, double* scratch, double* out15, void* stream) {
  const int nb = 64;
  k_fd_fit_rel<<<nb, 1024, 0, tg_stream(stream)>>>(diagC, mask, kd0, kd1, kd2, md0, md1, md2, n0,
                                                  n1, n2, scratch);
  TG_LAUNCH_CHECK();
  k_fd_fit_rel_final<<<1, 32, 0, tg_stream(stream)>>>(scratch, nb, out15);
  TG_LAUNCH_CHECK();
  return 0;
}

// Diagonal scaling of the FD preconditioner, z = S B^-1 S r with S = diag sqrt(B_ii / C_ii):
// B_ii = sigma m0 m1 m2 + sum_d c_d k_d m m is the diagonal of the surrogate.  It follows the
// smooth pointwise variation of the coefficients (weights of rational basis functions,
// Jacobian of a curved map) that the constant direction weights cannot; S = I where the fit
// is exact.  Constrained DoFs get 1.
__global__ void k_fd_diag_scale(const double* __restrict__ d, const uint8_t* __restrict__ mask,
                                const double* __restrict__ kd0, const double* __restrict__ kd1,
                                const double* __restrict__ kd2, const double* __restrict__ md0,
                                const double* __restrict__ md1, const double* __restrict__ md2,
                                double c0, double c1, double c2, double sigma, int n0, int n1,
                                int n2, double* __restrict__ out) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double v = 1.0;
    const double di = d[i];
    if (!(mask && mask[i]) && di > 0.0) {
      const int i0 = (int)(i % n0);
      const int64_t r = i / n0;
      const int i1 = (int)(r % n1), i2 = (int)(r / n1);
      const double m0 = md0[i0], m1 = md1 ? md1[i1] : 1.0, m2 = md2 ? md2[i2] : 1.0;
      double b = sigma * m0 * m1 * m2 + c0 * kd0[i0] * m1 * m2;
      if (md1) b += c1 * m0 * kd1[i1] * m2;
      if (md2) b += c2 * m0 * m1 * kd2[i2];
      if (b > 0.0) v = sqrt(b / di);
    }
    out[i] = v;
  }
}
extern "C" int tg_fd_diag_scale(const double* diagC, const uint8_t* mask, const double* kd0,
                                const double* kd1, const double* kd2, const double* md0,
                                const double* md1, const double* md2, double c0, double c1,
                                double c2, double sigma, int32_t n0, int32_t n1, int32_t n2,
                                double* out, void* stream) {
  const int64_t n = (int64_t)n0 * n1 * n2;
  if (n <= 0) return 0;
  k_fd_diag_scale<<<tg_vec_grid(n), 256, 0, tg_stream(stream)>>>(
      diagC, mask, kd0, kd1, kd2, md0, md1, md2, c0, c1, c2, sigma, n0, n1, n2, out);
  TG_LAUNCH_CHECK();
  return 0;
}

// y = x * s (element-wise; y may alias x)
__global__ void k_vmul(double* __restrict__ y, const double* __restrict__ x,
                       const double* __restrict__ s, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] = x[i] * s[i];
}
extern "C" int tg_vmul(double* y, const double* x, const double* s, int64_t n, void* stream) {
  if (n <= 0) return 0;
  k_vmul<<<tg_vec_grid(n), 256, 0, tg_stream(stream)>>>(y, x, s, n);
  TG_LAUNCH_CHECK();
  return 0;
}

// ---- vector kernels of the preconditioned CG driver -------------------------------------
// p = z + beta p
__global__ void k_xpby(double* __restrict__ p, double beta, const double* __restrict__ z,
                       int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = fma(beta, p[i], z[i]);
}
extern "C" int tg_xpby(double* p, double beta, const double* z, int64_t n, void* stream) {
  if (n <= 0) return 0;
  k_xpby<<<tg_vec_grid(n), 256, 0, tg_stream(stream)>>>(p, beta, z, n);
  TG_LAUNCH_CHECK();
  return 0;
}

// x += a p ; r -= a q ; part[b] = sum r*r   (two-stage, fixed grid: deterministic)
__global__ void __launch_bounds__(256)
k_pcg_update(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
             const double* __restrict__ q, double a, int64_t n, double* __restrict__ part) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = fma(a, p[i], x[i]);
    const double ri = fma(-a, q[i], r[i]);
    r[i] = ri;
    s = fma(ri, ri, s);
  }
  s = tg_warp_sum(s);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = s;
  __syncthreads();
  if (wid == 0) {
    double v = (lane < 8) ? sh[lane] : 0.0;
    v = tg_warp_sum(v);
    if (lane == 0) part[blockIdx.x] = v;
  }
}
__global__ void k_sum_parts(const double* __restrict__ part, int nb, double* __restrict__ out) {
  __shared__ double sh[32];
  double v = 0.0;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) v += part[b];
  v = tg_warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    r = tg_warp_sum(r);
    if (lane == 0) out[0] = r;
  }
}
// scratch: tg_cg_scratch_len() doubles ; out1: device double = r.r after the update
extern "C" int tg_pcg_update(double* x, double* r, const double* p, const double* q, double a,
                             int64_t n, double* scratch, double* out1, void* stream) {
  const int g = tg_vec_grid(n);
  k_pcg_update<<<g, 256, 0, tg_stream(stream)>>>(x, r, p, q, a, n, scratch);
  TG_LAUNCH_CHECK();
  k_sum_parts<<<1, 256, 0, tg_stream(stream)>>>(scratch, g, out1);
  TG_LAUNCH_CHECK();
  return 0;
}
