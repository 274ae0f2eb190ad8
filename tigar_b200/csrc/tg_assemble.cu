// Gauss-point assembly of element matrices/vectors on a tensor-product basis,
// written into a windowed CSR matrix by colouring (no atomics).
// Reference: dolfin::Assembler behind common.py:1215-1216 (matrix) and
// :1169 (vector).  With the extracted basis (tabN) this computes
// sum_e M_e^T K_e M_e directly (the element-fused M^T A M of SURVEY 7.1-5b).
#include "tg_common.cuh"

struct TgAlpha {
  int n;
  signed char al[TG_MAXJET][3];
};

struct TgColour {
  int k[3];       // first cell of this colour per direction
  int cnt[3];     // cells of this colour per direction
  int stride[3];
};

__device__ inline double tg_jet1(const TgBasis& B, const int* e, const int* q, const int* a,
                                 const signed char* al) {
  const int nd = B.nder + 1;
  double v = B.tab[0][(((int64_t)e[0] * B.nq[0] + q[0]) * B.nloc[0] + a[0]) * nd + al[0]];
  if (B.dim > 1) v *= B.tab[1][(((int64_t)e[1] * B.nq[1] + q[1]) * B.nloc[1] + a[1]) * nd + al[1]];
  if (B.dim > 2) v *= B.tab[2][(((int64_t)e[2] * B.nq[2] + q[2]) * B.nloc[2] + a[2]) * nd + al[2]];
  return v;
}

__device__ inline void tg_colour_cell(const TgColour& C, int dim, int64_t bid, int* e) {
  int m0 = (int)(bid % C.cnt[0]);
  int64_t r = bid / C.cnt[0];
  int m1 = (int)(r % C.cnt[1]);
  int m2 = (int)(r / C.cnt[1]);
  e[0] = C.k[0] + C.stride[0] * m0;
  e[1] = (dim > 1) ? C.k[1] + C.stride[1] * m1 : 0;
  e[2] = (dim > 2) ? C.k[2] + C.stride[2] * m2 : 0;
}

template <int NACC>
__global__ void __launch_bounds__(256)
k_assemble_matrix(TgBasis B, TgWin W, TgAlpha S, TgAlpha T, int sameST,
                  const double* __restrict__ coef, int64_t cell0, TgColour C, int QC,
                  double* __restrict__ vals) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nen = B.nloc[0] * B.nloc[1] * B.nloc[2];
  const int nqp = B.nq[0] * B.nq[1] * B.nq[2];
  const int nS = S.n, nT = T.n;
  const int tid = threadIdx.x, nth = blockDim.x;

  int e[3];
  tg_colour_cell(C, B.dim, blockIdx.x, e);
  const int64_t cell = e[0] + (int64_t)B.nel[0] * (e[1] + (int64_t)B.nel[1] * e[2]);
  const double* cc = coef + (cell - cell0) * (int64_t)nS * nT * nqp;

  // smem carve-up
  double* BJS = (double*)smem_raw;                 // [QC][nS][nen]
  double* BJT = sameST ? BJS : BJS + (size_t)QC * nS * nen;   // [QC][nT][nen]
  double* G = BJT + (size_t)QC * nT * nen;         // [QC][nS][nen]
  double* cs = G + (size_t)QC * nS * nen;          // [QC][nS*nT]
  long long* rbase = (long long*)(cs + (size_t)QC * nS * nT);   // [nen]
  int* gc = (int*)(rbase + nen);                   // [nen][3]
  int* rlo = gc + 3 * nen;                         // [nen][3]
  int* rlen = rlo + 3 * nen;                       // [nen][3]
  int* rstride = rlen + 3 * nen;                   // [nen]

  for (int a = tid; a < nen; a += nth) {
    int al[3];
    tg_decode(a, B.nloc, B.dim, al);
    int g[3] = {0, 0, 0}, rr[3] = {0, 0, 0};
    bool valid = true;
    for (int d = 0; d < B.dim; d++) {
      g[d] = B.idx[d][e[d] * B.nloc[d] + al[d]];
      rr[d] = g[d] - W.row0[d];
      valid = valid && rr[d] >= 0 && rr[d] < W.nr[d];
    }
    if (valid) {
      TgRowWin rw = tg_row_window(W, rr);
      TgRowAddr ra = tg_row_addr(W, rr, rw);
      rbase[a] = ra.base;
      rstride[a] = ra.stride;
      for (int d = 0; d < 3; d++) {
        rlo[3 * a + d] = rw.lo[d];
        rlen[3 * a + d] = rw.len[d];
      }
      rlo[3 * a + 0] = ra.lo0;
      rlen[3 * a + 0] = ra.len0;
    } else {
      rbase[a] = -1;
    }
    for (int d = 0; d < 3; d++) gc[3 * a + d] = g[d] - W.col0[d];
  }

  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = 0.0;

  for (int q0 = 0; q0 < nqp; q0 += QC) {
    const int nqc = min(QC, nqp - q0);
    __syncthreads();
    for (int i = tid; i < nqc * nS * nen; i += nth) {
      int a = i % nen;
      int t = i / nen;
      int s = t % nS;
      int ql = t / nS;
      int q[3], al[3];
      tg_decode(q0 + ql, B.nq, B.dim, q);
      tg_decode(a, B.nloc, B.dim, al);
      BJS[i] = tg_jet1(B, e, q, al, S.al[s]);
    }
    if (!sameST) {
      for (int i = tid; i < nqc * nT * nen; i += nth) {
        int a = i % nen;
        int t = i / nen;
        int s = t % nT;
        int ql = t / nT;
        int q[3], al[3];
        tg_decode(q0 + ql, B.nq, B.dim, q);
        tg_decode(a, B.nloc, B.dim, al);
        BJT[i] = tg_jet1(B, e, q, al, T.al[s]);
      }
    }
    for (int i = tid; i < nqc * nS * nT; i += nth) {
      int st = i % (nS * nT);
      int ql = i / (nS * nT);
      cs[i] = cc[(int64_t)st * nqp + q0 + ql];
    }
    __syncthreads();
    for (int i = tid; i < nqc * nS * nen; i += nth) {
      int b = i % nen;
      int t = i / nen;
      int s = t % nS;
      int ql = t / nS;
      double g = 0.0;
      for (int tt = 0; tt < nT; tt++)
        g += cs[(ql * nS + s) * nT + tt] * BJT[(ql * nT + tt) * nen + b];
      G[i] = g;
    }
    __syncthreads();
    const int nk = nqc * nS;
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      int pair = tid + i * nth;
      if (pair < nen * nen) {
        int a = pair / nen, b = pair - a * nen;
        double s = 0.0;
        for (int k = 0; k < nk; k++) s += BJS[k * nen + a] * G[k * nen + b];
        acc[i] += s;
      }
    }
  }

#pragma unroll
  for (int i = 0; i < NACC; i++) {
    int pair = tid + i * nth;
    if (pair < nen * nen) {
      int a = pair / nen, b = pair - a * nen;
      if (rbase[a] < 0) continue;
      int pos = ((gc[3 * b + 2] - rlo[3 * a + 2]) * rlen[3 * a + 1] +
                 (gc[3 * b + 1] - rlo[3 * a + 1])) * rlen[3 * a + 0] +
                (gc[3 * b + 0] - rlo[3 * a + 0]);
      vals[rbase[a] + (long long)pos * rstride[a]] += acc[i];
    }
  }
}

static int tg_colour_setup(const tg_basis* h_B, const int32_t* h_stride, int64_t cell0,
                           int64_t ncells, int* elo, int* ehi) {
  int dim = h_B->dim;
  int64_t slab = 1;
  for (int d = 0; d < dim - 1; d++) slab *= h_B->nel[d];
  TG_REQUIRE(cell0 % slab == 0 && ncells % slab == 0,
             "cell range must be whole slabs of the last direction");
  *elo = (int)(cell0 / slab);
  *ehi = (int)((cell0 + ncells) / slab);
  TG_REQUIRE(*ehi <= h_B->nel[dim - 1], "cell range exceeds patch");
  for (int d = 0; d < dim; d++) TG_REQUIRE(h_stride[d] >= 1, "colour stride");
  return 0;
}

// iterate colours; fn(colour) launches one kernel
template <class F>
static int tg_for_colours(const tg_basis* h_B, const int32_t* h_stride, int elo, int ehi, F fn) {
  int dim = h_B->dim;
  int st[3] = {1, 1, 1}, lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  for (int d = 0; d < dim; d++) {
    st[d] = h_stride[d];
    hi[d] = h_B->nel[d];
  }
  lo[dim - 1] = elo;
  hi[dim - 1] = ehi;
  for (int k2 = 0; k2 < st[2]; k2++)
    for (int k1 = 0; k1 < st[1]; k1++)
      for (int k0 = 0; k0 < st[0]; k0++) {
        TgColour C;
        int kk[3] = {k0, k1, k2};
        int64_t n = 1;
        for (int d = 0; d < 3; d++) {
          // first cell >= lo[d] congruent to kk[d] mod st[d]
          int first = lo[d] + ((kk[d] - lo[d]) % st[d] + st[d]) % st[d];
          C.k[d] = first;
          C.stride[d] = st[d];
          C.cnt[d] = (first < hi[d]) ? (hi[d] - first + st[d] - 1) / st[d] : 0;
          n *= C.cnt[d];
        }
        if (n == 0) continue;
        int rc = fn(C, n);
        if (rc) return rc;
      }
  return 0;
}

static int g_smem_optin = -1;
static int tg_smem_limit() {
  if (g_smem_optin < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  return g_smem_optin;
}

extern "C" int tg_assemble_matrix_ex(const tg_basis* h_B, const tg_win* h_W, int32_t nS,
                                     const int32_t* h_alphaS, int32_t nT,
                                     const int32_t* h_alphaT, const int32_t* h_stride,
                                     const double* coef, int64_t cell0, int64_t ncells,
                                     double* vals, void* stream) {
  TG_REQUIRE(nS >= 1 && nS <= TG_MAXJET && nT >= 1 && nT <= TG_MAXJET, "jet count");
  TgBasis B = tg_basis_dev(h_B);
  TgWin W = tg_win_dev(h_W);
  TgAlpha S, T;
  S.n = nS;
  T.n = nT;
  int same = (nS == nT);
  for (int s = 0; s < nS; s++)
    for (int d = 0; d < 3; d++) {
      S.al[s][d] = (signed char)h_alphaS[3 * s + d];
      TG_REQUIRE(h_alphaS[3 * s + d] <= h_B->nder, "derivative order not tabulated");
    }
  for (int s = 0; s < nT; s++)
    for (int d = 0; d < 3; d++) {
      T.al[s][d] = (signed char)h_alphaT[3 * s + d];
      TG_REQUIRE(h_alphaT[3 * s + d] <= h_B->nder, "derivative order not tabulated");
      if (same && T.al[s][d] != S.al[s][d]) same = 0;
    }
  int elo, ehi;
  int rc = tg_colour_setup(h_B, h_stride, cell0, ncells, &elo, &ehi);
  if (rc) return rc;
  const int nen = B.nloc[0] * B.nloc[1] * B.nloc[2];
  const int nqp = B.nq[0] * B.nq[1] * B.nq[2];
  const int nth = 256;
  // q-chunk so that smem fits (target <= ~100 KB to keep 2 CTAs/SM when possible)
  size_t fixed = (size_t)nen * (8 + 10 * 4) + 64;
  size_t perq = ((size_t)nS * nen * (same ? 2 : 1) + (same ? 0 : (size_t)nT * nen) +
                 (size_t)nS * nen * (same ? 0 : 1) + (size_t)nS * nT) * 8;
  // (same: BJS + G ; distinct: BJS + BJT + G)
  perq = ((size_t)nS * nen + (same ? 0 : (size_t)nT * nen) + (size_t)nS * nen + (size_t)nS * nT) * 8;
  size_t budget = 100 * 1024;
  int limit = tg_smem_limit();
  int QC = (int)((budget - fixed) / perq);
  if (QC < 1) {
    QC = (int)(((size_t)limit - fixed) / perq);
    TG_REQUIRE(QC >= 1, "element too large for shared memory");
  }
  if (QC > nqp) QC = nqp;
  size_t smem = fixed + perq * QC;
  int npair = nen * nen;
  int nacc = (npair + nth - 1) / nth;
  cudaStream_t s = tg_stream(stream);

#define TG_LAUNCH_ASM(N)                                                                      \
  {                                                                                           \
    TG_CHECK(cudaFuncSetAttribute(k_assemble_matrix<N>,                                       \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    rc = tg_for_colours(h_B, h_stride, elo, ehi, [&](const TgColour& C, int64_t n) -> int {   \
      k_assemble_matrix<N><<<(unsigned)n, nth, smem, s>>>(B, W, S, T, same, coef, cell0, C,   \
                                                          QC, vals);                          \
      TG_LAUNCH_CHECK();                                                                      \
      return 0;                                                                               \
    });                                                                                       \
  }
  if (nacc <= 1) TG_LAUNCH_ASM(1)
  else if (nacc <= 2) TG_LAUNCH_ASM(2)
  else if (nacc <= 4) TG_LAUNCH_ASM(4)
  else if (nacc <= 8) TG_LAUNCH_ASM(8)
  else if (nacc <= 16) TG_LAUNCH_ASM(16)
  else if (nacc <= 32) TG_LAUNCH_ASM(32)
  else if (nacc <= 64) TG_LAUNCH_ASM(64)
  else if (nacc <= 184) TG_LAUNCH_ASM(184)
  else {
    tg_set_error("element with %d local functions is too large", nen);
    return 2;
  }
#undef TG_LAUNCH_ASM
  return rc;
}

extern "C" int tg_assemble_matrix(const tg_basis* h_B, const tg_win* h_W, int32_t nS,
                                  const int32_t* h_alphaS, int32_t nT, const int32_t* h_alphaT,
                                  const double* coef, int64_t cell0, int64_t ncells,
                                  double* vals, void* stream) {
  int32_t stride[3];
  for (int d = 0; d < 3; d++) stride[d] = (d < h_B->dim) ? h_B->nloc[d] : 1;
  return tg_assemble_matrix_ex(h_B, h_W, nS, h_alphaS, nT, h_alphaT, stride, coef, cell0,
                               ncells, vals, stream);
}

// ---------------------------------------------------------------------------
// Sum-factorised element matrices, 3-D, NL local functions and NQ Gauss points
// per direction.  One CTA per cell, NL^4 threads: thread = (b1,a1,b2,a2) owns
// the NL x NL block of (a3,b3) entries.  Per term (D^s v, D^t u, coefficient G):
//   T1[a1,b1,q2,q3] = sum_q1 S1[q1,a1] T1[q1,b1] G[q1,q2,q3]      (shared)
//   t2[q3]          = sum_q2 S2[q2,a2] T2[q2,b2] T1[a1,b1,q2,q3]  (registers)
//   acc[a3][b3]    += sum_q3 S3[q3,a3] T3[q3,b3] t2[q3]
// 3*NL^2*NQ^2 ... ~100 DFMA per thread per term instead of NQ^3 per entry.
struct TgTerms {
  int n;
  int nslots;
  short slot[TG_MAXJET * TG_MAXJET];
  signed char aS[TG_MAXJET * TG_MAXJET][3];
  signed char aT[TG_MAXJET * TG_MAXJET][3];
};

__device__ __forceinline__ int l1i(int a, int NL) { return (a / NL) % NL; }
__device__ __forceinline__ int l2i(int a, int NL) { return a / (NL * NL); }

template <int NL, int NQ>
__global__ void __launch_bounds__(NL* NL* NL* NL, (NL == 4) ? 3 : 1)
k_assemble_matrix_sf3(TgBasis B, TgWin W, TgTerms TT, const double* __restrict__ coef,
                      int64_t cell0, TgColour C, double* __restrict__ vals) {
  constexpr int NT = NL * NL * NL * NL;
  constexpr int NQP = NQ * NQ * NQ;
  constexpr int NEN = NL * NL * NL;
  constexpr int MAXD = 5;                       // derivative orders 0..4
  extern __shared__ double Gall[];               // [nterms][NQP]
  __shared__ double T1[2 * NQ * NQ * NL * NL];  // 2 x [q3][q2][a1][b1]
  __shared__ double tabs[3][MAXD][NQ][NL];      // [d][k][q][a]
  __shared__ long long rbase[NEN];
  __shared__ int gidx[3][NL], rlo[3][NL], rlen[3][NL], ridx[3][NL];
  __shared__ int rstr;

  const int tid = threadIdx.x;
  const int b1 = tid % NL, a1 = (tid / NL) % NL, b2 = (tid / (NL * NL)) % NL,
            a2 = tid / (NL * NL * NL);
  int e[3];
  tg_colour_cell(C, 3, blockIdx.x, e);
  const int64_t cell = e[0] + (int64_t)B.nel[0] * (e[1] + (int64_t)B.nel[1] * e[2]);
  const double* cc = coef + (cell - cell0) * (int64_t)TT.nslots * NQP;
  const int nd = B.nder + 1;

  for (int i = tid; i < TT.n * NQP; i += NT) {
    int term = i / NQP;
    Gall[i] = __ldcs(cc + (int64_t)TT.slot[term] * NQP + (i - term * NQP));
  }
  for (int i = tid; i < 3 * nd * NQ * NL; i += NT) {
    int a = i % NL, t = i / NL, q = t % NQ;
    t /= NQ;
    int k = t % nd, d = t / nd;
    tabs[d][k][q][a] = B.tab[d][(((int64_t)e[d] * NQ + q) * NL + a) * nd + k];
  }
  if (tid < 3 * NL) {
    int d = tid / NL, a = tid % NL;
    int g = B.idx[d][e[d] * NL + a];
    gidx[d][a] = g - W.col0[d];              // column coordinate (local)
    int r = g - W.row0[d];                   // row coordinate (local), -1 if not owned
    if (r < 0 || r >= W.nr[d]) r = -1;
    ridx[d][a] = r;
    int lo = (r >= 0) ? W.lo[d][r] : 0;
    rlo[d][a] = lo;
    rlen[d][a] = (r >= 0) ? W.hi[d][r] - lo + 1 : 1;
  }
  __syncthreads();
  for (int a = tid; a < NEN; a += NT) {
    int l0 = a % NL, l1 = (a / NL) % NL, l2 = a / (NL * NL);
    int r0 = ridx[0][l0], r1 = ridx[1][l1], r2 = ridx[2][l2];
    if (r0 < 0 || r1 < 0 || r2 < 0) {
      rbase[a] = -1;
    } else {
      const long long l1 = rlen[1][l1i(a, NL)], l2 = rlen[2][l2i(a, NL)];
      const long long T1 = W.S[1][W.nr[1]];
      if (W.layout == 0) {
        // rowptr in closed form from the 1-D prefix sums (L1/L2 resident):
        // S0[r0]*len1*len2 + T0*(S1[r1]*len2 + T1*S2[r2])
        const long long T0 = W.S[0][W.nr[0]];
        rbase[a] = W.S[0][r0] * l1 * l2 + T0 * (W.S[1][r1] * l2 + T1 * W.S[2][r2]);
      } else {
        const int H = W.H, chunk = r0 / H;
        const long long nchunk = (W.nr[0] + H - 1) / H;
        rbase[a] = (long long)H * nchunk * W.w0max * (W.S[1][r1] * l2 + T1 * W.S[2][r2]) +
                   (long long)chunk * H * (W.w0max * l1 * l2) + (r0 - chunk * H);
      }
    }
  }
  if (tid == 0) rstr = (W.layout == 0) ? 1 : W.H;
  if (W.layout != 0 && tid < NL) {        // SELL: uniform band in the first direction
    const int r = ridx[0][tid];
    if (r >= 0) {
      rlo[0][tid] = W.bs0[r];
      rlen[0][tid] = W.w0max;
    }
  }

  double acc[NL][NL];
#pragma unroll
  for (int i = 0; i < NL; i++)
#pragma unroll
    for (int j = 0; j < NL; j++) acc[i][j] = 0.0;

  // Terms arrive sorted by their last-direction orders (aS[2], aT[2]); terms of
  // one group share the stage-3 tables, so their t2[] are summed first and the
  // most expensive contraction runs once per group instead of once per term.
  double t2[NQ];
#pragma unroll
  for (int q3 = 0; q3 < NQ; q3++) t2[q3] = 0.0;
  for (int term = 0; term < TT.n; term++) {
    const int s0 = TT.aS[term][0], s1 = TT.aS[term][1], s2 = TT.aS[term][2];
    const int t0 = TT.aT[term][0], t1 = TT.aT[term][1], t2o = TT.aT[term][2];
    const double* G = Gall + term * NQP;
    double* T1b = T1 + (term & 1) * (NQ * NQ * NL * NL);       // double buffered
    if (term == 0) __syncthreads();                            // Gall, tabs loaded
    for (int o = tid; o < NQ * NQ * NL * NL; o += NT) {
      int ob1 = o % NL, oa1 = (o / NL) % NL, q23 = o / (NL * NL);
      double v = 0.0;
#pragma unroll
      for (int q1 = 0; q1 < NQ; q1++)
        v += tabs[0][s0][q1][oa1] * tabs[0][t0][q1][ob1] * G[q23 * NQ + q1];
      T1b[o] = v;
    }
    __syncthreads();
    double w2[NQ];
#pragma unroll
    for (int q2 = 0; q2 < NQ; q2++) w2[q2] = tabs[1][s1][q2][a2] * tabs[1][t1][q2][b2];
#pragma unroll
    for (int q3 = 0; q3 < NQ; q3++) {
      double v = t2[q3];
#pragma unroll
      for (int q2 = 0; q2 < NQ; q2++) v += w2[q2] * T1b[((q3 * NQ + q2) * NL + a1) * NL + b1];
      t2[q3] = v;
    }
    const bool last = (term + 1 == TT.n) || TT.aS[term + 1][2] != s2 || TT.aT[term + 1][2] != t2o;
    if (!last) continue;
    double vv[NL][NQ];
#pragma unroll
    for (int b3 = 0; b3 < NL; b3++)
#pragma unroll
      for (int q3 = 0; q3 < NQ; q3++) vv[b3][q3] = tabs[2][t2o][q3][b3];
#pragma unroll
    for (int a3 = 0; a3 < NL; a3++) {
      double w[NQ];
#pragma unroll
      for (int q3 = 0; q3 < NQ; q3++) w[q3] = t2[q3] * tabs[2][s2][q3][a3];
#pragma unroll
      for (int b3 = 0; b3 < NL; b3++) {
        double v = acc[a3][b3];
#pragma unroll
        for (int q3 = 0; q3 < NQ; q3++) v += w[q3] * vv[b3][q3];
        acc[a3][b3] = v;
      }
    }
#pragma unroll
    for (int q3 = 0; q3 < NQ; q3++) t2[q3] = 0.0;
  }

  // scatter: rows a = (a1,a2,a3), columns b = (b1,b2,b3)
  const int d0 = gidx[0][b1] - rlo[0][a1];
  const int d1 = gidx[1][b2] - rlo[1][a2];
  const int len0 = rlen[0][a1], len1 = rlen[1][a2];
  double* ptr[NL][NL];
#pragma unroll
  for (int a3 = 0; a3 < NL; a3++) {
    const int64_t base = rbase[a1 + NL * (a2 + NL * a3)];
#pragma unroll
    for (int b3 = 0; b3 < NL; b3++) {
      const int d2 = gidx[2][b3] - rlo[2][a3];
      ptr[a3][b3] =
          (base >= 0) ? vals + base + (((int64_t)d2 * len1 + d1) * len0 + d0) * rstr : nullptr;
    }
  }
  // colouring guarantees exclusive ownership: all loads first, then all stores
#pragma unroll
  for (int a3 = 0; a3 < NL; a3++)
#pragma unroll
    for (int b3 = 0; b3 < NL; b3++)
      if (ptr[a3][b3]) acc[a3][b3] += *ptr[a3][b3];
#pragma unroll
  for (int a3 = 0; a3 < NL; a3++)
#pragma unroll
    for (int b3 = 0; b3 < NL; b3++)
      if (ptr[a3][b3]) *ptr[a3][b3] = acc[a3][b3];
}

extern "C" int tg_assemble_sf_supported(const tg_basis* h_B) {
  if (h_B->dim != 3 || h_B->nder > 4) return 0;
  int nl = h_B->nloc[0], nq = h_B->nq[0];
  for (int d = 1; d < 3; d++)
    if (h_B->nloc[d] != nl || h_B->nq[d] != nq) return 0;
  return (nl == nq && (nl == 3 || nl == 4 || nl == 5)) ? 1 : 0;
}

extern "C" int tg_assemble_matrix_terms(const tg_basis* h_B, const tg_win* h_W, int32_t nterms,
                                        const int32_t* h_terms, int32_t nslots,
                                        const int32_t* h_stride, const double* coef,
                                        int64_t cell0, int64_t ncells, double* vals,
                                        void* stream) {
  TG_REQUIRE(tg_assemble_sf_supported(h_B), "basis not supported by the sum-factorised kernel");
  TG_REQUIRE(nterms >= 0 && nterms <= TG_MAXJET * TG_MAXJET, "term count");
  if (nterms == 0 || ncells == 0) return 0;
  TgBasis B = tg_basis_dev(h_B);
  TgWin W = tg_win_dev(h_W);
  TgTerms TT;
  TT.n = nterms;
  TT.nslots = nslots;
  for (int i = 0; i < nterms; i++) {
    const int32_t* t = h_terms + 7 * i;
    TG_REQUIRE(t[0] >= 0 && t[0] < nslots, "term slot");
    TT.slot[i] = (short)t[0];
    for (int d = 0; d < 3; d++) {
      TG_REQUIRE(t[1 + d] >= 0 && t[1 + d] <= h_B->nder && t[4 + d] >= 0 && t[4 + d] <= h_B->nder,
                 "derivative order not tabulated");
      TT.aS[i][d] = (signed char)t[1 + d];
      TT.aT[i][d] = (signed char)t[4 + d];
    }
  }
  int elo, ehi;
  int rc = tg_colour_setup(h_B, h_stride, cell0, ncells, &elo, &ehi);
  if (rc) return rc;
  cudaStream_t s = tg_stream(stream);
  const int nl = h_B->nloc[0];
  const size_t gsm = (size_t)nterms * nl * nl * nl * sizeof(double);
  TG_REQUIRE(gsm <= 40 * 1024, "too many terms for the coefficient tile");
#define TG_LAUNCH_SF(N)                                                                        \
  rc = tg_for_colours(h_B, h_stride, elo, ehi, [&](const TgColour& C, int64_t n) -> int {      \
    k_assemble_matrix_sf3<N, N><<<(unsigned)n, N * N * N * N, gsm, s>>>(B, W, TT, coef, cell0,  \
                                                                        C, vals);              \
    TG_LAUNCH_CHECK();                                                                         \
    return 0;                                                                                  \
  });
  if (nl == 3) { TG_LAUNCH_SF(3) }
  else if (nl == 4) { TG_LAUNCH_SF(4) }
  else { TG_LAUNCH_SF(5) }
#undef TG_LAUNCH_SF
  return rc;
}

// element vector, sum-factorised through shared memory: one CTA per cell.
//   u1[a0,q1,q2] = sum_q0 tab0[q0][a0][s0] c[q0,q1,q2]
//   u2[a0,a1,q2] = sum_q1 tab1[q1][a1][s1] u1[a0,q1,q2]
//   b[a0,a1,a2] += sum_q2 tab2[q2][a2][s2] u2[a0,a1,q2]
struct TgSlots {
  int nslots;
  short slot[TG_MAXJET];
  int row0[3], nr[3];       // slab of the global vector that is written
};

__global__ void k_assemble_vector(TgBasis B, TgAlpha S, TgSlots SL, const double* __restrict__ coef,
                                  int64_t cell0, TgColour C, double* __restrict__ bvec) {
  extern __shared__ double smv[];
  const int n0 = B.nloc[0], n1 = B.nloc[1], n2 = B.nloc[2];
  const int q0n = B.nq[0], q1n = B.nq[1], q2n = B.nq[2];
  const int nen = n0 * n1 * n2;
  const int nqp = q0n * q1n * q2n;
  const int nd = B.nder + 1;
  const int tid = threadIdx.x, nth = blockDim.x;
  int e[3];
  tg_colour_cell(C, B.dim, blockIdx.x, e);
  const int64_t cell = e[0] + (int64_t)B.nel[0] * (e[1] + (int64_t)B.nel[1] * e[2]);
  const double* cc = coef + (cell - cell0) * (int64_t)SL.nslots * nqp;
  double* tb0 = smv;
  double* tb1 = tb0 + q0n * n0 * nd;
  double* tb2 = tb1 + q1n * n1 * nd;
  double* cq = tb2 + q2n * n2 * nd;        // [nqp]
  double* u1 = cq + nqp;                   // [q2][q1][a0]
  double* u2 = u1 + n0 * q1n * q2n;        // [q2][a1][a0]
  double* acc = u2 + n0 * n1 * q2n;        // [nen]
  for (int i = tid; i < q0n * n0 * nd; i += nth) tb0[i] = B.tab[0][(int64_t)e[0] * q0n * n0 * nd + i];
  for (int i = tid; i < q1n * n1 * nd; i += nth)
    tb1[i] = (B.dim > 1) ? B.tab[1][(int64_t)e[1] * q1n * n1 * nd + i] : 1.0;
  for (int i = tid; i < q2n * n2 * nd; i += nth)
    tb2[i] = (B.dim > 2) ? B.tab[2][(int64_t)e[2] * q2n * n2 * nd + i] : 1.0;
  for (int a = tid; a < nen; a += nth) acc[a] = 0.0;
  for (int s = 0; s < S.n; s++) {
    const int s0 = S.al[s][0], s1 = (B.dim > 1) ? S.al[s][1] : 0, s2 = (B.dim > 2) ? S.al[s][2] : 0;
    __syncthreads();
    for (int i = tid; i < nqp; i += nth) cq[i] = cc[(int64_t)SL.slot[s] * nqp + i];
    __syncthreads();
    for (int o = tid; o < n0 * q1n * q2n; o += nth) {
      int a0 = o % n0, r = o / n0;                     // r = q2*q1n + q1
      double v = 0.0;
      for (int q = 0; q < q0n; q++) v += tb0[(q * n0 + a0) * nd + s0] * cq[r * q0n + q];
      u1[o] = v;
    }
    __syncthreads();
    for (int o = tid; o < n0 * n1 * q2n; o += nth) {
      int a0 = o % n0, t = o / n0, a1 = t % n1, q2 = t / n1;
      double v = 0.0;
      for (int q = 0; q < q1n; q++) v += tb1[(q * n1 + a1) * nd + s1] * u1[(q2 * q1n + q) * n0 + a0];
      u2[o] = v;
    }
    __syncthreads();
    for (int a = tid; a < nen; a += nth) {
      int a01 = a % (n0 * n1), a2 = a / (n0 * n1);
      double v = 0.0;
      for (int q = 0; q < q2n; q++) v += tb2[(q * n2 + a2) * nd + s2] * u2[q * n0 * n1 + a01];
      acc[a] += v;
    }
  }
  __syncthreads();
  for (int a = tid; a < nen; a += nth) {
    int al[3];
    tg_decode(a, B.nloc, B.dim, al);
    int64_t g = 0, mul = 1;
    bool valid = true;
    for (int d = 0; d < B.dim; d++) {
      int r = B.idx[d][e[d] * B.nloc[d] + al[d]] - SL.row0[d];
      valid = valid && r >= 0 && r < SL.nr[d];
      g += mul * r;
      mul *= SL.nr[d];
    }
    if (valid) bvec[g] += acc[a];
  }
}

extern "C" int tg_assemble_vector_slots(const tg_basis* h_B, int32_t nS, const int32_t* h_alphaS,
                                        const int32_t* h_slots, int32_t nslots,
                                        const int32_t* h_stride, const double* coef,
                                        int64_t cell0, int64_t ncells, double* b, void* stream);

extern "C" int tg_assemble_vector_ex(const tg_basis* h_B, int32_t nS, const int32_t* h_alphaS,
                                     const int32_t* h_stride, const double* coef,
                                     int64_t cell0, int64_t ncells, double* b, void* stream) {
  int32_t slots[TG_MAXJET];
  for (int i = 0; i < TG_MAXJET; i++) slots[i] = i;
  return tg_assemble_vector_slots(h_B, nS, h_alphaS, slots, nS, h_stride, coef, cell0, ncells, b,
                                  stream);
}

extern "C" int tg_assemble_vector_slots(const tg_basis* h_B, int32_t nS, const int32_t* h_alphaS,
                                        const int32_t* h_slots, int32_t nslots,
                                        const int32_t* h_stride, const double* coef,
                                        int64_t cell0, int64_t ncells, double* b, void* stream) {
  int32_t row0[3] = {0, 0, 0}, nr[3] = {1, 1, 1};
  for (int d = 0; d < h_B->dim; d++) nr[d] = h_B->n[d];
  return tg_assemble_vector_part(h_B, nS, h_alphaS, h_slots, nslots, h_stride, row0, nr, coef,
                                 cell0, ncells, b, stream);
}

extern "C" int tg_assemble_vector_part(const tg_basis* h_B, int32_t nS, const int32_t* h_alphaS,
                                       const int32_t* h_slots, int32_t nslots,
                                       const int32_t* h_stride, const int32_t* h_row0,
                                       const int32_t* h_nr, const double* coef, int64_t cell0,
                                       int64_t ncells, double* b, void* stream) {
  TG_REQUIRE(nS >= 1 && nS <= TG_MAXJET, "jet count");
  TgSlots SL;
  SL.nslots = nslots;
  for (int d = 0; d < 3; d++) {
    SL.row0[d] = (d < h_B->dim) ? h_row0[d] : 0;
    SL.nr[d] = (d < h_B->dim) ? h_nr[d] : 1;
  }
  for (int i = 0; i < nS; i++) {
    TG_REQUIRE(h_slots[i] >= 0 && h_slots[i] < nslots, "slot index");
    SL.slot[i] = (short)h_slots[i];
  }
  TgBasis B = tg_basis_dev(h_B);
  TgAlpha S;
  S.n = nS;
  for (int s = 0; s < nS; s++)
    for (int d = 0; d < 3; d++) {
      S.al[s][d] = (signed char)h_alphaS[3 * s + d];
      TG_REQUIRE(h_alphaS[3 * s + d] <= h_B->nder, "derivative order not tabulated");
    }
  int elo, ehi;
  int rc = tg_colour_setup(h_B, h_stride, cell0, ncells, &elo, &ehi);
  if (rc) return rc;
  const int nen = B.nloc[0] * B.nloc[1] * B.nloc[2];
  int nth = nen < 32 ? 32 : (nen > 256 ? 256 : ((nen + 31) / 32) * 32);
  const int nd = B.nder + 1;
  size_t smem = 0;
  for (int d = 0; d < 3; d++) smem += (size_t)B.nq[d] * B.nloc[d] * nd;
  smem += (size_t)B.nq[0] * B.nq[1] * B.nq[2];
  smem += (size_t)B.nloc[0] * B.nq[1] * B.nq[2];
  smem += (size_t)B.nloc[0] * B.nloc[1] * B.nq[2];
  smem += (size_t)nen;
  smem *= sizeof(double);
  TG_REQUIRE(smem <= 48 * 1024, "element too large for the vector-assembly tiles");
  cudaStream_t s = tg_stream(stream);
  return tg_for_colours(h_B, h_stride, elo, ehi, [&](const TgColour& C, int64_t n) -> int {
    k_assemble_vector<<<(unsigned)n, nth, smem, s>>>(B, S, SL, coef, cell0, C, b);
    TG_LAUNCH_CHECK();
    return 0;
  });
}

extern "C" int tg_assemble_vector(const tg_basis* h_B, int32_t nS, const int32_t* h_alphaS,
                                  const double* coef, int64_t cell0, int64_t ncells, double* b,
                                  void* stream) {
  int32_t stride[3];
  for (int d = 0; d < 3; d++) stride[d] = (d < h_B->dim) ? h_B->nloc[d] : 1;
  return tg_assemble_vector_ex(h_B, nS, h_alphaS, stride, coef, cell0, ncells, b, stream);
}

__global__ void k_sum(const double* __restrict__ x, int64_t n, double* __restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    acc += x[i];
  acc = tg_warp_sum(acc);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    acc = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    acc = tg_warp_sum(acc);
    if (lane == 0) atomicAdd(out, acc);
  }
}

extern "C" int tg_sum(const double* x, int64_t n, double* out1, void* stream) {
  TG_CHECK(cudaMemsetAsync(out1, 0, sizeof(double), tg_stream(stream)));
  if (n == 0) return 0;
  int grid = (int)(tg_cdiv(n, 256) < 1184 ? tg_cdiv(n, 256) : 1184);
  k_sum<<<grid, 256, 0, tg_stream(stream)>>>(x, n, out1);
  TG_LAUNCH_CHECK();
  return 0;
}
