// M^T A M on windowed CSR operands (MatPtAP, common.py:1194-1195), two
// phases, each entry an exact box-intersection sparse dot product:
//   AP[I,j] = sum_{J in win_A(I) ^ supp_FE(j)} A[I,J] M[J,j]
//   C [i,j] = sum_{I in supp_FE(i) ^ colbox_AP(j)} M[I,i] AP[I,j]
// No atomics, fixed summation order (deterministic).  v0: one warp per output
// row, one lane per output entry.
#include "tg_common.cuh"

__device__ inline void tg_pos_decode(int pos, const TgRowWin& rw, int* c) {
  c[0] = rw.lo[0] + pos % rw.len[0];
  int t = pos / rw.len[0];
  c[1] = rw.lo[1] + t % rw.len[1];
  c[2] = rw.lo[2] + t / rw.len[1];
}

__global__ void k_ptap_ap(TgWin wA, const double* __restrict__ Av, TgWin wM,
                          const double* __restrict__ Mv, TgWin wMT, TgWin wP,
                          double* __restrict__ APv, int64_t nrows) {
  int64_t I = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (I >= nrows) return;
  int Ic[3];
  tg_decode(I, wA.nr, wA.dim, Ic);
  TgRowWin ra = tg_row_window(wA, Ic);
  TgRowWin rp = tg_row_window(wP, Ic);
  const int64_t baseA = wA.rowptr[I], baseP = wP.rowptr[I];
  const int tot = rp.len[0] * rp.len[1] * rp.len[2];
  const int dim = wA.dim;
  for (int pos = lane; pos < tot; pos += 32) {
    int j[3];
    tg_pos_decode(pos, rp, j);
    int lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      if (d < dim) {
        lo[d] = max(ra.lo[d], wMT.lo[d][j[d]]);
        hi[d] = min(ra.lo[d] + ra.len[d] - 1, wMT.hi[d][j[d]]);
      } else {
        lo[d] = 0;
        hi[d] = 0;
      }
    }
    double acc = 0.0;
    for (int J2 = lo[2]; J2 <= hi[2]; J2++) {
      int m2 = (dim > 2) ? j[2] - wM.lo[2][J2] : 0;
      for (int J1 = lo[1]; J1 <= hi[1]; J1++) {
        int m1 = 0, lenM1 = 1;
        if (dim > 1) {
          int l = wM.lo[1][J1];
          m1 = j[1] - l;
          lenM1 = wM.hi[1][J1] - l + 1;
        }
        int64_t rowJ = (int64_t)wM.nr[0] * (J1 + (int64_t)wM.nr[1] * J2);
        int64_t offA = baseA + ((int64_t)(J2 - ra.lo[2]) * ra.len[1] + (J1 - ra.lo[1])) * ra.len[0] - ra.lo[0];
        for (int J0 = lo[0]; J0 <= hi[0]; J0++) {
          int l0 = wM.lo[0][J0];
          int lenM0 = wM.hi[0][J0] - l0 + 1;
          int64_t pm = wM.rowptr[rowJ + J0] + ((int64_t)m2 * lenM1 + m1) * lenM0 + (j[0] - l0);
          acc += Av[offA + J0] * Mv[pm];
        }
      }
    }
    APv[baseP + pos] = acc;
  }
}

extern "C" int tg_ptap_ap(const tg_win* h_wA, const double* Avals, const tg_win* h_wM,
                          const double* Mvals, const tg_win* h_wMT, const tg_win* h_wP,
                          double* APvals, void* stream) {
  int64_t nrows = tg_win_nrows(h_wA);
  k_ptap_ap<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
      tg_win_dev(h_wA), Avals, tg_win_dev(h_wM), Mvals, tg_win_dev(h_wMT), tg_win_dev(h_wP),
      APvals, nrows);
  TG_LAUNCH_CHECK();
  return 0;
}

__global__ void k_ptap_c(TgWin wM, const double* __restrict__ Mv, TgWin wMT, TgWin wP,
                         const double* __restrict__ APv, TgWin wPT, TgWin wC,
                         double* __restrict__ Cv, int64_t nrows) {
  int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= nrows) return;
  int ic[3];
  tg_decode(i, wC.nr, wC.dim, ic);
  TgRowWin rc = tg_row_window(wC, ic);
  TgRowWin rt = tg_row_window(wMT, ic);   // FE support box of N_i
  const int64_t baseC = wC.rowptr[i];
  const int tot = rc.len[0] * rc.len[1] * rc.len[2];
  const int dim = wC.dim;
  for (int pos = lane; pos < tot; pos += 32) {
    int j[3];
    tg_pos_decode(pos, rc, j);
    int lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      if (d < dim) {
        lo[d] = max(rt.lo[d], wPT.lo[d][j[d]]);
        hi[d] = min(rt.lo[d] + rt.len[d] - 1, wPT.hi[d][j[d]]);
      } else {
        lo[d] = 0;
        hi[d] = 0;
      }
    }
    double acc = 0.0;
    for (int I2 = lo[2]; I2 <= hi[2]; I2++) {
      int mi2 = 0, pj2 = 0;
      if (dim > 2) {
        mi2 = ic[2] - wM.lo[2][I2];
        pj2 = j[2] - wP.lo[2][I2];
      }
      for (int I1 = lo[1]; I1 <= hi[1]; I1++) {
        int mi1 = 0, lenM1 = 1, pj1 = 0, lenP1 = 1;
        if (dim > 1) {
          int l = wM.lo[1][I1];
          mi1 = ic[1] - l;
          lenM1 = wM.hi[1][I1] - l + 1;
          int lp = wP.lo[1][I1];
          pj1 = j[1] - lp;
          lenP1 = wP.hi[1][I1] - lp + 1;
        }
        int64_t row = (int64_t)wM.nr[0] * (I1 + (int64_t)wM.nr[1] * I2);
        for (int I0 = lo[0]; I0 <= hi[0]; I0++) {
          int l0 = wM.lo[0][I0];
          int lenM0 = wM.hi[0][I0] - l0 + 1;
          int lp0 = wP.lo[0][I0];
          int lenP0 = wP.hi[0][I0] - lp0 + 1;
          int64_t pm = wM.rowptr[row + I0] + ((int64_t)mi2 * lenM1 + mi1) * lenM0 + (ic[0] - l0);
          int64_t pp = wP.rowptr[row + I0] + ((int64_t)pj2 * lenP1 + pj1) * lenP0 + (j[0] - lp0);
          acc += Mv[pm] * APv[pp];
        }
      }
    }
    Cv[baseC + pos] = acc;
  }
}

extern "C" int tg_ptap_c(const tg_win* h_wM, const double* Mvals, const tg_win* h_wMT,
                         const tg_win* h_wP, const double* APvals, const tg_win* h_wPT,
                         const tg_win* h_wC, double* Cvals, void* stream) {
  int64_t nrows = tg_win_nrows(h_wC);
  k_ptap_c<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
      tg_win_dev(h_wM), Mvals, tg_win_dev(h_wMT), tg_win_dev(h_wP), APvals, tg_win_dev(h_wPT),
      tg_win_dev(h_wC), Cvals, nrows);
  TG_LAUNCH_CHECK();
  return 0;
}


// ===========================================================================
// Kronecker-structured M^T A M.  For tensor-product bases M = M_2 (x) M_1 (x) M_0,
// so the global M never has to be read:
//   (1) tg_ptap_kron_ap : AP[I,:] = A[I,:] * M.  One warp per FE row; the row's
//       dense column box is pulled into shared memory once (coalesced) and
//       contracted direction by direction with the tiny per-row 1-D blocks
//       tab_d[I_d][J][j] = M_d[J,j]  (sum factorisation: ~3*7 FMAs per value
//       instead of one box-intersection loop per output entry).
//   (2) tg_win_rowcombine(d) : Y[(..i_d..),:] = sum_I M_d[I,i_d] X[(..I..),:],
//       applied for d = 0,1,2 turns the FE row grid of AP into the IGA row grid
//       of C one direction at a time (<= p(p+1)+1 input rows per output row).
// A is streamed exactly once; every intermediate is smaller than the previous.
#define TG_KB 10         // padded stride of the per-row 1-D blocks (window <= 10)
#define TG_KAP_WARPS 4

struct TgKronTabs {
  const double* tab[3];   // [n_d][TG_KB][TG_KB]: tab[I][J - loA(I)][j - loP(I)]
};

template <int KE>
__global__ void __launch_bounds__(TG_KAP_WARPS * 32)
k_ptap_kron_ap(TgWin wA, const double* __restrict__ Av, TgKronTabs T, TgWin wP,
               double* __restrict__ APv, int64_t nrows, int box) {
  extern __shared__ double smk[];           // [TG_KAP_WARPS][3*box + 3*KB*KB]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t I = blockIdx.x * (int64_t)TG_KAP_WARPS + wid;
  if (I >= nrows) return;
  int Ic[3];
  tg_decode(I, wA.nr, wA.dim, Ic);
  const TgRowWin ra = tg_row_window(wA, Ic);
  const TgRowWin rp = tg_row_window(wP, Ic);
  const int L0 = ra.len[0], L1 = ra.len[1], L2 = ra.len[2];
  const int K0 = rp.len[0], K1 = rp.len[1], K2 = rp.len[2];
  double* a = smk + (size_t)wid * (3 * box + 3 * TG_KB * TG_KB);
  double* s1 = a + box;
  double* s2 = s1 + box;
  double* tb = s2 + box;                    // the three per-row 1-D blocks
  const double* __restrict__ arow = Av + wA.rowptr[I];
  const int n = L0 * L1 * L2;
  for (int p = lane; p < n; p += 32) a[p] = __ldcs(arow + p);
  for (int d = 0; d < wA.dim; d++) {
    const double* t = T.tab[d] + (int64_t)Ic[d] * TG_KB * TG_KB;
    for (int p = lane; p < TG_KB * TG_KB; p += 32) tb[d * TG_KB * TG_KB + p] = __ldg(t + p);
  }
  const double* t0 = tb;
  const double* t1 = tb + TG_KB * TG_KB;
  const double* t2 = tb + 2 * TG_KB * TG_KB;
  __syncwarp();
  // stage 1: one lane per (J1,J2) pair t: s1[t*K0 + j0] = sum_J a[t*L0 + J] t0[J][j0]
  const int nt = L1 * L2;
  for (int t = lane; t < nt; t += 32) {
    double av[KE];
#pragma unroll
    for (int J = 0; J < KE; J++) av[J] = (J < L0) ? a[t * L0 + J] : 0.0;
    for (int j0 = 0; j0 < K0; j0++) {
      double v = 0.0;
#pragma unroll
      for (int J = 0; J < KE; J++) v += av[J] * t0[J * TG_KB + j0];
      s1[t * K0 + j0] = v;
    }
  }
  __syncwarp();
  double* out = APv + wP.rowptr[I];
  const int n1 = K0 * L1 * L2;
  if (wA.dim == 1) {
    for (int o = lane; o < n1; o += 32) out[o] = s1[o];
    return;
  }
  // stage 2: one lane per (j0,J2): s2[(J2*K1 + j1)*K0 + j0] = sum_J s1[(J2*L1 + J)*K0 + j0] t1[J][j1]
  const int n02 = K0 * L2;
  for (int o = lane; o < n02; o += 32) {
    const int J2 = o / K0, j0 = o - J2 * K0;
    double sv[KE];
#pragma unroll
    for (int J = 0; J < KE; J++) sv[J] = (J < L1) ? s1[(J2 * L1 + J) * K0 + j0] : 0.0;
    for (int j1 = 0; j1 < K1; j1++) {
      double v = 0.0;
#pragma unroll
      for (int J = 0; J < KE; J++) v += sv[J] * t1[J * TG_KB + j1];
      s2[(J2 * K1 + j1) * K0 + j0] = v;
    }
  }
  __syncwarp();
  const int n2 = K0 * K1 * L2;
  if (wA.dim == 2) {
    for (int o = lane; o < n2; o += 32) out[o] = s2[o];
    return;
  }
  // stage 3: one lane per (j0,j1): out[(j2*K1 + j1)*K0 + j0] = sum_J s2[J*K0*K1 + j01] t2[J][j2]
  const int k01 = K0 * K1;
  for (int j01 = lane; j01 < k01; j01 += 32) {
    double sv[KE];
#pragma unroll
    for (int J = 0; J < KE; J++) sv[J] = (J < L2) ? s2[J * k01 + j01] : 0.0;
    for (int j2 = 0; j2 < K2; j2++) {
      double v = 0.0;
#pragma unroll
      for (int J = 0; J < KE; J++) v += sv[J] * t2[J * TG_KB + j2];
      out[j2 * k01 + j01] = v;
    }
  }
}

extern "C" int tg_ptap_kron_ap(const tg_win* h_wA, const double* Avals,
                               const double* const* h_tabs, const tg_win* h_wP, double* APvals,
                               int32_t box, void* stream) {
  TG_REQUIRE(h_wA->w0max <= TG_KB && h_wP->w0max <= TG_KB, "window wider than the 1-D block");
  TG_REQUIRE(h_wA->maxrow > 0 && h_wP->maxrow > 0, "window descriptors lack maxrow");
  TG_REQUIRE(box >= h_wA->maxrow && box >= h_wP->maxrow, "box smaller than a row");
  size_t smem = (size_t)TG_KAP_WARPS * (3 * box + 3 * TG_KB * TG_KB) * sizeof(double);
  TG_REQUIRE(smem <= 200 * 1024, "row box too large for the shared-memory tile");
  int64_t nrows = tg_win_nrows(h_wA);
  if (nrows == 0) return 0;
  TgKronTabs T;
  for (int d = 0; d < 3; d++) T.tab[d] = (d < h_wA->dim) ? h_tabs[d] : nullptr;
  // widest FE-side window over all directions bounds the contraction length
  int wmax = h_wA->w0max;
  if (h_wA->maxrow > 0) {
    // maxrow = prod of per-direction maxima; the host passes box >= that, and every
    // per-direction length is <= TG_KB; use the conservative TG_KB unless the first
    // direction's width (equal in all directions for isotropic degrees) says less
    int iso = 1;
    for (int d = 0; d < h_wA->dim; d++) iso *= h_wA->w0max;
    if (iso != h_wA->maxrow) wmax = TG_KB;
  }
#define TG_KAP_LAUNCH(KE)                                                                      \
  {                                                                                            \
    TG_CHECK(cudaFuncSetAttribute(k_ptap_kron_ap<KE>,                                          \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    k_ptap_kron_ap<KE><<<(unsigned)tg_cdiv(nrows, TG_KAP_WARPS), TG_KAP_WARPS * 32, smem,      \
                         tg_stream(stream)>>>(tg_win_dev(h_wA), Avals, T, tg_win_dev(h_wP),    \
                                              APvals, nrows, box);                             \
  }
  if (wmax <= 4) TG_KAP_LAUNCH(4)
  else if (wmax <= 6) TG_KAP_LAUNCH(6)
  else if (wmax <= 8) TG_KAP_LAUNCH(8)
  else TG_KAP_LAUNCH(10)
#undef TG_KAP_LAUNCH
  TG_LAUNCH_CHECK();
  return 0;
}

struct TgRowComb {
  int d;                    // direction being transformed
  const int32_t* mfirst;    // 1-D extraction rows of direction d: first column of node I
  const double* mvals;      // [n_fe_d][np1]
  int np1;
  const int32_t* slo;       // FE support of IGA function i in direction d
  const int32_t* shi;
};

// one warp per output row; lanes over the output window.  The loop over the
// (<= p(p+1)+1) contributing input rows is OUTSIDE the loop over output
// entries: per input row the weight, window and base pointer are computed once
// (warp-uniform), per output entry only a range test, one gather and one FMA
// remain.  Accumulators live in registers (TG_RC_SLOTS entries per lane).
#define TG_RC_SLOTS 12
__global__ void __launch_bounds__(256)
k_win_rowcombine(TgWin wX, const double* __restrict__ Xv, TgWin wY, double* __restrict__ Yv,
                 TgRowComb R, int64_t nrowsY) {
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrowsY) return;
  int rc[3];
  tg_decode(row, wY.nr, wY.dim, rc);
  const TgRowWin ry = tg_row_window(wY, rc);
  const int d = R.d;
  const int i = rc[d];
  const int Ilo = R.slo[i], Ihi = R.shi[i];
  const int tot = ry.len[0] * ry.len[1] * ry.len[2];
  double* out = Yv + wY.rowptr[row];
  const int64_t sx[3] = {1, wX.nr[0], (int64_t)wX.nr[0] * wX.nr[1]};
  int xc[3] = {rc[0], rc[1], rc[2]};
  xc[d] = 0;
  const int64_t xrow0 = xc[0] * sx[0] + xc[1] * sx[1] + xc[2] * sx[2];
  // strides of the X row's window: only the transformed direction's length varies
  const int ylen_d = ry.len[d];
  // decompose each of this lane's output positions once: (pd = offset in direction d,
  // prest = offset contribution of the other directions, expressed with len_d factored out)
  // X position = A*pd' + B  where the split depends on d:
  //   d=0: pos = (p2*len1 + p1)*lenX0 + pd'      -> hi = (p2*len1+p1), lo part = pd'
  //   d=1: pos = (p2*lenX1 + pd')*len0 + p0
  //   d=2: pos = (pd'*len1 + p1)*len0 + p0
  double acc[TG_RC_SLOTS];
  int pd[TG_RC_SLOTS], pa[TG_RC_SLOTS], pb[TG_RC_SLOTS];
#pragma unroll
  for (int s = 0; s < TG_RC_SLOTS; s++) {
    acc[s] = 0.0;
    const int pos = lane + 32 * s;
    int p0 = 0, p1 = 0, p2 = 0;
    if (pos < tot) {
      p0 = pos % ry.len[0];
      const int t = pos / ry.len[0];
      p1 = t % ry.len[1];
      p2 = t / ry.len[1];
    }
    if (d == 0) { pd[s] = p0 + ry.lo[0]; pa[s] = p2 * ry.len[1] + p1; pb[s] = 0; }
    else if (d == 1) { pd[s] = p1 + ry.lo[1]; pa[s] = p2; pb[s] = p0; }
    else { pd[s] = p2 + ry.lo[2]; pa[s] = 0; pb[s] = p1 * ry.len[0] + p0; }
    if (pos >= tot) pd[s] = -(1 << 30);
  }
  (void)ylen_d;
  for (int I = Ilo; I <= Ihi; I++) {
    const int k = i - R.mfirst[I];
    if (k < 0 || k >= R.np1) continue;                       // warp-uniform
    const double wgt = R.mvals[I * R.np1 + k];
    const int lod = wX.lo[d][I], lend = wX.hi[d][I] - lod + 1;
    const double* __restrict__ xr = Xv + wX.rowptr[xrow0 + I * sx[d]];
    // pos_X = pa*mulA + cd*mulD + pb
    int mulA, mulD;
    if (d == 0) { mulA = lend; mulD = 1; }
    else if (d == 1) { mulA = lend * ry.len[0]; mulD = ry.len[0]; }
    else { mulA = 0; mulD = ry.len[1] * ry.len[0]; }
#pragma unroll
    for (int s = 0; s < TG_RC_SLOTS; s++) {
      const int cd = pd[s] - lod;
      if (cd >= 0 && cd < lend) acc[s] += wgt * xr[pa[s] * mulA + cd * mulD + pb[s]];
    }
  }
#pragma unroll
  for (int s = 0; s < TG_RC_SLOTS; s++) {
    const int pos = lane + 32 * s;
    if (pos < tot) out[pos] = acc[s];
  }
}

// generic variant for rows longer than 32*TG_RC_SLOTS entries
__global__ void __launch_bounds__(256)
k_win_rowcombine_big(TgWin wX, const double* __restrict__ Xv, TgWin wY, double* __restrict__ Yv,
                     TgRowComb R, int64_t nrowsY) {
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrowsY) return;
  int rc[3];
  tg_decode(row, wY.nr, wY.dim, rc);
  const TgRowWin ry = tg_row_window(wY, rc);
  const int d = R.d;
  const int i = rc[d];
  const int Ilo = R.slo[i], Ihi = R.shi[i];
  const int tot = ry.len[0] * ry.len[1] * ry.len[2];
  double* out = Yv + wY.rowptr[row];
  const int64_t sx[3] = {1, wX.nr[0], (int64_t)wX.nr[0] * wX.nr[1]};
  int xc[3] = {rc[0], rc[1], rc[2]};
  xc[d] = 0;
  const int64_t xrow0 = xc[0] * sx[0] + xc[1] * sx[1] + xc[2] * sx[2];
  for (int pos = lane; pos < tot; pos += 32) {
    int c[3];
    c[0] = ry.lo[0] + pos % ry.len[0];
    int t = pos / ry.len[0];
    c[1] = ry.lo[1] + t % ry.len[1];
    c[2] = ry.lo[2] + t / ry.len[1];
    double acc = 0.0;
    for (int I = Ilo; I <= Ihi; I++) {
      const int k = i - R.mfirst[I];
      if (k < 0 || k >= R.np1) continue;
      const double wgt = R.mvals[I * R.np1 + k];
      const int lod = wX.lo[d][I], lend = wX.hi[d][I] - lod + 1;
      const int cd = c[d] - lod;
      if (cd < 0 || cd >= lend) continue;
      int len0 = ry.len[0], len1 = ry.len[1];
      int p0 = c[0] - ry.lo[0], p1 = c[1] - ry.lo[1], p2 = c[2] - ry.lo[2];
      if (d == 0) { len0 = lend; p0 = cd; }
      else if (d == 1) { len1 = lend; p1 = cd; }
      else { p2 = cd; }
      const int64_t xr = xrow0 + I * sx[d];
      acc += wgt * Xv[wX.rowptr[xr] + ((int64_t)p2 * len1 + p1) * len0 + p0];
    }
    out[pos] = acc;
  }
}

extern "C" int tg_win_rowcombine(const tg_win* h_wX, const double* Xvals, const tg_win* h_wY,
                                 double* Yvals, int32_t d, const int32_t* mfirst,
                                 const double* mvals, int32_t np1, const int32_t* supp_lo,
                                 const int32_t* supp_hi, void* stream) {
  TG_REQUIRE(d >= 0 && d < h_wX->dim, "direction");
  int64_t nrows = tg_win_nrows(h_wY);
  if (nrows == 0) return 0;
  TgRowComb R;
  R.d = d;
  R.mfirst = mfirst;
  R.mvals = mvals;
  R.np1 = np1;
  R.slo = supp_lo;
  R.shi = supp_hi;
  if (h_wY->maxrow > 0 && h_wY->maxrow <= 32 * TG_RC_SLOTS)
    k_win_rowcombine<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
        tg_win_dev(h_wX), Xvals, tg_win_dev(h_wY), Yvals, R, nrows);
  else
    k_win_rowcombine_big<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
        tg_win_dev(h_wX), Xvals, tg_win_dev(h_wY), Yvals, R, nrows);
  TG_LAUNCH_CHECK();
  return 0;
}
