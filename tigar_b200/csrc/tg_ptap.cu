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
