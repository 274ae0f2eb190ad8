// M^T A M (MatPtAP, common.py:1194-1195) for tensor-product bases as three
// two-sided "march" passes, one per parametric direction:
//
//   Y[(..i..),(..j..)] = sum_{I,J} M_d[I,i] X[(..I..),(..J..)] M_d[J,j]
//
// i.e. direction d of BOTH the row and the column grid goes from FE nodes to IGA
// functions in one pass (a 1-D PtAP of a banded matrix for every fibre = every
// fixed choice of the other row/column coordinates).  Each pass shrinks the
// operand by nnz(A_1d)/nnz(C_1d) ~ (p+1)^2/(2p+1): A is read once, the two
// intermediates are 0.47x and 0.22x of A (p = 3), C is written once; the global M
// is never read (M = M_2 (x) M_1 (x) M_0).
//
// Kernel: a CTA owns a small tile of "lines" (fixed other row coordinates) and
// marches along direction d.  Per step the next FE row of every line -- a
// contiguous run of the value array -- is pulled into shared memory by 1-D
// bulk async copies (TMA, SASS UBLKCP) through an NS-deep mbarrier ring.  One
// thread per fibre keeps a sliding (p+1) x (2p+1) accumulator block in
// registers (the p+1 IGA rows the current FE row touches); finished rows go to
// a shared-memory tile and are written to HBM coalesced.  Deterministic (no
// atomics); FP64 throughout.
#include "tg_common.cuh"

#define TGM_THREADS 256

struct TgMarch {
  int d, KA, KAmax;              // march direction; q-stride of tabc; widest X window in d
  const int32_t* first;          // [n_fe_d]  first IGA function of FE row I
  const double* mrow;            // [n_fe_d][p+1]      M_d[I, first(I)+k]        (eps-filtered)
  const double* tabc;            // [n_fe_d][KA][TWP]  M_d[loX(I)+q, first(I)+m] (eps-filtered)
  const int32_t* slo;            // FE support of function i in direction d
  const int32_t* shi;
  const int32_t* ga;             // line-group boundaries in the two other directions
  const int32_t* gb;
  const int32_t* seg;            // output-row boundaries of the march segments
  int stage_doubles, out_doubles, maxlines;
};

__device__ __forceinline__ int tgm_len(const TgWin& w, int k, int r) {
  return (k < w.dim) ? (__ldg(w.hi[k] + r) - __ldg(w.lo[k] + r) + 1) : 1;
}
__device__ __forceinline__ long long tgm_S(const TgWin& w, int k, int r) {
  return (k < w.dim) ? (long long)__ldg(w.S[k] + r) : (long long)r;
}
__device__ __forceinline__ long long tgm_T(const TgWin& w, int k) {
  return (k < w.dim) ? (long long)__ldg(w.S[k] + w.nr[k]) : 1LL;
}

// rowptr(r) = c0 + c1*len_d(r_d) + c2*S_d[r_d]  for a line with other coordinates (ra, rb)
__device__ inline void tgm_line_consts(const TgWin& w, int d, int a, int b, int ra, int rb, int la,
                                       int lb, long long* c) {
  const long long T0 = tgm_T(w, 0), T1 = tgm_T(w, 1);
  if (d == 0) {
    c[0] = T0 * (tgm_S(w, 1, ra) * lb + T1 * tgm_S(w, 2, rb));
    c[1] = 0;
    c[2] = (long long)la * lb;
  } else if (d == 1) {
    c[0] = T0 * T1 * tgm_S(w, 2, rb);
    c[1] = (long long)lb * tgm_S(w, 0, ra);
    c[2] = (long long)lb * T0;
  } else {
    c[0] = 0;
    c[1] = tgm_S(w, 0, ra) * lb + T0 * tgm_S(w, 1, rb);
    c[2] = T0 * T1;
  }
  (void)a;
  (void)b;
}

template <int P, int NS>
__global__ void __launch_bounds__(TGM_THREADS, 2)
k_ptap_march(TgWin wX, const double* __restrict__ Xv, TgWin wY, double* __restrict__ Yv,
             TgMarch R) {
  constexpr int CW = 2 * P + 1, TW = P + 2, TWP = (TW + 1) & ~1;
  extern __shared__ __align__(128) unsigned char smraw[];
  double* stg = (double*)smraw;                                  // [NS][stage_doubles]
  double* osm = stg + (size_t)NS * R.stage_doubles;              // [out_doubles]
  long long* lc = (long long*)(osm + R.out_doubles);             // [maxlines][6]
  uint64_t* full = (uint64_t*)(lc + 6 * R.maxlines);             // [NS]
  int* li = (int*)(full + NS);                                   // [maxlines][3]
  int* par = li + 3 * R.maxlines;                                // [NS][maxlines]

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int dim = wX.dim, d = R.d;
  const int a = (d == 0) ? 1 : 0, b = (d == 2) ? 1 : 2;
  const int ra0 = __ldg(R.ga + blockIdx.x), ra1 = __ldg(R.ga + blockIdx.x + 1);
  const int rb0 = (b < dim) ? __ldg(R.gb + blockIdx.y) : 0;
  const int rb1 = (b < dim) ? __ldg(R.gb + blockIdx.y + 1) : 1;
  const int na = ra1 - ra0, nb = rb1 - rb0, nlines = na * nb;

  // ---- fibre of this thread ---------------------------------------------------
  int Fa = 0, Fb = 0;
  for (int r = ra0; r < ra1; r++) Fa += tgm_len(wX, a, r);
  for (int r = rb0; r < rb1; r++) Fb += tgm_len(wX, b, r);
  const bool active = tid < Fa * Fb;
  int ca = active ? tid % Fa : 0, cb = active ? tid / Fa : 0;
  int ia = 0, ib_ = 0;
  for (int r = ra0; r < ra1 - 1; r++) {
    const int L = tgm_len(wX, a, r);
    if (ca < L) break;
    ca -= L;
    ia++;
  }
  for (int r = rb0; r < rb1 - 1; r++) {
    const int L = tgm_len(wX, b, r);
    if (cb < L) break;
    cb -= L;
    ib_++;
  }
  const int line = ib_ * na + ia;
  const int la = tgm_len(wX, a, ra0 + ia), lb = tgm_len(wX, b, rb0 + ib_);
  int u, v, stride;
  if (d == 0) { u = 0; v = cb * la + ca; stride = 1; }
  else if (d == 1) { u = ca; v = cb * la; stride = la; }
  else { u = cb * la + ca; v = 0; stride = la * lb; }

  // ---- per-line address constants and shared-memory slots ------------------------
  for (int l = tid; l < nlines; l += TGM_THREADS) {
    const int ja = l % na, jb = l / na;
    const int lla = tgm_len(wX, a, ra0 + ja), llb = tgm_len(wX, b, rb0 + jb);
    tgm_line_consts(wX, d, a, b, ra0 + ja, rb0 + jb, lla, llb, lc + 6 * l);
    tgm_line_consts(wY, d, a, b, ra0 + ja, rb0 + jb, lla, llb, lc + 6 * l + 3);
    li[3 * l] = lla * llb;
  }
  if (tid == 0) {
    for (int s = 0; s < NS; s++) tg_mbar_init(&full[s], (uint32_t)nlines);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    int sx = 0, sy = 0;
    for (int l = 0; l < nlines; l++) {
      const int L = li[3 * l];
      li[3 * l + 1] = sx;
      li[3 * l + 2] = sy;
      sx += (R.KAmax * L + 3) & ~1;
      sy += CW * L;
    }
  }
  __syncthreads();
  const int slotX = li[3 * line + 1], slotY = li[3 * line + 2];

  // ---- march range ---------------------------------------------------------------
  const int i_lo = __ldg(R.seg + blockIdx.z), i_hi = __ldg(R.seg + blockIdx.z + 1);
  const int I_start = __ldg(R.slo + i_lo), I_end = __ldg(R.shi + i_hi - 1);
  const int nsteps = I_end - I_start + 1;

  auto issue = [&](int t) {
    const int s = t % NS;
    const int I = I_start + t;
    const int lenI = tgm_len(wX, d, I);
    const long long Sd = tgm_S(wX, d, I);
    for (int l = lane; l < nlines; l += 32) {
      const long long addr = lc[6 * l] + lc[6 * l + 1] * lenI + lc[6 * l + 2] * Sd;
      const int n = lenI * li[3 * l];
      const int off = (int)(addr & 1);
      const uint32_t bytes = (uint32_t)(((n + off + 1) & ~1) * 8);
      par[s * R.maxlines + l] = off;
      tg_mbar_expect_tx(&full[s], bytes);
      tg_bulk_g2s(stg + (size_t)s * R.stage_doubles + li[3 * l + 1], Xv + (addr - off), bytes,
                  &full[s]);
    }
  };
  if (wid == 0)
    for (int t = 0; t < NS && t < nsteps; t++) issue(t);

  double acc[P + 1][CW];
#pragma unroll
  for (int k = 0; k <= P; k++)
#pragma unroll
    for (int c = 0; c < CW; c++) acc[k][c] = 0.0;
  int ib = __ldg(R.first + I_start);

  auto emit_shift = [&]() {
    if (ib >= i_lo && ib < i_hi) {                     // CTA-uniform
      const int loY = __ldg(wY.lo[d] + ib);
      const int lenC = __ldg(wY.hi[d] + ib) - loY + 1;
      const int clo = loY - (ib - P);
      __syncthreads();                                 // previous tile fully written out
      if (active) {
        double* o = osm + slotY + u + v * lenC;
#pragma unroll
        for (int c = 0; c < CW; c++) {
          const int jj = c - clo;
          if (jj >= 0 && jj < lenC) o[jj * stride] = acc[0][c];
        }
      }
      __syncthreads();
      const long long SdY = tgm_S(wY, d, ib);
      for (int l = wid; l < nlines; l += TGM_THREADS / 32) {
        const long long addr = lc[6 * l + 3] + lc[6 * l + 4] * lenC + lc[6 * l + 5] * SdY;
        const int n = lenC * li[3 * l];
        const double* src = osm + li[3 * l + 2];
        double* dst = Yv + addr;
        for (int o = lane; o < n; o += 32) dst[o] = src[o];
      }
    }
#pragma unroll
    for (int k = 0; k < P; k++)
#pragma unroll
      for (int c = 0; c < CW; c++) acc[k][c] = acc[k + 1][c];
#pragma unroll
    for (int c = 0; c < CW; c++) acc[P][c] = 0.0;
    ib++;
  };

  for (int t = 0; t < nsteps; t++) {
    const int s = t % NS;
    const int I = I_start + t;
    const int lenI = tgm_len(wX, d, I);
    tg_mbar_wait(&full[s], (uint32_t)((t / NS) & 1));
    double tv[TWP];
#pragma unroll
    for (int m = 0; m < TWP; m++) tv[m] = 0.0;
    if (active) {
      const double* xs = stg + (size_t)s * R.stage_doubles + slotX + par[s * R.maxlines + line] +
                         u + v * lenI;
      const double2* mc = (const double2*)(R.tabc + (size_t)I * R.KA * TWP);
      for (int q = 0; q < lenI; q++) {
        const double x = xs[q * stride];
#pragma unroll
        for (int m2 = 0; m2 < TWP / 2; m2++) {
          const double2 cf = __ldg(mc + q * (TWP / 2) + m2);
          tv[2 * m2] += x * cf.x;
          tv[2 * m2 + 1] += x * cf.y;
        }
      }
    }
    __syncthreads();                                   // stage s consumed by every thread
    if (wid == 0 && t + NS < nsteps) issue(t + NS);
    const int f = __ldg(R.first + I);
    while (ib < f) emit_shift();
    const double* mrp = R.mrow + (size_t)I * (P + 1);
#pragma unroll
    for (int k = 0; k <= P; k++) {
      const double mr = __ldg(mrp + k);
#pragma unroll
      for (int m = 0; m < TW; m++) {
        const int c = m - k + P;
        if (c >= 0 && c < CW) acc[k][c] += mr * tv[m];
      }
    }
  }
  for (int k = 0; k <= P; k++) emit_shift();
}

// host: one pass.  All pointer members of h_R are device arrays; the group /
// segment boundary arrays are ALSO given on the host (h_*) for the grid size.
extern "C" int tg_ptap_march(const tg_win* h_wX, const double* Xvals, const tg_win* h_wY,
                             double* Yvals, int32_t d, int32_t p, int32_t KA, int32_t KAmax,
                             const int32_t* first, const double* mrow, const double* tabc,
                             const int32_t* slo, const int32_t* shi, const int32_t* ga,
                             int32_t nga, const int32_t* gb, int32_t ngb, const int32_t* seg,
                             int32_t nseg, int32_t stage_doubles, int32_t out_doubles,
                             int32_t maxlines, void* stream) {
  TG_REQUIRE(h_wX->dim >= 2 && h_wX->dim <= 3, "march PtAP needs a 2-D or 3-D patch");
  TG_REQUIRE(d >= 0 && d < h_wX->dim, "direction");
  TG_REQUIRE(p >= 1 && p <= 4, "degree 1..4");
  TG_REQUIRE(h_wX->layout == 0 && h_wY->layout == 0, "row-major windows only");
  TG_REQUIRE(nga >= 1 && ngb >= 1 && nseg >= 1, "empty grid");
  TG_REQUIRE(ngb <= 65535 && nseg <= 65535, "grid too large");
  TgMarch R;
  R.d = d;
  R.KA = KA;
  R.KAmax = KAmax;
  R.first = first;
  R.mrow = mrow;
  R.tabc = tabc;
  R.slo = slo;
  R.shi = shi;
  R.ga = ga;
  R.gb = gb;
  R.seg = seg;
  R.stage_doubles = stage_doubles;
  R.out_doubles = out_doubles;
  R.maxlines = maxlines;
  constexpr int NS = 4;
  size_t smem = ((size_t)NS * stage_doubles + out_doubles) * 8 + (size_t)maxlines * 6 * 8 +
                NS * 8 + (size_t)maxlines * 3 * 4 + (size_t)NS * maxlines * 4 + 16;
  TG_REQUIRE(smem <= 220 * 1024, "line tile too large for shared memory");
  dim3 grid((unsigned)nga, (unsigned)ngb, (unsigned)nseg);
#define TGM_LAUNCH(PP)                                                                          \
  {                                                                                             \
    TG_CHECK(cudaFuncSetAttribute(k_ptap_march<PP, NS>,                                         \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    k_ptap_march<PP, NS><<<grid, TGM_THREADS, smem, tg_stream(stream)>>>(                       \
        tg_win_dev(h_wX), Xvals, tg_win_dev(h_wY), Yvals, R);                                   \
  }
  switch (p) {
    case 1: TGM_LAUNCH(1) break;
    case 2: TGM_LAUNCH(2) break;
    case 3: TGM_LAUNCH(3) break;
    default: TGM_LAUNCH(4) break;
  }
#undef TGM_LAUNCH
  TG_LAUNCH_CHECK();
  return 0;
}
