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

__device__ __forceinline__ void tgm_cp_async8(uint32_t dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tgm_cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tg_smem_u32(bar))
               : "memory");
}

// TMA = true: rows staged by cp.async.bulk (one copy per line per step, issued by
// warp 0); false: by 8-byte cp.async (LDGSTS) issued by all warps, one line per warp.
template <int P, int NS, bool TMA>
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
    for (int s = 0; s < NS; s++) tg_mbar_init(&full[s], TMA ? (uint32_t)nlines : TGM_THREADS);
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
    if (!TMA) {
      for (int l = wid; l < nlines; l += TGM_THREADS / 32) {
        const long long addr = lc[6 * l] + lc[6 * l + 1] * lenI + lc[6 * l + 2] * Sd;
        const int n = lenI * li[3 * l];
        const double* src = Xv + addr;
        const uint32_t dst = tg_smem_u32(stg + (size_t)s * R.stage_doubles + li[3 * l + 1]);
        for (int o = lane; o < n; o += 32) tgm_cp_async8(dst + 8u * o, src + o);
      }
      tgm_cp_async_arrive(&full[s]);
      return;
    }
    if (wid != 0) return;
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
      const double* xs = stg + (size_t)s * R.stage_doubles + slotX +
                         (TMA ? par[s * R.maxlines + line] : 0) + u + v * lenI;
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
    if (t + NS < nsteps) issue(t + NS);
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
                             int32_t maxlines, int32_t variant, void* stream) {
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
#define TGM_LAUNCH2(PP, TT)                                                                     \
  {                                                                                             \
    TG_CHECK(cudaFuncSetAttribute(k_ptap_march<PP, NS, TT>,                                     \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    k_ptap_march<PP, NS, TT><<<grid, TGM_THREADS, smem, tg_stream(stream)>>>(                   \
        tg_win_dev(h_wX), Xvals, tg_win_dev(h_wY), Yvals, R);                                   \
  }
#define TGM_LAUNCH(PP)                                                                          \
  {                                                                                             \
    if (variant == 1) TGM_LAUNCH2(PP, true) else TGM_LAUNCH2(PP, false)                         \
  }
  switch (p) {
    case 1: TGM_LAUNCH(1) break;
    case 2: TGM_LAUNCH(2) break;
    case 3: TGM_LAUNCH(3) break;
    default: TGM_LAUNCH(4) break;
  }
#undef TGM_LAUNCH
#undef TGM_LAUNCH2
  TG_LAUNCH_CHECK();
  return 0;
}

// ===========================================================================
// Warp-independent march (the default).  ncu on the CTA-tiled kernel above showed
// ~370 instructions per warp and step for ~43 DFMAs, CTA barriers on the critical
// path and per-step table loads missing L1 (36 % issue utilisation, 43-60 %
// long-scoreboard stalls); an ablation of a first warp-level version showed the
// load, compute and store parts adding up instead of overlapping (in-order issue,
// 16 warps/SM at 128 registers).  Hence:
//  * every warp owns a "task" -- up to TGW_MAXSUB pieces, each a contiguous range
//    of fibres of one line -- and runs its own NS-deep cp.async (LDGSTS) ring in a
//    private shared-memory slice: no CTA-level synchronisation inside the march;
//  * the march advances by GROUPS of consecutive FE rows that share first(I)
//    (one knot span: <= TGW_RMAX rows).  One wait / warp-sync / prefetch issue and
//    (usually) one finished IGA row per group; the group's rows are straight-line
//    code with static register indices, so the loads of one row overlap the
//    DFMAs of the previous one;
//  * all per-node tables of the CTA's march segment live in shared memory: the
//    1-D extraction row of every FE node as a zero-padded vector
//    cpad[J] = [0, M_d[J,first(J)..first(J)+p], 0, 0], so that the column-side
//    contraction reads p+2 consecutive entries at offset 1 - (first(J)-first(I))
//    (static register indices, dynamic shared-memory address); the row-side
//    weights are cpad[I][1..p+1];
//  * each lane copies exactly len_d(I) values per row (its piece's sub-row,
//    strided by the piece's lane count: contiguous global reads per piece);
//    finished rows leave through a transpose tile that aliases the ring stage
//    about to be refilled (d = 0,1) or directly (d = 2, already coalesced).
#include <stdlib.h>
#include <type_traits>
#define TGW_MAXSUB 8
#define TGW_TSTRIDE (4 * TGW_MAXSUB + 4)
#define TGW_RMAX 4

struct TgMarchW {
  int GMAX, ntask, maxnodes, maxrows, maxgroups, stgpad, dbg;
  const int4* irec;        // [n_fe_d] {len_d(I) | lo_d(I) << 8 (X window), first(I), sbits, group(I)}
  const long long* Sx;     // [n_fe_d] S_d[I] of the X window
  const int4* jrec;        // [n_cp_d] {lo_d(i) of Y - (i-p), len_d(i) of Y, S_d[i] lo, hi}
  const double* cpad;      // [n_fe_d][p+4]
  const int32_t* grp;      // [ngroups+1] first FE row of every group
  const int32_t* slo;
  const int32_t* shi;
  const int32_t* tasks;    // [ntask][TGW_TSTRIDE]: npieces, -, -, -, then {ra, rb, cb0, ncb} each
  const int32_t* seg;
};

__device__ __forceinline__ void tgm_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tgm_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int P, int D, int NS, int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB)
k_ptap_march_w(TgWin wX, const double* __restrict__ Xv, TgWin wY, double* __restrict__ Yv,
               TgMarchW R) {
  constexpr int CW = 2 * P + 1, TW = P + 2, CPS = P + 4, NR = P + 1;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr bool TMA = (D != 2);                     // rows staged by bulk async copies
  // per-node tables of the segment in shared memory.  TS = false (read them through
  // L1 with ld.global.nc, no segment length limit) was measured 1.6x SLOWER on the
  // TMA-staged passes, where nothing else competes for L1: the ~90 warp-uniform
  // coefficient loads per group want the shared-memory broadcast path.
  constexpr bool TS = true;
  const int STG = 32 * R.GMAX + R.stgpad;                             // doubles per stage (even)
  const int WSM = (NS * STG + 2 * 32 + NS + NS * 16 + 1) & ~1;        // doubles per warp (even)
  double* ring0 = (double*)smraw;                                     // [WPC][WSM]
  int4* irec_sm = (int4*)(ring0 + (size_t)WPC * WSM);                 // [maxnodes]
  int4* jrec_sm = irec_sm + R.maxnodes;                               // [maxrows]
  double* cpad_sm = (double*)(jrec_sm + R.maxrows);                   // [maxnodes][CPS]
  unsigned* S_sm = (unsigned*)(cpad_sm + (size_t)R.maxnodes * CPS);   // [maxnodes] (fits 32 bits)
  int* gb_sm = (int*)(S_sm + R.maxnodes);                             // [maxgroups+1]

  // ---- march segment: whole groups covering the FE support of rows [i_lo, i_hi) -----
  const int i_lo = __ldg(R.seg + blockIdx.y), i_hi = __ldg(R.seg + blockIdx.y + 1);
  const int g0 = __ldg(R.irec + __ldg(R.slo + i_lo)).w;
  const int g1 = __ldg(R.irec + __ldg(R.shi + i_hi - 1)).w;
  const int ngroups = g1 - g0 + 1;
  const int rowA = __ldg(R.grp + g0), rowB = __ldg(R.grp + g1 + 1) - 1;
  const int J0 = __ldg(R.irec + rowA).x >> 8;
  if (TS) {
    const int xe = __ldg(R.irec + rowB).x;
    const int nn = (xe >> 8) + (xe & 255) - J0;                       // nodes J0 .. hi_d(rowB)
    for (int e = tid; e < nn * CPS; e += WPC * 32) cpad_sm[e] = __ldg(R.cpad + (size_t)J0 * CPS + e);
    for (int e = tid; e < nn; e += WPC * 32) {
      irec_sm[e] = __ldg(R.irec + J0 + e);
      S_sm[e] = (unsigned)__ldg(R.Sx + J0 + e);
    }
    for (int e = tid; e < i_hi - i_lo; e += WPC * 32) jrec_sm[e] = __ldg(R.jrec + i_lo + e);
    for (int e = tid; e <= ngroups; e += WPC * 32) gb_sm[e] = __ldg(R.grp + g0 + e) - J0;
  }
  // table accessors, indices relative to node J0 / row i_lo / group g0
  const int4* irec_g = R.irec + J0;
  const long long* S_g = R.Sx + J0;
  const int4* jrec_g = R.jrec + i_lo;
  const double* cpad_s = TS ? cpad_sm : R.cpad + (size_t)J0 * CPS;
  const int32_t* grp_g = R.grp + g0;
  auto ld_irec = [&](int n) -> int4 { return TS ? irec_sm[n] : __ldg(irec_g + n); };
  auto ld_S = [&](int n) -> unsigned { return TS ? S_sm[n] : (unsigned)__ldg(S_g + n); };
  auto ld_jrec = [&](int i) -> int4 { return TS ? jrec_sm[i] : __ldg(jrec_g + i); };
  auto ld_gb = [&](int k) -> int { return TS ? gb_sm[k] : __ldg(grp_g + k) - J0; };
  auto ld_c = [&](const double* q) -> double { return TS ? *q : __ldg(q); };
  __syncthreads();                                   // the only CTA barrier
  const int task = blockIdx.x * WPC + wid;
  if (task >= R.ntask) return;
  double* stg = ring0 + (size_t)wid * WSM;
  long long* lcs = (long long*)(stg + NS * STG) + lane;               // [2][32]: 64-bit row bases
  uint64_t* full = (uint64_t*)(stg + NS * STG + 64);                  // [NS] (TMA variant)
  int* par_s = (int*)(full + NS);                                     // [NS][TGW_MAXSUB][TGW_RMAX]
  constexpr int a = (D == 0) ? 1 : 0, b = (D == 2) ? 1 : 2;

  // ---- this lane's piece and fibre ---------------------------------------------------
  const int32_t* T = R.tasks + (size_t)task * TGW_TSTRIDE;
  const int npieces = __ldg(T);
  int f = lane, pre = 0, ra = 0, rb = 0, cb0 = 0, la = 1, np = 1, pidx = 0;
  bool active = false;
  for (int k = 0; k < npieces; k++) {
    const int ra_ = __ldg(T + 4 + 4 * k), rb_ = __ldg(T + 5 + 4 * k);
    const int cb0_ = __ldg(T + 6 + 4 * k), ncb_ = __ldg(T + 7 + 4 * k);
    const int la_ = tgm_len(wX, a, ra_);
    const int np_ = la_ * ncb_;
    if (!active) {
      ra = ra_; rb = rb_; cb0 = cb0_; la = la_; np = np_;
      if (f < np_) active = true;
      else { f -= np_; pre += np_; pidx++; }
    }
  }
  const int lb = tgm_len(wX, b, rb);
  const int ca = f % la, cbl = f / la, cb = cb0 + cbl;
  // fibre value q of a staged row: xa + xb*len + q*XS (piece-relative);
  // copy k of a row: row + sL*len + k*sB  ->  piece slot + f + k*np
  int xa, xb, sB;
  unsigned c1x, c2x, c1y, c2y;                       // < 2^32 (checked by the host)
  {
    long long cX[3], cY[3];
    tgm_line_consts(wX, D, a, b, ra, rb, la, lb, cX);
    tgm_line_consts(wY, D, a, b, ra, rb, la, lb, cY);
    if (D == 0) { xa = 0; xb = f; sB = np; cX[0] += f; cX[1] += cb0 * la; cY[0] += f; cY[1] += cb0 * la; }
    else if (D == 1) { xa = ca; xb = cbl * la; sB = np; cX[0] += f; cX[1] += cb0 * la; cY[0] += f; cY[1] += cb0 * la; }
    else { xa = f; xb = 0; sB = la * lb; cX[0] += cb * la + ca; cY[0] += cb * la + ca; }
    lcs[0] = cX[0];
    lcs[32] = cY[0];
    c1x = (unsigned)cX[1]; c2x = (unsigned)cX[2]; c1y = (unsigned)cY[1]; c2y = (unsigned)cY[2];
  }
  const int XS = (D == 0) ? 1 : (D == 1 ? la : np);
  // this piece's region of a stage; the TMA variant needs 16-byte aligned row slots
  // with room for one leading and one trailing element per row
  const int pslot = TMA ? (((pre * R.GMAX + 1) & ~1) + (2 * TGW_RMAX + 2) * pidx) : pre * R.GMAX;
  const bool leader = active && f == 0;
  if (TMA) {
    if (lane == 0) {
      for (int s_ = 0; s_ < NS; s_++) tg_mbar_init(&full[s_], (uint32_t)npieces);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  const uint32_t slot_u32 = tg_smem_u32(stg + pslot + f);
  const double* xslot = stg + pslot + xa;

  const int np8 = np * 8;
  int istage = 0;                                    // stage the next issue() fills
  auto issue = [&](int gk) {
    if (TMA) {
      if (gk < ngroups && leader && !(R.dbg & 1)) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int n0 = ld_gb(gk), n1 = ld_gb(gk + 1);
        const long long base = lcs[0];
        long long addr[TGW_RMAX];
        int cnt[TGW_RMAX];
        int tot = 0;
#pragma unroll
        for (int r = 0; r < TGW_RMAX; r++) {
          cnt[r] = 0;
          if (n0 + r < n1) {
            const unsigned lenI = (unsigned)ld_irec(n0 + r).x & 255u;
            addr[r] = base + (unsigned long long)c1x * lenI + (unsigned long long)c2x * ld_S(n0 + r);
            const int off = (int)(addr[r] & 1);
            cnt[r] = ((int)lenI * np + off + 1) & ~1;
            par_s[(istage * TGW_MAXSUB + pidx) * TGW_RMAX + r] = tot + off;
            tot += cnt[r];
          }
        }
        tg_mbar_expect_tx(&full[istage], (uint32_t)(tot * 8));
        double* dst = stg + istage * STG + pslot;
#pragma unroll
        for (int r = 0; r < TGW_RMAX; r++) {
          if (n0 + r < n1) {
            tg_bulk_g2s(dst, Xv + (addr[r] & ~1LL), (uint32_t)(cnt[r] * 8), &full[istage]);
            dst += cnt[r];
          }
        }
      }
    } else {
      if (gk < ngroups && active && !(R.dbg & 1)) {
        uint32_t dst = slot_u32 + (uint32_t)(istage * STG * 8);
        const int n1 = ld_gb(gk + 1);
        const long long base = lcs[0];
        for (int n = ld_gb(gk); n < n1; n++) {
          const unsigned lenI = (unsigned)ld_irec(n).x & 255u;
          const double* src = Xv + (base + (unsigned long long)c1x * lenI +
                                    (unsigned long long)c2x * ld_S(n));
#pragma unroll
          for (int k = 0; k < 2 * P + 1; k++)
            if (k < (int)lenI) tgm_cp_async8(dst + (uint32_t)(k * np8), src + (unsigned)(k * sB));
          for (int k = 2 * P + 1; k < (int)lenI; k++)      // FE degree above the spline degree
            tgm_cp_async8(dst + (uint32_t)(k * np8), src + (unsigned)(k * sB));
          dst += lenI * (unsigned)np8;
        }
      }
      tgm_cp_commit();
    }
    istage = (istage + 1 == NS) ? 0 : istage + 1;
  };
  for (int g = 0; g < NS - 1; g++) issue(g);

  double acc[NR][CW];
#pragma unroll
  for (int k = 0; k < NR; k++)
#pragma unroll
    for (int c = 0; c < CW; c++) acc[k][c] = 0.0;
  int ib = ld_irec(ld_gb(0)).y;
  int cstage = 0;                                    // stage the next group consumes
  uint32_t cphase = 0;

  // emit the finished IGA row ib, held in physical accumulator row ROT, and clear it (it
  // becomes the newest row of the sliding block); `tile` is a ring stage nobody is
  // reading (the one the next issue() refills)
  auto emit = [&](auto rc, double* tile) {
    constexpr int ROT = decltype(rc)::value;
    if (ib >= i_lo && ib < i_hi && !(R.dbg & 4)) {   // warp-uniform
      const int4 jr = ld_jrec(ib - i_lo);
      const int clo = jr.x, lenC = jr.y;
      const long long SdY = ((long long)(unsigned)jr.z) | ((long long)jr.w << 32);
      double* yrow = Yv + (lcs[32] + (unsigned long long)c1y * (unsigned)lenC + c2y * SdY);
      if (D == 2) {
        if (active) {
#pragma unroll
          for (int c = 0; c < CW; c++) {
            const int jj = c - clo;
            if (jj >= 0 && jj < lenC) yrow[(size_t)jj * sB] = acc[ROT][c];
          }
        }
      } else {
        double* oslot = tile + pslot;
        __syncwarp();                                // tile free: its readers are done
        if (active) {
          double* o = oslot + xa + xb * lenC;
          if (lenC == CW) {
#pragma unroll
            for (int c = 0; c < CW; c++) o[c * XS] = acc[ROT][c];
          } else {
#pragma unroll
            for (int c = 0; c < CW; c++) {
              const int jj = c - clo;
              if (jj >= 0 && jj < lenC) o[jj * XS] = acc[ROT][c];
            }
          }
        }
        __syncwarp();
        if (active) {
          const double* src = oslot + f;
#pragma unroll
          for (int k = 0; k < CW; k++)
            if (k < lenC) yrow[(unsigned)(k * np)] = src[k * np];
        }
      }
    }
    // slide the block up (register moves: rotating the block by code specialisation
    // -- NR copies of the group code -- was measured 2x SLOWER, instruction cache)
#pragma unroll
    for (int k = 0; k < NR - 1; k++)
#pragma unroll
      for (int c = 0; c < CW; c++) acc[k][c] = acc[k + 1][c];
#pragma unroll
    for (int c = 0; c < CW; c++) acc[NR - 1][c] = 0.0;
  };

  // the rows of one group: column-side contraction then the sliding row-side update;
  // ROT = physical accumulator row of IGA row `ib`
  auto rows = [&](auto rc, int n0, int n1) {
    constexpr int ROT = decltype(rc)::value;
    const double* xrow = xslot + cstage * STG;
    const int* pr = par_s + (cstage * TGW_MAXSUB + pidx) * TGW_RMAX;
#pragma unroll
    for (int r = 0; r < TGW_RMAX; r++) {
      if (n0 + r < n1) {
        const int4 ir = ld_irec(n0 + r);
        const int lenI = ir.x & 255;
        const double* xs_ = (TMA ? xslot + cstage * STG + pr[r] : xrow) + xb * lenI;
        xrow += np * lenI;
        const double* cp = cpad_s + ((ir.x >> 8) - J0) * CPS + 2;
        const unsigned sb = (unsigned)ir.z;
        double tv[TW];
        auto col = [&](int q, bool first_) {
          const double x = xs_[q * XS];
          const double* c = cp + q * CPS - (int)((sb >> (2 * q)) & 3u);
#pragma unroll
          for (int m = 0; m < TW; m++) tv[m] = first_ ? x * ld_c(c + m) : fma(x, ld_c(c + m), tv[m]);
        };
        col(0, true);
        if (lenI == P + 1) {
#pragma unroll
          for (int q = 1; q < P + 1; q++) col(q, false);
        } else if (lenI == 2 * P + 1) {
#pragma unroll
          for (int q = 1; q < 2 * P + 1; q++) col(q, false);
        } else {
          for (int q = 1; q < lenI; q++) col(q, false);
        }
        const double* mrp = cpad_s + (n0 + r) * CPS + 1;
#pragma unroll
        for (int k = 0; k < NR; k++) {
          const double mr = ld_c(mrp + k);
#pragma unroll
          for (int m = 0; m < TW; m++) {
            const int c = m - k + P;
            if (c >= 0 && c < CW) acc[(k + ROT) % NR][c] += mr * tv[m];
          }
        }
      }
    }
  };

  const std::integral_constant<int, 0> rot0{};
  for (int gk = 0; gk < ngroups; gk++) {
    const int n0 = ld_gb(gk), n1 = ld_gb(gk + 1);
    const int F = ld_irec(n0).y;
    while (ib < F) {
      emit(rot0, stg + istage * STG);
      ib++;
    }
    if (TMA) {
      if (!(R.dbg & 1)) tg_mbar_wait(&full[cstage], cphase);
    } else {
      tgm_cp_wait<NS - 2>();
    }
    __syncwarp();
    issue(gk + NS - 1);
    if (active && !(R.dbg & 2)) {
      rows(rot0, n0, n1);
    }
    if (++cstage == NS) {
      cstage = 0;
      cphase ^= 1u;
    }
  }
  if (!TMA) tgm_cp_wait<0>();
  for (int k = 0; k < NR; k++) {
    emit(rot0, stg);
    ib++;
  }
}

extern "C" int tg_ptap_march_w(const tg_win* h_wX, const double* Xvals, const tg_win* h_wY,
                               double* Yvals, int32_t d, int32_t p, int32_t GMAX,
                               const void* irec, const void* Sx, const void* jrec,
                               const double* cpad, const int32_t* grp, const int32_t* slo,
                               const int32_t* shi, const int32_t* tasks, int32_t ntask,
                               const int32_t* seg, int32_t nseg, int32_t maxnodes,
                               int32_t maxrows, int32_t maxgroups, int32_t maxpieces,
                               int32_t wpc, void* stream) {
  TG_REQUIRE(h_wX->dim >= 2 && h_wX->dim <= 3, "march PtAP needs a 2-D or 3-D patch");
  TG_REQUIRE(d >= 0 && d < h_wX->dim, "direction");
  TG_REQUIRE(p >= 1 && p <= 4, "degree 1..4");
  TG_REQUIRE(h_wX->layout == 0 && h_wY->layout == 0, "row-major windows only");
  TG_REQUIRE(ntask >= 1 && nseg >= 1 && nseg <= 65535, "grid");
  TG_REQUIRE(GMAX >= 2 * p + 1, "stage must hold one output row");
  TgMarchW R;
  R.GMAX = GMAX;
  R.ntask = ntask;
  R.maxnodes = maxnodes;
  R.maxrows = maxrows;
  R.maxgroups = maxgroups;
  TG_REQUIRE(maxpieces >= 1 && maxpieces <= TGW_MAXSUB, "pieces per task");
  R.stgpad = (d != 2) ? (2 * TGW_RMAX + 2) * maxpieces + 2 : (GMAX & 1) * 0;
  if ((32 * GMAX + R.stgpad) & 1) R.stgpad++;
  {
    const char* e = getenv("TIGAR_B200_MARCH_DBG");   // profiling experiments only
    R.dbg = e ? atoi(e) : 0;
  }
  R.irec = (const int4*)irec;
  R.Sx = (const long long*)Sx;
  R.jrec = (const int4*)jrec;
  R.cpad = cpad;
  R.grp = grp;
  R.slo = slo;
  R.shi = shi;
  R.tasks = tasks;
  R.seg = seg;
  constexpr int NS = 3;
  // warps per CTA: 8 (two CTAs per SM) or 16 (one CTA per SM: the segment tables are
  // shared by twice as many warps, so segments can be about twice as long)
  TG_REQUIRE(p >= 4 ? wpc == 4 : (wpc == 8 || wpc == 16), "warps per CTA: 4 (p = 4), 8 or 16");
  const bool wide = wpc == 16;
  const int WPC = wpc;
  size_t smem = (size_t)WPC * ((NS * (32 * GMAX + R.stgpad) + 2 * 32 + NS + NS * 16 + 1) & ~1) * 8 +
                (size_t)maxnodes * ((p + 4) * 8 + 4 + 16) + (size_t)maxrows * 16 +
                (size_t)(maxgroups + 1) * 4 + 16;
  TG_REQUIRE(smem <= 220 * 1024, "stage ring + tables too large for shared memory");
  dim3 grid((unsigned)tg_cdiv(ntask, WPC), (unsigned)nseg, 1);
#define TGW_LAUNCH3(PP, DD, W_, MB_)                                                            \
  {                                                                                             \
    TG_CHECK(cudaFuncSetAttribute(k_ptap_march_w<PP, DD, NS, W_, MB_>,                          \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    k_ptap_march_w<PP, DD, NS, W_, MB_><<<grid, W_ * 32, smem, tg_stream(stream)>>>(            \
        tg_win_dev(h_wX), Xvals, tg_win_dev(h_wY), Yvals, R);                                   \
  }
#define TGW_LAUNCH2(PP, DD)                                                                     \
  {                                                                                             \
    if (PP >= 4) TGW_LAUNCH3(PP, DD, 4, 3)                                                      \
    else if (wide) TGW_LAUNCH3(PP, DD, 16, 1)                                                   \
    else TGW_LAUNCH3(PP, DD, 8, 2)                                                              \
  }
#define TGW_LAUNCH(PP)                                                                          \
  {                                                                                             \
    if (d == 0) TGW_LAUNCH2(PP, 0) else if (d == 1) TGW_LAUNCH2(PP, 1) else TGW_LAUNCH2(PP, 2)  \
  }
  switch (p) {
    case 1: TGW_LAUNCH(1) break;
    case 2: TGW_LAUNCH(2) break;
    case 3: TGW_LAUNCH(3) break;
    default: TGW_LAUNCH(4) break;
  }
#undef TGW_LAUNCH
#undef TGW_LAUNCH2
#undef TGW_LAUNCH3
  TG_LAUNCH_CHECK();
  return 0;
}
