// Global sum-factorised assembly of the extracted system (element-fused path):
//   C[(i0,i1,i2),(j0,j1,j2)] = sum_terms sum_q  c_t(q) * prod_d D^{a_d}N_{i_d}(q_d) D^{b_d}N_{j_d}(q_d)
// (dolfin.assemble + MatPtAP of common.py:1215-1216, 1194-1195 for a tensor-product spline, with
// A_FE and M never formed).  Instead of building 64x64 element matrices and scattering them
// (coloured read-modify-write: 23x the compulsory traffic), the quadrature
// sum is contracted ONE DIRECTION AT A TIME over the whole patch:
//
//   X0[slot][q0,q1,q2]  --d=0-->  Y1[k1][(i0,j0)][q1,q2]  --d=1-->  Y2[k2][(i0,j0)][(i1,j1)][q2]
//                       --d=2-->  C
//
// Each stage is the same "march": a thread owns one inner index (everything except the marched
// direction), walks the cells of the direction in order, keeps the (p+1)^2 partial sums of the
// basis-function PAIRS of the current cell in registers, and when the march leaves the support
// of a pair writes the finished sum -- every output is written exactly once, in a fixed order,
// without atomics or colours.  6x fewer flops than element-wise sum factorisation (1.3 vs 7.7
// TFLOP at 256^3 cubic) and streaming, fully coalesced traffic:
//   every array is blocked as [kind][cell of the marched direction][inner][NQ], the NQ Gauss
//   points of a cell contiguous (one 32-byte vector load per thread and cell), consecutive
//   threads consecutive inner indices; the writer of a stage lays its output out for the next.
// Terms that agree in the derivative orders of the remaining directions are summed as soon as
// they meet ("kinds": 9 -> 9 -> 4 -> 1 for the Laplacian).  The last direction is processed in
// chunks of cell layers (bounded intermediates); partial sums that cross a chunk boundary are
// added to C by the next chunk (flag per pair, uniform over the CTA).
// The load vector takes the same route with (p+1) sums per thread.
//
// Input staging: a CTA's 256 inner indices x NQ values of one (cell, input) item are ONE
// contiguous 8 KB run, so (NQ even, 16-byte aligned strides) thread 0 streams the items through a
// 4-slot shared-memory ring with 1-D bulk async copies (TMA, cp.async.bulk + mbarrier
// full/empty pairs), three items ahead of the contraction: memory latency is covered by bytes in
// flight in shared memory (96 KB per SM) instead of by registers (the march holds (p+1)^2 FP64
// sums per thread and runs at 24 warps/SM).  For the last stage to see contiguous items too, the
// stage before it writes its output already in the order of the rows of C (perm_* arguments).
// Fallback (odd NQ or unaligned): one-ahead register prefetch.
#include "tg_common.cuh"
#include <string.h>
#include <stdlib.h>

#ifndef GSF_CB
#define GSF_CB 8
#endif
#ifndef GSF_THREADS
#define GSF_THREADS 256
#endif
#define GSF_MAXIN 64        // inputs summed into one output kind (plan row length)
#ifndef GSF_NS
#define GSF_NS 4
#endif
// slots of the bulk-copy ring (each GSF_THREADS * NQ doubles): GSF_NS for NQ <= 4
__host__ __device__ constexpr int gsf_ring_slots(int nq) { return nq <= 4 ? GSF_NS : 3; }

struct GsfArgs {
  const double* X;
  long long skin, scell;   // X[k*skin + (e - cbase)*scell + inner*NQ + q]
  int cbase;               // global cell index of X's first cell
  int c0, c1;              // cells marched: [c0, c1)
  int nel;                 // cells of this direction
  long long ninner;
  int nv, nw;              // inner = (u*nv + v)*nw + w
  double* Y;
  long long skout, so_f, so_u, so_v;   // Y[k*skout + f*so_f + u*so_u + v*so_v + w]
  const double* tab;       // [nel][NQ][NL][nd]
  int nd;
  const int* idx;          // [nel][NL] ; first[e] = idx[e*NL]
  const int64_t* rowbase;  // pairs: f(i,j) = rowbase[i] + j
  const int* plan;         // [nout][1 + 3*maxin] : nin, (kin, al, be)*
  int maxin;
  // last stage only -------------------------------------------------------
  int dim;                 // 2 or 3
  int n0, n1;              // rows of the plane directions (n1 = 1 in 2-D)
  const int64_t* S0;     // [n0+1] prefix sums of the first-direction window lengths
  const int64_t* S1;     // [n1+1] (3-D)
  long long F0;
  const int64_t* rowptr; // local rows of C
  const int* loL;          // [nrL] window start of the last direction (local columns)
  int row0L, nrL, col0L;   // owned rows of the last direction / column shift
  double* out;             // C values or the load vector
  // permuted output of the stage before the last (3-D pairs): position of (f1, f0) in the
  // thread order of the last stage, u = S1[i1]*F0 + S0[i0]*len1 + dj1*len0 + dj0
  int perm_out, perm_in;
  const int64_t* pS0;      // [n0+1] (also used by perm_out)
  const int64_t* pS1;      // [n1+1]
  const int32_t* pLo1;     // [n1]
  const longlong2* prow;   // [n1] packed {S1[i]*F0, len1[i] | lo1[i] << 32} or NULL
  int pn0;
  long long pF0;
};

template <int NL, bool PAIR>
struct GsfAcc {
  double v[PAIR ? NL * NL : NL];
};

template <int NL, int NQ, bool PAIR, bool LAST, int MINB, bool TMA>
__global__ void __launch_bounds__(GSF_THREADS, ((PAIR && NL >= 5) ? 2 : MINB) * (256 / GSF_THREADS))
k_gsf(const __grid_constant__ GsfArgs A) {
  constexpr int NLP = (NL + 1) & ~1;
  constexpr int NS = TMA ? gsf_ring_slots(NQ) : 1;                 // ring slots
  extern __shared__ __align__(128) unsigned char gsf_dyn[];        // TMA ring (dynamic)
  double* const ring = reinterpret_cast<double*>(gsf_dyn);
  __shared__ __align__(8) uint64_t fullb[NS], emptyb[NS];
  __shared__ __align__(16) double stab[GSF_CB * NQ * 3 * NLP];   // [c][q][k][a]
  __shared__ int sfirst[GSF_CB + 1];
  const int tid = threadIdx.x;
  const int ko = blockIdx.y;
  const int* plan = A.plan + (long long)ko * (1 + 3 * A.maxin);
  const int nin = plan[0];
  // the plan of this output kind, one packed word per input (kin | al << 16 | be << 20), in
  // shared memory: the item loop reads it once per item instead of three global loads
  __shared__ int splan[GSF_MAXIN];
  if (tid < nin) splan[tid] = plan[1 + 3 * tid] | (plan[2 + 3 * tid] << 16) | (plan[3 + 3 * tid] << 20);
  __syncthreads();
  const long long t = blockIdx.x * (long long)GSF_THREADS + tid;

  // ---- what this thread owns -------------------------------------------------------
  bool active;
  long long inner = 0, uoff = 0;
  int rowoff = 0, inrow = 0, len01 = 1;              // LAST && PAIR
  int pS0i0 = 0, pl0 = 1, pdj0 = 0;                   // perm_out: this thread's (i0, dj0)
  if (!LAST) {
    active = t < A.ninner;
    if (active) {
      inner = t;
      const long long w = t % A.nw, r = t / A.nw;
      const long long v = r % A.nv, u = r / A.nv;
      uoff = u * A.so_u + v * A.so_v + w;
      if (PAIR && A.perm_out) {
        // u = f0 = S0[i0] + dj0 : find i0
        int lo = 0, hi = A.pn0;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (A.pS0[mid] <= u) lo = mid; else hi = mid;
        }
        pS0i0 = (int)A.pS0[lo];
        pl0 = (int)A.pS0[lo + 1] - pS0i0;
        pdj0 = (int)u - pS0i0;
        uoff = v * A.so_v + w;                         // the f-dependent part is added per store
      }
    }
  } else if (!PAIR) {
    active = t < A.ninner;
    inner = t;
  } else {
    const long long F1 = (A.dim == 3) ? A.S1[A.n1] : 1;
    active = t < A.F0 * F1;
    if (active) {
      long long rem = t;
      int i1 = 0;
      long long l1 = 1, f1 = 0;
      if (A.dim == 3) {
        int lo = 0, hi = A.n1;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (A.S1[mid] * A.F0 <= t) lo = mid; else hi = mid;
        }
        i1 = lo;
        rem = t - A.S1[i1] * A.F0;
        l1 = A.S1[i1 + 1] - A.S1[i1];
      }
      int lo = 0, hi = A.n0;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (A.S0[mid] * l1 <= rem) lo = mid; else hi = mid;
      }
      const int i0 = lo;
      const long long rem2 = rem - A.S0[i0] * l1;
      const long long l0 = A.S0[i0 + 1] - A.S0[i0];
      const long long dj1 = rem2 / l0, dj0 = rem2 - dj1 * l0;
      f1 = (A.dim == 3) ? A.S1[i1] + dj1 : 0;
      inner = A.perm_in ? t : f1 * A.F0 + A.S0[i0] + dj0;
      rowoff = i1 * A.n0 + i0;
      inrow = (int)rem2;
      len01 = (int)(l0 * l1);
    }
  }
  const int nplane = A.n0 * A.n1;

  GsfAcc<NL, PAIR> acc;
#pragma unroll
  for (int i = 0; i < (PAIR ? NL * NL : NL); i++) acc.v[i] = 0.0;

  // pairs that were already in flight before c0 (chunked last direction): their sums are ADDED
  unsigned carried = 0;
  if (LAST && A.c0 > 0) {
    const int s = A.idx[(long long)A.c0 * NL] - A.idx[(long long)(A.c0 - 1) * NL];
#pragma unroll
    for (int a = 0; a < NL; a++)
#pragma unroll
      for (int b = 0; b < (PAIR ? NL : 1); b++)
        if (a + s < NL && (!PAIR || b + s < NL)) carried |= 1u << (a * NL + b);
  }

  // Output addressing, hoisted per row i of the marched direction: every finished sum of row i
  // goes to  rowp(i) + j * jstride  (jstride is a per-thread constant), so a flush costs one
  // address computation per row, not one per entry.
  int jstride;
  double* ybase = nullptr;
  if (!LAST) {
    ybase = A.Y + ko * A.skout + uoff;
    jstride = (int)((PAIR && A.perm_out) ? pl0 * A.so_f : A.so_f);
  } else {
    jstride = PAIR ? len01 : 0;
  }
  auto rowp = [&](int i, bool& ok) -> double* {
    ok = true;
    if (!LAST) {
      if (!PAIR) return ybase + (long long)i * A.so_f;            // + 0 * jstride (j unused)
      if (A.perm_out) {
        if (A.prow) {          // one 16-byte load and 32-bit products per row
          const longlong2 rc = __ldg(A.prow + i);
          const int l1 = (int)(rc.y & 0xffffffffll), lo1 = (int)(rc.y >> 32);
          return ybase + (rc.x + (long long)(pS0i0 * l1 + pdj0 - lo1 * pl0)) * A.so_f;
        }
        const long long s1 = A.pS1[i], l1 = A.pS1[i + 1] - s1;
        return ybase + (s1 * A.pF0 + (long long)pS0i0 * l1 + pdj0 -
                        (long long)A.pLo1[i] * pl0) * A.so_f;
      }
      return ybase + A.rowbase[i] * A.so_f;
    }
    const int r = i - A.row0L;
    ok = (r >= 0 && r < A.nrL);
    if (!ok) return nullptr;
    if (PAIR)
      return A.out + A.rowptr[(long long)r * nplane + rowoff] + inrow -
             (long long)(A.col0L + A.loL[r]) * len01;
    return A.out + (long long)r * nplane + inner;
  };

  // ---- input staging ---------------------------------------------------------------------
  // TMA: items (cell, input) stream through the ring, thread 0 issues NS-1 items ahead.
  // otherwise: one-ahead software prefetch into registers.
  const long long inner0 = blockIdx.x * (long long)GSF_THREADS;
  const int nvalid = (int)min((long long)GSF_THREADS, A.ninner - inner0);
  const int nitems = (A.c1 - A.c0) * nin;
  auto issue = [&](int it) {                // thread 0 only
    const int slot = it % NS;
    const int e = A.c0 + it / nin, in = it % nin;
    const int kin = splan[in] & 0xffff;
    const double* src = A.X + kin * A.skin + (long long)(e - A.cbase) * A.scell + inner0 * NQ;
    const uint32_t bytes = (uint32_t)nvalid * NQ * 8u;
    tg_mbar_expect_tx(&fullb[slot], bytes);
    tg_bulk_g2s(&ring[slot * GSF_THREADS * NQ], src, bytes, &fullb[slot]);
  };
  auto load_item = [&](int e, int in, double* x) {
    const int kin = splan[in] & 0xffff;
    const double* xp = A.X + kin * A.skin + (long long)(e - A.cbase) * A.scell + inner * NQ;
    if constexpr (NQ % 2 == 0) {
#pragma unroll
      for (int q = 0; q < NQ; q += 2) {
        const double2 v = __ldcs(reinterpret_cast<const double2*>(xp + q));
        x[q] = v.x;
        x[q + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int q = 0; q < NQ; q++) x[q] = __ldcs(xp + q);
    }
  };
  double xn[NQ];
#pragma unroll
  for (int q = 0; q < NQ; q++) xn[q] = 0.0;
  int en = A.c0, inn = 0;
  int item = 0;
  if (TMA) {
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < NS; k++) {
        tg_mbar_init(&fullb[k], 1);
        tg_mbar_init(&emptyb[k], GSF_THREADS);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
      for (int it = 0; it < NS - 1 && it < nitems; it++) issue(it);
  } else {
    if (active && nin > 0 && en < A.c1) load_item(en, inn, xn);
  }

  for (int cb = A.c0; cb < A.c1; cb += GSF_CB) {
    const int ncb = min(GSF_CB, A.c1 - cb);
    __syncthreads();
    for (int e = tid; e < ncb * NQ * A.nd * NL; e += GSF_THREADS) {
      // global [c][q][a][k] -> shared [c][q][k][a]
      const int k = e % A.nd;
      int r = e / A.nd;
      const int a = r % NL;
      r /= NL;                                  // r = c*NQ + q
      stab[(r * 3 + k) * NLP + a] = A.tab[(long long)cb * NQ * NL * A.nd + e];
    }
    if (tid <= ncb) {
      const int e = cb + tid;
      sfirst[tid] = (e < A.nel) ? A.idx[(long long)e * NL] : (A.idx[(long long)(A.nel - 1) * NL] + NL);
    }
    __syncthreads();
    for (int c = 0; c < ncb; c++) {
      const int e = cb + c;
      for (int in = 0; in < nin; in++) {
        const int pk = splan[in];
        const int al = (pk >> 16) & 15, be = (pk >> 20) & 15;
        double x[NQ];
        if (!TMA) {
          if (!active) continue;
#pragma unroll
          for (int q = 0; q < NQ; q++) x[q] = xn[q];
          if (++inn == nin) {
            inn = 0;
            en++;
          }
          if (en < A.c1) load_item(en, inn, xn);
        } else {
          // every thread of the CTA takes part in the ring protocol, active or not
          const int slot = item % NS;
          if (tid == 0) {
            const int nx = item + NS - 1;                   // refill the slot freed by item-1
            if (nx < nitems) {
              if (nx >= NS) tg_mbar_wait(&emptyb[nx % NS], (uint32_t)((nx / NS - 1) & 1));
              issue(nx);
            }
          }
          tg_mbar_wait(&fullb[slot], (uint32_t)((item / NS) & 1));
#pragma unroll
          for (int q = 0; q < NQ; q += 2) {
            const double2 v =
                *reinterpret_cast<const double2*>(&ring[(slot * GSF_THREADS + tid) * NQ + q]);
            x[q] = v.x;
            x[q + 1] = v.y;
          }
          tg_mbar_arrive(&emptyb[slot]);
          item++;
          if (!active) continue;
        }
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const double* ta = &stab[((c * NQ + q) * 3 + al) * NLP];
          if (PAIR) {
            const double* tb = &stab[((c * NQ + q) * 3 + be) * NLP];
            double y[NL];
#pragma unroll
            for (int b = 0; b < NL; b++) y[b] = tb[b] * x[q];
#pragma unroll
            for (int a = 0; a < NL; a++) {
              const double ta_a = ta[a];
#pragma unroll
              for (int b = 0; b < NL; b++) acc.v[a * NL + b] = fma(ta_a, y[b], acc.v[a * NL + b]);
            }
          } else {
#pragma unroll
            for (int a = 0; a < NL; a++) acc.v[a] = fma(ta[a], x[q], acc.v[a]);
          }
        }
      }
      // pairs / functions whose support ends with this cell are complete.  The shift s is
      // uniform over the CTA: one static code path per value (no per-entry predicates)
      const int first = sfirst[c];
      int s = sfirst[c + 1] - first;
      if (s > NL || e + 1 == A.c1) s = NL;      // end of the direction or of the chunk: flush all
      unsigned nc = 0;
#pragma unroll
      for (int ss = 1; ss <= NL; ss++) {
        if (s != ss) continue;
        if (active) {
          const long long fj = PAIR ? (long long)first * jstride : 0;   // one wide multiply
#pragma unroll
          for (int a = 0; a < NL; a++) {
            if (!(a < ss || PAIR)) continue;          // vector: only the leaving functions
            bool ok;
            double* rp = rowp(first + a, ok);
            if (!ok) continue;
#pragma unroll
            for (int b = 0; b < (PAIR ? NL : 1); b++) {
              if (!(a < ss || (PAIR && b < ss))) continue;
              double* q = PAIR ? rp + fj + (long long)b * jstride : rp;
              const double val = acc.v[PAIR ? a * NL + b : a];
              if (LAST && ((carried >> (a * NL + b)) & 1u)) *q += val;
              else *q = val;
            }
          }
        }
        // shift the window of partial sums by ss functions
#pragma unroll
        for (int a = 0; a < NL; a++)
#pragma unroll
          for (int b = 0; b < (PAIR ? NL : 1); b++) {
            const bool keep = (a + ss < NL) && (!PAIR || b + ss < NL);
            const int src = PAIR ? (a + ss) * NL + (b + ss) : a + ss;
            acc.v[PAIR ? a * NL + b : a] = keep ? acc.v[keep ? src : 0] : 0.0;
            if (keep && ((carried >> (keep ? (a + ss) * NL + (PAIR ? b + ss : 0) : 0)) & 1u))
              nc |= 1u << (a * NL + b);
          }
      }
      carried = nc;
    }
  }
}

template <int NL, int NQ>
static int gsf_launch2(const GsfArgs& A, int pair, int last, int tma, long long nthreads, int nout,
                       cudaStream_t st) {
  dim3 grid((unsigned)tg_cdiv(nthreads, GSF_THREADS), (unsigned)nout);
  // TIGAR_B200_GSF_MINB=2: two resident CTAs per SM (up to 128 registers, no spills) for the
  // TMA-staged matrix kernels, whose memory latency is covered by the ring, not by occupancy
  static const int minb2 = getenv("TIGAR_B200_GSF_MINB") ? atoi(getenv("TIGAR_B200_GSF_MINB")) == 2 : 0;
  if constexpr (NQ % 2 == 0) {
    if (tma) {
      constexpr int ring_bytes = gsf_ring_slots(NQ) * GSF_THREADS * NQ * 8;
#define GSF_GO(P_, L_, M_)                                                                      \
  do {                                                                                          \
    if (ring_bytes + 4096 > 48 * 1024)                                                              \
      TG_CHECK(cudaFuncSetAttribute(k_gsf<NL, NQ, P_, L_, M_, true>,                            \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes));  \
    k_gsf<NL, NQ, P_, L_, M_, true><<<grid, GSF_THREADS, ring_bytes, st>>>(A);                  \
  } while (0)
      if (pair && minb2) {
        if (last) GSF_GO(true, true, 2);
        else GSF_GO(true, false, 2);
      } else if (pair && last) GSF_GO(true, true, 3);
      else if (pair) GSF_GO(true, false, 3);
      else if (last) GSF_GO(false, true, 3);
      else GSF_GO(false, false, 3);
#undef GSF_GO
      TG_LAUNCH_CHECK();
      return 0;
    }
  }
  if (pair && last) k_gsf<NL, NQ, true, true, 3, false><<<grid, GSF_THREADS, 0, st>>>(A);
  else if (pair) k_gsf<NL, NQ, true, false, 3, false><<<grid, GSF_THREADS, 0, st>>>(A);
  else if (last) k_gsf<NL, NQ, false, true, 3, false><<<grid, GSF_THREADS, 0, st>>>(A);
  else k_gsf<NL, NQ, false, false, 3, false><<<grid, GSF_THREADS, 0, st>>>(A);
  TG_LAUNCH_CHECK();
  return 0;
}

extern "C" int tg_gsf_supported(int32_t nloc, int32_t nq) {
  return (nloc >= 2 && nloc <= 5 && (nq == nloc || nq == nloc + 1)) ? 1 : 0;
}

// One march stage.  See the file header for the layouts; h_* are host scalars.
//   X            input  [nin_kinds][cells c >= cbase][ninner][nq]  (kind stride skin, cell stride
//                scell = ninner*nq unless the caller says otherwise)
//   pair != 0    matrix (pairs of functions), else load vector
//   last != 0    writes C (windowed CSR h_W, local rows) / the vector, else Y with
//                Y[k*skout + f*so_f + u*so_u + v*so_v + w], inner = (u*nv + v)*nw + w
//   tab/idx      1-D tables of the direction ([nel][nq][nloc][nd], [nel][nloc]); rowbase[n_d]
//                = S_d[i] - lo_d[i] of the (global) C window: f(i,j) = rowbase[i] + j
//   plan         device int32 [nout][1 + 3*maxin]
extern "C" int tg_gsf_stage(const double* X, int64_t skin, int64_t scell, int32_t cbase,
                            int32_t c0, int32_t c1, int32_t nel, int32_t nloc, int32_t nq,
                            int32_t nd, const double* tab, const int32_t* idx,
                            const int64_t* rowbase, const int32_t* plan, int32_t nout,
                            int32_t maxin, int32_t pair, int64_t ninner, int32_t nv, int32_t nw,
                            double* Y, int64_t skout, int64_t so_f, int64_t so_u, int64_t so_v,
                            int32_t last, const tg_win* h_W, int64_t h_F0, int32_t vec_row0,
                            int32_t vec_nr, double* out, int32_t perm, const tg_win* h_Wperm,
                            const int64_t* perm_rows, void* stream) {
  TG_REQUIRE(tg_gsf_supported(nloc, nq), "gsf: unsupported (nloc, nq)");
  TG_REQUIRE(nd >= 1 && nd <= 3, "gsf: tables hold derivative orders 0..2");
  TG_REQUIRE(maxin >= 1 && maxin <= GSF_MAXIN, "gsf: too many inputs per output kind");
  GsfArgs A;
  memset(&A, 0, sizeof(A));
  A.X = X; A.skin = skin; A.scell = scell; A.cbase = cbase; A.c0 = c0; A.c1 = c1; A.nel = nel;
  A.ninner = ninner; A.nv = nv > 0 ? nv : 1; A.nw = nw > 0 ? nw : 1;
  A.Y = Y; A.skout = skout; A.so_f = so_f; A.so_u = so_u; A.so_v = so_v;
  A.tab = tab; A.nd = nd; A.idx = idx; A.rowbase = rowbase; A.plan = plan; A.maxin = maxin;
  long long nthreads = ninner;
  A.n0 = 1; A.n1 = 1; A.dim = 2;
  if (last) {
    A.out = out;
    if (pair) {
      TG_REQUIRE(h_W && h_W->layout == 0 && (h_W->dim == 2 || h_W->dim == 3),
                 "gsf: last stage needs a 2-D/3-D row-major window");
      const int L = h_W->dim - 1;
      A.dim = h_W->dim;
      A.n0 = h_W->nr[0];
      A.n1 = (h_W->dim == 3) ? h_W->nr[1] : 1;
      A.S0 = h_W->S[0];
      A.S1 = (h_W->dim == 3) ? h_W->S[1] : nullptr;
      A.rowptr = h_W->rowptr;
      A.loL = h_W->lo[L];
      A.row0L = h_W->row0[L]; A.nrL = h_W->nr[L]; A.col0L = h_W->col0[L];
      A.F0 = h_F0;             // = S0[n0] (the prefix sums live on the device)
      nthreads = ninner;       // = F0 * F1, computed by the caller
    } else {
      A.row0L = vec_row0; A.nrL = vec_nr;
      A.n0 = (int)ninner;      // plane size as n0*n1 with n1 = 1
    }
  }
  // perm: this stage writes (perm, not last) / reads (perm, last) the pair index (f1, f0) in the
  // thread order of the last stage, which makes the last stage's items contiguous
  if (perm && pair) {
    TG_REQUIRE(h_Wperm && h_Wperm->dim == 3, "gsf: perm layout is for 3-D windows");
    A.pS0 = h_Wperm->S[0]; A.pS1 = h_Wperm->S[1]; A.pLo1 = h_Wperm->lo[1];
    A.pn0 = h_Wperm->nr[0]; A.pF0 = h_F0;
    A.prow = (!last) ? reinterpret_cast<const longlong2*>(perm_rows) : nullptr;
    if (last) A.perm_in = 1; else A.perm_out = 1;
  }
  if (nthreads <= 0 || c1 <= c0) return 0;
  // bulk-copy ring: even NQ, 16-byte aligned strides, contiguous items per CTA
  int tma = (nq % 2 == 0) && ((uintptr_t)X % 16 == 0) && (skin % 2 == 0) && (scell % 2 == 0) &&
            (!last || !pair || h_W->dim == 2 || A.perm_in);
  static const int no_tma = getenv("TIGAR_B200_GSF_TMA") ? !atoi(getenv("TIGAR_B200_GSF_TMA")) : 0;
  if (no_tma) tma = 0;
  cudaStream_t st = tg_stream(stream);
#define GSF_CASE(NL_, NQ_) \
  if (nloc == NL_ && nq == NQ_) return gsf_launch2<NL_, NQ_>(A, pair, last, tma, nthreads, nout, st);
  GSF_CASE(2, 2) GSF_CASE(2, 3) GSF_CASE(3, 3) GSF_CASE(3, 4) GSF_CASE(4, 4) GSF_CASE(4, 5)
  GSF_CASE(5, 5) GSF_CASE(5, 6)
#undef GSF_CASE
  tg_set_error("gsf: no instantiation for nloc=%d nq=%d", nloc, nq);
  return 2;
}
