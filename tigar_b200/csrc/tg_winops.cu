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
#include <stdlib.h>

static inline int64_t tg_win_lines(const tg_win* w);
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
// streaming load of a matrix value: LD = 0 ld.global.cs (evict-first), 1 L1::no_allocate,
// 2 L1::no_allocate + 256-byte L2 prefetch
template <int LD>
__device__ __forceinline__ double tg_ld_stream(const double* p) {
  double v;
  if (LD == 1)
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (LD == 2)
    asm volatile("ld.global.L1::no_allocate.L2::256B.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else
    v = __ldcs(p);
  return v;
}

template <bool DOT, int U, int LD = 0>
__global__ void __launch_bounds__(TG_WS_BLOCK, (U <= 4) ? 4 : 3)
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
        // U coalesced 256-byte value loads in flight per warp per trip
        double accs[U];
#pragma unroll
        for (int u = 0; u < U; u++) accs[u] = 0.0;
        int p = lane;
        for (; p + 32 * (U - 1) < n; p += 32 * U) {
          double a[U], xv[U];
#pragma unroll
          for (int u = 0; u < U; u++) a[u] = tg_ld_stream<LD>(av + p + 32 * u);
#pragma unroll
          for (int u = 0; u < U; u++) xv[u] = xr[xtab[p + 32 * u]];
#pragma unroll
          for (int u = 0; u < U; u++) accs[u] += a[u] * xv[u];
        }
        {   // remainder (< U chunks): predicated, still issued together
          double a[U], xv[U];
#pragma unroll
          for (int u = 0; u < U - 1; u++)
            a[u] = (p + 32 * u < n) ? tg_ld_stream<LD>(av + p + 32 * u) : 0.0;
#pragma unroll
          for (int u = 0; u < U - 1; u++) xv[u] = (p + 32 * u < n) ? xr[xtab[p + 32 * u]] : 0.0;
#pragma unroll
          for (int u = 0; u < U - 1; u++) accs[u] += a[u] * xv[u];
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc += accs[u];
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

// ---------------------------------------------------------------------------
// Direct kernel with a per-warp row prefetch: while a warp reduces row r out of
// its private shared-memory buffer, the NEXT row it will process (r+8, or its
// first row of the CTA's next item) is already streaming in by 8-byte cp.async
// (LDGSTS).  Bytes in flight per SM go from 32 warps x 1 KB (U=4 chunk loads) to
// 32 warps x 2.7 KB, with no cross-warp synchronisation.
__device__ __forceinline__ void tg_cp8(uint32_t dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

template <bool DOT>
__global__ void __launch_bounds__(TG_WS_BLOCK, 4)
k_win_spmv_pf(TgWin w, const double* __restrict__ vals, const double* __restrict__ x,
              int64_t xoff, double* __restrict__ y, int nchunk, int nitems, int w0max, int RB,
              double* __restrict__ part) {
  extern __shared__ __align__(16) double rowbuf[];            // [8 warps][2][RB]
  __shared__ double sh[32];
  __shared__ int xtab[TG_WS_MAXTAB];
  __shared__ int tab_len1, tab_len2;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nr0 = w.nr[0], nr1 = w.nr[1];
  const int nc0 = w.nc[0];
  const int64_t pl = (int64_t)nc0 * w.nc[1];
  double* mybuf = rowbuf + (size_t)wid * 2 * RB;
  const uint32_t mybuf_u32 = tg_smem_u32(mybuf);
  if (threadIdx.x == 0) {
    tab_len1 = -1;
    tab_len2 = -1;
  }
  __syncthreads();

  // value-array address and length of row r0 of item `item`; false if the item has no such row
  auto locate = [&](int item, int r0, const double*& av, int& n) -> bool {
    const int line = item / nchunk;
    const int ch = item - line * nchunk;
    const int rend = min(nr0, ch * TG_WS_ROWS + TG_WS_ROWS);
    r0 += ch * TG_WS_ROWS;
    if (r0 >= rend) return false;
    const int r2 = line / nr1, r1 = line - r2 * nr1;
    int len1 = 1, len2 = 1;
    int64_t linebase = 0;
    if (w.dim > 1) len1 = __ldg(w.hi[1] + r1) - __ldg(w.lo[1] + r1) + 1;
    if (w.dim > 2) len2 = __ldg(w.hi[2] + r2) - __ldg(w.lo[2] + r2) + 1;
    if (w.dim > 1) {
      const int64_t T0 = __ldg(w.S[0] + nr0);
      int64_t inner = __ldg(w.S[1] + r1) * len2;
      if (w.dim > 2) inner += __ldg(w.S[1] + nr1) * __ldg(w.S[2] + r2);
      linebase = T0 * inner;
    }
    const int tile = len1 * len2;
    av = vals + linebase + __ldg(w.S[0] + r0) * tile;
    n = (__ldg(w.hi[0] + r0) - __ldg(w.lo[0] + r0) + 1) * tile;
    return true;
  };
  // the row this warp processes after (item, local row k): (item, k+8) or the first row
  // it owns in a later item of this CTA
  auto advance = [&](int& item, int& k, const double*& av, int& n) -> bool {
    k += TG_WS_BLOCK / 32;
    if (k < TG_WS_ROWS && locate(item, k, av, n)) return true;
    for (item += gridDim.x; item < nitems; item += gridDim.x) {
      k = wid;
      if (locate(item, k, av, n)) return true;
    }
    return false;
  };
  auto prefetch = [&](const double* av, int n, int buf) {
    const uint32_t dst = mybuf_u32 + (uint32_t)(buf * RB * 8);
    for (int p = lane; p < n; p += 32) tg_cp8(dst + 8u * p, av + p);
  };

  // prime the pipeline with this warp's first row
  int pitem = blockIdx.x - gridDim.x, pk = TG_WS_ROWS;        // "before the first item"
  const double* pav = nullptr;
  int pn = 0;
  bool pvalid = advance(pitem, pk, pav, pn);
  if (pvalid) prefetch(pav, pn, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  int cbuf = 0;

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
    for (int r0 = ch * TG_WS_ROWS + wid; r0 < rend; r0 += TG_WS_BLOCK / 32) {
      // (item, r0) is the row primed in buffer cbuf; start the following one
      pvalid = advance(pitem, pk, pav, pn);
      __syncwarp();                                   // everyone done with buffer cbuf^1
      if (pvalid) prefetch(pav, pn, cbuf ^ 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncwarp();
      const double* sv = mybuf + cbuf * RB;
      const int64_t row = r0 + (int64_t)nr0 * line;
      const int lo0 = __ldg(w.lo[0] + r0);
      const int len0 = __ldg(w.hi[0] + r0) - lo0 + 1;
      const double* xr = xb + lo0;
      const int n = len0 * tile;
      double acc = 0.0;
      if (usetab && len0 == w0max) {
        double acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        int p = lane;
        for (; p + 96 < n; p += 128) {
          const double x0 = xr[xtab[p]], x1 = xr[xtab[p + 32]], x2 = xr[xtab[p + 64]],
                       x3 = xr[xtab[p + 96]];
          acc += sv[p] * x0;
          acc1 += sv[p + 32] * x1;
          acc2 += sv[p + 64] * x2;
          acc3 += sv[p + 96] * x3;
        }
        {
          const bool b0 = p < n, b1 = p + 32 < n, b2 = p + 64 < n;
          const double x0 = b0 ? xr[xtab[p]] : 0.0, x1 = b1 ? xr[xtab[p + 32]] : 0.0,
                       x2 = b2 ? xr[xtab[p + 64]] : 0.0;
          if (b0) acc += sv[p] * x0;
          if (b1) acc1 += sv[p + 32] * x1;
          if (b2) acc2 += sv[p + 64] * x2;
        }
        acc += (acc1 + acc2) + acc3;
      } else {
        for (int p = lane; p < n; p += 32) {
          int c0 = p % len0, t = p / len0, c1 = t % len1, c2 = t / len1;
          acc += sv[p] * xr[c0 + nc0 * c1 + pl * c2];
        }
      }
      cbuf ^= 1;
      acc = tg_warp_sum(acc);
      if (lane == 0) {
        y[row] = acc;
        if (DOT) dot += x[xoff + row] * acc;
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (DOT) {
    dot = tg_block_sum_ws(dot, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = dot;
  }
}

// ---------------------------------------------------------------------------
// TMA-staged variant (the default for row-major windows that fit): the rows of
// one (r1,r2) line are one contiguous run of the value array, so a producer warp
// pulls stages of TG_TMA_ROWS consecutive rows through a ring of shared-memory
// buffers with 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) completing on
// mbarriers, and eight consumer warps each reduce one row per stage.  The x
// entries a stage touches -- the union of its rows' column boxes, (rows+w0-1) x
// len1 x len2 values -- are staged in shared memory too (register-prefetched
// one stage ahead, double buffered), so the consumers never wait on a global
// load: the first TMA version gathered x through L1/L2 and was latency-bound
// there (0.34 of HBM peak, profiles/r1_ncu_spmv_tma.txt).  NS x 22 KB per CTA in
// flight, independent of occupancy.
#define TG_TMA_ROWS 8
#define TG_TMA_CONSUMERS (TG_TMA_ROWS * 32)
#define TG_TMA_THREADS (TG_TMA_CONSUMERS + 32)
#define TG_TMA_XW 16          // stride of the x tile's first direction (>= rows + w0max - 1)
#define TG_TMA_XPT 4          // x-tile elements prefetched per consumer thread

struct TgStage {              // one stage = rows [r0a, r0b) of line (r1, r2)
  int ch, r1, r2, r0a, r0b, len1, len2, tile, c0lo, cw;
  long long line, xline, linebase, start, end;
};

// line-dependent part (recomputed only when the stage sequence enters a new line)
__device__ __forceinline__ void tg_stage_line(const TgWin& w, TgStage& d, long long pl) {
  const int nr0 = w.nr[0], nr1 = w.nr[1];
  int lo1 = 0, lo2 = 0;
  d.len1 = 1;
  d.len2 = 1;
  d.linebase = 0;
  if (w.dim > 1) {
    lo1 = __ldg(w.lo[1] + d.r1);
    d.len1 = __ldg(w.hi[1] + d.r1) - lo1 + 1;
  }
  if (w.dim > 2) {
    lo2 = __ldg(w.lo[2] + d.r2);
    d.len2 = __ldg(w.hi[2] + d.r2) - lo2 + 1;
  }
  d.tile = d.len1 * d.len2;
  if (w.dim > 1) {
    const long long T0 = __ldg(w.S[0] + nr0);
    long long inner = __ldg(w.S[1] + d.r1) * d.len2;
    if (w.dim > 2) inner += __ldg(w.S[1] + nr1) * __ldg(w.S[2] + d.r2);
    d.linebase = T0 * inner;
  }
  d.xline = (long long)w.nc[0] * lo1 + pl * lo2;
}
// first-direction window data of every row of a line, staged in shared memory once
struct TgDir0 {
  const int* lo;            // [nr0]
  const int* len;           // [nr0]
  const long long* S;       // [nr0 + 1]
};
// chunk-dependent part
__device__ __forceinline__ void tg_stage_chunk(const TgWin& w, TgStage& d, const TgDir0& z) {
  d.r0a = d.ch * TG_TMA_ROWS;
  d.r0b = min(w.nr[0], d.r0a + TG_TMA_ROWS);
  d.c0lo = z.lo[d.r0a];
  d.cw = z.lo[d.r0b - 1] + z.len[d.r0b - 1] - d.c0lo;
  d.start = d.linebase + z.S[d.r0a] * d.tile;
  d.end = d.linebase + z.S[d.r0b] * d.tile;
}
__device__ __forceinline__ TgStage tg_stage_first(const TgWin& w, long long g, int spl,
                                                  long long pl, const TgDir0& z) {
  TgStage d;
  d.line = g / spl;
  d.ch = (int)(g - d.line * spl);
  d.r2 = (int)(d.line / w.nr[1]);
  d.r1 = (int)(d.line - (long long)d.r2 * w.nr[1]);
  tg_stage_line(w, d, pl);
  tg_stage_chunk(w, d, z);
  return d;
}
// returns true if the line changed
__device__ __forceinline__ bool tg_stage_next(const TgWin& w, TgStage& d, int spl, long long pl,
                                              const TgDir0& z) {
  bool newline = false;
  if (++d.ch == spl) {
    d.ch = 0;
    d.line++;
    if (++d.r1 == w.nr[1]) {
      d.r1 = 0;
      d.r2++;
    }
    tg_stage_line(w, d, pl);
    newline = true;
  }
  tg_stage_chunk(w, d, z);
  return newline;
}

template <bool DOT>
__global__ void __launch_bounds__(TG_TMA_THREADS)
k_win_spmv_tma(TgWin w, const double* __restrict__ vals, const double* __restrict__ x,
               int64_t xoff, double* __restrict__ y, long long nstage_tot, int spl, int w0max,
               int maxtile, int NS, int stage_doubles, double* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char tsm[];
  double* bufs = (double*)tsm;                                       // [NS][stage_doubles]
  const int XT = TG_TMA_XW * TG_TMA_XPT * 16;                        // 64 groups x 16 columns
  double* xs = bufs + (size_t)NS * stage_doubles;                    // [2][XT]
  uint64_t* full = (uint64_t*)(xs + 2 * XT);
  uint64_t* empty = full + NS;
  long long* gtab = (long long*)(empty + NS);                        // [2][64] x offset of group g
  long long* S0s = gtab + 2 * 64;                                    // [nr0 + 1]
  int* lo0s = (int*)(S0s + w.nr[0] + 1);                             // [nr0]
  int* len0s = lo0s + w.nr[0];                                       // [nr0]
  int* xtab = len0s + w.nr[0];                                       // [w0max * maxtile]
  __shared__ double sh[32];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nr0 = w.nr[0];
  const int nc0 = w.nc[0];
  const long long pl = (long long)nc0 * w.nc[1];
  // contiguous range of stages of this CTA
  const long long st0 = nstage_tot * blockIdx.x / gridDim.x;
  const long long st1 = nstage_tot * (blockIdx.x + 1) / gridDim.x;
  const int nst = (int)(st1 - st0);

  if (tid == 0) {
    for (int s = 0; s < NS; s++) {
      tg_mbar_init(&full[s], 1);
      tg_mbar_init(&empty[s], TG_TMA_ROWS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // x-tile offset of window position p of a full-width row (relative to the row's lo0)
  for (int p = tid; p < w0max * maxtile; p += TG_TMA_THREADS)
    xtab[p] = (p / w0max) * TG_TMA_XW + p % w0max;
  for (int r = tid; r <= w.nr[0]; r += TG_TMA_THREADS) {
    S0s[r] = __ldg(w.S[0] + r);
    if (r < w.nr[0]) {
      const int l = __ldg(w.lo[0] + r);
      lo0s[r] = l;
      len0s[r] = __ldg(w.hi[0] + r) - l + 1;
    }
  }
  __syncthreads();
  TgDir0 z;
  z.lo = lo0s;
  z.len = len0s;
  z.S = S0s;

  double dot = 0.0;
  if (wid == TG_TMA_ROWS) {
    // ---------------- producer warp (one lane) ----------------
    if (lane == 0 && nst > 0) {
      TgStage d = tg_stage_first(w, st0, spl, pl, z);
      for (int t = 0; t < nst; t++) {
        const int s = t % NS;
        const int k = t / NS;
        if (k > 0) tg_mbar_wait(&empty[s], (uint32_t)((k - 1) & 1));
        const long long astart = d.start & ~1LL;                  // 16-byte aligned
        const uint32_t bytes = (uint32_t)((((d.end - astart) * 8) + 15) & ~15LL);
        tg_mbar_expect_tx(&full[s], bytes);
        tg_bulk_g2s(bufs + (size_t)s * stage_doubles, vals + astart, bytes, &full[s]);
        if (t + 1 < nst) tg_stage_next(w, d, spl, pl, z);
      }
    }
  } else if (nst > 0) {
    // ---------------- consumer warps: one row per stage each ----------------
    // x tile: thread tid owns tile slots tid + 256 j = (group tid/16 + 16 j, column tid%16)
    const int xg = tid >> 4, xc = tid & 15;
    double xr_[TG_TMA_XPT];
    // gtab[g] = x offset of group g = (c1, c2) of the line's column box
    auto build_gtab = [&](const TgStage& d, long long* gt) {
      if (tid < 64) {
        const int c2 = tid / d.len1, c1 = tid - c2 * d.len1;
        gt[tid] = (tid < d.tile) ? (long long)nc0 * c1 + pl * c2 : -1;
      }
    };
    auto xload = [&](const TgStage& d, const long long* gt) {
      const double* xb = x + d.xline + d.c0lo + xc;
#pragma unroll
      for (int j = 0; j < TG_TMA_XPT; j++) {
        const long long go = gt[xg + 16 * j];
        xr_[j] = (go >= 0 && xc < d.cw) ? xb[go] : 0.0;
      }
    };
    TgStage cur = tg_stage_first(w, st0, spl, pl, z), nxt;
    int gsel = 0;                                  // gtab buffer of the current line
    build_gtab(cur, gtab);
    asm volatile("bar.sync 1, %0;" ::"n"(TG_TMA_CONSUMERS) : "memory");
    xload(cur, gtab);
#pragma unroll
    for (int j = 0; j < TG_TMA_XPT; j++) xs[tid + TG_TMA_CONSUMERS * j] = xr_[j];
    asm volatile("bar.sync 1, %0;" ::"n"(TG_TMA_CONSUMERS) : "memory");
    for (int t = 0; t < nst; t++) {
      const int s = t % NS;
      const int k = t / NS;
      const double* xt = xs + (t & 1) * XT;
      const bool more = t + 1 < nst;
      if (more) {
        nxt = cur;
        if (tg_stage_next(w, nxt, spl, pl, z)) {   // new line: its group table (other buffer)
          gsel ^= 1;
          build_gtab(nxt, gtab + 64 * gsel);
          asm volatile("bar.sync 1, %0;" ::"n"(TG_TMA_CONSUMERS) : "memory");
        }
        xload(nxt, gtab + 64 * gsel);
      }
      tg_mbar_wait(&full[s], (uint32_t)(k & 1));
      const int r0 = cur.r0a + wid;
      if (r0 < cur.r0b) {
        const int lo0 = lo0s[r0];
        const int len0 = len0s[r0];
        const long long rstart = cur.start + (S0s[r0] - S0s[cur.r0a]) * cur.tile;
        const double* sv = bufs + (size_t)s * stage_doubles + (rstart - (cur.start & ~1LL));
        const int n = len0 * cur.tile;
        const long long row = r0 + (long long)nr0 * cur.line;
        double acc = 0.0;
        if (cur.cw <= TG_TMA_XW) {
          const double* xq = xt + (lo0 - cur.c0lo);
          if (len0 == w0max) {
            double acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
            int p = lane;
            for (; p + 96 < n; p += 128) {
              acc += sv[p] * xq[xtab[p]];
              acc1 += sv[p + 32] * xq[xtab[p + 32]];
              acc2 += sv[p + 64] * xq[xtab[p + 64]];
              acc3 += sv[p + 96] * xq[xtab[p + 96]];
            }
            for (; p < n; p += 32) acc += sv[p] * xq[xtab[p]];
            acc += (acc1 + acc2) + acc3;
          } else {
            for (int p = lane; p < n; p += 32) {
              const int g = p / len0, c0 = p - g * len0;
              acc += sv[p] * xq[g * TG_TMA_XW + c0];
            }
          }
        } else {                       // window union wider than the tile: gather from global
          const double* xg_ = x + cur.xline + lo0;
          for (int p = lane; p < n; p += 32) {
            const int g = p / len0, c0 = p - g * len0;
            const int c2 = g / cur.len1, c1 = g - c2 * cur.len1;
            acc += sv[p] * xg_[c0 + (long long)nc0 * c1 + pl * c2];
          }
        }
        acc = tg_warp_sum(acc);
        if (lane == 0) {
          y[row] = acc;
          if (DOT) dot += x[xoff + row] * acc;
        }
      }
      __syncwarp();
      if (lane == 0) tg_mbar_arrive(&empty[s]);
      if (more) {
        double* dst = xs + ((t + 1) & 1) * XT;
#pragma unroll
        for (int j = 0; j < TG_TMA_XPT; j++) dst[tid + TG_TMA_CONSUMERS * j] = xr_[j];
        cur = nxt;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TG_TMA_CONSUMERS) : "memory");
    }
  }
  if (DOT) {
    dot = tg_block_sum_ws(dot, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = dot;
  }
}

static int g_tma_smem_ok = -1;
static int g_last_spmv_kind = 0;      // 0 direct-load, 1 TMA-staged, 2 SELL
extern "C" int tg_last_spmv_kind(void) { return g_last_spmv_kind; }

// returns 0; *launched = 1 if the staged kernel took the call
static int tg_win_spmv_tma_try(const tg_win* h_w, const double* vals, const double* x,
                               int64_t xoff, double* y, double* part, cudaStream_t st,
                               int* launched) {
  *launched = 0;
  // Opt-in (TIGAR_B200_TMA_SPMV=1): 0.39-0.41 of HBM peak on B200 against 0.70 for the
  // direct-load kernel -- with 2 CTAs/SM in lock step the per-stage chain (x-tile
  // loads, mbarrier wait, reduction, CTA barrier) is not hidden (profiles/r1_spmv_variants.txt)
  const char* e = getenv("TIGAR_B200_TMA_SPMV");
  if (!e || atoi(e) == 0) return 0;
  if (h_w->layout != 0 || h_w->maxrow <= 0 || h_w->w0max <= 0) return 0;
  if ((((uintptr_t)vals) & 15) != 0) return 0;
  const int w0max = h_w->w0max;
  const int maxtile = h_w->maxrow / w0max;
  if (TG_TMA_ROWS + w0max - 1 > TG_TMA_XW) return 0;
  if (maxtile > 16 * TG_TMA_XPT) return 0;             // x tile: 64 groups x 16 columns
  const int64_t nlines = tg_win_nrows(h_w) / h_w->nr[0];
  const int spl = (int)tg_cdiv(h_w->nr[0], TG_TMA_ROWS);
  const long long nstage_tot = (long long)nlines * spl;
  if (nstage_tot < 1) return 0;
  int stage_doubles = (int)(((int64_t)TG_TMA_ROWS * h_w->maxrow + 2 + 15) & ~15);
  size_t stage_bytes = (size_t)stage_doubles * 8;
  if (h_w->nr[0] > 1024) return 0;                     // first-direction tables in smem
  const size_t fixed = (size_t)2 * TG_TMA_XW * TG_TMA_XPT * 16 * 8 + 2 * 64 * 8 +
                       (size_t)w0max * maxtile * 4 + 2 * 8 * 8 + 16 * (size_t)h_w->nr[0] + 8 + 256;
  const size_t budget = 112 * 1024;
  if (fixed + 2 * stage_bytes > budget) return 0;
  int NS = (int)((budget - fixed) / stage_bytes);
  if (NS > 6) NS = 6;
  size_t smem = (size_t)NS * stage_bytes + fixed;
  if (g_tma_smem_ok < 0) {
    cudaError_t e1 = cudaFuncSetAttribute(k_win_spmv_tma<true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    cudaError_t e2 = cudaFuncSetAttribute(k_win_spmv_tma<false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    g_tma_smem_ok = (e1 == cudaSuccess && e2 == cudaSuccess) ? 1 : 0;
  }
  if (!g_tma_smem_ok) return 0;
  long long g = tg_ws_grid_size() / 2;          // 2 CTAs per SM
  if (g > nstage_tot) g = nstage_tot;
  TgWin w = tg_win_dev(h_w);
  if (part) {
    // unused partial sums of the wider reduction grid must be zero
    TG_CHECK(cudaMemsetAsync(part, 0, sizeof(double) * tg_ws_grid_size(), st));
    k_win_spmv_tma<true><<<(unsigned)g, TG_TMA_THREADS, smem, st>>>(
        w, vals, x, xoff, y, nstage_tot, spl, w0max, maxtile, NS, stage_doubles, part);
  } else {
    k_win_spmv_tma<false><<<(unsigned)g, TG_TMA_THREADS, smem, st>>>(
        w, vals, x, xoff, y, nstage_tot, spl, w0max, maxtile, NS, stage_doubles, nullptr);
  }
  TG_LAUNCH_CHECK();
  *launched = 1;
  g_last_spmv_kind = 1;
  return 0;
}

// ---------------------------------------------------------------------------
// SELL-H layout (tg_win.layout == 1): lanes = rows.  A warp owns one
// (line, chunk) item = H consecutive rows; per slot the 32 lanes read 32
// consecutive values (one coalesced line of the stream) and 32 consecutive x
// entries.  No reduction, no per-row index math, W0 independent loads in flight
// per lane per inner iteration.
template <int W0, bool DOT, int G>
__global__ void __launch_bounds__(256)
k_sell_spmv(TgWin w, const double* __restrict__ vals, const double* __restrict__ x, int64_t xoff,
            double* __restrict__ y, int nchunk, int64_t nitems, double* __restrict__ part) {
  // One CTA per (line, chunk) item = H consecutive rows; its 8 warps split the
  // (c1,c2) slot groups, so the CTA streams one contiguous block of values and
  // all warps gather from the same small x neighbourhood (stays in L1).
  __shared__ double sh[32];
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int H = w.H;
  const int w0 = (W0 > 0) ? W0 : w.w0max;
  const int nr0 = w.nr[0], nr1 = w.nr[1];
  const int64_t nc0 = w.nc[0];
  const int64_t pl = nc0 * w.nc[1];
  const int64_t T1 = (w.dim > 1) ? w.S[1][nr1] : 1;
  const int64_t linemul = (int64_t)H * nchunk * w0;
  double dot = 0.0;
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t line = item / nchunk;
    const int ch = (int)(item - line * nchunk);
    const int r2 = (int)(line / nr1);
    const int r1 = (int)(line - (int64_t)r2 * nr1);
    int lo1 = 0, len1 = 1, lo2 = 0, len2 = 1;
    int64_t inner = 0;
    if (w.dim > 1) {
      lo1 = __ldg(w.lo[1] + r1);
      len1 = __ldg(w.hi[1] + r1) - lo1 + 1;
    }
    if (w.dim > 2) {
      lo2 = __ldg(w.lo[2] + r2);
      len2 = __ldg(w.hi[2] + r2) - lo2 + 1;
    }
    if (w.dim > 1) inner = __ldg(w.S[1] + r1) * len2;
    if (w.dim > 2) inner += T1 * __ldg(w.S[2] + r2);
    const int ngroups = len1 * len2;
    const int64_t slots = (int64_t)w0 * ngroups;
    const int r0 = ch * H + lane;
    const bool valid = lane < H && r0 < nr0;
    double acc0 = 0.0, acc1 = 0.0;
    if (valid) {
      const double* __restrict__ abase = vals + linemul * inner + (int64_t)ch * H * slots + lane;
      const double* xb = x + __ldg(w.bs0 + r0) + nc0 * lo1 + pl * lo2;
      if (W0 > 0) {
        // G slot groups per trip: G*W0 value loads and G*W0 x loads in flight
        for (int g0 = wid * G; g0 < ngroups; g0 += 8 * G) {
          double av[G][W0 > 0 ? W0 : 1], xv[G][W0 > 0 ? W0 : 1];
#pragma unroll
          for (int u = 0; u < G; u++) {
            const int g = min(g0 + u, ngroups - 1);
            const double* __restrict__ a = abase + (int64_t)g * W0 * H;
#pragma unroll
            for (int k = 0; k < W0; k++) av[u][k] = __ldcs(a + (int64_t)k * H);
          }
#pragma unroll
          for (int u = 0; u < G; u++) {
            const int g = min(g0 + u, ngroups - 1);
            const int c2 = g / len1, c1 = g - c2 * len1;
            const double* xp = xb + nc0 * c1 + pl * c2;
#pragma unroll
            for (int k = 0; k < W0; k++) xv[u][k] = xp[k];
          }
#pragma unroll
          for (int u = 0; u < G; u++) {
            if (g0 + u < ngroups) {
#pragma unroll
              for (int k = 0; k < W0; k++) {
                if (k & 1) acc1 += av[u][k] * xv[u][k];
                else acc0 += av[u][k] * xv[u][k];
              }
            }
          }
        }
      } else {
        for (int g = wid; g < ngroups; g += 8) {
          const int c2 = g / len1, c1 = g - c2 * len1;
          const double* __restrict__ a = abase + (int64_t)g * w0 * H;
          const double* xp = xb + nc0 * c1 + pl * c2;
          for (int k = 0; k < w0; k++) acc0 += __ldcs(a + (int64_t)k * H) * xp[k];
        }
      }
    }
    red[wid][lane] = acc0 + acc1;
    __syncthreads();
    if (wid == 0 && valid) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 8; j++) acc += red[j][lane];
      const int64_t row = r0 + (int64_t)nr0 * line;
      y[row] = acc;
      if (DOT) dot += x[xoff + row] * acc;
    }
    __syncthreads();
  }
  if (DOT) {
    dot = tg_block_sum_ws(dot, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = dot;
  }
}

template <bool DOT, int G>
static void tg_sell_launch_g(int w0, int g, cudaStream_t st, const TgWin& w, const double* vals,
                             const double* x, int64_t xoff, double* y, int nchunk,
                             int64_t nitems, double* part) {
  switch (w0) {
    case 3: k_sell_spmv<3, DOT, G><<<g, 256, 0, st>>>(w, vals, x, xoff, y, nchunk, nitems, part); break;
    case 5: k_sell_spmv<5, DOT, G><<<g, 256, 0, st>>>(w, vals, x, xoff, y, nchunk, nitems, part); break;
    case 7: k_sell_spmv<7, DOT, G><<<g, 256, 0, st>>>(w, vals, x, xoff, y, nchunk, nitems, part); break;
    case 9: k_sell_spmv<9, DOT, G><<<g, 256, 0, st>>>(w, vals, x, xoff, y, nchunk, nitems, part); break;
    default: k_sell_spmv<0, DOT, 1><<<g, 256, 0, st>>>(w, vals, x, xoff, y, nchunk, nitems, part); break;
  }
}

static int g_sell_G = 0;
template <bool DOT>
static void tg_sell_launch(int w0, int g, cudaStream_t st, const TgWin& w, const double* vals,
                           const double* x, int64_t xoff, double* y, int nchunk, int64_t nitems,
                           double* part) {
  if (!g_sell_G) {
    const char* e = getenv("TIGAR_B200_SELL_G");
    g_sell_G = e ? atoi(e) : 2;
    if (g_sell_G != 1 && g_sell_G != 2 && g_sell_G != 4) g_sell_G = 2;
  }
  if (g_sell_G == 1) tg_sell_launch_g<DOT, 1>(w0, g, st, w, vals, x, xoff, y, nchunk, nitems, part);
  else if (g_sell_G == 2) tg_sell_launch_g<DOT, 2>(w0, g, st, w, vals, x, xoff, y, nchunk, nitems, part);
  else tg_sell_launch_g<DOT, 4>(w0, g, st, w, vals, x, xoff, y, nchunk, nitems, part);
}

static inline int64_t tg_win_lines(const tg_win* w) {
  int64_t n = 1;
  for (int d = 1; d < w->dim; d++) n *= w->nr[d];
  return n;
}

int tg_win_spmv_launch(const tg_win* h_w, const double* vals, const double* x, int64_t xoff,
                       double* y, double* part, cudaStream_t st) {
  TG_REQUIRE(h_w->dim >= 1 && h_w->dim <= 3, "dim");
  if (h_w->layout == 1) {
    TG_REQUIRE(h_w->H >= 1 && h_w->H <= 32 && h_w->bs0 != nullptr, "SELL descriptor");
    int nchunk = (int)tg_cdiv(h_w->nr[0], h_w->H);
    int64_t nitems = tg_win_lines(h_w) * nchunk;
    if (nitems == 0) return 0;
    int g = tg_ws_grid_size() * 2;          // 8 CTAs of 256 threads per SM
    if ((int64_t)g > nitems) g = (int)nitems;
    TgWin w = tg_win_dev(h_w);
    if (part) {
      TG_CHECK(cudaMemsetAsync(part, 0, sizeof(double) * tg_ws_grid_size() * 2, st));
      tg_sell_launch<true>(h_w->w0max, g, st, w, vals, x, xoff, y, nchunk, nitems, part);
    } else {
      tg_sell_launch<false>(h_w->w0max, g, st, w, vals, x, xoff, y, nchunk, nitems, nullptr);
    }
    TG_LAUNCH_CHECK();
    g_last_spmv_kind = 2;
    return 0;
  }
  {
    int launched = 0;
    int rc = tg_win_spmv_tma_try(h_w, vals, x, xoff, y, part, st, &launched);
    if (rc || launched) return rc;
  }
  int nchunk = (int)tg_cdiv(h_w->nr[0], TG_WS_ROWS);
  int64_t nitems = tg_win_lines(h_w) * nchunk;
  TG_REQUIRE(nitems < (int64_t)2147483647, "too many row items");
  if (nitems == 0) return 0;
  int g = tg_ws_grid_size();
  TgWin w = tg_win_dev(h_w);
  TG_REQUIRE(h_w->w0max >= 1, "window descriptor lacks w0max");
  static int U = 0;
  if (!U) {
    const char* e = getenv("TIGAR_B200_SPMV_U");
    U = e ? atoi(e) : 4;
    if (U != 4 && U != 6 && U != 8) U = 4;
  }
  static int PF = -1;
  if (PF < 0) {
    // opt-in: measured 0.46-0.50 of HBM peak against 0.70 without the prefetch (8-byte
    // cp.async has no evict-first / no-allocate form, so the streamed rows evict the x
    // entries the gathers want from L1; profiles/r1_spmv_variants.txt)
    const char* e = getenv("TIGAR_B200_SPMV_PF");
    PF = e ? atoi(e) : 0;
  }
  if (PF && h_w->maxrow > 0 && h_w->maxrow <= 704) {
    // per-warp double-buffered row prefetch: 8 warps x 2 x RB doubles per CTA, 4 CTAs/SM
    const int RB = (h_w->maxrow + 1) & ~1;
    const size_t smem = (size_t)(TG_WS_BLOCK / 32) * 2 * RB * 8;
    static int attr_ok = 0;
    if (!attr_ok) {
      TG_CHECK(cudaFuncSetAttribute(k_win_spmv_pf<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      TG_CHECK(cudaFuncSetAttribute(k_win_spmv_pf<false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_ok = 1;
    }
    if (part)
      k_win_spmv_pf<true><<<g, TG_WS_BLOCK, smem, st>>>(w, vals, x, xoff, y, nchunk, (int)nitems,
                                                        h_w->w0max, RB, part);
    else
      k_win_spmv_pf<false><<<g, TG_WS_BLOCK, smem, st>>>(w, vals, x, xoff, y, nchunk,
                                                         (int)nitems, h_w->w0max, RB, nullptr);
    TG_LAUNCH_CHECK();
    g_last_spmv_kind = 3;
    return 0;
  }
  if (U > 4) {                       // 3 resident CTAs per SM: keep it one wave
    g = (g / 4) * 3;
    if (part) TG_CHECK(cudaMemsetAsync(part, 0, sizeof(double) * tg_ws_grid_size(), st));
  }
#define TG_SPMV_LAUNCH(UU)                                                                    \
  if (part)                                                                                   \
    k_win_spmv<true, UU><<<g, TG_WS_BLOCK, 0, st>>>(w, vals, x, xoff, y, nchunk, (int)nitems, \
                                                    h_w->w0max, part);                        \
  else                                                                                        \
    k_win_spmv<false, UU><<<g, TG_WS_BLOCK, 0, st>>>(w, vals, x, xoff, y, nchunk,             \
                                                     (int)nitems, h_w->w0max, nullptr);
  static int LDM = -1;
  if (LDM < 0) {
    const char* e = getenv("TIGAR_B200_SPMV_LD");
    LDM = e ? atoi(e) : 0;
  }
  if (U == 4 && LDM == 1) {
    if (part)
      k_win_spmv<true, 4, 1><<<g, TG_WS_BLOCK, 0, st>>>(w, vals, x, xoff, y, nchunk, (int)nitems,
                                                       h_w->w0max, part);
    else
      k_win_spmv<false, 4, 1><<<g, TG_WS_BLOCK, 0, st>>>(w, vals, x, xoff, y, nchunk,
                                                        (int)nitems, h_w->w0max, nullptr);
  } else if (U == 4 && LDM == 2) {
    if (part)
      k_win_spmv<true, 4, 2><<<g, TG_WS_BLOCK, 0, st>>>(w, vals, x, xoff, y, nchunk, (int)nitems,
                                                       h_w->w0max, part);
    else
      k_win_spmv<false, 4, 2><<<g, TG_WS_BLOCK, 0, st>>>(w, vals, x, xoff, y, nchunk,
                                                        (int)nitems, h_w->w0max, nullptr);
  } else if (U == 4) { TG_SPMV_LAUNCH(4) }
  else if (U == 6) { TG_SPMV_LAUNCH(6) }
  else { TG_SPMV_LAUNCH(8) }
#undef TG_SPMV_LAUNCH
  TG_LAUNCH_CHECK();
  g_last_spmv_kind = 0;
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
  // layout 0: one warp per row, lanes over its entries.
  // layout 1: one thread per row (consecutive threads = consecutive rows of a
  //           chunk, so every slot access is coalesced).
  int64_t r;
  int lane, step;
  if (w.layout == 0) {
    r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    lane = threadIdx.x & 31;
    step = 32;
  } else {
    r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    lane = 0;
    step = 1;
  }
  if (r >= nrows) return;
  int rc[3];
  tg_decode(r, w.nr, w.dim, rc);
  TgRowWin rw = tg_row_window(w, rc);
  TgRowAddr ra = tg_row_addr(w, rc, rw);
  int cc[3] = {rc[0], rc[1], rc[2]};
  cc[w.dim - 1] += col_shift;
  const int64_t mycol = cc[0] + (int64_t)w.nc[0] * (cc[1] + (int64_t)w.nc[1] * cc[2]);
  const bool mr = rowmask[r] != 0;
  const int tot = ra.len0 * rw.len[1] * rw.len[2];
  for (int pos = lane; pos < tot; pos += step) {
    int c0 = pos % ra.len0;
    int t = pos / ra.len0;
    int c1 = t % rw.len[1];
    int c2 = t / rw.len[1];
    int64_t col = (ra.lo0 + c0) +
                  (int64_t)w.nc[0] * ((rw.lo[1] + c1) + (int64_t)w.nc[1] * (rw.lo[2] + c2));
    if (mr || colmask[col]) vals[ra.base + (int64_t)pos * ra.stride] = (mr && col == mycol) ? diag : 0.0;
  }
}

extern "C" int tg_win_zero_rows_cols(const tg_win* h_w, double* vals, const uint8_t* rowmask,
                                     const uint8_t* colmask, double diag, int32_t col_shift,
                                     void* stream) {
  int64_t nrows = tg_win_nrows(h_w);
  if (nrows == 0) return 0;
  const int64_t nthreads = (h_w->layout == 0) ? nrows * 32 : nrows;
  k_win_zero_rows_cols<<<(unsigned)tg_cdiv(nthreads, 256), 256, 0, tg_stream(stream)>>>(
      tg_win_dev(h_w), vals, nrows, rowmask, colmask, diag, col_shift);
  TG_LAUNCH_CHECK();
  return 0;
}

// Same operation when the constrained set is a union of whole hyperplanes (side DoFs of
// getSideDofs, BSplines.py:599-649, the usual case): hp[d][c] != 0 marks a constrained
// hyperplane c of direction d (GLOBAL coordinates).  A row is touched only if it is constrained
// itself or its window reaches a constrained hyperplane in some direction, so interior rows
// leave after a handful of 1-byte loads instead of testing every column against the mask.
__global__ void k_win_zero_rows_cols_hp(TgWin w, double* __restrict__ vals, int64_t nrows,
                                        const uint8_t* __restrict__ hp0,
                                        const uint8_t* __restrict__ hp1,
                                        const uint8_t* __restrict__ hp2, double diag, int dsel,
                                        const int32_t* __restrict__ sel, int nsel) {
  // dsel >= 0: only the rows whose coordinate in direction dsel is one of sel[0..nsel) (local
  // row coordinates) are visited: warp k handles the k-th row of that sub-grid
  const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  int rc[3];
  if (dsel < 0) {
    if (wid >= nrows) return;
    tg_decode(wid, w.nr, w.dim, rc);
  } else {
    int n[3] = {w.nr[0], w.nr[1], w.nr[2]};
    n[dsel] = nsel;
    if (wid >= (int64_t)n[0] * n[1] * n[2]) return;
    tg_decode(wid, n, w.dim, rc);
    rc[dsel] = sel[rc[dsel]];
  }
  const TgRowWin rw = tg_row_window(w, rc);
  const uint8_t* hp[3] = {hp0, hp1, hp2};
  bool mr = false, touch = false;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (d >= w.dim) continue;
    mr = mr || hp[d][rc[d] + w.row0[d]] != 0;
    for (int c = 0; c < rw.len[d]; c++) touch = touch || hp[d][rw.lo[d] + c + w.col0[d]] != 0;
  }
  if (!mr && !touch) return;
  const TgRowAddr ra = tg_row_addr(w, rc, rw);
  const int tot = rw.len[0] * rw.len[1] * rw.len[2];
  for (int pos = lane; pos < tot; pos += 32) {
    const int c0 = pos % rw.len[0];
    const int t = pos / rw.len[0];
    const int c1 = t % rw.len[1], c2 = t / rw.len[1];
    int cc[3] = {rw.lo[0] + c0 + w.col0[0], rw.lo[1] + c1 + w.col0[1], rw.lo[2] + c2 + w.col0[2]};
    bool mc = false, own = true;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (d >= w.dim) continue;
      mc = mc || hp[d][cc[d]] != 0;
      own = own && cc[d] == rc[d] + w.row0[d];
    }
    if (mr || mc) vals[ra.base + pos] = (mr && own) ? diag : 0.0;
  }
}

// h_sel[d] / h_nsel[d]: device list of the LOCAL row coordinates of direction d whose window
// reaches a constrained hyperplane (or that are constrained themselves); the union over d of
// these sub-grids contains every row the operation touches (it is idempotent, overlaps are
// harmless).  h_sel == NULL: scan all rows.
extern "C" int tg_win_zero_rows_cols_hp(const tg_win* h_w, double* vals, const uint8_t* hp0,
                                        const uint8_t* hp1, const uint8_t* hp2, double diag,
                                        const int32_t* const* h_sel, const int32_t* h_nsel,
                                        void* stream) {
  TG_REQUIRE(h_w->layout == 0, "hyperplane BC kernel needs the row-major window layout");
  int64_t nrows = tg_win_nrows(h_w);
  if (nrows == 0) return 0;
  if (!h_sel) {
    k_win_zero_rows_cols_hp<<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
        tg_win_dev(h_w), vals, nrows, hp0, hp1, hp2, diag, -1, nullptr, 0);
    TG_LAUNCH_CHECK();
    return 0;
  }
  for (int d = 0; d < h_w->dim; d++) {
    if (h_nsel[d] <= 0) continue;
    int64_t n = h_nsel[d];
    for (int e = 0; e < h_w->dim; e++)
      if (e != d) n *= h_w->nr[e];
    k_win_zero_rows_cols_hp<<<(unsigned)tg_cdiv(n * 32, 256), 256, 0, tg_stream(stream)>>>(
        tg_win_dev(h_w), vals, nrows, hp0, hp1, hp2, diag, d, h_sel[d], h_nsel[d]);
    TG_LAUNCH_CHECK();
  }
  return 0;
}

__global__ void k_win_diag_inv(TgWin w, const double* __restrict__ vals, int64_t nrows,
                               int col_shift, double* __restrict__ dinv) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  int rc[3];
  tg_decode(r, w.nr, w.dim, rc);
  TgRowWin rw = tg_row_window(w, rc);
  TgRowAddr ra = tg_row_addr(w, rc, rw);
  int cc[3] = {rc[0], rc[1], rc[2]};
  cc[w.dim - 1] += col_shift;
  const int64_t pos = ((int64_t)(cc[2] - rw.lo[2]) * rw.len[1] + (cc[1] - rw.lo[1])) * ra.len0 +
                      (cc[0] - ra.lo0);
  double d = vals[ra.base + pos * ra.stride];
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


// ---- layout conversion: exact row-major CSR order <-> the window's layout -----
template <bool EXPORT>
__global__ void k_win_convert(TgWin w, const double* __restrict__ src, double* __restrict__ dst,
                              int64_t nrows) {
  int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  int rc[3];
  tg_decode(r, w.nr, w.dim, rc);
  TgRowWin rw = tg_row_window(w, rc);
  TgRowAddr ra = tg_row_addr(w, rc, rw);
  const int64_t base = w.rowptr[r];
  const int tot = rw.len[0] * rw.len[1] * rw.len[2];
  for (int pos = lane; pos < tot; pos += 32) {
    int c0 = pos % rw.len[0];
    int t = pos / rw.len[0];           // = c2*len1 + c1
    int64_t p = ra.base + ((int64_t)t * ra.len0 + (rw.lo[0] + c0 - ra.lo0)) * ra.stride;
    if (EXPORT) dst[base + pos] = src[p];
    else dst[p] = src[base + pos];
  }
}

extern "C" int64_t tg_win_storage(const tg_win* h_w, const int64_t* h_T) {
  // h_T: totals of the per-direction window lengths (exact nnz = T0*T1*T2)
  int64_t T[3] = {1, 1, 1};
  for (int d = 0; d < h_w->dim; d++) T[d] = h_T[d];
  if (h_w->layout == 0) return T[0] * T[1] * T[2];
  return (int64_t)h_w->H * tg_cdiv(h_w->nr[0], h_w->H) * h_w->w0max * T[1] * T[2];
}

extern "C" int tg_win_export_vals(const tg_win* h_w, const double* vals, double* out_csr,
                                  void* stream) {
  int64_t nrows = tg_win_nrows(h_w);
  if (nrows == 0) return 0;
  k_win_convert<true><<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
      tg_win_dev(h_w), vals, out_csr, nrows);
  TG_LAUNCH_CHECK();
  return 0;
}

extern "C" int tg_win_import_vals(const tg_win* h_w, const double* in_csr, double* vals,
                                  void* stream) {
  int64_t nrows = tg_win_nrows(h_w);
  if (nrows == 0) return 0;
  k_win_convert<false><<<(unsigned)tg_cdiv(nrows * 32, 256), 256, 0, tg_stream(stream)>>>(
      tg_win_dev(h_w), in_csr, vals, nrows);
  TG_LAUNCH_CHECK();
  return 0;
}
