// Shared helpers for the tigar_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tigar_b200.h"

void tg_set_error(const char* fmt, ...);

#define TG_CHECK(call)                                                        \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      tg_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call,                \
                   cudaGetErrorString(e__));                                  \
      return 1;                                                               \
    }                                                                         \
  } while (0)

void tg_count_launch();
#define TG_LAUNCH_CHECK()            \
  do {                               \
    tg_count_launch();               \
    TG_CHECK(cudaGetLastError());    \
  } while (0)

#define TG_REQUIRE(cond, msg)                                                 \
  do {                                                                        \
    if (!(cond)) {                                                            \
      tg_set_error("%s:%d: requirement failed: %s (%s)", __FILE__, __LINE__,  \
                   #cond, msg);                                               \
      return 2;                                                               \
    }                                                                         \
  } while (0)

static inline cudaStream_t tg_stream(void* s) { return (cudaStream_t)s; }

static inline int64_t tg_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device copies of the descriptors (passed by value to kernels) --------
struct TgWin {
  int dim;
  int nr[3], nc[3];
  const int32_t* lo[3];
  const int32_t* hi[3];
  const int64_t* rowptr;
  const int64_t* S[3];
  int row0[3], col0[3];
  int layout, H, w0max;
  const int32_t* bs0;
};

static inline TgWin tg_win_dev(const tg_win* w) {
  TgWin d;
  d.dim = w->dim;
  for (int k = 0; k < 3; k++) {
    d.nr[k] = (k < w->dim) ? w->nr[k] : 1;
    d.nc[k] = (k < w->dim) ? w->nc[k] : 1;
    d.lo[k] = (k < w->dim) ? w->lo[k] : nullptr;
    d.hi[k] = (k < w->dim) ? w->hi[k] : nullptr;
    d.S[k] = (k < w->dim) ? w->S[k] : nullptr;
    d.row0[k] = (k < w->dim) ? w->row0[k] : 0;
    d.col0[k] = (k < w->dim) ? w->col0[k] : 0;
  }
  d.rowptr = w->rowptr;
  d.layout = w->layout;
  d.H = w->H;
  d.w0max = w->w0max;
  d.bs0 = w->bs0;
  return d;
}

static inline int64_t tg_win_nrows(const tg_win* w) {
  int64_t n = 1;
  for (int k = 0; k < w->dim; k++) n *= w->nr[k];
  return n;
}

struct TgBasis {
  int dim;
  int n[3], nel[3], nloc[3], nq[3];
  int nder;
  const double* tab[3];
  const int32_t* idx[3];
  const double* wq[3];
  const double* xq[3];
};

static inline TgBasis tg_basis_dev(const tg_basis* b) {
  TgBasis d;
  d.dim = b->dim;
  d.nder = b->nder;
  for (int k = 0; k < 3; k++) {
    bool in = k < b->dim;
    d.n[k] = in ? b->n[k] : 1;
    d.nel[k] = in ? b->nel[k] : 1;
    d.nloc[k] = in ? b->nloc[k] : 1;
    d.nq[k] = in ? b->nq[k] : 1;
    d.tab[k] = in ? b->tab[k] : nullptr;
    d.idx[k] = in ? b->idx[k] : nullptr;
    d.wq[k] = in ? b->wq[k] : nullptr;
    d.xq[k] = in ? b->xq[k] : nullptr;
  }
  return d;
}

// row -> (r0,r1,r2), first direction fastest (BSplines.py:354-370)
__host__ __device__ inline void tg_decode(int64_t idx, const int* n, int dim, int* c) {
  c[0] = (int)(idx % n[0]);
  c[1] = 0;
  c[2] = 0;
  if (dim > 1) {
    int64_t r = idx / n[0];
    c[1] = (int)(r % n[1]);
    if (dim > 2) c[2] = (int)(r / n[1]);
  }
}

// window of row r: per-direction lo/len
struct TgRowWin {
  int lo[3];
  int len[3];
};

__device__ inline TgRowWin tg_row_window(const TgWin& w, const int* r) {
  TgRowWin rw;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (d < w.dim) {
      rw.lo[d] = w.lo[d][r[d]];
      rw.len[d] = w.hi[d][r[d]] - rw.lo[d] + 1;
    } else {
      rw.lo[d] = 0;
      rw.len[d] = 1;
    }
  }
  return rw;
}

// Address of entry (d0,d1,d2) of a row (offsets inside its window):
//   base + ((d2*len1 + d1)*len0 + d0) * stride,  d0 = c0 - lo0
struct TgRowAddr {
  long long base;
  int stride, len0, lo0;
};

__device__ inline TgRowAddr tg_row_addr(const TgWin& w, const int* rc, const TgRowWin& rw) {
  TgRowAddr a;
  if (w.layout == 0) {
    const long long row = rc[0] + (long long)w.nr[0] * (rc[1] + (long long)w.nr[1] * rc[2]);
    a.base = w.rowptr[row];
    a.stride = 1;
    a.len0 = rw.len[0];
    a.lo0 = rw.lo[0];
  } else {
    const int H = w.H;
    const int chunk = rc[0] / H, lane = rc[0] - chunk * H;
    const long long nchunk = (w.nr[0] + H - 1) / H;
    const long long slots = (long long)w.w0max * rw.len[1] * rw.len[2];
    long long inner = 0;
    if (w.dim > 1) inner = w.S[1][rc[1]] * rw.len[2];
    if (w.dim > 2) inner += w.S[1][w.nr[1]] * w.S[2][rc[2]];
    a.base = (long long)H * nchunk * w.w0max * inner + (long long)chunk * H * slots + lane;
    a.stride = H;
    a.len0 = w.w0max;
    a.lo0 = w.bs0[rc[0]];
  }
  return a;
}

// position of column c inside the window (caller guarantees containment)
__device__ inline int tg_win_pos(const TgRowWin& rw, const int* c) {
  return ((c[2] - rw.lo[2]) * rw.len[1] + (c[1] - rw.lo[1])) * rw.len[0] + (c[0] - rw.lo[0]);
}

__device__ inline double tg_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier / 1-D bulk async copy (TMA, SASS UBLKCP) helpers ------------
__device__ __forceinline__ uint32_t tg_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tg_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tg_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tg_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tg_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tg_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t a = tg_smem_u32(bar);
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tg_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(tg_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(tg_smem_u32(bar))
      : "memory");
}

