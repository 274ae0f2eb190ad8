// Gauss-point coefficient evaluation: jets of FE/IGA functions + a small
// register-machine program per quadrature point.
// Stands in for the FFC-generated per-cell tabulate_tensor coefficient code
// behind dolfin.assemble (common.py:1169,1215-1216); the geometry algebra it
// evaluates is built on the host from calculusUtils.py:18-24,56-69,255-276.
#include "tg_common.cuh"

#define TG_MAXQJ 96
#define TG_MAXOUT 256

enum {
  OP_NOP = 0, OP_CONST = 1, OP_MOV = 2, OP_ADD = 3, OP_SUB = 4, OP_MUL = 5, OP_DIV = 6,
  OP_NEG = 7, OP_SIN = 8, OP_COS = 9, OP_EXP = 10, OP_LOG = 11, OP_SQRT = 12, OP_POW = 13,
  OP_ABS = 14, OP_TAN = 15, OP_TANH = 16, OP_MAX = 17, OP_MIN = 18, OP_SINH = 19,
  OP_COSH = 20, OP_ATAN = 21, OP_GT = 22, OP_SEL = 23
};

struct TgJetSpec {
  int njets;
  const double* coef[TG_MAXQJ];
  short ncomp[TG_MAXQJ];
  short comp[TG_MAXQJ];
  signed char al[TG_MAXQJ][3];
};

struct TgOutSpec {
  int nout;
  int reg[TG_MAXOUT];
};

// One CTA per cell, one thread per Gauss point.  Jets are sum-factorised
// through shared memory (3 small contractions per jet instead of a
// nen-term sum per point), then every thread interprets the program.
template <int NREG>
__global__ void k_qp_eval(TgBasis B, TgJetSpec J, const int4* __restrict__ prog, int nprog,
                          const double* __restrict__ consts, TgOutSpec O, int64_t cell0,
                          int64_t ncells, int nqp, double* __restrict__ out) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int64_t cl = blockIdx.x;
  const int n0 = B.nloc[0], n1 = B.nloc[1], n2 = B.nloc[2];
  const int q0n = B.nq[0], q1n = B.nq[1], q2n = B.nq[2];
  const int nen = n0 * n1 * n2;
  const int nd = B.nder + 1;
  int e[3];
  tg_decode(cell0 + cl, B.nel, B.dim, e);

  // smem: tabs[3][nq][nloc][nd] | c[nen] | s1[q0n*n1*n2] | s2[q0n*q1n*n2]
  double* tb0 = sm;
  double* tb1 = tb0 + q0n * n0 * nd;
  double* tb2 = tb1 + q1n * n1 * nd;
  double* cf = tb2 + q2n * n2 * nd;
  double* s1 = cf + nen;
  double* s2 = s1 + q0n * n1 * n2;
  for (int i = tid; i < q0n * n0 * nd; i += nth) tb0[i] = B.tab[0][(int64_t)e[0] * q0n * n0 * nd + i];
  for (int i = tid; i < q1n * n1 * nd; i += nth)
    tb1[i] = (B.dim > 1) ? B.tab[1][(int64_t)e[1] * q1n * n1 * nd + i] : 1.0;
  for (int i = tid; i < q2n * n2 * nd; i += nth)
    tb2[i] = (B.dim > 2) ? B.tab[2][(int64_t)e[2] * q2n * n2 * nd + i] : 1.0;

  const bool active = tid < nqp;
  int q[3] = {0, 0, 0};
  if (active) tg_decode(tid, B.nq, B.dim, q);

  double R[NREG];
  if (active) {
    double wq = 1.0;
    for (int d = 0; d < B.dim; d++) {
      R[d] = B.xq[d][e[d] * B.nq[d] + q[d]];
      wq *= B.wq[d][e[d] * B.nq[d] + q[d]];
    }
    R[B.dim] = wq;
  }
  const int r0 = B.dim + 1;
  const double* lastc = nullptr;
  int lastcomp = -1;
  for (int j = 0; j < J.njets; j++) {
    const int a0 = J.al[j][0], a1 = (B.dim > 1) ? J.al[j][1] : 0, a2 = (B.dim > 2) ? J.al[j][2] : 0;
    __syncthreads();
    if (J.coef[j] != lastc || J.comp[j] != lastcomp) {
      lastc = J.coef[j];
      lastcomp = J.comp[j];
      for (int a = tid; a < nen; a += nth) {
        int l0 = a % n0, t = a / n0, l1 = t % n1, l2 = t / n1;
        int64_t g = B.idx[0][e[0] * n0 + l0];
        if (B.dim > 1) g += (int64_t)B.n[0] * B.idx[1][e[1] * n1 + l1];
        if (B.dim > 2) g += (int64_t)B.n[0] * B.n[1] * B.idx[2][e[2] * n2 + l2];
        cf[a] = lastc[g * J.ncomp[j] + lastcomp];
      }
      __syncthreads();
    }
    // s1[(l2*n1 + l1)*q0n + q0] = sum_l0 c[l0,l1,l2] tab0[q0][l0][a0]
    for (int o = tid; o < q0n * n1 * n2; o += nth) {
      int qq = o % q0n, r = o / q0n;
      double acc = 0.0;
      for (int l0 = 0; l0 < n0; l0++) acc += cf[r * n0 + l0] * tb0[(qq * n0 + l0) * nd + a0];
      s1[o] = acc;
    }
    __syncthreads();
    // s2[(l2*q1n + q1)*q0n + q0] = sum_l1 s1[l2,l1,q0] tab1[q1][l1][a1]
    for (int o = tid; o < q0n * q1n * n2; o += nth) {
      int qq0 = o % q0n, t = o / q0n, qq1 = t % q1n, l2 = t / q1n;
      double acc = 0.0;
      for (int l1 = 0; l1 < n1; l1++)
        acc += s1[(l2 * n1 + l1) * q0n + qq0] * tb1[(qq1 * n1 + l1) * nd + a1];
      s2[o] = acc;
    }
    __syncthreads();
    if (active) {
      double acc = 0.0;
      for (int l2 = 0; l2 < n2; l2++)
        acc += s2[(l2 * q1n + q[1]) * q0n + q[0]] * tb2[(q[2] * n2 + l2) * nd + a2];
      R[r0 + j] = acc;
    }
  }
  if (!active) return;

  for (int pc = 0; pc < nprog; pc++) {
    int4 in = __ldg(&prog[pc]);
    double a = 0.0, b = 0.0, r;
    if (in.x != OP_CONST) {
      a = R[in.z];
      b = R[in.w];
    }
    switch (in.x) {
      case OP_CONST: r = consts[in.z]; break;
      case OP_MOV: r = a; break;
      case OP_ADD: r = a + b; break;
      case OP_SUB: r = a - b; break;
      case OP_MUL: r = a * b; break;
      case OP_DIV: r = a / b; break;
      case OP_NEG: r = -a; break;
      case OP_SIN: r = sin(a); break;
      case OP_COS: r = cos(a); break;
      case OP_EXP: r = exp(a); break;
      case OP_LOG: r = log(a); break;
      case OP_SQRT: r = sqrt(a); break;
      case OP_POW: r = pow(a, b); break;
      case OP_ABS: r = fabs(a); break;
      case OP_TAN: r = tan(a); break;
      case OP_TANH: r = tanh(a); break;
      case OP_MAX: r = fmax(a, b); break;
      case OP_MIN: r = fmin(a, b); break;
      case OP_SINH: r = sinh(a); break;
      case OP_COSH: r = cosh(a); break;
      case OP_ATAN: r = atan(a); break;
      case OP_GT: r = (a > b) ? 1.0 : 0.0; break;
      case OP_SEL: r = (a != 0.0) ? b : 0.0; break;
      default: r = 0.0; break;
    }
    R[in.y] = r;
  }
  double* o = out + cl * (int64_t)O.nout * nqp + tid;
  for (int s = 0; s < O.nout; s++) o[(int64_t)s * nqp] = R[O.reg[s]];
}

extern "C" int tg_qp_eval(const tg_basis* h_B, int32_t nfun, const double* const* h_coefs,
                          const int32_t* h_ncomp, int32_t njets, const int32_t* h_jets,
                          const int32_t* prog, int32_t nprog, const double* consts,
                          int32_t nreg, int32_t nout, const int32_t* h_outregs, int64_t cell0,
                          int64_t ncells, double* out, void* stream) {
  TG_REQUIRE(njets <= TG_MAXQJ, "too many jets");
  TG_REQUIRE(nout <= TG_MAXOUT, "too many output slots");
  TgBasis B = tg_basis_dev(h_B);
  TgJetSpec J;
  J.njets = njets;
  for (int j = 0; j < njets; j++) {
    int f = h_jets[5 * j + 0];
    TG_REQUIRE(f >= 0 && f < nfun, "jet function index");
    J.coef[j] = h_coefs[f];
    J.ncomp[j] = (short)h_ncomp[f];
    J.comp[j] = (short)h_jets[5 * j + 1];
    for (int d = 0; d < 3; d++) {
      int a = h_jets[5 * j + 2 + d];
      TG_REQUIRE(a >= 0 && a <= h_B->nder, "jet derivative order exceeds tabulated order");
      J.al[j][d] = (signed char)a;
    }
  }
  TgOutSpec O;
  O.nout = nout;
  for (int s = 0; s < nout; s++) O.reg[s] = h_outregs[s];
  int nqp = B.nq[0] * B.nq[1] * B.nq[2];
  if (ncells == 0) return 0;
  TG_REQUIRE(nqp <= 1024, "too many Gauss points per cell");
  TG_REQUIRE(ncells < (int64_t)2147483647, "too many cells per launch");
  // jets of one function must be adjacent so its coefficient tile is loaded once
  int bs = ((nqp + 31) / 32) * 32;
  unsigned grid = (unsigned)ncells;
  const int nd = B.nder + 1;
  size_t smem = 0;
  for (int d = 0; d < 3; d++) smem += (size_t)B.nq[d] * B.nloc[d] * nd;
  smem += (size_t)B.nloc[0] * B.nloc[1] * B.nloc[2];
  smem += (size_t)B.nq[0] * B.nloc[1] * B.nloc[2];
  smem += (size_t)B.nq[0] * B.nq[1] * B.nloc[2];
  smem *= sizeof(double);
  TG_REQUIRE(smem <= 48 * 1024, "element too large for the qp-eval shared-memory tiles");
  cudaStream_t s = tg_stream(stream);
  const int4* p4 = (const int4*)prog;
  if (nreg <= 32)
    k_qp_eval<32><<<grid, bs, smem, s>>>(B, J, p4, nprog, consts, O, cell0, ncells, nqp, out);
  else if (nreg <= 64)
    k_qp_eval<64><<<grid, bs, smem, s>>>(B, J, p4, nprog, consts, O, cell0, ncells, nqp, out);
  else if (nreg <= 128)
    k_qp_eval<128><<<grid, bs, smem, s>>>(B, J, p4, nprog, consts, O, cell0, ncells, nqp, out);
  else if (nreg <= 512)
    k_qp_eval<512><<<grid, bs, smem, s>>>(B, J, p4, nprog, consts, O, cell0, ncells, nqp, out);
  else if (nreg <= 2048)
    k_qp_eval<2048><<<grid, bs, smem, s>>>(B, J, p4, nprog, consts, O, cell0, ncells, nqp, out);
  else {
    tg_set_error("qp program needs %d registers (max 2048)", nreg);
    return 2;
  }
  TG_LAUNCH_CHECK();
  return 0;
}
