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

template <int NREG>
__global__ void k_qp_eval(TgBasis B, TgJetSpec J, const int4* __restrict__ prog, int nprog,
                          const double* __restrict__ consts, TgOutSpec O, int64_t cell0,
                          int64_t ncells, int nqp, double* __restrict__ out) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= ncells * nqp) return;
  int64_t cl = t / nqp;
  int qp = (int)(t - cl * nqp);
  int e[3], q[3];
  tg_decode(cell0 + cl, B.nel, B.dim, e);
  tg_decode(qp, B.nq, B.dim, q);

  double R[NREG];
  double wq = 1.0;
  const int nd = B.nder + 1;
  const double* tb[3];
  const int32_t* ix[3];
  for (int d = 0; d < 3; d++) {
    if (d < B.dim) {
      R[d] = B.xq[d][e[d] * B.nq[d] + q[d]];
      wq *= B.wq[d][e[d] * B.nq[d] + q[d]];
      tb[d] = B.tab[d] + ((int64_t)e[d] * B.nq[d] + q[d]) * B.nloc[d] * nd;
      ix[d] = B.idx[d] + e[d] * B.nloc[d];
    } else {
      tb[d] = nullptr;
      ix[d] = nullptr;
    }
  }
  R[B.dim] = wq;
  const int r0 = B.dim + 1;
  for (int j = 0; j < J.njets; j++) R[r0 + j] = 0.0;

  // jets: sum over local basis functions
  const int n0 = B.nloc[0], n1 = B.nloc[1], n2 = B.nloc[2];
  for (int a2 = 0; a2 < n2; a2++) {
    int64_t g2 = (B.dim > 2) ? ix[2][a2] : 0;
    for (int a1 = 0; a1 < n1; a1++) {
      int64_t g1 = (B.dim > 1) ? ix[1][a1] : 0;
      for (int a0 = 0; a0 < n0; a0++) {
        int64_t g = ix[0][a0] + (int64_t)B.n[0] * (g1 + (int64_t)B.n[1] * g2);
        for (int j = 0; j < J.njets; j++) {
          double wgt = tb[0][a0 * nd + J.al[j][0]];
          if (B.dim > 1) wgt *= tb[1][a1 * nd + J.al[j][1]];
          if (B.dim > 2) wgt *= tb[2][a2 * nd + J.al[j][2]];
          R[r0 + j] += wgt * J.coef[j][g * J.ncomp[j] + J.comp[j]];
        }
      }
    }
  }

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
      default: r = 0.0; break;
    }
    R[in.y] = r;
  }
  double* o = out + cl * (int64_t)O.nout * nqp + qp;
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
  int64_t nt = ncells * nqp;
  if (nt == 0) return 0;
  int bs = 128;
  unsigned grid = (unsigned)tg_cdiv(nt, bs);
  cudaStream_t s = tg_stream(stream);
  const int4* p4 = (const int4*)prog;
  if (nreg <= 32)
    k_qp_eval<32><<<grid, bs, 0, s>>>(B, J, p4, nprog, consts, O, cell0, ncells, nqp, out);
  else if (nreg <= 128)
    k_qp_eval<128><<<grid, bs, 0, s>>>(B, J, p4, nprog, consts, O, cell0, ncells, nqp, out);
  else if (nreg <= 512)
    k_qp_eval<512><<<grid, bs, 0, s>>>(B, J, p4, nprog, consts, O, cell0, ncells, nqp, out);
  else if (nreg <= 2048)
    k_qp_eval<2048><<<grid, bs, 0, s>>>(B, J, p4, nprog, consts, O, cell0, ncells, nqp, out);
  else {
    tg_set_error("qp program needs %d registers (max 2048)", nreg);
    return 2;
  }
  TG_LAUNCH_CHECK();
  return 0;
}
