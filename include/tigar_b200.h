/*
 * tigar_b200.h -- C-ABI of the B200-native tIGAr hot path
 * (extraction -> Gauss-point assembly -> M^T A M / M^T b -> BCs -> CG).
 *
 * The reference (david-kamensky/tIGAr) has no FFI of its own: its hot path is
 * Python calling DOLFIN/PETSc.  Each entry point below names the reference
 * interface it replaces (file:line into the reference tree).  The Python
 * classes in tigar_b200/ (same names as the reference's) are the only callers;
 * INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - plain C, no torch types.  All pointers are DEVICE pointers unless the
 *    parameter name starts with h_ (host).  Memory is caller-owned.
 *  - every call enqueues on the caller's CUDA stream (void* = cudaStream_t,
 *    NULL = default stream) and returns without synchronising unless stated.
 *  - return 0 on success, non-zero on error; tg_last_error() gives the text.
 *  - FP64 values, int32 column indices, int64 row pointers.
 *  - tensor-product index convention of the reference: first parametric
 *    direction fastest (BSplines.py:354-358).
 */
#ifndef TIGAR_B200_H
#define TIGAR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_MAXDIM 3
#define TG_MAXJET 16   /* max distinct derivative multi-indices per form side */

/* Tensor-product basis tabulated at the Gauss points of every element of the
 * one-cell-per-knot-span mesh (BSplines.py:505-569).  Used both for the
 * extracted spline basis N = M_e^T phi and for the Lagrange FE basis phi.   */
typedef struct {
  int32_t dim;
  int32_t n[TG_MAXDIM];     /* global basis functions per direction          */
  int32_t nel[TG_MAXDIM];   /* elements (non-degenerate spans) per direction */
  int32_t nloc[TG_MAXDIM];  /* local functions per element per direction     */
  int32_t nq[TG_MAXDIM];    /* Gauss points per direction                    */
  int32_t nder;             /* derivative orders 0..nder are tabulated       */
  const double*  tab[TG_MAXDIM]; /* [nel][nq][nloc][nder+1]                  */
  const int32_t* idx[TG_MAXDIM]; /* [nel][nloc] global 1-D index             */
  const double*  wq[TG_MAXDIM];  /* [nel][nq]  Gauss weight * span length    */
  const double*  xq[TG_MAXDIM];  /* [nel][nq]  parametric coordinate         */
} tg_basis;

/* CSR matrix whose row patterns are tensor-product windows: row (r1,r2,r3)
 * holds the columns  lo[d][r_d] <= c_d <= hi[d][r_d], stored c1 fastest.   */
typedef struct {
  int32_t dim;
  int32_t nr[TG_MAXDIM];
  int32_t nc[TG_MAXDIM];
  const int32_t* lo[TG_MAXDIM];
  const int32_t* hi[TG_MAXDIM];
  const int64_t* rowptr;     /* [nrows+1]                                    */
  int32_t w0max;             /* max over rows of the first-direction window length */
  const int64_t* S[TG_MAXDIM]; /* exclusive prefix sums of the window lengths,
                                  [nr_d+1]: rowptr in closed form             */
  int32_t row0[TG_MAXDIM];   /* row-distributed blocks: global coordinate of local row /    */
  int32_t col0[TG_MAXDIM];   /* column 0 per direction (0 for a whole matrix).  Assembly
                                skips rows outside [row0, row0+nr).                        */
  int32_t layout;            /* 0: rows contiguous (rowptr[row] + pos).
                                1: SELL-H: rows of one (r1,r2) line are grouped H at a time
                                and stored slot-major, entry (row, slot) at
                                  linebase(r1,r2) + (r0/H)*H*slots + slot*H + r0%H,
                                slot = ((c2-lo2)*len1 + (c1-lo1))*w0max + (c0 - bs0[r0]),
                                slots = w0max*len1*len2,
                                linebase = H*ceil(nr0/H)*w0max*(S1[r1]*len2 + T1*S2[r2]);
                                first-direction windows are padded to the uniform band
                                [bs0[r0], bs0[r0]+w0max) with zeros.  Used for the IGA
                                system matrix: lanes = rows, fully coalesced SpMV.       */
  int32_t H;
  const int32_t* bs0;        /* [nr0] band start (layout 1)                              */
  int32_t maxrow;            /* longest row (values); 0 = unknown: the TMA-staged
                                SpMV is then not used.  The value array must be
                                readable up to the next 16-byte boundary past
                                its end (any separate allocation is).         */
} tg_win;

const char* tg_last_error(void);
int tg_version(void);
/* sizeof(tg_win) / sizeof(tg_basis) as compiled into the library: lets a binding (ctypes, cgo,
 * ...) verify its own struct definition before it passes a pointer                         */
int64_t tg_sizeof_win(void);
int64_t tg_sizeof_basis(void);
/* kernels launched by this library so far (bench.py's gpu_launches claim) */
int64_t tg_launch_count(void);
/* number of SMs of the current device (grid sizing) */
int tg_device_sm_count(void);

/* ---- (i) extraction ---------------------------------------------------- */

/* The reference's one native routine, basisFuncsInner(ghostKnots,nGhost,u,pl,i,ndu,left,right,
 * ders) (BSplines.py:73-120, bound at :135-145): Piegl-Tiller A2.2 with the caller's index
 * i = span+1, batched over n points.  ghostKnots/u/i/ders are device pointers;
 * ders[n*(p+1)].  Same operation order, no FMA contraction: bit-exact.              */
int tg_basis_funcs_inner(const double* ghostKnots, int32_t nGhost, int32_t p,
                         const double* u, const int32_t* i, int64_t n, double* ders,
                         void* stream);

/* Batched span search + Cox-de Boor: BSpline1.getKnotSpan / getNodes /
 * basisFuncs -> basisFuncsInner (BSplines.py:285-351, 73-120).  Bit-exact
 * with the reference recurrence (no FMA contraction).
 * out: span[n], nodes[n*(p+1)] (= (span-p+i) mod ncp), vals[n*(p+1)].      */
int tg_bspline_eval_batch(const double* knots, int32_t nk,
                          const double* ghostKnots, int32_t nGhost,
                          int32_t p, int32_t ncp, int32_t mult0, int32_t multLast,
                          const double* u, int64_t n,
                          int32_t* span, int32_t* nodes, double* vals,
                          void* stream);

/* One column of the homogeneous control net of a tensor-product control mesh on the device:
 * out[i] = g[i_d] (g = per-direction values, e.g. Greville abscissae; NULL: out[i] = cval).
 * ExplicitBSplineControlMesh.getHomogeneousCoordinate (BSplines.py:935-960) for all control
 * points at once (the per-point Python loop of common.py:373-375).                          */
int tg_tensor_column(double* out, const double* g, int32_t n0, int32_t n1, int32_t n2,
                     int32_t d, double cval, void* stream);

/* FE node coordinates of the CG Q_pf mesh in one direction
 * (DOLFIN tabulate_dof_coordinates, common.py:1471-1472): node e*pf+a.     */
int tg_fe_nodes_1d(const double* uniqueKnots, int32_t nel, int32_t pf,
                   double* x, void* stream);

/* Per-element 1-D tables in one direction:
 *   Me[e][a][i]   = N_{first(e)+i}(x_{e,a})   1-D element extraction block
 *   tabN[e][q][i][k] = d^k N_i/dxi^k (xi_q) = sum_a Me[e][a][i] lag[q][a][k]/h^k
 *   tabL[e][q][a][k] = lag[q][a][k]/h^k       (Lagrange FE basis)
 *   idxN[e][i] = (espan[e]-p+i) mod ncp ; idxL[e][a] = e*pf+a
 *   wq[e][q] = h_wq[q]*h ; xq[e][q] = uk[e]+tq[q]*h
 * lag/tq/gw are reference-element constants computed by the host.
 * Replaces the per-node Python loop of common.py:1497-1509 for tensor-
 * product B-splines (together with tg_m_fill).                              */
int tg_tabulate_1d(const double* ghostKnots, int32_t nGhost, int32_t p, int32_t ncp,
                   const double* uniqueKnots, const int32_t* espan, int32_t nel,
                   int32_t pf, int32_t nq, int32_t nder,
                   const double* lag, const double* tq, const double* gw,
                   double* Me, double* tabN, int32_t* idxN,
                   double* tabL, int32_t* idxL, double* wq, double* xq,
                   void* stream);

/* windowed-CSR pattern helpers */
int tg_win_rowlen(const tg_win* h_w, int64_t* rowlen, void* stream);
/* rowptr[nrows+1] from the exclusive prefix sums S_d[nr_d+1] (device, int64)
 * of the per-direction window lengths (PETSc MatSetPreallocation equivalent,
 * common.py:1483-1492): closed form, no scan.                                */
int tg_win_rowptr(const tg_win* h_w, const int64_t* const* h_S, int64_t* rowptr,
                  void* stream);
int tg_win_fill_cols(const tg_win* h_w, int32_t* cols, void* stream);

/* Global extraction operator values, M = M_w (x) M_v (x) M_u on its window
 * (AbstractCoordinateChartSpline.generateM, common.py:1516-1578).
 * mfirst[d][I_d] = first 1-D column of node I_d (span-p), mvals[d][I_d*(p_d+1)+i]. */
int tg_m_fill(const tg_win* h_wM, const int32_t* const* h_mfirst,
              const double* const* h_mvals, const int32_t* h_p,
              double* vals, void* stream);

/* y = A x for a general CSR matrix (M*U of common.py:379,1259; C*p in CG)  */
int tg_spmv(const int64_t* rowptr, const int32_t* cols, const double* vals,
            const double* x, double* y, int64_t nrows, void* stream);

/* out = M^T b without forming M^T (multTranspose, common.py:97-109):
 * gather over the support box of each IGA function; h_wT holds the
 * transposed ranges (rows = IGA functions, cols = FE nodes).               */
int tg_mt_vec(const tg_win* h_wM, const tg_win* h_wT, const double* Mvals,
              const double* b, double* out, void* stream);

/* ---- (ii) Gauss-point assembly ----------------------------------------- */

/* Evaluate per-Gauss-point coefficient slots with a small register-machine
 * program (stands in for the FFC-generated tabulate_tensor of
 * common.py:1215-1216; geometry per calculusUtils.py:18-24,56-69,255-276).
 * Inputs of the program (registers 0..): xi_1..xi_dim, wq, then one register
 * per requested jet (function f, multi-index alpha).
 * h_coefs[f]: device pointer of function f's coefficients, ncomp[f] components
 * interleaved ([n_global][ncomp]); h_jets: njets x (f, comp, a1, a2, a3).
 * prog: nprog x (op,dst,a,b) int32 on device.  out[cell-cell0][slot][qp].   */
int tg_qp_eval(const tg_basis* h_B, int32_t nfun, const double* const* h_coefs,
               const int32_t* h_ncomp, int32_t njets, const int32_t* h_jets,
               const int32_t* prog, int32_t nprog, const double* consts,
               int32_t nreg, int32_t nout, const int32_t* h_outregs,
               int64_t cell0, int64_t ncells, double* out, void* stream);

/* Run-time compiled Gauss-point kernels (FFC/dijitso's role behind
 * dolfin.assemble, common.py:1215-1216): CUDA C source -> NVRTC -> sm_100a.
 * tg_jit_check only compiles (no GPU needed); tg_jit_launch passes ONE POD
 * parameter block by value.                                                  */
int tg_jit_check(const char* src, int64_t* cubin_bytes);
int tg_jit_compile(const char* src, const char* kernel_name, void** handle);
int tg_jit_launch(void* handle, int64_t grid, int32_t block, int32_t smem_bytes,
                  const void* param, int32_t param_bytes, void* stream);
int tg_jit_free(void* handle);

/* A[I,J] += sum_q sum_{s,t} coef[cell][s*nT+t][q] D^{aS_s}psi_I D^{aT_t}psi_J
 * element by element, colours processed one launch each, no atomics
 * (dolfin::Assembler behind common.py:1215-1216).  Row = test function.
 * With B = Lagrange basis this is A_FE; with B = extracted basis it is
 * sum_e M_e^T K_e M_e written straight into C.                               */
int tg_assemble_matrix(const tg_basis* h_B, const tg_win* h_W,
                       int32_t nS, const int32_t* h_alphaS,
                       int32_t nT, const int32_t* h_alphaT,
                       const double* coef, int64_t cell0, int64_t ncells,
                       double* vals, void* stream);

/* same with explicit colour strides per direction (2 suffices for the CG
 * Lagrange basis; nloc is always safe and is the default above).            */
int tg_assemble_matrix_ex(const tg_basis* h_B, const tg_win* h_W,
                          int32_t nS, const int32_t* h_alphaS,
                          int32_t nT, const int32_t* h_alphaT,
                          const int32_t* h_stride,
                          const double* coef, int64_t cell0, int64_t ncells,
                          double* vals, void* stream);

/* Sum-factorised variant of tg_assemble_matrix_ex for 3-D bases with the same
 * nloc = nq in {3,4,5} per direction (tg_assemble_sf_supported).  Only the
 * listed terms are processed: h_terms = nterms x (slot, aS[3], aT[3]) where
 * slot indexes coef[cell][slot][qp] (nslots slots per cell).                 */
int tg_assemble_sf_supported(const tg_basis* h_B);
int tg_assemble_matrix_terms(const tg_basis* h_B, const tg_win* h_W, int32_t nterms,
                             const int32_t* h_terms, int32_t nslots,
                             const int32_t* h_stride, const double* coef,
                             int64_t cell0, int64_t ncells, double* vals, void* stream);

/* b[I] += sum_q sum_s coef[cell][s][q] D^{aS_s}psi_I  (common.py:1169)       */
int tg_assemble_vector(const tg_basis* h_B, int32_t nS, const int32_t* h_alphaS,
                       const double* coef, int64_t cell0, int64_t ncells,
                       double* b, void* stream);

int tg_assemble_vector_ex(const tg_basis* h_B, int32_t nS, const int32_t* h_alphaS,
                          const int32_t* h_stride, const double* coef,
                          int64_t cell0, int64_t ncells, double* b, void* stream);
/* same, reading term s from coef[cell][h_slots[s]][qp] with nslots slots per
 * cell (lets the matrix and the vector of one linear system share a single
 * Gauss-point pass: assembleLinearSystem, common.py:1223-1234).             */
int tg_assemble_vector_slots(const tg_basis* h_B, int32_t nS, const int32_t* h_alphaS,
                             const int32_t* h_slots, int32_t nslots,
                             const int32_t* h_stride, const double* coef,
                             int64_t cell0, int64_t ncells, double* b, void* stream);

/* vector assembly into a slab of the global vector: only entries whose
 * coordinate in direction d lies in [h_row0[d], h_row0[d]+h_nr[d]) are written,
 * at local index (coordinate - row0), local sizes h_nr.                      */
int tg_assemble_vector_part(const tg_basis* h_B, int32_t nS, const int32_t* h_alphaS,
                            const int32_t* h_slots, int32_t nslots,
                            const int32_t* h_stride, const int32_t* h_row0,
                            const int32_t* h_nr, const double* coef,
                            int64_t cell0, int64_t ncells, double* b, void* stream);

/* sum over all entries (functional assembly, poisson.py:132); result on device */
int tg_sum(const double* x, int64_t n, double* out1, void* stream);

/* ---- (iii) triple product, BCs, solve ---------------------------------- */

/* AP = A M on its window (first half of MatPtAP, common.py:1194-1195).
 * h_wMT: transposed ranges of M.                                            */
int tg_ptap_ap(const tg_win* h_wA, const double* Avals,
               const tg_win* h_wM, const double* Mvals, const tg_win* h_wMT,
               const tg_win* h_wP, double* APvals, void* stream);
/* C = M^T (AP).  h_wPT: transposed ranges of AP.                             */
int tg_ptap_c(const tg_win* h_wM, const double* Mvals, const tg_win* h_wMT,
              const tg_win* h_wP, const double* APvals, const tg_win* h_wPT,
              const tg_win* h_wC, double* Cvals, void* stream);

/* Kronecker-structured M^T A M for tensor-product bases (M = M_2 (x) M_1 (x) M_0;
 * the global M is never read).  h_tabs[d]: device array [n_fe_d][10][10] with
 * tab[I][J-loA_d(I)][j-loP_d(I)] = M_d[J,j]; box >= prod_d max(lenA_d, lenP_d)
 * (shared-memory tile per row).  AP = A*M, one pass over A.                  */
int tg_ptap_kron_ap(const tg_win* h_wA, const double* Avals, const double* const* h_tabs,
                    const tg_win* h_wP, double* APvals, int32_t box, void* stream);
/* Y[(..i_d..),:] = sum_I M_d[I,i_d] X[(..I_d..),:] : turns direction d of the row
 * grid from FE nodes into IGA functions.  mfirst/mvals: the 1-D extraction
 * rows of direction d ([n_fe_d] and [n_fe_d][np1]); supp_lo/hi[i]: FE support
 * of function i.  Apply for d = 0,1,2 to get C = M^T (AP).                   */
int tg_win_rowcombine(const tg_win* h_wX, const double* Xvals, const tg_win* h_wY,
                      double* Yvals, int32_t d, const int32_t* mfirst, const double* mvals,
                      int32_t np1, const int32_t* supp_lo, const int32_t* supp_hi,
                      void* stream);

/* One two-sided pass of the Kronecker-structured M^T A M (MatPtAP,
 * common.py:1194-1195): direction d of both the row and the column grid goes
 * from FE nodes to IGA functions,
 *   Y[(..i..),(..j..)] = sum_{I,J} M_d[I,i] X[(..I..),(..J..)] M_d[J,j];
 * applied for d = 0,1,2 it turns A_FE into C = M^T A M reading A once (the
 * global M is never formed).  h_wX / h_wY differ only in direction d (FE-FE
 * window -> IGA-IGA window, which must lie inside [i-p, i+p]).  Device arrays:
 *   first[n_fe_d], mrow[n_fe_d][p+1]   1-D extraction rows (eps-filtered),
 *   tabc[n_fe_d][KA][TWP]              M_d[loX_d(I)+q, first(I)+m], TWP = (p+3)&~1,
 *   slo/shi[n_cp_d]                    FE support of function i,
 *   ga[nga+1], gb[ngb+1]               groups of consecutive row coordinates of the
 *                                      two other directions handled by one CTA
 *                                      (sum of window lengths a x b <= 256),
 *   seg[nseg+1]                        output-row boundaries of the march segments.
 * stage_doubles (even) / out_doubles / maxlines: shared-memory sizing, maxima
 * over the CTAs of  sum_l ((KAmax*L_l + 3) & ~1),  sum_l (2p+1)*L_l  and the
 * number of lines, L_l = product of the line's other-direction window lengths.
 * variant 0: rows staged by 8-byte cp.async (LDGSTS) from all warps; variant 1:
 * by 1-D bulk async copies (TMA) -- then the value array of X must be readable
 * up to the next 16-byte boundary past its end.                               */
int tg_ptap_march(const tg_win* h_wX, const double* Xvals, const tg_win* h_wY, double* Yvals,
                  int32_t d, int32_t p, int32_t KA, int32_t KAmax, const int32_t* first,
                  const double* mrow, const double* tabc, const int32_t* slo,
                  const int32_t* shi, const int32_t* ga, int32_t nga, const int32_t* gb,
                  int32_t ngb, const int32_t* seg, int32_t nseg, int32_t stage_doubles,
                  int32_t out_doubles, int32_t maxlines, int32_t variant, void* stream);

/* Same pass, warp-independent variant (the default): every warp owns a task --
 * up to 8 pieces, each a contiguous range of fibres of one line -- and runs its
 * own 3-stage ring in a private shared-memory slice (1-D bulk async copies = TMA
 * with per-warp mbarriers for d = 0,1; 8-byte cp.async for d = 2, whose sub-rows
 * are strided; no CTA barriers inside the march); the march advances by groups
 * of <= 4 consecutive FE rows sharing
 * first(I); the per-node tables of a CTA's march segment live in shared memory.
 *   irec[n_fe_d] int32x4 {len_d(I) | lo_d(I) << 8 of X, first(I), sbits, group(I)};
 *        sbits: 2 bits per column q of the row's window, first(lo+q) - first(I) + 1
 *   Sx[n_fe_d]   int64   S_d[I] of X (exclusive prefix sum of the window lengths)
 *   jrec[n_cp_d] int32x4 {lo_d(i) of Y - (i-p), len_d(i) of Y, S_d[i] lo32, hi32}
 *   cpad[n_fe_d][p+4]    {0, M_d[I, first(I)+k] (k = 0..p, eps-filtered), 0, 0}
 *   grp[ngroups+1]       first FE row of every group
 *   tasks[ntask][36] int32 {npieces,0,0,0, {ra, rb, cb0, ncb} x npieces}: piece =
 *        fibres (all ca, cb0 <= cb < cb0+ncb) of line (ra, rb); <= 32 fibres/task
 *   seg[nseg+1]  output-row boundaries of the march segments (grid.y)
 * GMAX: most window entries (sum of len_d) of one group, >= 2p+1; maxnodes /
 * maxrows / maxgroups: most FE nodes, output rows and groups of one segment
 * (a segment marches over the whole groups covering its rows' FE support);
 * maxpieces: most pieces of one task; wpc: warps (tasks) per CTA -- 8 (two CTAs
 * per SM), 16 (one CTA per SM, the segment tables are shared by twice as many
 * warps: longer segments) or 4 (required for p = 4).  For d = 0,1 the value
 * array of X must be readable up to the next 16-byte boundary past its end.   */
int tg_ptap_march_w(const tg_win* h_wX, const double* Xvals, const tg_win* h_wY, double* Yvals,
                    int32_t d, int32_t p, int32_t GMAX, const void* irec, const void* Sx,
                    const void* jrec, const double* cpad, const int32_t* grp,
                    const int32_t* slo, const int32_t* shi, const int32_t* tasks,
                    int32_t ntask, const int32_t* seg, int32_t nseg, int32_t maxnodes,
                    int32_t maxrows, int32_t maxgroups, int32_t maxpieces, int32_t wpc,
                    void* stream);

/* ---- windowed-CSR operators (no column array: 8 B per non-zero) ---------- */
/* y = C x  (MatMult inside KSP, common.py:1255-1258; M*U, common.py:379,1259) */
int tg_win_spmv(const tg_win* h_w, const double* vals, const double* x, double* y,
                void* stream);
/* which kernel the last windowed SpMV used: 0 direct-load (k_win_spmv), 1 TMA-staged
 * rows + shared-memory x tiles (k_win_spmv_tma), 2 SELL layout (k_sell_spmv),
 * 3 direct-load with per-warp cp.async row prefetch (k_win_spmv_pf).  0 is the
 * default; 1 and 3 are opt-in (TIGAR_B200_TMA_SPMV=1, TIGAR_B200_SPMV_PF=1)       */
int tg_last_spmv_kind(void);
/* y = C x and out1[0] = sum_r x[xoff+r] y[r]; scratch: tg_cg_scratch_len()   */
int tg_win_spmv_dot(const tg_win* h_w, const double* vals, const double* x, int64_t xoff,
                    double* y, double* scratch, double* out1, void* stream);
/* zeroRowsColumns on the window pattern; rowmask over local rows, colmask over
 * columns; col_shift = offset of the row's own column in the last direction
 * (0 unless the block is a slab of a row-distributed matrix).                */
int tg_win_zero_rows_cols(const tg_win* h_w, double* vals, const uint8_t* rowmask,
                          const uint8_t* colmask, double diag, int32_t col_shift,
                          void* stream);
/* same, for a constrained set that is a union of whole hyperplanes (the side DoFs of
 * getSideDofs, BSplines.py:599-649): hp_d[c] != 0 marks hyperplane c of direction d (global
 * coordinates; h_w's row0/col0 place a slab-local block).  Rows whose window reaches no
 * constrained hyperplane are skipped after a few byte loads.  h_sel[d] (device) / h_nsel[d]:
 * the LOCAL row coordinates of direction d that can be affected (window reaches a constrained
 * hyperplane): only those sub-grids are visited, one launch per direction (NULL: all rows).
 * Row-major layout only.                                                                     */
int tg_win_zero_rows_cols_hp(const tg_win* h_w, double* vals, const uint8_t* hp0,
                             const uint8_t* hp1, const uint8_t* hp2, double diag,
                             const int32_t* const* h_sel, const int32_t* h_nsel, void* stream);
int tg_win_diag_inv(const tg_win* h_w, const double* vals, int32_t col_shift, double* dinv,
                    void* stream);
/* Jacobi-CG on a windowed matrix; same contract as tg_solve_cg.             */
int tg_win_solve_cg(const tg_win* h_w, const double* vals, const double* b, double* x,
                    double rtol, double atol, int32_t maxit, int32_t check_every,
                    double* work, int32_t* h_iters, double* h_relres, void* stream);

/* layout conversion of a windowed matrix's values: exact row-major CSR order
 * (rowptr) <-> the window's own layout.  export: out_csr[rowptr[r]+pos] =
 * vals(r,pos); import: the inverse, padding zero-filled.                     */
int tg_win_export_vals(const tg_win* h_w, const double* vals, double* out_csr, void* stream);
int tg_win_import_vals(const tg_win* h_w, const double* in_csr, double* vals, void* stream);
/* number of doubles the value array of a window occupies in its layout       */
int64_t tg_win_storage(const tg_win* h_w, const int64_t* h_T);

/* zeroRowsColumns(zeroDofs, diag) (common.py:1199-1200); mask[i]!=0 marks a
 * constrained DoF.                                                          */
int tg_zero_rows_cols(const int64_t* rowptr, const int32_t* cols, double* vals,
                      int64_t nrows, const uint8_t* mask, double diag, void* stream);
/* b[zeroDofs] = 0 (common.py:1154-1158) */
int tg_zero_entries(double* b, const uint8_t* mask, int64_t n, void* stream);
/* mask[idx[k]] = 1 for the zeroDofs list (device int64 indices, duplicates allowed;
 * mask must be zero-initialised): the row/column masks of the calls above        */
int tg_mask_set(uint8_t* mask, const int64_t* idx, int64_t nidx, int64_t n, void* stream);

/* dinv[i] = 1/C[i,i] */
int tg_diag_inv(const int64_t* rowptr, const int32_t* cols, const double* vals,
                int64_t nrows, int64_t row0, double* dinv, void* stream);

/* Jacobi-preconditioned CG, all iteration state on the device, one host
 * check every `check_every` iterations (solve() of common.py:1255-1258).
 * work: 4*n + tg_cg_scratch_len() + 8 doubles.  Synchronises before
 * returning.  Reductions are two-stage on a fixed grid: deterministic.     */
int tg_cg_scratch_len(void);
int tg_solve_cg(const int64_t* rowptr, const int32_t* cols, const double* vals,
                const double* b, double* x, int64_t n,
                double rtol, double atol, int32_t maxit, int32_t check_every,
                double* work, int32_t* h_iters, double* h_relres, void* stream);

/* Live timing of the SpMV launches inside the CG drivers (CUDA events on the
 * solver's stream, collected at the host checks): the roofline numerator of
 * bench.py.  enable(1) resets the counters.                                  */
void tg_prof_enable(int on);
void tg_prof_get(double* spmv_ms, int64_t* spmv_launches);

/* y += a x.  Block-row accumulation y_i = sum_j C_ij x_j of an equal-order
 * multi-field system (field blocks of common.py:337-351, 1546-1573): one
 * windowed SpMV per block, summed with this.                                */
int tg_axpy(double* y, double a, const double* x, int64_t n, void* stream);

/* CG building blocks for the row-distributed multi-GPU solver (the host
 * interleaves torch.distributed halo exchange / all-reduce between them).
 * scratch: tg_cg_scratch_len() doubles.
 *   spmv_dot : y = A x (local rows; cols index the extended vector x),
 *              out1[0] = sum_r x[xoff+r]*y[r]
 *   init     : r = b - y ; p = dinv r ; out2 = {r.dinv.r, r.r}
 *   axpy_dot : a = num[0]/den[0]; x += a p; r -= a q; out2 = {r.dinv.r, r.r}
 *   xpby     : p = dinv*r + (num[0]/den[0]) p
 *   dot      : out1[0] = a.b                                               */
int tg_cg_spmv_dot(const int64_t* rowptr, const int32_t* cols, const double* vals,
                   const double* x, int64_t xoff, double* y, int64_t nrows,
                   double* scratch, double* out1, void* stream);
int tg_cg_init(const double* b, const double* y, const double* dinv, double* r,
               double* p, int64_t n, double* scratch, double* out2, void* stream);
int tg_cg_axpy_dot(double* x, double* r, const double* p, const double* q,
                   const double* dinv, int64_t n, const double* num, const double* den,
                   double* scratch, double* out2, void* stream);
int tg_cg_xpby(double* p, const double* r, const double* dinv, int64_t n,
               const double* num, const double* den, void* stream);
int tg_dot(const double* a, const double* b, int64_t n, double* scratch,
           double* out1, void* stream);

/* ---- dense FP64 building blocks of the solve stage (tg_dense.cu) -------------------------- */
/* C = alpha op(A) op(B) + beta C, column-major, strided batch (blockIdx.z); op = transpose when
 * the flag is non-zero.  Hand-written DFMA kernel (64x64x16 tiles, 4x4 per thread).          */
int tg_dgemm_batched(int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K,
                     double alpha, const double* A, int32_t lda, int64_t strideA,
                     const double* B, int32_t ldb, int64_t strideB, double beta,
                     double* C, int32_t ldc, int64_t strideC, int32_t batch, void* stream);
/* measured DFMA peak of the current device in TFLOP/s (best of 5 after warm-up; synchronises):
 * the roofline denominator of the FP64-bound kernels.  scratch1: one device double.          */
int tg_fp64_peak(double* scratch1, double* h_tflops, void* stream);

/* Fast-diagonalisation preconditioner for the KSP solve (solve(), common.py:1255-1258): the
 * tensor-product operator  sigma M(x)M(x)M + sum_d c_d K_d (x) M (x) M  is inverted exactly in the
 * generalised eigenbasis K_d U_d = M_d U_d Lambda_d (1-D, host setup): three mode products with
 * U_d^T (tg_dgemm_batched), this scaling, three mode products with U_d.
 *   t[pl, i2] /= (sigma + l0[i0] + l1[i1] + l2[i2])^pw   (0 where the sum is not finite and
 *   positive: constrained hyperplanes carry +inf) for the mq plane entries pl = q0 .. q0+mq-1
 *   (i0 = pl % n0, i1 = pl / n0; whole tensor: q0 = 0, mq = n0*n1; a chunk of planes on each
 *   rank of a row-distributed solve).  l1 / l2 may be NULL (1-D / 2-D).                       */
int tg_fd_scale(double* t, const double* l0, const double* l1, const double* l2,
                int32_t n0, int32_t n1, int32_t n2, int64_t q0, int64_t mq, double sigma,
                int32_t pw, void* stream);
/* dst = mask ? 0 : src ;  z = mask ? r*cinv : z  (constrained rows are diag*identity,
 * zeroRowsColumns(zeroDofs, diag), common.py:1199-1200)                                       */
int tg_masked_copy(double* dst, const double* src, const uint8_t* mask, int64_t n, void* stream);
int tg_masked_fix(double* z, const double* r, const uint8_t* mask, double cinv, int64_t n,
                  void* stream);
/* out4[a] = sum over unconstrained DoFs of diagC[i] * b_a[i], b_a the Kronecker products of the
 * 1-D diagonals (a < 3: k_a (x) m (x) m ; a = 3: m (x) m (x) m): right-hand side of the
 * least-squares fit of c_d and sigma to diag(C).  scratch: 256 doubles.                       */
int tg_fd_fit(const double* diagC, const uint8_t* mask, const double* kd0, const double* kd1,
              const double* kd2, const double* md0, const double* md1, const double* md2,
              int32_t n0, int32_t n1, int32_t n2, double* scratch, double* out4, void* stream);

/* Relative-error variant: minimises sum_i (sum_a x_a col_a(i)/d_i - 1)^2 (robust on curved /
 * rational geometry, where the absolute fit is dominated by the largest diagonal entries).
 * scratch: 15*64 doubles; out15 (device): the 10 upper-triangular Gram sums (G00 G01 G02 G03
 * G11 G12 G13 G22 G23 G33; index 3 = mass column), the 4 right-hand sides, the DoF count.    */
int tg_fd_fit_rel(const double* diagC, const uint8_t* mask, const double* kd0, const double* kd1,
                  const double* kd2, const double* md0, const double* md1, const double* md2,
                  int32_t n0, int32_t n1, int32_t n2, double* scratch, double* out15,
                  void* stream);

/* s_i = sqrt(B_ii / C_ii), B_ii the diagonal of the FD surrogate with weights (c_d, sigma);
 * 1 on constrained DoFs: the diagonal scaling of z = S B^-1 S r.                            */
int tg_fd_diag_scale(const double* diagC, const uint8_t* mask, const double* kd0,
                     const double* kd1, const double* kd2, const double* md0, const double* md1,
                     const double* md2, double c0, double c1, double c2, double sigma,
                     int32_t n0, int32_t n1, int32_t n2, double* out, void* stream);
/* y = x * s element-wise (y may alias x) */
int tg_vmul(double* y, const double* x, const double* s, int64_t n, void* stream);
/* preconditioned CG, vector kernels (driver: tigar_b200/solvers.py):
 *   xpby       : p = z + beta p
 *   pcg_update : x += a p ; r -= a q ; out1[0] = r.r   (scratch: tg_cg_scratch_len())         */
int tg_xpby(double* p, double beta, const double* z, int64_t n, void* stream);
int tg_pcg_update(double* x, double* r, const double* p, const double* q, double a, int64_t n,
                  double* scratch, double* out1, void* stream);

/* ---- direct solve: band Cholesky on the device (tg_band.cu) -------------------------------- */
/* The reference's default solve() is a sparse direct LU (common.py:1255-1256).  In the
 * reference's own DoF numbering the IGA matrix of a 2-D (or small 3-D) patch is banded;
 * LAPACK lower band storage AB[(i-j) + j*ldab], ldab >= bw + 32, zero-initialised.
 * Equal-order multi-field systems (field blocks of common.py:337-351): block (fr, fc) of the
 * nf x nf grid is written in the node-major interleaved numbering row' = nf*row + fr (band
 * nf*bw + nf - 1); nf = 1, fr = fc = 0 for a scalar system.
 * info: device int32, zero-initialised; from_win sets -1 if a non-zero lies outside the band,
 * cholesky sets k+1 if the pivot block at column k is not positive definite.                  */
int tg_band_from_win(const tg_win* h_w, const double* vals, int32_t bw, int32_t ldab,
                     double* AB, int32_t* info, int32_t nf, int32_t fr, int32_t fc, void* stream);
int tg_band_cholesky(int64_t n, int32_t bw, int32_t ldab, double* AB, int32_t* info,
                     void* stream);
/* L L^T x = b ; b is overwritten (work), work: n doubles, x must not alias b.                 */
int tg_band_solve(int64_t n, int32_t bw, int32_t ldab, const double* AB, double* b,
                  double* work, double* x, void* stream);
/* out2 = {max |a_ij - a_ji|, max |a_ij|} over a whole square windowed matrix (zero-initialise
 * out2): symmetry test in front of Cholesky / CG.                                             */
int tg_win_asym(const tg_win* h_w, const double* vals, double* out2, void* stream);

/* ---- global sum-factorised assembly of the extracted system (tg_gsf.cu) --------------------- */
/* dolfin.assemble + MatPtAP (common.py:1215-1216, 1194-1195) for a tensor-product spline, A_FE
 * and M never formed: the quadrature sum is contracted one direction at a time over the whole
 * patch; every stage is one "march" launch per output kind.  Arrays are blocked as
 * [kind][cell of the marched direction][inner][nq]; see the header of tg_gsf.cu.
 *   X, skin, scell, cbase  input, X[k*skin + (e - cbase)*scell + inner*nq + q]
 *   c0, c1, nel            cells marched [c0, c1) of nel (chunks of the last direction)
 *   nloc, nq, nd, tab, idx 1-D tables of the direction ([nel][nq][nloc][nd], [nel][nloc])
 *   rowbase[n_d]           S_d[i] - lo_d[i] of the system window: f(i,j) = rowbase[i] + j
 *   plan (device int32)    [nout][1 + 3*maxin]: nin, then (input kind, test order, trial order)
 *   pair                   1: matrix (pairs of functions), 0: load vector (f = i)
 *   ninner, nv, nw         inner = (u*nv + v)*nw + w
 *   Y, skout, so_*         output of a non-last stage, Y[k*skout + f*so_f + u*so_u + v*so_v + w]
 *   last                   1: writes `out` = values of the windowed matrix h_W (local rows; rows
 *                          outside [row0, row0+nr) of the last direction are skipped) or, for
 *                          pair = 0, the vector slab [vec_row0, vec_row0+vec_nr) of the last
 *                          direction; ninner = F0*F1 (pair) or the plane size; h_F0 = total 1-D
 *                          window length of the first direction.
 *   perm_rows              optional (device, [nr1][2], 16-byte aligned): per row i1 of h_Wperm
 *                          {S1[i1]*F0, len1[i1] | lo1[i1] << 32} -- the writer of the permuted
 *                          layout then needs one load per row; NULL: computed from h_Wperm.
 *   perm, h_Wperm          3-D matrices: the stage before the last (perm = 1, last = 0) writes the
 *                          pair index (f1, f0) in the thread order of the last stage
 *                          (u = S1[i1]*F0 + S0[i0]*len1 + dj1*len0 + dj0, windows of h_Wperm,
 *                          h_F0 = F0) and the last stage (perm = 1) reads it so: every (cell,
 *                          input) item of a CTA is then one contiguous run, staged by 1-D bulk
 *                          async copies (TMA) through a shared-memory ring (even nq).
 * tg_gsf_supported(nloc, nq): instantiated for 2 <= nloc <= 5, nq in {nloc, nloc+1}.           */
int tg_gsf_supported(int32_t nloc, int32_t nq);
int tg_gsf_stage(const double* X, int64_t skin, int64_t scell, int32_t cbase,
                 int32_t c0, int32_t c1, int32_t nel, int32_t nloc, int32_t nq,
                 int32_t nd, const double* tab, const int32_t* idx,
                 const int64_t* rowbase, const int32_t* plan, int32_t nout,
                 int32_t maxin, int32_t pair, int64_t ninner, int32_t nv, int32_t nw,
                 double* Y, int64_t skout, int64_t so_f, int64_t so_u, int64_t so_v,
                 int32_t last, const tg_win* h_W, int64_t h_F0, int32_t vec_row0,
                 int32_t vec_nr, double* out, int32_t perm, const tg_win* h_Wperm,
                 const int64_t* perm_rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif
