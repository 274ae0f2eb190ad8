"""
ctypes binding of the C-ABI in include/tigar_b200.h (libtigar_b200.so).

The shared library is the product: there is no CPU fallback.  Import of this
module fails loudly if the library has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C tigar_b200/csrc``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TIGAR_B200_LIB") or os.path.join(_HERE, "libtigar_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "tigar_b200: %s not found -- the CUDA extension must be built "
        "(make -C tigar_b200/csrc); there is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

TG_MAXDIM = 3
c_i32, c_i64, c_dbl, c_vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p


class tg_basis(C.Structure):
    _fields_ = [("dim", c_i32), ("n", c_i32 * 3), ("nel", c_i32 * 3),
                ("nloc", c_i32 * 3), ("nq", c_i32 * 3), ("nder", c_i32),
                ("tab", c_vp * 3), ("idx", c_vp * 3), ("wq", c_vp * 3),
                ("xq", c_vp * 3)]


class tg_win(C.Structure):
    _fields_ = [("dim", c_i32), ("nr", c_i32 * 3), ("nc", c_i32 * 3),
                ("lo", c_vp * 3), ("hi", c_vp * 3), ("rowptr", c_vp), ("w0max", c_i32), ("S", c_vp * 3), ("row0", c_i32 * 3), ("col0", c_i32 * 3),
                ("layout", c_i32), ("H", c_i32), ("bs0", c_vp), ("maxrow", c_i32)]


PB = C.POINTER(tg_basis)
PW = C.POINTER(tg_win)
PI32 = C.POINTER(c_i32)
PVP = C.POINTER(c_vp)

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "tg_last_error": [],
    "tg_version": [],
    "tg_device_sm_count": [],
    "tg_sizeof_win": [],
    "tg_sizeof_basis": [],
    "tg_launch_count": [],
    "tg_last_spmv_kind": [],
    "tg_basis_funcs_inner": [c_vp, c_i32, c_i32, c_vp, c_vp, c_i64, c_vp, c_vp],
    "tg_bspline_eval_batch": [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32,
                              c_vp, c_i64, c_vp, c_vp, c_vp, c_vp],
    "tg_fe_nodes_1d": [c_vp, c_i32, c_i32, c_vp, c_vp],
    "tg_tabulate_1d": [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32,
                       c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "tg_win_rowlen": [PW, c_vp, c_vp],
    "tg_win_fill_cols": [PW, c_vp, c_vp],
    "tg_win_rowptr": [PW, PVP, c_vp, c_vp],
    "tg_win_spmv": [PW, c_vp, c_vp, c_vp, c_vp],
    "tg_win_spmv_dot": [PW, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp],
    "tg_win_zero_rows_cols": [PW, c_vp, c_vp, c_vp, c_dbl, c_i32, c_vp],
    "tg_win_zero_rows_cols_hp": [PW, c_vp, c_vp, c_vp, c_vp, c_dbl, PVP, PI32, c_vp],
    "tg_win_diag_inv": [PW, c_vp, c_i32, c_vp, c_vp],
    "tg_win_solve_cg": [PW, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_i32, c_i32, c_vp, PI32,
                        C.POINTER(c_dbl), c_vp],
    "tg_m_fill": [PW, PVP, PVP, PI32, c_vp, c_vp],
    "tg_spmv": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp],
    "tg_mt_vec": [PW, PW, c_vp, c_vp, c_vp, c_vp],
    "tg_qp_eval": [PB, c_i32, PVP, PI32, c_i32, PI32, c_vp, c_i32, c_vp, c_i32, c_i32,
                   PI32, c_i64, c_i64, c_vp, c_vp],
    "tg_jit_check": [C.c_char_p, C.POINTER(c_i64)],
    "tg_jit_compile": [C.c_char_p, C.c_char_p, C.POINTER(c_vp)],
    "tg_jit_launch": [c_vp, c_i64, c_i32, c_i32, c_vp, c_i32, c_vp],
    "tg_jit_free": [c_vp],
    "tg_assemble_matrix": [PB, PW, c_i32, PI32, c_i32, PI32, c_vp, c_i64, c_i64, c_vp, c_vp],
    "tg_assemble_matrix_ex": [PB, PW, c_i32, PI32, c_i32, PI32, PI32, c_vp, c_i64, c_i64,
                              c_vp, c_vp],
    "tg_assemble_sf_supported": [PB],
    "tg_assemble_matrix_terms": [PB, PW, c_i32, PI32, c_i32, PI32, c_vp, c_i64, c_i64, c_vp,
                                 c_vp],
    "tg_assemble_vector": [PB, c_i32, PI32, c_vp, c_i64, c_i64, c_vp, c_vp],
    "tg_assemble_vector_ex": [PB, c_i32, PI32, PI32, c_vp, c_i64, c_i64, c_vp, c_vp],
    "tg_assemble_vector_slots": [PB, c_i32, PI32, PI32, c_i32, PI32, c_vp, c_i64, c_i64, c_vp,
                                 c_vp],
    "tg_assemble_vector_part": [PB, c_i32, PI32, PI32, c_i32, PI32, PI32, PI32, c_vp, c_i64,
                                c_i64, c_vp, c_vp],
    "tg_sum": [c_vp, c_i64, c_vp, c_vp],
    "tg_ptap_ap": [PW, c_vp, PW, c_vp, PW, PW, c_vp, c_vp],
    "tg_ptap_c": [PW, c_vp, PW, PW, c_vp, PW, PW, c_vp, c_vp],
    "tg_ptap_kron_ap": [PW, c_vp, PVP, PW, c_vp, c_i32, c_vp],
    "tg_win_rowcombine": [PW, c_vp, PW, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp],
    "tg_ptap_march": [PW, c_vp, PW, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                      c_vp, c_i32, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "tg_ptap_march_w": [PW, c_vp, PW, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                        c_vp, c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "tg_win_export_vals": [PW, c_vp, c_vp, c_vp],
    "tg_win_import_vals": [PW, c_vp, c_vp, c_vp],
    "tg_win_storage": [PW, C.POINTER(c_i64)],
    "tg_zero_rows_cols": [c_vp, c_vp, c_vp, c_i64, c_vp, c_dbl, c_vp],
    "tg_zero_entries": [c_vp, c_vp, c_i64, c_vp],
    "tg_mask_set": [c_vp, c_vp, c_i64, c_i64, c_vp],
    "tg_diag_inv": [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp],
    "tg_cg_scratch_len": [],
    "tg_solve_cg": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_i32, c_i32,
                    c_vp, PI32, C.POINTER(c_dbl), c_vp],
    "tg_prof_enable": [c_i32],
    "tg_prof_get": [C.POINTER(c_dbl), C.POINTER(c_i64)],
    "tg_cg_spmv_dot": [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp],
    "tg_cg_init": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp],
    "tg_cg_axpy_dot": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp],
    "tg_cg_xpby": [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp],
    "tg_dot": [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp],
    "tg_axpy": [c_vp, c_dbl, c_vp, c_i64, c_vp],
    "tg_dgemm_batched": [c_i32, c_i32, c_i32, c_i32, c_i32, c_dbl, c_vp, c_i32, c_i64,
                         c_vp, c_i32, c_i64, c_dbl, c_vp, c_i32, c_i64, c_i32, c_vp],
    "tg_fp64_peak": [c_vp, C.POINTER(c_dbl), c_vp],
    "tg_fd_scale": [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, c_i64, c_dbl, c_i32, c_vp],
    "tg_masked_copy": [c_vp, c_vp, c_vp, c_i64, c_vp],
    "tg_masked_fix": [c_vp, c_vp, c_vp, c_dbl, c_i64, c_vp],
    "tg_fd_fit": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32,
                  c_vp, c_vp, c_vp],
    "tg_fd_fit_rel": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32,
                  c_vp, c_vp, c_vp],
    "tg_fd_diag_scale": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_dbl,
                         c_dbl, c_i32, c_i32, c_i32, c_vp, c_vp],
    "tg_vmul": [c_vp, c_vp, c_vp, c_i64, c_vp],
    "tg_xpby": [c_vp, c_dbl, c_vp, c_i64, c_vp],
    "tg_pcg_update": [c_vp, c_vp, c_vp, c_vp, c_dbl, c_i64, c_vp, c_vp, c_vp],
    "tg_band_from_win": [PW, c_vp, c_i32, c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp],
    "tg_band_cholesky": [c_i64, c_i32, c_i32, c_vp, c_vp, c_vp],
    "tg_band_solve": [c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "tg_win_asym": [PW, c_vp, c_vp, c_vp],
    "tg_tensor_column": [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_dbl, c_vp],
    "tg_gsf_supported": [c_i32, c_i32],
    "tg_gsf_stage": [c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp,
                     c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_i32, c_vp, c_i64,
                     c_i64, c_i64, c_i64, c_i32, PW, c_i64, c_i32, c_i32, c_vp, c_i32, PW, c_vp,
                     c_vp],
}
_RESTYPES = {"tg_sizeof_win": c_i64, "tg_sizeof_basis": c_i64, "tg_last_error": C.c_char_p, "tg_prof_enable": None, "tg_prof_get": None, "tg_launch_count": c_i64, "tg_win_storage": c_i64}

for _name, _args in SIGNATURES.items():
    _f = getattr(lib, _name)          # AttributeError if a symbol is missing
    _f.argtypes = _args
    _f.restype = _RESTYPES.get(_name, C.c_int)


if lib.tg_sizeof_win() != C.sizeof(tg_win) or lib.tg_sizeof_basis() != C.sizeof(tg_basis):
    raise ImportError("tigar_b200: descriptor structs of the ctypes binding (%d, %d bytes) do not "
                      "match libtigar_b200.so (%d, %d): rebuild the library"
                      % (C.sizeof(tg_win), C.sizeof(tg_basis), lib.tg_sizeof_win(),
                         lib.tg_sizeof_basis()))


class TigarError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise TigarError("tigar_b200: " + lib.tg_last_error().decode())


def i32arr(seq):
    seq = [int(x) for x in seq]
    return (c_i32 * max(len(seq), 1))(*seq)


def vparr(ptrs):
    return (c_vp * max(len(ptrs), 1))(*[int(p) for p in ptrs])
