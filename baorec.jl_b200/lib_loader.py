"""ctypes binding of libbaorec_b200.so (the C ABI declared in include/baorec_b200.h).

The library is mandatory: there is no CPU or PyTorch fallback.  If it has not
been built (python baorec.jl_b200/build.py) importing this module raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "lib" / "libbaorec_b200.so"

OK = 0
ERR_INVALID, ERR_CUDA, ERR_CUFFT, ERR_NCCL, ERR_OUT_OF_BOX, ERR_NOT_PLANNED, ERR_NOMEM, ERR_OUT_OF_RANGE, ERR_IO = -1, -2, -3, -4, -5, -6, -7, -8, -9
DTYPE_F32, DTYPE_F64 = 4, 8
MAS_CIC, MAS_TSC, MAS_PCS = 0, 1, 2
FIELD_DISP, FIELD_RSD, FIELD_SUM = 0, 1, 2
ITERATIVE, MULTIGRID = 0, 1
FIELDS = {"disp": FIELD_DISP, "rsd": FIELD_RSD, "sum": FIELD_SUM}
MAS = {"cic": MAS_CIC, "tsc": MAS_TSC, "pcs": MAS_PCS}


class BaorecError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libbaorec_b200 error {code}: {msg}")
        self.code = code


class OutOfBoxError(BaorecError):
    pass


class OutOfRangeError(BaorecError):
    """A redshift / distance outside the cosmology tables (reference: Interpolations.jl BoundsError)."""


class CatalogIOError(BaorecError, OSError):
    """A catalog file that cannot be opened, read, parsed or written (BAOREC_ERR_IO)."""


class Params(C.Structure):
    """struct baorec_params."""
    _fields_ = [
        ("bias", C.c_float), ("f", C.c_float), ("smoothing_radius", C.c_float), ("beta", C.c_float),
        ("n_iter", C.c_int32), ("has_los", C.c_int32), ("los", C.c_float * 3),
        ("jacobi_damping_factor", C.c_float), ("jacobi_niterations", C.c_int32),
        ("vcycle_niterations", C.c_int32), ("mas", C.c_int32), ("ran_min", C.c_float), ("box_pad", C.c_float),
    ]


class CosmologyParams(C.Structure):
    """struct baorec_cosmology."""
    _fields_ = [("h", C.c_double), ("Omega_b0", C.c_double), ("Omega_c0", C.c_double), ("Omega_nu0", C.c_double),
                ("Omega_g0", C.c_double), ("Omega_k0", C.c_double), ("Omega_L0", C.c_double), ("w0", C.c_double),
                ("wa", C.c_double), ("z_tab_min", C.c_double), ("z_tab_max", C.c_double), ("z_tab_num", C.c_int64)]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_f3 = C.POINTER(C.c_float)
_pp = C.POINTER(Params)

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/baorec_b200.h
SIGNATURES = {
    "baorec_version": [],
    "baorec_last_error": [],
    "baorec_create": [_i, C.POINTER(_vp)],
    "baorec_destroy": [_vp],
    "baorec_plan": [_vp, _i, _i, _i, _f3, _f3],
    "baorec_set_box": [_vp, _f3, _f3],
    "baorec_set_option": [_vp, C.c_char_p, _i64],
    "baorec_scratch_bytes": [_vp],
    "baorec_sort_reuse_count": [_vp],
    "baorec_launch_counts": [_vp, C.POINTER(_i64), C.POINTER(_i64)],
    "baorec_last_stage_ms": [_vp, _f3, _i],
    "baorec_profile_enable": [_vp, _i],
    "baorec_profile_read": [_vp, C.c_char_p, _i, _f3, C.POINTER(C.c_int32), _i],
    "baorec_comm_unique_id": [_vp],
    "baorec_comm_init": [_vp, _i, _i, _vp],
    "baorec_plan_dist": [_vp, _i, _i, _i, _f3, _f3],
    "baorec_dist_ipc_close": [_vp],
    "baorec_dist_ipc_export": [_vp, _vp],
    "baorec_dist_ipc_open": [_vp, _vp, _i],
    "baorec_dist_exchange_mode": [_vp],
    "baorec_slab_range": [_vp, C.POINTER(_i), C.POINTER(_i)],
    "baorec_shard_catalog_f32": [_vp, _i, _vp, _vp, _vp, _vp, _i64] + [C.POINTER(_vp)] * 4 + [C.POINTER(_i64), _vp],
    "baorec_unshard_f32": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "baorec_reconstruct_dist_f32": [_vp, _pp, _i, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp],
    "baorec_reconstruct_dist_host_f32": [_vp, _pp, _i, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp],
    "baorec_slab_owner_f32": [_vp, _vp, _i64, _vp, _vp],
    "baorec_dist_r2c_f32": [_vp, _vp, _vp, _vp],
    "baorec_dist_c2r_f32": [_vp, _vp, _vp, _vp],
    "baorec_run_dist_f32": [_vp, _pp, _i, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp],
    "baorec_read_shifts_dist_f32": [_vp, _pp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp],
    "baorec_cic_scatter_f32": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp],
    "baorec_cic_cells_f32": [_vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp],
    "baorec_gather_cells_f32": [_vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp],
    "baorec_gather_f32": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i, _vp],
    "baorec_setup_box_f32": [_vp, _vp, _vp, _vp, _i64, _f, _f3, _f3, _vp],
    "baorec_smooth_f32": [_vp, _vp, _f, _vp],
    "baorec_setup_overdensity_f32": [_vp, _pp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i, _vp],
    "baorec_iterate_f32": [_vp, _vp, _vp, _i, _f, _f3, _vp],
    "baorec_reconstructed_overdensity_f32": [_vp, _pp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp],
    "baorec_mg_jacobi_f32": [_vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _f3, _vp],
    "baorec_mg_residual_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f3, _vp],
    "baorec_mg_restrict_f32": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "baorec_mg_prolong_f32": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "baorec_mg_vcycle_f32": [_vp, _vp, _vp, _f, _f, _i, _f3, _vp],
    "baorec_mg_fmg_f32": [_vp, _vp, _vp, _f, _f, _i, _i, _f3, _vp],
    "baorec_reconstructed_potential_f32": [_vp, _pp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp],
    "baorec_compute_displacements_f32": [_vp, _vp, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, _vp],
    "baorec_read_shifts_f32": [_vp, _pp, _i, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp],
    "baorec_reconstructed_positions_f32": [_vp, _pp, _i, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp],
    "baorec_read_result_cache_f32": [_vp, _pp, _i, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp],
    "baorec_displacement_meshes_f32": [_vp, _vp, _i, _vp, _vp, _vp, _vp],
    "baorec_run_host_f32": [_vp, _pp, _i, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _f3, _f3],
    "baorec_read_host_f32": [_vp, _pp, _i, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp],
    "baorec_batch_host_f32": [_vp, _pp, _i, _i] + [C.POINTER(_vp)] * 4 + [C.POINTER(_i64), _i, _i] + [C.POINTER(_vp)] * 3,
    "baorec_batch_files_f32": [_vp, _pp, _i, _i, C.POINTER(C.c_char_p), C.c_char, C.POINTER(_i), _i, _i, C.POINTER(C.c_char_p), _i,
                               C.POINTER(_i64), C.POINTER(C.c_double)],
    "baorec_result_cache": [_vp],
    "baorec_cosmo_set": [_vp, C.POINTER(CosmologyParams)],
    "baorec_cosmo_build_table": [C.POINTER(CosmologyParams), C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "baorec_cosmo_tables": [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), _i64],
    "baorec_sky_to_cartesian_f32": [_vp, _vp, _vp, _vp, _i64, _f, _vp, _vp, _vp, _vp],
    "baorec_cartesian_to_sky_f32": [_vp, _vp, _vp, _vp, _i64, _f, _vp, _vp, _vp, _vp],
    "baorec_fkp_weights_f32": [_vp, _vp, _i64, _f, _vp, _vp],
    "baorec_wrap_positions_f32": [_vp, _vp, _vp, _vp, _i64, _f3, _f3, _vp],
    "baorec_power_multipoles_f32": [_vp, _vp, _vp, _f3, C.c_double, C.c_double, _i, _i, C.c_double] + [C.POINTER(C.c_double)] * 5 + [_vp],
    "baorec_power_multipoles_interlaced_f32": [_vp, _vp, _vp, _vp, _vp, _f3, C.c_double, C.c_double, _i, _i, C.c_double] + [C.POINTER(C.c_double)] * 5 + [_vp],
    "baorec_interlace_positions_f32": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "baorec_compute_auto_box_f32": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i, _i, _f3, C.c_double, C.c_double, _i, C.c_double] + [C.POINTER(C.c_double)] * 5 + [_vp],
    "baorec_text_catalog_scan": [C.c_char_p, C.c_char, C.POINTER(_i64), C.POINTER(_i), _i],
    "baorec_text_catalog_read_f32": [C.c_char_p, C.c_char, _i, C.POINTER(_i), C.POINTER(_vp), _i64, C.POINTER(_i64), _i],
    "baorec_npy_info": [C.c_char_p, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i64), C.POINTER(_i64)],
    "baorec_npy_read_columns_f32": [C.c_char_p, _i, C.POINTER(_i), C.POINTER(_vp), _i64, C.POINTER(_i64), _i],
    "baorec_npy_write_columns_f32": [C.c_char_p, _i, C.POINTER(_vp), _i64],
    "baorec_catalog_select_f32": [_i, C.POINTER(_vp), _i64, _i, _f, _f, C.POINTER(_i64), _i],
    "baorec_host_alloc": [C.POINTER(_vp), _i64],
    "baorec_host_free": [_vp],
}
_RESTYPES = {"baorec_last_error": C.c_char_p, "baorec_scratch_bytes": C.c_int64, "baorec_sort_reuse_count": C.c_int64, "baorec_result_cache": C.c_void_p}

_lib = None


def load() -> C.CDLL:
    """Loads the shared library (once).  Raises if it is missing -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python baorec.jl_b200/build.py` "
            "(there is no CPU fallback for the B200 engine)")
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(code: int):
    if code == OK:
        return
    msg = load().baorec_last_error().decode("utf-8", "replace")
    if code == ERR_OUT_OF_BOX:
        raise OutOfBoxError(code, msg)
    if code == ERR_OUT_OF_RANGE:
        raise OutOfRangeError(code, msg)
    if code == ERR_IO:
        raise CatalogIOError(code, msg)
    raise BaorecError(code, msg)


def f3(values):
    return (C.c_float * 3)(*[float(v) for v in values])
