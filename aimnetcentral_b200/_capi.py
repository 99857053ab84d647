"""ctypes binding of libaimnet2_b200.so (include/aimnet2_b200.h).  No CPU fallback: if the library cannot be loaded the
import of the compute path fails loudly."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

AIMNET_OK = 0
AIMNET_NEIGHBOR_OVERFLOW = 1
COULOMB = {None: 0, "none": 0, "simple": 1, "dsf": 2, "ewald": 3}
WANT_FORCES = 1
WANT_STRESS = 2

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class Weights(C.Structure):
    _fields_ = [
        ("num_charge_channels", C.c_int),
        ("afv", C.c_void_p), ("agh_a", C.c_void_p), ("agh_q", C.c_void_p), ("shifts_s", C.c_void_p),
        ("eta_s", C.c_float), ("rc_s", C.c_float),
        ("n_layers", C.c_int * 3),
        ("layer_dims", C.c_void_p * 3),
        ("mlp_w", C.c_void_p * 3),
        ("mlp_b", C.c_void_p * 3),
        ("head_w", C.c_void_p * 3),
        ("head_b", C.c_void_p * 3),
        ("sae", C.c_void_p),
        ("sr_rc", C.c_float), ("sr_envelope", C.c_int),
        ("d3_c6ref", C.c_void_p), ("d3_cnref", C.c_void_p), ("d3_rcov", C.c_void_p), ("d3_r4r2", C.c_void_p),
    ]


class Options(C.Structure):
    _fields_ = [
        ("coulomb_method", C.c_int), ("dsf_alpha", C.c_float), ("dsf_rc", C.c_float), ("ewald_accuracy", C.c_float),
        ("dispersion", C.c_int), ("d3_s6", C.c_float), ("d3_s8", C.c_float), ("d3_a1", C.c_float),
        ("d3_a2", C.c_float), ("d3_cutoff", C.c_float), ("d3_smoothing", C.c_float), ("sr_cutoff", C.c_float),
        ("neighbor_skin", C.c_float),
    ]


class System(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int), ("n_mol", C.c_int),
        ("coord", C.c_void_p), ("numbers", C.c_void_p), ("mol_idx", C.c_void_p), ("charge", C.c_void_p),
        ("mult", C.c_void_p), ("cell", C.c_void_p), ("host_cell", C.c_void_p), ("n_cells", C.c_int),
        ("pbc_host", C.c_void_p), ("nbmat", C.c_void_p), ("shifts", C.c_void_p), ("nb_width", C.c_int),
    ]


class Result(C.Structure):
    _fields_ = [
        ("energy", C.c_void_p), ("charges", C.c_void_p), ("spin_charges", C.c_void_p), ("forces", C.c_void_p),
        ("stress", C.c_void_p), ("nbmat_out", C.c_void_p), ("shifts_out", C.c_void_p), ("nbmat_out_width", C.c_int),
    ]


_LIB = None
ABI_VERSION = 2   # include/aimnet2_b200.h AIMNET2_ABI_VERSION

EXPORTS = [
    "aimnet2_last_error", "aimnet2_abi_version", "aimnet2_abi_struct_sizes", "aimnet2_neighbor_matrix", "aimnet2_wrap_positions",
    "aimnet2_conv_sv_2d_sp_fwd", "aimnet2_conv_sv_2d_sp_bwd", "aimnet2_engine_create", "aimnet2_engine_destroy",
    "aimnet2_engine_set_options", "aimnet2_engine_set_gemm_backend", "aimnet2_engine_set_small_m_rows", "aimnet2_engine_set_conv_impl", "aimnet2_engine_conv_mode", "aimnet2_engine_set_dense_min_molecules", "aimnet2_engine_debug_poison", "aimnet2_engine_debug_layout", "aimnet2_engine_debug_read_workspace", "aimnet2_engine_set_deterministic", "aimnet2_engine_eval", "aimnet2_engine_eval_host",
    "aimnet2_engine_enable_cuda_graph", "aimnet2_engine_graph_stats", "aimnet2_engine_last_launches", "aimnet2_engine_info", "aimnet2_engine_skin_stats", "aimnet2_engine_enable_timing",
    "aimnet2_engine_last_timing", "aimnet2_gemm_nt", "aimnet2_gemm_set_trace", "aimnet2_engine_neighbor_caps", "aimnet2_engine_set_species_first_pass",
    "aimnet2_dsf_coulomb", "aimnet2_dftd3", "aimnet2_ewald_summation", "aimnet2_estimate_ewald_parameters",
]


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the sources changed and nvcc is present) the C-ABI library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if os.path.exists(_build.NVCC):
        path = _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and nvcc is not available: the CUDA extension is required "
                           "(there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.aimnet2_last_error.restype = C.c_char_p
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    lib.aimnet2_neighbor_matrix.argtypes = [vp, ci, cf, vp, vp, vp, ci, vp, ci, ci, ci, ci, vp, vp, vp, c_int_p, vp]
    lib.aimnet2_wrap_positions.argtypes = [vp, vp, ci, vp, ci, vp, vp, vp]
    lib.aimnet2_conv_sv_2d_sp_fwd.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp]
    lib.aimnet2_conv_sv_2d_sp_bwd.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]
    lib.aimnet2_engine_create.argtypes = [C.POINTER(vp), C.POINTER(Weights), ci]
    lib.aimnet2_engine_destroy.argtypes = [vp]
    lib.aimnet2_engine_set_options.argtypes = [vp, C.POINTER(Options)]
    lib.aimnet2_engine_set_gemm_backend.argtypes = [vp, ci]
    lib.aimnet2_engine_set_deterministic.argtypes = [vp, ci]
    lib.aimnet2_engine_set_small_m_rows.argtypes = [vp, ci]
    lib.aimnet2_engine_set_conv_impl.argtypes = [vp, ci]
    lib.aimnet2_engine_conv_mode.argtypes = [vp, c_int_p, c_int_p, c_int_p]
    lib.aimnet2_engine_set_dense_min_molecules.argtypes = [vp, ci]
    lib.aimnet2_engine_debug_poison.argtypes = [vp, ci]
    lib.aimnet2_engine_debug_layout.argtypes = [vp, C.c_char_p, ci]
    lib.aimnet2_engine_debug_read_workspace.argtypes = [vp, vp, C.c_int64, C.c_int64]
    lib.aimnet2_engine_eval.argtypes = [vp, C.POINTER(System), C.POINTER(Result), ci, vp]
    lib.aimnet2_engine_eval_host.argtypes = [vp, C.POINTER(System), C.POINTER(Result), ci]
    lib.aimnet2_engine_enable_cuda_graph.argtypes = [vp, ci]
    lib.aimnet2_engine_graph_stats.argtypes = [vp, c_int_p, c_int_p, c_int_p]
    lib.aimnet2_engine_last_launches.argtypes = [vp]
    lib.aimnet2_engine_info.argtypes = [vp, c_int_p, c_int_p, C.POINTER(C.c_int64)]
    lib.aimnet2_engine_neighbor_caps.argtypes = [vp, c_int_p, c_int_p]
    lib.aimnet2_engine_set_species_first_pass.argtypes = [vp, C.c_int]
    lib.aimnet2_engine_skin_stats.argtypes = [vp, c_int_p, c_int_p]
    lib.aimnet2_engine_enable_timing.argtypes = [vp, ci]
    lib.aimnet2_engine_last_timing.argtypes = [vp, c_float_p, ci]
    lib.aimnet2_gemm_nt.argtypes = [vp, ci, vp, ci, vp, vp, ci, vp, ci, ci, ci, ci, ci, ci, vp]
    lib.aimnet2_gemm_set_trace.argtypes = [vp]
    cd = C.c_double
    lib.aimnet2_dsf_coulomb.argtypes = [vp, vp, ci, cf, cf, vp, ci, vp, ci, vp, vp, ci, ci, vp, vp, vp, vp, vp]
    lib.aimnet2_dftd3.argtypes = [vp, vp, ci, cf, cf, cf, cf, cf, cf, vp, vp, vp, vp, vp, ci, vp, ci, vp, vp, ci, ci, vp, vp, vp,
                                  vp, vp]
    lib.aimnet2_ewald_summation.argtypes = [vp, vp, ci, vp, vp, vp, vp, ci, vp, vp, ci, ci, cd, vp, vp, vp, vp, vp]
    lib.aimnet2_estimate_ewald_parameters.argtypes = [c_float_p, ci, cd, C.POINTER(cd), C.POINTER(cd), C.POINTER(cd)]
    for name in EXPORTS:
        if name != "aimnet2_last_error":
            getattr(lib, name).restype = ci
    lib.aimnet2_abi_struct_sizes.argtypes = [c_int_p, c_int_p, c_int_p, c_int_p]
    # the ctypes mirrors above must describe the structs this library was compiled with
    sizes = [C.c_int() for _ in range(4)]
    lib.aimnet2_abi_struct_sizes(*[C.byref(x) for x in sizes])
    mine = [C.sizeof(Weights), C.sizeof(Options), C.sizeof(System), C.sizeof(Result)]
    if lib.aimnet2_abi_version() != ABI_VERSION or [x.value for x in sizes] != mine:
        raise RuntimeError(f"{path}: ABI mismatch (library version {lib.aimnet2_abi_version()}, struct sizes "
                           f"{[x.value for x in sizes]}; binding version {ABI_VERSION}, sizes {mine}) - rebuild the library")
    _LIB = lib
    return lib


class NeighborOverflowError(Exception):
    """Mirror of nvalchemiops.neighbors.NeighborOverflowError (aimnet/calculators/neighbors.py:16)."""


def check(rc: int, what: str = ""):
    if rc == AIMNET_OK:
        return
    msg = load().aimnet2_last_error().decode(errors="replace")
    if rc == AIMNET_NEIGHBOR_OVERFLOW:
        raise NeighborOverflowError(what or "max_neighbors too small")
    if rc == -1:
        raise ValueError(f"{what}: {msg}" if what else msg)
    raise RuntimeError(f"{what}: {msg}" if what else msg)
