"""ctypes binding of libopmb200.so -- one prototype per declaration of include/opmb200.h.

The library is the ONLY compute path of this package: if it cannot be loaded, or if it finds no
CUDA device, the caller gets an exception -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# OPMB200_LIB: a differently built copy of the SAME library (design experiments, e.g. -fmad=false)
LIB_PATH = os.environ.get("OPMB200_LIB") or os.path.join(HERE, "libopmb200.so")

SUCCESS, INVALID_ARGUMENT, BAD_OPTIONS, MATRIX_BLOCK_ERROR, SOLVER_ABORT, CUDA_ERROR, NCCL_ERROR, \
    DIAGONAL_MISSING, NOT_PREPARED = range(9)


class Result(C.Structure):
    """== Dune::InverseOperatorResult"""
    _fields_ = [("iterations", C.c_int), ("reduction", C.c_double), ("converged", C.c_int),
                ("conv_rate", C.c_double), ("elapsed", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Halo(C.Structure):
    _fields_ = [("n_neighbors", C.c_int), ("neighbor_rank", C.POINTER(C.c_int)),
                ("send_ptr", C.POINTER(C.c_int)), ("send_rows", C.POINTER(C.c_int)),
                ("recv_ptr", C.POINTER(C.c_int)), ("recv_rows", C.POINTER(C.c_int))]


class Info(C.Structure):
    _fields_ = [("block_size", C.c_int), ("n_rows", C.c_int64), ("n_interior", C.c_int64), ("nnzb", C.c_int64),
                ("n_levels", C.c_int), ("n_slices", C.c_int), ("padded_blocks", C.c_int64),
                ("structurally_symmetric", C.c_int), ("preconditioner", C.c_int), ("relaxation", C.c_double),
                ("tol", C.c_double), ("maxiter", C.c_int), ("n_ranks", C.c_int), ("t_analysis_s", C.c_double),
                ("t_update_ms", C.c_double), ("t_solve_ms", C.c_double), ("kernel_launches", C.c_int64),
                ("schedule", C.c_int), ("n_chunks", C.c_int), ("chunk_rows", C.c_int), ("est_steps", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_vp = C.c_void_p  # double* that may be a host or a device pointer

# name -> (restype, argtypes); must list every symbol of include/opmb200.h
PROTOTYPES = {
    "opmb200_version": (C.c_int, []),
    "opmb200_last_error": (C.c_char_p, []),
    "opmb200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "opmb200_set_device": (C.c_int, [C.c_int]),
    "opmb200_row_coloring": (C.c_int, [C.c_int64, _i32p, _i32p, C.c_int, _i32p, _i32p, _i32p, C.POINTER(C.c_int32)]),
    "opmb200_plan_schedule": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _i32p, _i32p, C.c_int64, C.c_int, C.c_int,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_double), _vp, _vp, _vp]),
    "opmb200_plan_tiles": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _i32p, _i32p, C.c_int64, C.c_int, C.c_int,
                                      _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "opmb200_partition_simple": (C.c_int, [C.c_int32, C.c_int32, _i32p]),
    "opmb200_localize": (C.c_int, [C.c_int64, _i32p, _i32p, _i32p, C.c_int32, C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int64), _vp, _vp, _vp, _vp]),
    "opmb200_comm_unique_id": (C.c_int, [_vp]),
    "opmb200_comm_create": (C.c_int, [C.c_int, C.c_int, _vp, C.POINTER(_vp)]),
    "opmb200_comm_destroy": (C.c_int, [_vp]),
    "opmb200_create": (C.c_int, [C.c_char_p, C.c_int, C.c_int64, C.c_int64, _i32p, _i32p, C.c_int64, _vp,
                                  C.POINTER(Halo), C.POINTER(_vp)]),
    "opmb200_destroy": (C.c_int, [_vp]),
    "opmb200_update_values": (C.c_int, [_vp, _vp]),
    "opmb200_precond_apply": (C.c_int, [_vp, _vp, _vp]),
    "opmb200_op_apply": (C.c_int, [_vp, _vp, _vp]),
    "opmb200_op_applyscaleadd": (C.c_int, [_vp, C.c_double, _vp, _vp]),
    "opmb200_dot": (C.c_int, [_vp, _vp, _vp, C.POINTER(C.c_double)]),
    "opmb200_solve": (C.c_int, [_vp, _vp, _vp, C.c_double, C.POINTER(Result)]),
    "opmb200_get_info": (C.c_int, [_vp, C.POINTER(Info)]),
    "opmb200_get_levels": (C.c_int, [_vp, _vp, _vp]),
    "opmb200_get_reorder": (C.c_int, [_vp, _vp, _vp]),
    "opmb200_get_dinv": (C.c_int, [_vp, _vp]),
    "opmb200_get_ilu0": (C.c_int, [_vp, _vp]),
    "opmb200_get_history": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_int)]),
    "opmb200_time_kernel": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "opmb200_p2p_export": (C.c_int, [_vp, _vp]),
    "opmb200_p2p_import": (C.c_int, [_vp, _vp]),
    "opmb200_timer_start": (C.c_int, [_vp]),
    "opmb200_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "opmb200_set_wells": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "opmb200_cpr_quasi_impes_weights": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "opmb200_cpr_coarse_entries": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "opmb200_cpr_restrict": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "opmb200_cpr_prolongate": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
}

_lib = None


class B200Error(RuntimeError):
    """carries the opmb200_status; the subclasses mirror the exceptions the Dune adapter throws"""

    def __init__(self, status, message):
        super().__init__(f"[opmb200 status {status}] {message}")
        self.status = status
        self.message = message


class InvalidArgument(B200Error, ValueError):
    """std::invalid_argument (unknown solver / preconditioner type, malformed options)"""


class MatrixBlockError(B200Error):
    """Dune::MatrixBlockError"""


class SolverAbort(B200Error):
    """Dune::SolverAbort"""


def _preload_nccl():
    """libopmb200 links libnccl.so.2.  PyTorch ships its own (newer) copy under the same SONAME; if
    the system copy were loaded first, a later `import torch` in the same process would bind to
    it and miss symbols.  Load torch's copy first when it exists so that both agree."""
    import glob
    import sys
    for base in sys.path:
        for cand in glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
            except OSError:
                pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        _preload_nccl()
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status == SUCCESS:
        return
    msg = lib().opmb200_last_error().decode(errors="replace")
    if status in (BAD_OPTIONS, INVALID_ARGUMENT):
        raise InvalidArgument(status, msg)
    if status in (MATRIX_BLOCK_ERROR, DIAGONAL_MISSING):
        raise MatrixBlockError(status, msg)
    if status == SOLVER_ABORT:
        raise SolverAbort(status, msg)
    raise B200Error(status, msg)


def ptr(a) -> C.c_void_p:
    """host numpy array, torch tensor (host or cuda), raw int address or None -> void*"""
    if a is None:
        return C.c_void_p(None)
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        if not a.flags.c_contiguous:
            raise ValueError("array must be C-contiguous")
        if a.dtype != np.float64:
            raise TypeError(f"the C ABI takes double*: got a {a.dtype} array")
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch tensor
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        if "float64" not in str(a.dtype):
            raise TypeError(f"the C ABI takes double*: got a {a.dtype} tensor")
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")
