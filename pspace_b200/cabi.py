"""ctypes binding of the C-ABI in include/pspace_cuda.h (lib/libpspace_cuda.so).

Plumbing only: argument marshalling, error -> exception.  There is no fallback of any kind: if the
shared library is missing or no CUDA device is usable, calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSPACE_CUDA_LIB") or os.path.join(_HERE, "lib", "libpspace_cuda.so")   # override: kernel experiments

_vp, _i, _ll, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t
_ull, _dbl = ctypes.c_ulonglong, ctypes.c_double


class PspaceCudaError(RuntimeError):
    pass


class ProjectArgs(ctypes.Structure):
    _fields_ = [("k0", _i), ("k1", _i), ("j0", _i), ("j1", _i), ("q0", _ll), ("q1", _ll),
                ("ncols", _i), ("accumulate", _i), ("ksplit", _i), ("variant", _i), ("d_pair_mask", _vp), ("symmetric", _i)]


# name -> (restype, argtypes); every symbol declared in include/pspace_cuda.h
SIGNATURES = {
    "pspace_cuda_last_error": (ctypes.c_char_p, []),
    "pspace_cuda_version": (_i, []),
    "pspace_cuda_device_count": (_i, []),
    "pspace_cuda_create": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "pspace_cuda_destroy": (_i, [_vp]),
    "pspace_cuda_basis_degrees": (_i, [_i, _vp, _i, ctypes.POINTER(_i), _vp, _ll, _vp]),
    "pspace_cuda_basis_count": (_i, [_i, _vp, _i, ctypes.POINTER(_ll)]),
    "pspace_cuda_initialize_basis": (_i, [_vp, _vp, _vp]),
    "pspace_cuda_initialize_quadrature": (_i, [_vp, _vp, _vp]),
    "pspace_cuda_initialize": (_i, [_vp, _vp]),
    "pspace_cuda_num_basis": (_i, [_vp]),
    "pspace_cuda_num_params": (_i, [_vp]),
    "pspace_cuda_num_quad": (_ll, [_vp]),
    "pspace_cuda_is_complex": (_i, [_vp]),
    "pspace_cuda_get_degrees": (_i, [_vp, _vp]),
    "pspace_cuda_get_rules": (_i, [_vp, _i, _vp, _vp, _vp]),
    "pspace_cuda_get_table": (_i, [_vp, _i, _vp]),
    "pspace_cuda_get_weights": (_i, [_vp, _ll, _ll, _vp]),
    "pspace_cuda_tensor_grid": (_i, [_vp, _ll, _ll, _vp, _vp, _vp, _vp]),
    "pspace_cuda_eval_basis_grid": (_i, [_vp, _i, _i, _ll, _ll, _vp, _vp]),
    "pspace_cuda_eval_basis_points": (_i, [_vp, _i, _i, _ll, _vp, _vp, _vp]),
    "pspace_cuda_evaluate_expansion": (_i, [_vp, _i, _i, _ll, _ll, _i, _i, _vp, _vp, _vp]),
    "pspace_cuda_scatter_vector_tacs": (_i, [_i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "pspace_cuda_gather_vector_tacs": (_i, [_i, _i, _i, _i, _vp, _vp, _vp]),
    "pspace_cuda_scatter_matrix_tacs": (_i, [_i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "pspace_cuda_point_quantity_layout": (_i, [_i, _i, _i, _vp, _vp, _vp]),
    "pspace_cuda_pair_mask": (_i, [_vp, _vp, _vp, _vp]),
    "pspace_cuda_moments": (_i, [_vp, _ll, _ll, _i, _vp, _vp, _vp, _vp, _vp]),
    "pspace_cuda_project_workspace": (_sz, [_vp, _i, ctypes.POINTER(ProjectArgs)]),
    "pspace_cuda_project_vector": (_i, [_vp, ctypes.POINTER(ProjectArgs), _vp, _vp, _vp, _sz, _vp]),
    "pspace_cuda_project_matrix": (_i, [_vp, ctypes.POINTER(ProjectArgs), _vp, _vp, _vp, _sz, _vp]),
    "pspace_cuda_shard_bounds": (_i, [_ll, _i, _i, ctypes.POINTER(_ll), ctypes.POINTER(_ll)]),
    "pspace_cuda_row_block": (_i, [_i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "pspace_cuda_nccl_allreduce": (_i, [_vp, _sz, _vp, _vp]),
    "pspace_cuda_project_matrix_allreduce": (_i, [_vp, ctypes.POINTER(ProjectArgs), _vp, _vp, _vp, _sz, _vp, _vp]),
    "pspace_cuda_project_vector_allreduce": (_i, [_vp, ctypes.POINTER(ProjectArgs), _vp, _vp, _vp, _sz, _vp, _vp]),
    "pspace_cuda_project_matrix_reduce_scatter": (_i, [_vp, ctypes.POINTER(ProjectArgs), _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "pspace_cuda_peer_layout": (_i, [_vp, ctypes.POINTER(ProjectArgs), _i, ctypes.POINTER(_ll)]),
    "pspace_cuda_project_matrix_peer": (_i, [_vp, ctypes.POINTER(ProjectArgs), _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "pspace_cuda_peer_reduce_gather": (_i, [_vp, ctypes.POINTER(ProjectArgs), _vp, _vp, _i, _i, _vp]),
    "pspace_cuda_project_matrix_batched": (_i, [_vp, ctypes.POINTER(ProjectArgs), _i, _vp, _ll, _vp, _ll, _vp, _sz, _vp]),
    "pspace_cuda_project_vector_host": (_i, [_vp, ctypes.POINTER(ProjectArgs), _vp, _vp]),
    "pspace_cuda_project_matrix_host": (_i, [_vp, ctypes.POINTER(ProjectArgs), _vp, _vp]),
    "pspace_cuda_project_matrix_hostin": (_i, [_vp, ctypes.POINTER(ProjectArgs), _vp, _vp]),
    "pspace_cuda_adjoint_states": (_i, [_vp, _i, _i, _ll, _ll, _vp, _vp, _vp, _vp]),
    "pspace_cuda_adjoint_product": (_i, [_vp, _i, _vp, _ll, _ll, _vp, _vp, _i, _vp]),
    "pspace_cuda_activate": (_i, [_vp]),
    "pspace_cuda_launch_count": (_ll, [_vp]),
    "pspace_cuda_set_option": (_i, [_vp, ctypes.c_char_p, _i]),
    "pspace_cuda_last_plan": (_i, [_vp, ctypes.c_char_p, _i]),
    "pspace_cuda_fp64_peak": (_i, [_i, _i, _dbl, ctypes.POINTER(_dbl)]),
    "pspace_cuda_fill_synthetic": (_i, [_vp, _ll, _ull, _ull, _vp]),
}

_lib = None


def lib():
    """Load libpspace_cuda.so (once).  Raises if it has not been built — there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PspaceCudaError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(make -C pspace_b200/csrc). The pspace GPU path has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PspaceCudaError(f"pspace_cuda error {rc}: {lib().pspace_cuda_last_error().decode()}")


def np_ptr(a):
    return a.ctypes.data if a is not None else None
