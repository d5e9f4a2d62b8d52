"""ctypes binding of ``libcirq_b200.so`` (C-ABI in ``include/cirq_b200.h``).

The product path has no CPU fallback: if the shared library is missing or a
call fails, a ``B200Error`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_uint8, c_uint64, c_void_p

import numpy as np

C64 = 0
C128 = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcirq_b200.so')


class B200Error(RuntimeError):
    """Raised when the CUDA library is missing or reports an error."""


_lib = None

# name -> (restype, argtypes); every symbol declared in include/cirq_b200.h
SIGNATURES = {
    'b2q_version': (c_int, []),
    'b2q_last_error': (c_char_p, []),
    'b2q_device_info': (c_int, [POINTER(c_int), POINTER(c_uint64), POINTER(c_int)]),
    'b2q_launch_count': (c_uint64, []),
    'b2q_sv_init_basis': (c_int, [c_void_p, c_int, c_int, c_uint64, c_void_p]),
    'b2q_sv_scale': (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_void_p]),
    'b2q_sv_apply_matrix': (
        c_int,
        [c_void_p, c_int, c_int, c_void_p, POINTER(c_int), c_int, c_void_p, c_void_p],
    ),
    'b2q_sv_apply_batch': (
        c_int,
        [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), c_void_p, c_void_p, c_void_p],
    ),
    'b2q_sv_apply_diagonal': (
        c_int,
        [c_void_p, c_int, c_int, c_void_p, POINTER(c_int), c_int, c_void_p],
    ),
    'b2q_sv_norm2': (c_int, [c_void_p, c_int, c_int, POINTER(c_double), c_void_p]),
    'b2q_sv_gather': (c_int, [c_void_p, c_int, c_int, c_void_p, c_uint64, c_void_p, c_void_p]),
    'b2q_sv_marginal_probs': (
        c_int,
        [c_void_p, c_int, c_int, POINTER(c_int), c_int, c_void_p, c_void_p, c_void_p],
    ),
    'b2q_sv_sample': (
        c_int,
        [c_void_p, c_int, c_int, c_void_p, c_uint64, c_void_p, c_void_p, c_uint64, c_void_p],
    ),
    'b2q_sv_sample_workspace_bytes': (c_uint64, [c_int, c_uint64]),
    'b2q_cdf_sample': (c_int, [c_void_p, c_uint64, c_void_p, c_uint64, c_void_p, c_void_p]),
    'b2q_unpack_bits': (c_int, [c_void_p, c_uint64, POINTER(c_int), c_int, c_void_p, c_void_p]),
    'b2q_sv_collapse': (
        c_int,
        [c_void_p, c_int, c_int, POINTER(c_int), POINTER(c_int), c_int, c_double, c_void_p],
    ),
    'b2q_sv_pauli_expectation': (
        c_int,
        [c_void_p, c_int, c_int, c_uint64, c_uint64, POINTER(c_double), c_void_p],
    ),
    'b2q_sv_pauli_expectation_multi': (
        c_int, [c_void_p, c_int, c_int, c_uint64, c_void_p, c_int, c_void_p, c_void_p]),
    'b2q_sv_kron': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'b2q_sv_permute_bits': (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_int), c_void_p]),
    'b2q_sv_permute_bits_inplace': (c_int, [c_void_p, c_int, c_int, POINTER(c_int), POINTER(c_int), c_void_p]),
    'b2q_sv_argmax_abs': (c_int, [c_void_p, c_int, c_int, POINTER(c_uint64), c_void_p]),
    'b2q_sv_kron_allclose': (
        c_int,
        [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_double, c_double, POINTER(c_int), c_void_p],
    ),
    'b2q_sv_allclose': (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_double, POINTER(c_int), c_void_p]
    ),
    'b2q_dm_partial_trace': (
        c_int, [c_void_p, c_int, c_int, POINTER(c_int), c_int, c_void_p, c_void_p]
    ),
    'b2q_dm_diagonal': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'b2q_probs_marginal': (c_int, [c_void_p, c_int, POINTER(c_int), c_int, c_void_p, c_void_p]),
    'b2q_dm_pauli_expectation': (c_int, [c_void_p, c_int, c_int, c_uint64, c_uint64, POINTER(c_double), c_void_p]),
    'b2q_dm_trace': (c_int, [c_void_p, c_int, c_int, POINTER(c_double), c_void_p]),
    'b2q_dm_collapse': (
        c_int,
        [c_void_p, c_int, c_int, POINTER(c_int), POINTER(c_int), c_int, c_double, c_void_p],
    ),
    'b2q_dist_alloc': (c_int, [c_uint64, POINTER(c_void_p)]),
    'b2q_dist_free': (c_int, [c_void_p]),
    'b2q_dist_ipc_get': (c_int, [c_void_p, c_void_p]),
    'b2q_dist_ipc_open': (c_int, [c_void_p, POINTER(c_void_p)]),
    'b2q_dist_ipc_close': (c_int, [c_void_p]),
    'b2q_dist_swap_bit': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'b2q_dist_pack': (c_int, [c_void_p, c_int, c_int, POINTER(c_int), c_int, c_void_p, c_void_p]),
    'b2q_dist_apply_exchange': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                        POINTER(c_int), c_int, c_int, c_int, c_void_p]),
    'b2q_bsv_apply_select': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, POINTER(c_int), c_int,
                                     c_void_p, c_void_p, c_int, c_void_p]),
    'b2q_host_left_apply': (c_int, [c_void_p, c_int, c_void_p, POINTER(c_int), c_int]),
    'b2q_host_compose': (c_int, [c_void_p, c_int, c_int, POINTER(c_int), POINTER(c_int), c_void_p]),
    'b2q_host_compose_diag': (c_int, [c_void_p, c_int, c_int, POINTER(c_int), POINTER(c_int), c_void_p]),
    'b2q_schedule_op_bytes': (c_int, []),
    'b2q_run_schedule': (c_int, [c_int, c_int, c_void_p, POINTER(c_int), c_void_p, c_int, c_void_p, POINTER(c_int),
                                 c_void_p]),
    'b2q_sv_reduced_density_matrix': (c_int, [c_void_p, c_int, c_int, POINTER(c_int), c_int, c_void_p, c_void_p]),
    'b2q_bsv_apply_select_multi': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, POINTER(c_int), c_int,
                                           c_void_p, c_int, c_void_p]),
    'b2q_bsv_kraus_weights': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, POINTER(c_int), c_int,
                                      c_void_p, c_void_p]),
    'b2q_bsv_collapse': (c_int, [c_void_p, c_int, c_int, c_int, c_uint64, c_void_p, c_void_p, c_void_p]),
    'b2q_dist_unpack': (c_int, [c_void_p, c_int, c_int, POINTER(c_int), c_int, c_void_p, c_void_p]),
    'b2q_dist_swap_bits': (c_int, [c_void_p, POINTER(c_void_p), c_int, c_int, POINTER(c_int), c_int, c_int,
                                   c_void_p]),
    'b2q_tile_blocks_feasible': (c_int, [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    'b2q_sv_apply_tile_blocks': (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int),
                                         c_void_p, c_void_p]),
}

# Host-only helpers exported for unit tests (not part of the product ABI).
DEBUG_SIGNATURES = {
    'b2q_set_lane_mode': (c_int, [c_int]),
    'b2q_set_vec_mode': (c_int, [c_int]),
    'b2q_set_tc_mode': (c_int, [c_int]),
    'b2q_set_tc_stage_mode': (c_int, [c_int]),
    'b2q_set_tc_stage_opts': (c_int, [c_int, c_int]),
    'b2q_debug_tc_stage_plan': (c_int, [c_int, POINTER(c_int), c_int, POINTER(ctypes.c_int64)]),
    'b2q_debug_plan': (c_int, [c_int, c_int, POINTER(c_int), c_int, POINTER(c_int)]),
    'b2q_debug_permute_plan': (c_int, [c_int, c_int, POINTER(c_int), c_int, POINTER(c_int), POINTER(c_int)]),
    'b2q_debug_tile_plan': (c_int, [c_int, c_int, POINTER(c_int), POINTER(ctypes.c_int64)]),
    'b2q_debug_permute_matrix': (c_int, [c_void_p, POINTER(c_int), c_int, c_void_p]),
    'b2q_debug_rdm_plan': (c_int, [c_int, POINTER(c_int), c_int, POINTER(ctypes.c_int64)]),
    'b2q_debug_pauli_plan': (c_int, [c_int, POINTER(ctypes.c_int64)]),
}


def load():
    """Loads the shared library (once) and declares all signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; '
            'g.build()"` (or `make -C cirq_b200/csrc`). There is no CPU fallback.'
        )
    lib = ctypes.CDLL(LIB_PATH)
    for table in (SIGNATURES, DEBUG_SIGNATURES):
        for name, (restype, argtypes) in table.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
    mode = os.environ.get('CIRQ_B200_LANE_MODE')
    if mode is not None:
        lib.b2q_set_lane_mode(int(mode))
    mode = os.environ.get('CIRQ_B200_TC_MODE')
    if mode is not None:
        lib.b2q_set_tc_mode(int(mode))
    mode = os.environ.get('CIRQ_B200_TC_STAGE_MODE')
    if mode is not None:
        lib.b2q_set_tc_stage_mode(int(mode))
    mode = os.environ.get('CIRQ_B200_TC_STAGE_OPTS')
    if mode is not None:
        early, ahead = (int(v) for v in mode.split(','))
        lib.b2q_set_tc_stage_opts(early, ahead)
    mode = os.environ.get('CIRQ_B200_VEC_MODE')
    if mode is not None:
        lib.b2q_set_vec_mode(int(mode))
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b2q_last_error().decode('utf-8', 'replace')
        raise B200Error(f'cirq_b200 error {rc}: {msg}')


def int_array(values):
    values = [int(v) for v in values]
    return (c_int * max(1, len(values)))(*values)


def dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.complex64:
        return C64
    if dt == np.complex128:
        return C128
    raise ValueError(f'dtype must be complex64 or complex128 but was {dt}')


def as_c128_buffer(matrix) -> np.ndarray:
    """Row-major complex128 copy of a matrix, kept alive by the caller."""
    return np.ascontiguousarray(np.asarray(matrix, dtype=np.complex128))
