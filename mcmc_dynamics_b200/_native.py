"""ctypes binding of the C ABI declared in ``include/mcd_b200.h`` (``libmcd_b200.so``) and loader
of the torch operator library built on top of it (``libmcd_torch.so``).

There is no CPU fallback: if the library is missing, or no CUDA device is visible when a handle is
created, the call raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
#: MCD_B200_LIB selects another build of the same library (A/B runs of kernel tuning knobs)
LIB_PATH = os.environ.get('MCD_B200_LIB') or os.path.join(_HERE, '_lib', 'libmcd_b200.so')
TORCH_LIB_PATH = os.path.join(_HERE, '_lib', 'libmcd_torch.so')

ABI_VERSION = 3
NPARAM = 11
MAX_THETA = 16

ROT_CONSTANT, ROT_RADIAL = 0, 1
BG_NONE, BG_FIXED_PMEMBER, BG_FIXED_DENSITY, BG_GAUSSIAN = 0, 1, 2, 3
MATH_FAST, MATH_PLAIN = 0, 1

#: parameter slots, in the order of the MCD_P_* enum
PARAM_SLOTS = ('v_sys', 'sigma_max', 'v_maxx', 'v_maxy', 'ra_center', 'dec_center', 'a', 'r_peak', 'v_back',
               'sigma_back', 'f_back')

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)


class PackDesc(ctypes.Structure):
    """``mcd_pack_desc``."""
    _fields_ = [
        ('rotation', ctypes.c_int32), ('background', ctypes.c_int32), ('n_theta', ctypes.c_int32),
        ('math_mode', ctypes.c_int32), ('n_stars', ctypes.c_int64),
        ('ra', _c_double_p), ('dec', _c_double_p), ('v', _c_double_p), ('verr', _c_double_p),
        ('pmember', _c_double_p), ('density', _c_double_p), ('lnlike_background', _c_double_p),
        ('slot', ctypes.c_int32 * NPARAM), ('fixed_value', ctypes.c_double * NPARAM),
        ('unit_scale', ctypes.c_double * NPARAM),
        ('lower', ctypes.c_double * MAX_THETA), ('upper', ctypes.c_double * MAX_THETA),
        ('fixed_prior_ok', ctypes.c_int32), ('device', ctypes.c_int32), ('n_stars_total', ctypes.c_int64),
        ('n_segments', ctypes.c_int32), ('segment_offsets', _c_int64_p),
    ]


class Info(ctypes.Structure):
    """``mcd_info``."""
    _fields_ = [
        ('n_stars', ctypes.c_int64), ('n_theta', ctypes.c_int32), ('n_columns', ctypes.c_int32),
        ('bytes_per_star', ctypes.c_int32), ('flops_per_term', ctypes.c_int32), ('free_centre', ctypes.c_int32),
        ('sm_count', ctypes.c_int32), ('last_grid_x', ctypes.c_int32), ('last_grid_y', ctypes.c_int32),
        ('last_block', ctypes.c_int32), ('last_walker_tile', ctypes.c_int32), ('launches', ctypes.c_int64),
        ('n_segments', ctypes.c_int32),
    ]


#: every symbol include/mcd_b200.h declares: name -> (restype, argtypes)
_vp = ctypes.c_void_p
SYMBOLS = {
    'mcd_abi_version': (ctypes.c_int, []),
    'mcd_last_error': (ctypes.c_char_p, []),
    'mcd_pack_create': (ctypes.c_int, [ctypes.POINTER(PackDesc), ctypes.POINTER(_vp)]),
    'mcd_pack_reconfigure': (ctypes.c_int, [_vp, ctypes.POINTER(PackDesc)]),
    'mcd_destroy': (None, [_vp]),
    'mcd_get_info': (ctypes.c_int, [_vp, ctypes.POINTER(Info)]),
    # the two per-call entry points take raw addresses (ndarray.ctypes.data): no pointer object per call
    'mcd_lnlike': (ctypes.c_int, [_vp, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    'mcd_lnlike_device': (ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, _vp]),
    'mcd_lnprob': (ctypes.c_int, [_vp, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    'mcd_lnprob_device': (ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, _vp]),
    'mcd_lnprob_partial_device': (ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, _vp]),
    'mcd_exchange_bytes': (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, _c_int64_p]),
    'mcd_exchange_attach': (ctypes.c_int, [_vp, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint64),
                                           ctypes.c_int32]),
    'mcd_lnprob_allreduce_device': (ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, _vp]),
    'mcd_lnprob_allreduce': (ctypes.c_int, [_vp, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    'mcd_exchange_status': (ctypes.c_int, [_vp]),
    'mcd_lnlike_per_star': (ctypes.c_int, [_vp, _c_double_p, _c_double_p]),
    'mcd_lnlike_per_star_device': (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    'mcd_membership_per_star': (ctypes.c_int, [_vp, _c_double_p, _c_double_p]),
    'mcd_membership_per_star_device': (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    'mcd_model_per_star': (ctypes.c_int, [_vp, _c_double_p, _c_double_p, _c_double_p]),
    'mcd_model_per_star_device': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp]),
    'mcd_calculate_lnlike': (ctypes.c_int, [_vp, _c_double_p, _c_double_p, _c_double_p]),
    'mcd_single_stars_lnlike': (ctypes.c_int, [ctypes.c_int32, _c_double_p, ctypes.c_int64, _c_double_p, _c_double_p,
                                               ctypes.c_int64, ctypes.c_double, _c_double_p]),
    'mcd_single_stars_lnlike_device': (ctypes.c_int, [ctypes.c_int32, _vp, ctypes.c_int64, _vp, _vp, ctypes.c_int64,
                                                      ctypes.c_double, _vp, _vp]),
    'mcd_gaussian_lnlike': (ctypes.c_int, [ctypes.c_int32, _c_double_p, _c_double_p, ctypes.c_int64, ctypes.c_double,
                                           ctypes.c_double, _c_double_p]),
    'mcd_ensemble_create': (ctypes.c_int, [_vp, ctypes.c_int32, ctypes.c_uint64, ctypes.c_double,
                                           ctypes.POINTER(_vp)]),
    'mcd_ensemble_destroy': (None, [_vp]),
    'mcd_ensemble_set_state': (ctypes.c_int, [_vp, _c_double_p]),
    'mcd_ensemble_run': (ctypes.c_int, [_vp, ctypes.c_int32, _c_double_p, _c_double_p, _c_int64_p]),
    'mcd_ensemble_get_state': (ctypes.c_int, [_vp, _c_double_p, _c_double_p]),
    'mcd_ensemble_engine': (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    'mcd_measure_fp64_peak': (ctypes.c_int, [ctypes.c_int32, _c_double_p, _c_double_p]),
    'mcd_measure_read_bandwidth': (ctypes.c_int, [ctypes.c_int32, ctypes.c_int64, _c_double_p]),
}

_lib = None
_torch_ops_loaded = False


class NativeError(RuntimeError):
    """An entry point of libmcd_b200.so returned an error code."""


def load_library(path=None):
    """Load ``libmcd_b200.so`` and bind every declared symbol.  Raises if the library has not been
    built (``python __graft_entry__.py``) or lacks a symbol."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise NativeError("{0} is missing: build it with `python __graft_entry__.py`. There is no CPU "
                          "fallback for the likelihood path.".format(path))
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError names the missing symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.mcd_abi_version() != ABI_VERSION:
        raise NativeError('ABI version mismatch: library {0}, binding {1}'.format(lib.mcd_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def load_torch_ops():
    """Register ``torch.ops.mcd_b200.*`` (device-tensor entry points on the current CUDA stream)."""
    global _torch_ops_loaded
    import torch
    if not _torch_ops_loaded:
        load_library()
        if not os.path.exists(TORCH_LIB_PATH):
            raise NativeError("{0} is missing: build it with `python __graft_entry__.py`.".format(TORCH_LIB_PATH))
        torch.ops.load_library(TORCH_LIB_PATH)
        _torch_ops_loaded = True
    return torch.ops.mcd_b200


def check(rc):
    if rc != 0:
        raise NativeError('libmcd_b200 error {0}: {1}'.format(rc, load_library().mcd_last_error().decode()))


def as_double_ptr(array):
    return array.ctypes.data_as(_c_double_p)


_addressof, _c_char = ctypes.addressof, ctypes.c_char


def address(array):
    """Address of a C-contiguous ndarray's data for the raw-address entry points: through the buffer protocol
    (0.7 us) where the array is writable and not empty, ``ndarray.ctypes.data`` (2 us) otherwise."""
    try:
        return _addressof(_c_char.from_buffer(array))
    except (TypeError, ValueError):
        return array.ctypes.data


def contiguous(values):
    return np.ascontiguousarray(values, dtype=np.float64)
