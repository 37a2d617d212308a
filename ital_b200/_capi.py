"""ctypes binding of include/ital_b200.h.  There is no fallback: a missing library is an ImportError-like
RuntimeError, a missing GPU surfaces as the library's own ITAL_ECUDA error on the first ital_create."""
import ctypes
import os

import numpy as np

from .build import LIB_PATH

ITAL_F32, ITAL_F64 = 0, 1
RECORD_HEADER = 8

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)
_shard_p = ctypes.c_void_p

# name -> (restype, argtypes): one entry per function declared in include/ital_b200.h
SIGNATURES = {
    'ital_last_error': (ctypes.c_char_p, []),
    'ital_version': (ctypes.c_int, []),
    'ital_create': (ctypes.c_int, [ctypes.POINTER(_shard_p), ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                   ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    'ital_destroy': (ctypes.c_int, [_shard_p]),
    'ital_set_stream': (ctypes.c_int, [_shard_p, ctypes.c_void_p]),
    'ital_reset': (ctypes.c_int, [_shard_p]),
    'ital_record_doubles': (ctypes.c_int64, [_shard_p]),
    'ital_width_cap': (ctypes.c_int64, [_shard_p]),
    'ital_width': (ctypes.c_int64, [_shard_p]),
    'ital_export_points': (ctypes.c_int, [_shard_p, ctypes.c_int, _c_int64_p, _c_double_p]),
    'ital_add_labelled': (ctypes.c_int, [_shard_p, _c_double_p, ctypes.c_double]),
    'ital_add_labelled_many': (ctypes.c_int, [_shard_p, ctypes.c_int, _c_double_p, _c_double_p]),
    'ital_update_labelled': (ctypes.c_int, [_shard_p, ctypes.c_int, _c_int64_p, _c_double_p]),
    'ital_mark_seen': (ctypes.c_int, [_shard_p, ctypes.c_int64, _c_int64_p]),
    'ital_restrict_candidates': (ctypes.c_int, [_shard_p, ctypes.c_int64, _c_int64_p]),
    'ital_restrict_top': (ctypes.c_int, [_shard_p, ctypes.c_int64]),
    'ital_fetch_propose_dev': (ctypes.c_int, [_shard_p, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]),
    'ital_fetch_commit_dev': (ctypes.c_int, [_shard_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    'ital_fetch_result': (ctypes.c_int, [_shard_p, ctypes.c_int, _c_int64_p, _c_double_p]),
    'ital_fetch_begin': (ctypes.c_int, [_shard_p, ctypes.c_double, ctypes.c_double]),
    'ital_fetch_propose': (ctypes.c_int, [_shard_p, ctypes.c_double, ctypes.c_int, _c_double_p]),
    'ital_fetch_commit': (ctypes.c_int, [_shard_p, _c_double_p]),
    'ital_variance_propose': (ctypes.c_int, [_shard_p, ctypes.c_int, ctypes.c_int, _c_double_p]),
    'ital_set_sub_mode': (ctypes.c_int, [_shard_p, ctypes.c_int]),
    'ital_set_clip_cov': (ctypes.c_int, [_shard_p, ctypes.c_double]),
    'ital_fetch_propose_sub': (ctypes.c_int, [_shard_p, ctypes.c_int, ctypes.c_int64, _c_double_p]),
    'ital_fetch_end': (ctypes.c_int, [_shard_p]),
    'ital_fetch': (ctypes.c_int, [_shard_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                  _c_int64_p, _c_double_p]),
    'ital_peer_export': (ctypes.c_int, [_shard_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]),
    'ital_peer_connect': (ctypes.c_int, [_shard_p, ctypes.c_void_p, ctypes.c_int64]),
    'ital_peer_disconnect': (ctypes.c_int, [_shard_p]),
    'ital_peer_slot_doubles': (ctypes.c_int64, [_shard_p]),
    'ital_fetch_peer': (ctypes.c_int, [_shard_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                       _c_int64_p, _c_double_p]),
    'ital_set_lazy_rows': (ctypes.c_int, [_shard_p, ctypes.c_int]),
    'ital_set_bulk_stream': (ctypes.c_int, [_shard_p, ctypes.c_int]),
    'ital_set_label_estimation': (ctypes.c_int, [_shard_p, ctypes.c_int]),
    'ital_set_fused': (ctypes.c_int, [_shard_p, ctypes.c_int]),
    'ital_fused_trace': (ctypes.c_int64, [_shard_p, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int64]),
    'ital_fetch_stats': (ctypes.c_int, [_shard_p, _c_double_p]),
    'ital_last_scores': (ctypes.c_int, [_shard_p, _c_double_p]),
    'ital_rel_mean': (ctypes.c_int, [_shard_p, _c_double_p]),
    'ital_rel_var': (ctypes.c_int, [_shard_p, _c_double_p]),
    'ital_top_results': (ctypes.c_int64, [_shard_p, ctypes.c_int64, _c_int64_p, _c_double_p]),
    'ital_predict': (ctypes.c_int, [_shard_p, _c_double_p, ctypes.c_int64, _c_double_p, _c_double_p]),
    'ital_predict_proj': (ctypes.c_int, [_shard_p, _c_double_p, ctypes.c_int64, _c_double_p, _c_double_p]),
    'ital_profile_enable': (ctypes.c_int, [_shard_p, ctypes.c_int]),
    'ital_profile_read': (ctypes.c_int, [_shard_p, _c_double_p, _c_int64_p, _c_double_p]),
    'ital_launch_count': (ctypes.c_int64, [_shard_p]),
    'ital_transfer_bytes': (ctypes.c_int, [_shard_p, _c_int64_p, _c_int64_p]),
    'ital_snq_nodes': (ctypes.c_int64, [ctypes.c_int, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                        _c_int32_p, _c_double_p]),
    'ital_snq_order': (ctypes.c_int, [ctypes.c_int]),
    'ital_h_table': (ctypes.c_int64, [_c_double_p, ctypes.c_int64]),
    'ital_phi_table': (ctypes.c_int64, [_c_double_p, ctypes.c_int64]),
    'ital_snq_sub': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_double, _c_int64_p,
                                    _c_double_p, _c_double_p, _c_int32_p, _c_double_p]),
    'ital_snq_general': (ctypes.c_int, [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_double, _c_int64_p,
                                        _c_double_p, _c_double_p, _c_int32_p, _c_double_p, _c_int32_p, _c_int32_p]),
}

_lib = None


class ItalError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library and type its entry points.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get('ITAL_B200_LIB', LIB_PATH)     # (override: tuning experiments with variant builds)
        if not os.path.exists(path):
            raise ItalError('%s is missing: build it with `python -m ital_b200.build` '
                            '(or __graft_entry__.build()); this package has no CPU fallback' % path)
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc < 0:
        raise ItalError('ital_b200: %s (code %d)' % (load().ital_last_error().decode(), rc))
    return rc


def dptr(a):
    return a.ctypes.data_as(_c_double_p)


def i64ptr(a):
    return a.ctypes.data_as(_c_int64_p)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)
