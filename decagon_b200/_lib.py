"""ctypes binding of libdecagon_b200.so (C ABI declared in include/decagon_b200.h).

There is no fallback: if the shared library is missing or no CUDA device is visible the
product raises -- nothing in ``decagon_b200`` routes through numpy / torch arithmetic.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DGN_LIB_PATH') or os.path.join(_HERE, 'libdecagon_b200.so')  # DGN_LIB_PATH: kernel experiments

DEC_KINDS = {'innerproduct': 0, 'distmult': 1, 'bilinear': 2, 'dedicom': 3}
LOSS_KINDS = {'hinge': 0, 'xent': 1}
PARAM_W1, PARAM_W2, PARAM_DEC_GLOBAL, PARAM_DEC_LOCAL = 0, 1, 2, 3
TENSOR_HIDDEN1, TENSOR_EMBEDDINGS, TENSOR_LAYER1_GROUP, TENSOR_LAYER2_GROUP, TENSOR_GRAD_EMBEDDINGS = range(5)

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_graph = ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol include/decagon_b200.h declares
SIGNATURES = {
    'dgn_last_error': (ctypes.c_char_p, []),
    'dgn_version': (ctypes.c_int, []),
    'dgn_device_count': (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    'dgn_csr_from_coo': (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, c_i32p, c_i32p, c_f32p,
                                        c_i32p, c_i32p, c_f32p]),
    'dgn_sampler_thresholds': (ctypes.c_int, [c_f64p, ctypes.c_int32, c_u32p]),
    'dgn_graph_create': (ctypes.c_int, [ctypes.POINTER(c_graph), ctypes.c_int, ctypes.c_int, c_i32p, c_i32p,
                                        ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int, ctypes.c_int]),
    'dgn_graph_destroy': (ctypes.c_int, [c_graph]),
    'dgn_graph_set_relation': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                              c_i32p, c_i32p, c_f32p]),
    'dgn_graph_set_features': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                              c_i32p, c_i32p, c_f32p]),
    'dgn_sampler_set_degrees': (ctypes.c_int, [c_graph, ctypes.c_int, c_f64p, ctypes.c_int32]),
    'dgn_graph_finalize': (ctypes.c_int, [c_graph]),
    'dgn_graph_relation_nnz': (ctypes.c_int, [c_graph, ctypes.c_int, c_i64p]),
    'dgn_graph_get_csr': (ctypes.c_int, [c_graph, ctypes.c_int, c_i32p, c_i32p, c_f32p]),
    'dgn_params_set': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f32p, ctypes.c_int64]),
    'dgn_params_get': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f32p, ctypes.c_int64]),
    'dgn_params_count': (ctypes.c_int, [c_graph, c_i64p]),
    'dgn_optimizer_reset': (ctypes.c_int, [c_graph, ctypes.c_float, ctypes.c_float, ctypes.c_float]),
    'dgn_encoder_forward': (ctypes.c_int, [c_graph, ctypes.c_float, ctypes.c_uint64, ctypes.c_uint32]),
    'dgn_train_step': (ctypes.c_int, [c_graph, ctypes.c_int, c_i32p, ctypes.c_int32, c_i64p, ctypes.c_int,
                                      ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_uint64,
                                      ctypes.c_uint32, ctypes.c_int, c_f32p]),
    'dgn_last_batch_outputs': (ctypes.c_int, [c_graph, c_f32p, c_f32p, c_i64p, ctypes.c_int32]),
    'dgn_grads_get': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f32p, ctypes.c_int64]),
    'dgn_keep_gradients': (ctypes.c_int, [c_graph, ctypes.c_int]),
    'dgn_predict_all_pairs': (ctypes.c_int, [c_graph, ctypes.c_int, c_f32p]),
    'dgn_predict_relations_dev': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    'dgn_predict_edges': (ctypes.c_int, [c_graph, ctypes.c_int, c_i32p, ctypes.c_int32, ctypes.c_int, c_f32p]),
    'dgn_evaluate_edges': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_int64, c_i32p, c_i32p,
                                          ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, c_f32p, c_f64p, c_f64p]),
    'dgn_rank_edges': (ctypes.c_int, [c_graph, ctypes.c_int, c_i32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_i32p, c_f32p]),
    'dgn_tensor_get': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_int, c_f32p, ctypes.c_int64]),
    'dgn_tensor_set': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_int, c_f32p, ctypes.c_int64]),
    'dgn_relation_matrices': (ctypes.c_int, [c_graph, ctypes.c_int, c_f32p, c_f32p]),
    'dgn_sync': (ctypes.c_int, [c_graph]),
    'dgn_timing_enable': (ctypes.c_int, [c_graph, ctypes.c_int]),
    'dgn_timing_reset': (ctypes.c_int, [c_graph]),
    'dgn_timing_get': (ctypes.c_int, [c_graph, ctypes.c_char_p, c_f64p, c_i64p]),
    'dgn_timeline_get': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                        c_f64p, c_f64p]),
    'dgn_launch_count': (ctypes.c_int, [c_graph, c_i64p]),
    'dgn_counters': (ctypes.c_int, [c_graph, c_i64p, c_i64p]),
    'dgn_timer_start': (ctypes.c_int, [c_graph]),
    'dgn_timer_stop': (ctypes.c_int, [c_graph, c_f64p]),
    'dgn_memory_bytes': (ctypes.c_int, [c_graph, c_i64p, c_i64p]),
    'dgn_comm_init': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.c_int]),
    'dgn_comm_handle': (ctypes.c_int, [c_graph, ctypes.c_void_p]),
    'dgn_comm_connect': (ctypes.c_int, [c_graph, ctypes.c_void_p]),
    'dgn_relation_owner': (ctypes.c_int, [c_graph, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    'dgn_partition_relations': (ctypes.c_int, [c_i64p, ctypes.c_int32, ctypes.c_int32, c_i32p]),
}


class DecagonB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load (once) and return the shared library; raises when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DecagonB200Error(
            'libdecagon_b200.so is not built (%s). Run `python -m decagon_b200.build` (needs nvcc); '
            'there is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError when the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().dgn_last_error().decode('utf-8', 'replace')
        if rc == -1:
            raise ValueError(msg)
        if rc == -4:
            raise NotImplementedError(msg)
        raise DecagonB200Error('decagon_b200 error %d: %s' % (rc, msg))


def ptr(a, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def device_count():
    n = ctypes.c_int(0)
    check(load().dgn_device_count(ctypes.byref(n)))
    return n.value


def csr_from_coo(n_rows, n_cols, rows, cols, vals):
    """Host-only canonicalisation used by the device path (exposed for CPU tests)."""
    rows, cols, vals = as_i32(rows), as_i32(cols), as_f32(vals)
    nnz = len(vals)
    rowptr = np.empty(n_rows + 1, dtype=np.int32)
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float32)
    check(load().dgn_csr_from_coo(n_rows, n_cols, nnz, ptr(rows, ctypes.c_int32), ptr(cols, ctypes.c_int32),
                                  ptr(vals, ctypes.c_float), ptr(rowptr, ctypes.c_int32), ptr(col, ctypes.c_int32),
                                  ptr(val, ctypes.c_float)))
    return rowptr, col, val


def partition_relations(weights, world):
    """Host-only: owner rank of every relation (longest-processing-time, the library's own rule)."""
    w = np.ascontiguousarray(weights, dtype=np.int64)
    out = np.empty(len(w), dtype=np.int32)
    check(load().dgn_partition_relations(ptr(w, ctypes.c_int64), len(w), int(world), ptr(out, ctypes.c_int32)))
    return out


def sampler_thresholds(degrees):
    d = np.ascontiguousarray(degrees, dtype=np.float64)
    out = np.empty(len(d), dtype=np.uint32)
    check(load().dgn_sampler_thresholds(ptr(d, ctypes.c_double), len(d), ptr(out, ctypes.c_uint32)))
    return out
