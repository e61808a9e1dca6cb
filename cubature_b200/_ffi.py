"""ctypes binding of include/cubature_cuda.h (the C ABI of libcubature_cuda.so).

There is no fallback: if the shared library (built in-tree by cubature_b200/csrc/Makefile or
__graft_entry__.build()) is missing, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcubature_cuda.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class CubTrace(C.Structure):
    _fields_ = [
        ("batch_cap", C.c_int64),
        ("batches", c_int64_p),
        ("n_batches", C.c_int64),
        ("region_cap", C.c_int64),
        ("centers", c_double_p),
        ("halfwidths", c_double_p),
        ("split_dim", c_int32_p),
        ("val", c_double_p),
        ("err", c_double_p),
        ("n_regions", C.c_int64),
    ]


class CubOptions(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("select", C.c_int32),
        ("device", C.c_int32),
        ("strict_all", C.c_int32),
        ("pool_capacity", C.c_int64),
        ("trace", C.POINTER(CubTrace)),
        ("comm", C.c_void_p),
        ("shard_min_regions", C.c_int64),
        ("checkpoint_save", C.c_char_p),
        ("checkpoint_load", C.c_char_p),
    ]


class CubCheckpointInfo(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("fdim", C.c_int32),
        ("family", C.c_int32),
        ("reserved", C.c_int32),
        ("n_regions", C.c_int64),
        ("num_eval", C.c_int64),
        ("iterations", C.c_int64),
        ("transform", C.c_int32 * 12),
        ("shift", C.c_double * 12),
    ]


class CubStats(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("converged", C.c_int32),
        ("num_eval", C.c_int64),
        ("num_regions", C.c_int64),
        ("iterations", C.c_int64),
        ("regions_evaluated", C.c_int64),
        ("ambiguous_decisions", C.c_int64),
        ("gpu_launches", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("device_ms", C.c_double),
        ("rule_ms", C.c_double),
        ("wall_ms", C.c_double),
        ("jit_ms", C.c_double),
        ("message", C.c_char * 256),
        ("rebalances", C.c_int64),
    ]


# every symbol include/cubature_cuda.h declares: (restype, argtypes)
SIGNATURES = {
    "cubature_cuda_version": (C.c_int, []),
    "cubature_cuda_strerror": (C.c_char_p, [C.c_int]),
    "cubature_cuda_points_per_region": (C.c_int64, [C.c_int]),
    "cubature_cuda_converged": (C.c_int, [C.c_int, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int]),
    "cubature_cuda_rule_points": (C.c_int, [C.c_int, c_double_p, c_double_p, c_double_p]),
    "cubature_cuda_checkpoint_info": (C.c_int, [C.c_char_p, C.POINTER(CubCheckpointInfo)]),
    "cubature_cuda_checkpoint_regions": (C.c_int64, [C.c_char_p, C.c_int64, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "cubature_cuda_integrand_builtin": (C.c_int, [C.c_int, C.c_int, C.c_int, c_double_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "cubature_cuda_integrand_nvrtc": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, c_double_p, C.c_size_t,
                                                C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]),
    "cubature_cuda_integrand_free": (None, [C.c_void_p]),
    "cubature_cuda_integrand_cubin": (C.c_int64, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "cubature_cuda_integrate": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_int_p, C.c_double, C.c_double, C.c_int,
                                          C.c_int64, C.POINTER(CubOptions), c_double_p, C.POINTER(CubStats)]),
    "cubature_cuda_eval_regions": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_double_p, C.c_int64, c_double_p, c_double_p,
                                             c_double_p, c_double_p, c_int32_p, C.POINTER(CubStats)]),
    "cubature_cuda_eval_points": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_double_p, c_double_p]),
    "cubature_cuda_bench_rule": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_int64, C.c_int, c_double_p, c_double_p]),
    "cubature_cuda_release_workspace": (None, []),
    "cubature_cuda_fp64_peak": (C.c_int, [C.c_int, c_double_p, c_double_p]),
    "cubature_cuda_fp64_peak_sustained": (C.c_int, [C.c_int, C.c_double, c_double_p, c_double_p]),
    "cubature_cuda_comm_unique_id": (C.c_int, [C.c_void_p]),
    "cubature_cuda_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "cubature_cuda_shard_count": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "cubature_cuda_tie_share": (C.c_int64, [C.c_int64, c_int64_p, C.c_int, C.c_int]),
    "cubature_cuda_comm_init_local": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "cubature_cuda_comm_free": (None, [C.c_void_p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libcubature_cuda.so is not built: run `make -C cubature_b200/csrc` or __graft_entry__.build() "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
