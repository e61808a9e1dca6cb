"""ctypes wrapper around liboracle.so (the CPU restatement of the reference, cubature_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (cubature_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INDIVIDUAL, PAIRED, L2, L1, LINF = 0, 1, 2, 3, 4


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "cubature_oracle.cpp")
    hdrs = [os.path.join(_HERE, "..", "cubature_b200", "csrc", h) for h in ("integrands.h", "detmath.h")]
    stale = not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_integrate_builtin.restype = C.c_void_p
        L.orc_integrate_builtin.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, dp, dp, ip, C.c_double, C.c_double,
                                            C.c_int, C.c_longlong, C.c_int]
        L.orc_integrate_user.restype = C.c_void_p
        L.orc_integrate_user.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.c_int, dp, dp, ip, C.c_double, C.c_double,
                                         C.c_int, C.c_longlong]
        for name in ("orc_status", "orc_converged"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("orc_num_eval", "orc_num_regions", "orc_num_iterations"):
            getattr(L, name).restype = C.c_longlong
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_get_result.argtypes = [C.c_void_p, dp]
        L.orc_get_batches.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.orc_get_regions.argtypes = [C.c_void_p, dp, dp, ip, dp, dp, dp]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_exact_sums.argtypes = [C.c_int]
        L.orc_eval_regions.restype = C.c_int
        L.orc_eval_regions.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, ip, dp, C.c_longlong, dp, dp, dp, dp, ip]
        for name in ("orc_dm_exp", "orc_dm_sin", "orc_dm_cos", "orc_dm_log"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double]
        L.orc_dm_pow.restype = C.c_double
        L.orc_dm_pow.argtypes = [C.c_double, C.c_double]
        L.orc_converged_vec.restype = C.c_int
        L.orc_converged_vec.argtypes = [C.c_int, dp, dp, C.c_double, C.c_double, C.c_int]
        L.orc_eval_family_points.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_longlong, dp, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleResult:
    """Everything the restated Rule.cubature produced: [2][fdim] result plus the full trace."""

    def __init__(self, handle, dim, fdim, want_regions=True):
        L = lib()
        self.status = L.orc_status(handle)
        self.num_eval = L.orc_num_eval(handle)
        self.converged = bool(L.orc_converged(handle))
        self.num_regions = L.orc_num_regions(handle)
        self.iterations = L.orc_num_iterations(handle)
        out = np.zeros(2 * fdim)
        L.orc_get_result(handle, _dp(out))
        self.val, self.err = out[:fdim].copy(), out[fdim:].copy()
        b = np.zeros(max(self.iterations, 1), dtype=np.int64)
        L.orc_get_batches(handle, b.ctypes.data_as(C.POINTER(C.c_longlong)))
        self.batches = b[: self.iterations].tolist()
        n = self.num_regions
        if want_regions:
            self.centers = np.zeros((n, dim))
            self.halfwidths = np.zeros((n, dim))
            self.split_dim = np.zeros(n, dtype=np.int32)
            self.region_val = np.zeros((n, fdim))
            self.region_err = np.zeros((n, fdim))
            self.errmax = np.zeros(n)
            L.orc_get_regions(handle, _dp(self.centers), _dp(self.halfwidths), _ip(self.split_dim), _dp(self.region_val),
                              _dp(self.region_err), _dp(self.errmax))
        L.orc_free(handle)

    def split_hist(self, dim):
        return np.bincount(self.split_dim, minlength=dim).tolist()


def set_threads(n):
    """Threads evaluating one integrand batch (the reference's vectorized interface parallelised by the user, README.md:
    129-163); 1 = the reference's own single-threaded behaviour.  Everything else stays serial."""
    lib().orc_set_threads(int(n))


def set_exact_sums(on):
    """Variant (NOT reference behaviour): carry the heap's running sums and the batch-pop residual rounding-free
    (binary128).  Shows where the reference's own accumulated rounding decides a batch size; the device-side selection
    of libcubature_cuda takes exactly these decisions."""
    lib().orc_set_exact_sums(1 if on else 0)


def integrate(family, dim, fdim, params, xmin, xmax, rel_tol, abs_tol, norm=INDIVIDUAL, max_eval=0, transform=None,
              use_libm=False, want_regions=True):
    L = lib()
    params = np.ascontiguousarray(params, dtype=np.float64)
    xmin = np.ascontiguousarray(xmin, dtype=np.float64)
    xmax = np.ascontiguousarray(xmax, dtype=np.float64)
    tr = None if transform is None else np.ascontiguousarray(transform, dtype=np.int32)
    h = L.orc_integrate_builtin(family, dim, fdim, _dp(params), len(params), _dp(xmin), _dp(xmax),
                                None if tr is None else _ip(tr), rel_tol, abs_tol, norm, int(max_eval), int(use_libm))
    if not h:
        raise RuntimeError("oracle: bad arguments")
    return OracleResult(h, dim, fdim, want_regions)


def integrate_user(point_fn_addr, dim, fdim, params, xmin, xmax, rel_tol, abs_tol, norm=INDIVIDUAL, max_eval=0,
                   transform=None):
    """point_fn_addr: address of `void f(const double* x, double* f, const double* p)` compiled by g++."""
    L = lib()
    params = np.ascontiguousarray(params, dtype=np.float64)
    xmin = np.ascontiguousarray(xmin, dtype=np.float64)
    xmax = np.ascontiguousarray(xmax, dtype=np.float64)
    tr = None if transform is None else np.ascontiguousarray(transform, dtype=np.int32)
    h = L.orc_integrate_user(point_fn_addr, dim, fdim, _dp(params), len(params), _dp(xmin), _dp(xmax),
                             None if tr is None else _ip(tr), rel_tol, abs_tol, norm, int(max_eval))
    if not h:
        raise RuntimeError("oracle: bad arguments")
    return OracleResult(h, dim, fdim)


def eval_regions(family, dim, fdim, params, centers, halfwidths, transform=None, shift=None):
    L = lib()
    params = np.ascontiguousarray(params, dtype=np.float64)
    centers = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, dim)
    halfwidths = np.ascontiguousarray(halfwidths, dtype=np.float64).reshape(-1, dim)
    n = centers.shape[0]
    val = np.zeros((n, fdim))
    err = np.zeros((n, fdim))
    sd = np.zeros(n, dtype=np.int32)
    tr = None if transform is None else np.ascontiguousarray(transform, dtype=np.int32)
    sh = None if shift is None else np.ascontiguousarray(shift, dtype=np.float64)
    L.orc_eval_regions(family, dim, fdim, _dp(params), len(params), None if tr is None else _ip(tr),
                       None if sh is None else _dp(sh), n, _dp(centers), _dp(halfwidths), _dp(val), _dp(err), _ip(sd))
    return val, err, sd


def eval_points(family, dim, fdim, params, x):
    """x: [dim][n] -> [fdim][n] integrand values (no rule), for bit comparisons with the device."""
    L = lib()
    params = np.ascontiguousarray(params, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(dim, -1)
    n = x.shape[1]
    out = np.zeros((fdim, n))
    L.orc_eval_family_points(family, dim, fdim, _dp(params), len(params), n, _dp(x), _dp(out))
    return out


def converged(val, err, abs_tol, rel_tol, norm):
    val = np.ascontiguousarray(val, dtype=np.float64)
    err = np.ascontiguousarray(err, dtype=np.float64)
    return lib().orc_converged_vec(len(val), _dp(val), _dp(err), abs_tol, rel_tol, norm)
