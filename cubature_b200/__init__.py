"""cubature_b200 -- host-side mirror of the reference's API for the h-adaptive integration loop.

The reference (jonathanschilling/Cubature, Java) exposes one entry point,
``Cubature.integrate(integrand, xmin, xmax, relTol, absTol, norm, maxEval) -> double[2][fdim]``
(Cubature.java:108-170) and one enum ``CubatureError`` (CubatureError.java:9-29).  This package keeps
those names, argument meanings and error behaviour (a ``RuntimeError`` where the reference throws a
``RuntimeException``) and forwards to the C ABI of ``libcubature_cuda.so`` -- the same ABI the Java
Panama-FFM shim in ``java/`` binds.  The integrand is a device plugin (``Integrand.builtin`` /
``Integrand.from_source``) instead of a host lambda: it has to run on the GPU.

No CPU path exists here: without the CUDA library or without a GPU every call raises.
"""
import ctypes as C
import enum
import os

import numpy as np

from . import _ffi


class CubatureError(enum.IntEnum):
    """Error norms, ordinals as in the reference (CubatureError.java:9-29)."""
    INDIVIDUAL = 0
    PAIRED = 1
    L2 = 2
    L1 = 3
    LINF = 4


Error = CubatureError  # the README's name for it (README.md:76-81)


class Family(enum.IntEnum):
    GENZ_OSCILLATORY = 1
    GENZ_PRODUCT_PEAK = 2
    GENZ_CORNER_PEAK = 3
    GENZ_GAUSSIAN = 4
    GENZ_C0 = 5
    GENZ_DISCONTINUOUS = 6
    GAUSS_ISO = 10
    REFTEST = 20
    OSC_VEC = 30
    CEXP = 31


class Select(enum.IntEnum):
    DEVICE = 0
    HEAP_EXACT = 1


class Transform(enum.IntEnum):
    NONE = 0
    SEMI_INF_UP = 1
    SEMI_INF_DOWN = 2
    INF = 3


def _dp(a):
    return a.ctypes.data_as(_ffi.c_double_p)


def _check(status, detail=b""):
    if status != 0:
        msg = _ffi.lib().cubature_cuda_strerror(status).decode()
        if detail:
            msg += ": " + (detail.decode(errors="replace") if isinstance(detail, bytes) else str(detail))
        raise RuntimeError("cubature_cuda: %s (status %d)" % (msg, status))


def points_per_region(dim):
    n = _ffi.lib().cubature_cuda_points_per_region(dim)
    if n < 0:
        _check(int(n))
    return int(n)


def rule_points(center, halfwidth):
    """The rule points of one box in the reference's order -> [P][dim] (for plotting evaluation positions)."""
    c = np.ascontiguousarray(center, dtype=np.float64)
    h = np.ascontiguousarray(halfwidth, dtype=np.float64)
    out = np.zeros((points_per_region(len(c)), len(c)))
    _check(_ffi.lib().cubature_cuda_rule_points(len(c), _dp(c), _dp(h), _dp(out)))
    return out


class Integrand:
    """A device-side integrand plugin: dim inputs -> fdim outputs (Cubature.java:98-99, moved to the GPU)."""

    def __init__(self, handle, dim, fdim):
        self._h = handle
        self.dim = dim
        self.fdim = fdim

    @classmethod
    def builtin(cls, family, dim, fdim=1, params=()):
        L = _ffi.lib()
        p = np.ascontiguousarray(params, dtype=np.float64)
        h = C.c_void_p()
        st = L.cubature_cuda_integrand_builtin(int(family), dim, fdim, _dp(p) if p.size else None, p.size, C.byref(h))
        _check(st)
        return cls(h, dim, fdim)

    @classmethod
    def from_source(cls, cuda_src, entry, dim, fdim=1, params=()):
        """CUDA C source defining ``__device__ void entry(const double* x, double* f, const double* p)``,
        compiled for sm_100a with NVRTC (needs no GPU)."""
        L = _ffi.lib()
        p = np.ascontiguousarray(params, dtype=np.float64)
        h = C.c_void_p()
        log = C.create_string_buffer(16384)
        st = L.cubature_cuda_integrand_nvrtc(cuda_src.encode(), entry.encode(), dim, fdim, _dp(p) if p.size else None, p.size,
                                             C.byref(h), log, len(log))
        _check(st, log.value)
        return cls(h, dim, fdim)

    def cubin(self, has_transform=False):
        L = _ffi.lib()
        n = L.cubature_cuda_integrand_cubin(self._h, int(has_transform), None, 0)
        if n < 0:
            _check(int(n))
        buf = C.create_string_buffer(int(n))
        L.cubature_cuda_integrand_cubin(self._h, int(has_transform), buf, int(n))
        return buf.raw

    def eval_points(self, x):
        """x[dim][n] -> f[fdim][n] on the GPU (integrand alone)."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.dim, -1)
        n = x.shape[1]
        out = np.zeros((self.fdim, n))
        _check(_ffi.lib().cubature_cuda_eval_points(self._h, self.dim, n, _dp(x), _dp(out)))
        return out

    def eval_regions(self, centers, halfwidths, transform=None, shift=None):
        """Rule.evalError on explicit boxes -> (val[n][fdim], err[n][fdim], split_dim[n], stats)."""
        c = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, self.dim)
        h = np.ascontiguousarray(halfwidths, dtype=np.float64).reshape(-1, self.dim)
        n = c.shape[0]
        val = np.zeros((n, self.fdim))
        err = np.zeros((n, self.fdim))
        sd = np.zeros(n, dtype=np.int32)
        tr = None if transform is None else np.ascontiguousarray(transform, dtype=np.int32)
        sh = None if shift is None else np.ascontiguousarray(shift, dtype=np.float64)
        st = _ffi.CubStats()
        st.struct_size = C.sizeof(_ffi.CubStats)
        rc = _ffi.lib().cubature_cuda_eval_regions(self._h, self.dim, None if tr is None else tr.ctypes.data_as(_ffi.c_int_p),
                                                   None if sh is None else _dp(sh), n, _dp(c), _dp(h), _dp(val), _dp(err),
                                                   sd.ctypes.data_as(_ffi.c_int32_p), C.byref(st))
        _check(rc, st.message)
        return val, err, sd, st

    def close(self):
        if self._h:
            _ffi.lib().cubature_cuda_integrand_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Stats:
    def __init__(self, st):
        for name, _ in _ffi.CubStats._fields_:
            v = getattr(st, name)
            setattr(self, name, v.decode(errors="replace") if isinstance(v, bytes) else v)

    def __repr__(self):
        return "Stats(%s)" % ", ".join("%s=%r" % kv for kv in self.__dict__.items() if kv[0] != "struct_size")


class Trace:
    def __init__(self, batches, centers, halfwidths, split_dim, val, err):
        self.batches, self.centers, self.halfwidths, self.split_dim, self.val, self.err = batches, centers, halfwidths, split_dim, val, err

    def split_hist(self, dim):
        return np.bincount(self.split_dim, minlength=dim).tolist()


class Comm:
    """One rank of a multi-GPU communicator (NCCL over NVLink)."""

    def __init__(self, handle, nranks, rank):
        self._h, self.nranks, self.rank = handle, nranks, rank

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        _check(_ffi.lib().cubature_cuda_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def init(cls, unique_id, nranks, rank, device=-1):
        h = C.c_void_p()
        _check(_ffi.lib().cubature_cuda_comm_init(unique_id, nranks, rank, device, C.byref(h)))
        return cls(h, nranks, rank)

    @classmethod
    def init_local(cls, group_key, nranks, rank, device=-1):
        """Ranks as threads of one process (exchange through host memory)."""
        h = C.c_void_p()
        _check(_ffi.lib().cubature_cuda_comm_init_local(group_key, nranks, rank, device, C.byref(h)))
        return cls(h, nranks, rank)

    def close(self):
        if self._h:
            _ffi.lib().cubature_cuda_comm_free(self._h)
            self._h = None


class Cubature:
    """Mirror of de.labathome.cubature.Cubature (Cubature.java:108-170)."""

    @staticmethod
    def integrate(integrand, xmin, xmax, relTol, absTol, norm=CubatureError.INDIVIDUAL, maxEval=0, **kw):
        """-> numpy [2][fdim]: row 0 values, row 1 error estimates (Cubature.java:169)."""
        return Cubature.integrate_ex(integrand, xmin, xmax, relTol, absTol, norm, maxEval, **kw)[0]

    @staticmethod
    def integrate_ex(integrand, xmin, xmax, relTol, absTol, norm=CubatureError.INDIVIDUAL, maxEval=0, transform=None,
                     select=Select.DEVICE, device=-1, pool_capacity=0, trace_regions=0, trace_batches=0, comm=None,
                     shard_min_regions=0, strict_all=False, checkpoint_save=None, checkpoint_load=None):
        """-> (result[2][fdim], Stats, Trace or None)."""
        L = _ffi.lib()
        if xmin is None or len(xmin) == 0:
            raise RuntimeError("xmin cannot be null or empty")  # Cubature.java:119-123
        if xmax is None or len(xmax) != len(xmin):
            raise RuntimeError("xmin and xmax must have the same length but have %d and %d" %
                               (len(xmin), 0 if xmax is None else len(xmax)))  # Cubature.java:125-128
        dim = len(xmin)
        xmin = np.ascontiguousarray(xmin, dtype=np.float64)
        xmax = np.ascontiguousarray(xmax, dtype=np.float64)
        fdim = integrand.fdim
        out = np.zeros((2, fdim))
        opt = _ffi.CubOptions()
        opt.struct_size = C.sizeof(_ffi.CubOptions)
        opt.select = int(select)
        opt.device = device
        opt.pool_capacity = pool_capacity
        opt.comm = comm._h if comm is not None else None
        opt.shard_min_regions = shard_min_regions
        opt.strict_all = 1 if strict_all else 0
        opt.checkpoint_save = None if checkpoint_save is None else os.fsencode(checkpoint_save)
        opt.checkpoint_load = None if checkpoint_load is None else os.fsencode(checkpoint_load)
        tr = None
        keep = []
        if trace_regions or trace_batches:
            tr = _ffi.CubTrace()
            nb = max(int(trace_batches), 1)
            nr = max(int(trace_regions), 1)
            batches = np.zeros(nb, dtype=np.int64)
            centers = np.zeros((nr, dim))
            halfw = np.zeros((nr, dim))
            sd = np.zeros(nr, dtype=np.int32)
            rv = np.zeros((nr, fdim))
            re_ = np.zeros((nr, fdim))
            keep = [batches, centers, halfw, sd, rv, re_]
            tr.batch_cap = nb
            tr.batches = batches.ctypes.data_as(_ffi.c_int64_p)
            tr.region_cap = nr if trace_regions else 0
            tr.centers, tr.halfwidths = _dp(centers), _dp(halfw)
            tr.split_dim = sd.ctypes.data_as(_ffi.c_int32_p)
            tr.val, tr.err = _dp(rv), _dp(re_)
            opt.trace = C.pointer(tr)
        st = _ffi.CubStats()
        st.struct_size = C.sizeof(_ffi.CubStats)
        trf = None if transform is None else np.ascontiguousarray([int(t) for t in transform], dtype=np.int32)
        rc = L.cubature_cuda_integrate(integrand._h, dim, _dp(xmin), _dp(xmax), None if trf is None else trf.ctypes.data_as(_ffi.c_int_p),
                                       float(relTol), float(absTol), int(norm), int(maxEval), C.byref(opt), _dp(out), C.byref(st))
        _check(rc, st.message)
        trace = None
        if tr is not None:
            nb = min(int(tr.n_batches), len(keep[0]))
            nr = min(int(tr.n_regions), len(keep[3])) if trace_regions else 0
            trace = Trace(keep[0][:nb].tolist(), keep[1][:nr], keep[2][:nr], keep[3][:nr], keep[4][:nr], keep[5][:nr])
            trace.n_batches, trace.n_regions = int(tr.n_batches), int(tr.n_regions)
        return out, Stats(st), trace


def checkpoint_info(path):
    """Header of a pool checkpoint (no GPU needed) -> dict."""
    info = _ffi.CubCheckpointInfo()
    _check(_ffi.lib().cubature_cuda_checkpoint_info(os.fsencode(path), C.byref(info)))
    return {"dim": info.dim, "fdim": info.fdim, "family": info.family, "n_regions": info.n_regions, "num_eval": info.num_eval,
            "iterations": info.iterations, "transform": list(info.transform)[:info.dim], "shift": list(info.shift)[:info.dim]}


def checkpoint_regions(path, first=0, count=None):
    """Regions [first, first + count) of a pool checkpoint (no GPU needed) -> (centers, halfwidths, val, err, split_dim)."""
    info = checkpoint_info(path)
    n = info["n_regions"] - first if count is None else count
    n = max(0, min(n, info["n_regions"] - first))
    c, h = np.zeros((n, info["dim"])), np.zeros((n, info["dim"]))
    v, e = np.zeros((n, info["fdim"])), np.zeros((n, info["fdim"]))
    sd = np.zeros(n, dtype=np.int32)
    got = _ffi.lib().cubature_cuda_checkpoint_regions(os.fsencode(path), int(first), int(n), _dp(c), _dp(h), _dp(v), _dp(e),
                                                      sd.ctypes.data_as(_ffi.c_int32_p))
    if got < 0:
        _check(int(got))
    return c[:got], h[:got], v[:got], e[:got], sd[:got]


def fp64_peak_tflops(device=-1):
    t, ms = C.c_double(), C.c_double()
    _check(_ffi.lib().cubature_cuda_fp64_peak(device, C.byref(t), C.byref(ms)))
    return t.value


def fp64_peak_sustained_tflops(device=-1, window_ms=400.0):
    """(burst, sustained) FP64 FMA peak: the DFMA probe launched back to back for window_ms (power-capped clocks)."""
    b, s = C.c_double(), C.c_double()
    _check(_ffi.lib().cubature_cuda_fp64_peak_sustained(device, float(window_ms), C.byref(b), C.byref(s)))
    return b.value, s.value


def bench_rule(integrand, xmin, xmax, n_regions, repeats=5):
    """Rule kernel alone on n_regions boxes resident in HBM -> (ms per launch, checksum)."""
    xmin = np.ascontiguousarray(xmin, dtype=np.float64)
    xmax = np.ascontiguousarray(xmax, dtype=np.float64)
    ms, cs = C.c_double(), C.c_double()
    _check(_ffi.lib().cubature_cuda_bench_rule(integrand._h, len(xmin), _dp(xmin), _dp(xmax), int(n_regions), int(repeats),
                                               C.byref(ms), C.byref(cs)))
    return ms.value, cs.value
