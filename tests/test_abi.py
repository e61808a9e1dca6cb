"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports every symbol that
include/cubature_cuda.h declares, compiles integrands for sm_100a with NVRTC (no GPU needed), reports
the reference's argument errors, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAN = float("nan")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cubature_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cubature_cuda_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported(cb):
    L = cb._ffi.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n
        assert n in cb._ffi.SIGNATURES, "binding missing for " + n
    assert L.cubature_cuda_version() // 1000 == 1


def test_library_is_self_contained(cb):
    out = subprocess.check_output(["ldd", cb._ffi.LIB_PATH]).decode()
    for forbidden in ("libcudart", "libnvrtc", "libnccl", "libcuda.so", "liboracle", "libtorch"):
        assert forbidden not in out, out  # cudart is static; NVRTC / NCCL are dlopen'ed on demand


def test_no_oracle_in_product_tree():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cubature_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")) and "embedded_sources" not in f:
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_points_per_region(cb):
    assert [cb.points_per_region(d) for d in (1, 2, 3, 4, 5, 6)] == [15, 17, 33, 57, 93, 149]  # README.md:147-153
    with pytest.raises(RuntimeError):
        cb.points_per_region(40)


def test_strerror(cb):
    L = cb._ffi.lib()
    assert L.cubature_cuda_strerror(0) == b"ok"
    assert b"NaN" in L.cubature_cuda_strerror(-2)


def test_nvrtc_compiles_builtin_without_gpu(cb):
    ig = cb.Integrand.builtin(cb.Family.GENZ_CORNER_PEAK, 6, 1, list(np.linspace(0.5, 1.5, 6)) + [0.5] * 6)
    cubin = ig.cubin()
    assert cubin[:4] == b"\x7fELF" and b"cub_gm_scalar" in cubin and b"cub_rule_staged" in cubin
    ig1 = cb.Integrand.builtin(cb.Family.OSC_VEC, 1, 64)
    assert b"cub_gk_separable" in ig1.cubin()


def test_nvrtc_user_source_and_error_log(cb):
    src = """
    __device__ void myf(const double* x, double* f, const double* p) {
        double s = 0.0;
        for (int i = 0; i < CUB_DIM; ++i) s = dm_fma(p[i], x[i], s);
        f[0] = dm_exp(-s);
        f[1] = dm_cos(s);
    }"""
    ig = cb.Integrand.from_source(src, "myf", 3, 2, [1.0, 2.0, 3.0])
    assert ig.cubin(True)[:4] == b"\x7fELF"
    with pytest.raises(RuntimeError) as e:
        cb.Integrand.from_source("__device__ void bad(const double* x, double* f, const double* p) { f[0] = nope(x[0]); }",
                                 "bad", 2, 1)
    assert "nope" in str(e.value)


def test_bad_arguments_mirror_reference_exceptions(cb):
    ig = cb.Integrand.builtin(cb.Family.GAUSS_ISO, 3, 1, [0.5])
    with pytest.raises(RuntimeError, match="xmin cannot be null or empty"):      # Cubature.java:119-123
        cb.Cubature.integrate(ig, [], [], 1e-4, 0.0)
    with pytest.raises(RuntimeError, match="same length"):                        # Cubature.java:125-128
        cb.Cubature.integrate(ig, [0, 0, 0], [1, 1], 1e-4, 0.0)
    with pytest.raises(RuntimeError, match="NaN"):                                # Rule.java:164-167
        cb.Cubature.integrate(ig, [0] * 3, [1] * 3, NAN, NAN)
    with pytest.raises(RuntimeError):                                             # dim mismatch with the plugin
        cb.Cubature.integrate(ig, [0] * 2, [1] * 2, 1e-3, NAN)
    with pytest.raises(RuntimeError):                                             # Rule.java:273-276
        cb.Cubature.integrate(ig, [0] * 3, [1] * 3, 1e-3, NAN, norm=7)
    with pytest.raises(RuntimeError):
        cb.Integrand.builtin(99, 3, 1, [0.5])
    with pytest.raises(RuntimeError):
        cb.Integrand.builtin(cb.Family.GENZ_GAUSSIAN, 3, 1, [0.5])               # wrong parameter count
    with pytest.raises(RuntimeError):
        cb.Integrand.builtin(cb.Family.GAUSS_ISO, 40, 1, [0.5])                  # Rule75GenzMalik.java:35-42


def test_no_cpu_fallback(cb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ig = cb.Integrand.builtin(cb.Family.GAUSS_ISO, 3, 1, [0.5])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        cb.Cubature.integrate(ig, [-2] * 3, [2] * 3, 1e-4, 0.0)
    # every other compute entry point fails the same way: rule evaluation, rule benchmark, FP64 probes
    with pytest.raises(RuntimeError, match="no CUDA device"):
        ig.eval_regions([[0.0] * 3], [[1.0] * 3])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        cb.bench_rule(ig, [-2] * 3, [2] * 3, 16, 1)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        cb.fp64_peak_tflops()
    with pytest.raises(RuntimeError, match="no CUDA device"):
        cb.fp64_peak_sustained_tflops(-1, 10.0)


def test_convergence_predicate_matches_oracle(cb, oracle):
    """Host logic: Rule.converged (Rule.java:162-279) in the product vs the restatement, random vectors."""
    L = cb._ffi.lib()
    rng = np.random.default_rng(11)
    dp = cb._ffi.c_double_p
    n = 0
    for trial in range(4000):
        f = int(rng.integers(1, 9))
        val = rng.normal(0, 1, f) * 10.0 ** rng.integers(-3, 3)
        err = np.abs(rng.normal(0, 1, f)) * 10.0 ** rng.integers(-8, 1)
        if trial % 7 == 0:
            val[rng.integers(0, f)] = 0.0
        if trial % 11 == 0:
            err[rng.integers(0, f)] = 0.0
        ab = [NAN, 0.0, 1e-4, 1e-2][trial % 4]
        rel = [1e-3, NAN, 1e-6, 1e-1][(trial // 4) % 4]
        if ab != ab and rel != rel:
            continue
        for norm in (0, 1, 2, 3, 4, 8, 9):  # 8, 9: the strict-ALL extension of INDIVIDUAL / PAIRED (section 8f-3)
            a = L.cubature_cuda_converged(f, val.ctypes.data_as(dp), err.ctypes.data_as(dp), ab, rel, norm)
            b = oracle.converged(val, err, ab, rel, norm)
            assert a == b, (f, norm, ab, rel, val, err)
            n += 1
    assert n > 10000
    v = np.zeros(2)
    assert L.cubature_cuda_converged(2, v.ctypes.data_as(dp), v.ctypes.data_as(dp), 1e-3, NAN, 10) == -4  # L2|strict: no such norm
    assert L.cubature_cuda_converged(2, v.ctypes.data_as(dp), v.ctypes.data_as(dp), NAN, NAN, 0) == -2


def test_rule_points_follow_the_reference_order(cb):
    """Point order of SURVEY.md Appendix A (Rule75GenzMalik.java:237-326, RuleGaussKronrod_1d.java:60-84)."""
    pts = cb.rule_points([0.0, 0.0, 0.0], [1.0, 2.0, 4.0])
    l2, l4, l5 = 0.3585685828003180919906451539079374954541, 0.9486832980505137995996680633298155601160, 0.6882472016116852977216287342936235251269
    assert pts.shape == (33, 3)
    assert pts[0].tolist() == [0, 0, 0]
    assert pts[1].tolist() == [-l2, 0, 0] and pts[2].tolist() == [l2, 0, 0] and pts[3].tolist() == [-l4, 0, 0] and pts[4].tolist() == [l4, 0, 0]
    assert pts[5].tolist() == [0, -2 * l2, 0]
    # (lambda4, lambda4, 0) orbit: (-,-) (+,-) (+,+) (-,+) for the pair (0,1)
    assert [p[:2].tolist() for p in pts[13:17]] == [[-l4, -2 * l4], [l4, -2 * l4], [l4, 2 * l4], [-l4, 2 * l4]]
    # Gray code: +++, -++, --+, +-+, +--, ---, -+-, ++-
    signs = np.sign(pts[25:33]).astype(int).tolist()
    assert signs == [[1, 1, 1], [-1, 1, 1], [-1, -1, 1], [1, -1, 1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1], [1, 1, -1]]
    assert np.allclose(np.abs(pts[25:33]), [l5, 2 * l5, 4 * l5])
    gk = cb.rule_points([0.5], [0.5])
    assert gk.shape == (15, 1) and gk[0, 0] == 0.5 and gk[1, 0] < 0.5 < gk[2, 0]
    assert abs(gk[7, 0] - (0.5 - 0.5 * 0.991455371120812639)) < 1e-16
