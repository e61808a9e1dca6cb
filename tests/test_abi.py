"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports every symbol that
include/cubature_cuda.h declares, compiles integrands for sm_100a with NVRTC (no GPU needed), reports
the reference's argument errors, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

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
    # the largest dimensions: the scalar kernel's geometry cells no longer fit static shared memory and are compiled out
    for dim in (10, 11, 12):
        ig = cb.Integrand.builtin(cb.Family.GENZ_CORNER_PEAK, dim, 1, [0.3] * dim + [0.5] * dim)
        assert b"cub_gm_scalar" in ig.cubin()


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


def test_struct_size_versioning(cb):
    """cub_options_t / cub_stats_t are versioned by struct_size: the library reads / writes only the bytes the caller
    declares (no GPU needed: the checks come before the device is opened)."""
    import ctypes as C
    L = cb._ffi.lib()
    ig = cb.Integrand.builtin(cb.Family.GAUSS_ISO, 3, 1, [0.5])
    xmin = (C.c_double * 3)(-2, -2, -2)
    xmax = (C.c_double * 3)(2, 2, 2)
    out = (C.c_double * 2)()
    dp = cb._ffi.c_double_p

    def call(opt, st):
        return L.cubature_cuda_integrate(ig._h, 3, C.cast(xmin, dp), C.cast(xmax, dp), None, 1e-4, 0.0, 0, 0,
                                         None if opt is None else C.byref(opt), C.cast(out, dp), None if st is None else C.byref(st))

    # a caller built against an older, shorter cub_stats_t: bytes beyond its struct_size stay untouched
    class Guarded(C.Structure):
        _fields_ = [("st", cb._ffi.CubStats), ("canary", C.c_uint8 * 64)]
    g = Guarded()
    C.memset(C.byref(g), 0xAB, C.sizeof(g))
    small = 24
    g.st.struct_size = small
    rc = call(None, g.st)
    raw = bytes(C.string_at(C.byref(g), C.sizeof(g)))
    assert raw[small:] == b"\xab" * (C.sizeof(g) - small), "bytes beyond struct_size were written"
    assert g.st.struct_size == small
    assert rc in (0, -5)  # CUB_OK on a GPU box, CUB_ERR_NO_DEVICE here
    # struct_size missing / nonsensical: rejected before anything is read or written
    bad = cb._ffi.CubStats()
    bad.struct_size = 0
    assert call(None, bad) == -1  # CUB_ERR_BAD_ARGS
    opt = cb._ffi.CubOptions()
    opt.struct_size = 0
    st = cb._ffi.CubStats()
    st.struct_size = C.sizeof(cb._ffi.CubStats)
    assert call(opt, st) == -1
    assert b"struct_size" in bytes(st.message)
    # a shorter options struct: fields beyond it take their defaults (select = DEVICE), the call goes through
    opt.struct_size = 8
    opt.select = 0
    opt.device = 12345  # beyond the declared size: must not be read
    rc = call(opt, st)
    assert rc in (0, -5), (rc, bytes(st.message))


def test_java_shim_descriptors_match_header():
    """The Panama FFM shim (java/.../Native.java, shipped as source: there is no JVM in the image) against
    include/cubature_cuda.h: every downcall handle names an exported function, and its FunctionDescriptor has the C
    prototype's arity and argument classes (pointer / 32-bit int / 64-bit int / double)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "cubature_cuda.h")).read()
    java = open(os.path.join(root, "java", "de", "labathome", "cubature", "Native.java")).read()
    header_nc = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(cubature_cuda_\w+)\s*\(([^;{]*?)\)\s*;", header_nc):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        protos[name] = (ret, [] if args in ("", "void") else [a.strip() for a in args.split(",")])

    def c_class(decl):
        decl = decl.strip()
        if "*" in decl:
            return "ADDRESS"
        base = re.sub(r"\b(const|unsigned|signed)\b", "", decl)
        base = re.sub(r"\b[a-z_]\w*$", "", base.strip()).strip() or decl  # drop the parameter name
        if re.search(r"\bdouble\b", decl):
            return "JAVA_DOUBLE"
        if re.search(r"\b(int64_t|size_t|long long|uint64_t)\b", decl):
            return "JAVA_LONG"
        if re.search(r"\b(int|int32_t|uint32_t)\b", decl):
            return "JAVA_INT"
        raise AssertionError("unclassified C type: %r" % decl)

    handles = re.findall(r'h\("(cubature_cuda_\w+)",\s*FunctionDescriptor\.(ofVoid|of)\(([^;]*?)\)\);', java, flags=re.S)
    assert len(handles) >= 5
    for name, kind, layouts in handles:
        assert name in protos, "%s is not declared in include/cubature_cuda.h" % name
        ret, args = protos[name]
        lay = [t.strip() for t in layouts.replace("\n", " ").split(",") if t.strip()]
        if kind == "of":
            assert lay[0] == c_class(ret + " x" if "*" not in ret else ret), (name, "return", lay[0], ret)
            lay = lay[1:]
        else:
            assert ret == "void", (name, ret)
        assert len(lay) == len(args), (name, lay, args)
        for j, (l, a) in enumerate(zip(lay, args)):
            assert l == c_class(a), (name, j, l, a)


def _ckpt_mix(h, data):
    """checkpoint.cpp: ckpt_mix (word-wise multiply-xorshift)."""
    M = (1 << 64) - 1
    n = len(data)
    i = 0
    while i + 8 <= n:
        w = int.from_bytes(data[i:i + 8], "little")
        h = ((h ^ w) * 0x9e3779b97f4a7c15) & M
        h ^= h >> 29
        i += 8
    tail = int.from_bytes(data[i:], "little") if i < n else 0
    h = ((h ^ tail ^ n) * 0xbf58476d1ce4e5b9) & M
    return h ^ (h >> 31)


def test_checkpoint_readers_on_a_hand_written_file(cb, tmp_path):
    """Pool checkpoint format (checkpoint.h): a file assembled here byte by byte is read back by the host-only readers
    cubature_cuda_checkpoint_info / _regions; a damaged header is rejected."""
    import struct
    dim, fdim, n = 2, 3, 5
    rng = np.random.default_rng(3)
    centers, halfw = rng.random((dim, n)), rng.random((dim, n))
    val, err = rng.standard_normal((fdim, n)), rng.random((fdim, n))
    errmax = err.max(axis=0)
    split = np.arange(n, dtype=np.int32) % dim
    payload = b"".join(a.astype("<f8").tobytes() for a in (centers, halfw, val, err, errmax)) + split.astype("<i4").tobytes()
    head = b"CUBCKPT1" + struct.pack("<4i3q", 1, dim, fdim, 4, n, 12345, 7) + struct.pack("<12i", *([1, 0] + [0] * 10))
    head += struct.pack("<12d", *([0.5] + [0.0] * 11)) + struct.pack("<12d", *([0.0] * 12)) + struct.pack("<12d", *([1.0] * 12))
    head += struct.pack("<2Q", 0xabcdef, 0x1234)
    head += struct.pack("<Q", _ckpt_mix(0x13198a2e03707344, head))
    path = tmp_path / "pool.ckpt"
    path.write_bytes(head + payload)
    info = cb.checkpoint_info(str(path))
    assert (info["dim"], info["fdim"], info["n_regions"], info["num_eval"], info["iterations"]) == (dim, fdim, n, 12345, 7)
    assert info["transform"] == [1, 0] and info["shift"][0] == 0.5
    c, h, v, e, sd = cb.checkpoint_regions(str(path), 1, 3)
    assert np.array_equal(c, centers.T[1:4]) and np.array_equal(h, halfw.T[1:4])
    assert np.array_equal(v, val.T[1:4]) and np.array_equal(e, err.T[1:4]) and np.array_equal(sd, split[1:4])
    # tools/eval_positions.py: the rule's points of every checkpointed box (examples/PlotEvalPositions.java:28-54)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import eval_positions
    pts, reg = eval_positions.eval_positions(str(path))
    assert pts.shape == (n * cb.points_per_region(dim), dim) and reg[0] == 0 and reg[-1] == n - 1
    for i in range(n):
        q = pts[reg == i]
        assert np.all(np.abs(q - centers.T[i]) <= halfw.T[i] * (1 + 1e-15)) and np.array_equal(q[0], centers.T[i])
    bad = bytearray(head + payload)
    bad[20] ^= 1
    path.write_bytes(bytes(bad))
    with pytest.raises(Exception):
        cb.checkpoint_info(str(path))
