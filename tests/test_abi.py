"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol that
include/deepfit_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="deepfit_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"DFB_API\s+[\w\s\*]+?\b(dfb_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ["dfb_grid_create", "dfb_grid_destroy", "dfb_knn", "dfb_patch_pca", "dfb_model_create",
              "dfb_model_forward", "dfb_model_destroy", "dfb_wls_fit", "dfb_principal_curvatures",
              "dfb_last_error_string", "dfb_abi_version", "dfb_device_check"]:
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol():
    from deepfit_b200 import _lib
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.lib_path())
    for s in declared_symbols():
        assert hasattr(raw, s), "missing export %s" % s
    assert lib.dfb_abi_version() == 2
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_library_exports_nothing_undeclared():
    """Every dynamic symbol of the .so is declared in include/: the boundary in deepfit_b200.h, the process-global
    test / profiling hooks (dfb_dbg_*) in deepfit_b200_debug.h and nowhere else."""
    import subprocess
    from deepfit_b200 import _lib
    _lib.load()
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and not l.split()[-1].startswith("_"))
    debug = declared_symbols("deepfit_b200_debug.h")
    assert debug and all(s.startswith("dfb_dbg_") for s in debug)
    assert not [s for s in declared_symbols() if s.startswith("dfb_dbg_")]
    assert exported == sorted(declared_symbols() + debug)


def test_precision_codes():
    """dfb_model_desc.precision: the mode in the low byte, a DFB_MIX_* layer mask above it (f16mix only)."""
    from deepfit_b200 import _lib
    header = open(os.path.join(ROOT, "include", "deepfit_b200.h")).read()
    for name, code in _lib.PRECISION.items():
        assert re.search(r"#define DFB_PRECISION_%s %d\b" % (name.upper(), code), header), name
    for name, bit in _lib.MIX_BITS.items():
        assert re.search(r"#define DFB_MIX_%s %d\b" % (name.upper(), bit), header), name
    assert _lib.precision_code("f16x3") == 3 and _lib.precision_code("f16mix") == 4
    assert _lib.precision_code("f16mix:qstn_max+stn_max") == 4 | (10 << 8)
    for bad in ("f16x3:qstn_max", "f16mix:head", "fp32"):
        with pytest.raises(ValueError):
            _lib.precision_code(bad)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from deepfit_b200 import _lib, ops
    with pytest.raises(_lib.DeepFitB200Error):
        _lib.require_device()
    with pytest.raises(ValueError):
        ops.wls_fit(torch.zeros(1, 3, 32), None, 2)      # CPU tensors are rejected, not silently computed
    from deepfit_b200 import DeepFit as D
    m = D.DeepFit(k=1, num_points=256, use_point_stn=True, use_feat_stn=True, jet_order=3, weight_mode="sigmoid")
    with pytest.raises(ValueError):
        m.forward(torch.zeros(1, 3, 256))


def test_usage_errors_match_reference_conventions():
    import torch
    from deepfit_b200 import DeepFit as D
    # reference: ValueError for an unsupported order (models/DeepFit.py:78) and for curvature with M < 5
    with pytest.raises(ValueError):
        D.fit_Wjet(torch.zeros(1, 3, 32), torch.ones(1, 32), order=7)
    with pytest.raises(ValueError):
        D.compute_principal_curvatures(torch.zeros(4, 3))
    with pytest.raises(NotImplementedError):
        D.DeepFit(k=1, num_points=256, use_point_stn=True, use_feat_stn=True, weight_mode="softmax")
    with pytest.raises(NotImplementedError):
        D.DeepFit(k=1, num_points=256, use_point_stn=True, use_feat_stn=True, arch="3dmfv")


def test_state_dict_contract():
    """125 entries, the reference's key set."""
    from deepfit_b200 import DeepFit as D
    from oracle import deepfit_oracle as O
    m = D.DeepFit(k=1, num_points=256, use_point_stn=True, use_feat_stn=True, jet_order=3, weight_mode="sigmoid")
    sd = O.seeded_state_dict(0)
    assert len(sd) == 125
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd, strict=True)
    layers = D.fold_batchnorm(m.state_dict())
    assert len(layers) == 20
    assert tuple(layers[16][0].shape) == (512, 1088) and tuple(layers[13][0].shape) == (4096, 256)


def test_text_io_is_byte_compatible_with_numpy(tmp_path):
    """the multi-threaded savetxt / loadtxt replacements equal numpy's defaults."""
    import numpy as np
    from deepfit_b200 import tutorial_utils as tu
    rng = np.random.default_rng(0)
    for shape, dtype in (((5000, 3), np.float32), ((7001, 2), np.float64), ((3, 10), np.float32), ((1, 1), np.float64)):
        a = (rng.normal(size=shape) * 10.0 ** rng.integers(-8, 8, size=shape)).astype(dtype)
        a.flat[0] = 0.0
        if a.size > 2:
            a.flat[1] = -1.0
            a.flat[2] = np.float32(1e-40) if dtype == np.float32 else 5e-324
        tu.savetxt(tmp_path / "ours.txt", a)
        np.savetxt(tmp_path / "numpy.txt", a)
        assert open(tmp_path / "ours.txt", "rb").read() == open(tmp_path / "numpy.txt", "rb").read()
    pts = (rng.normal(size=(6000, 3)) * 3).round(6)
    np.savetxt(tmp_path / "cloud.xyz", pts, fmt="%.6f")
    with open(tmp_path / "cloud.xyz", "a") as fh:
        fh.write("# trailing comment\n\n")
    ref = np.loadtxt(tmp_path / "cloud.xyz").astype("float32")
    got = tu.loadtxt_f32(tmp_path / "cloud.xyz")
    assert got.shape == ref.shape and np.array_equal(got, ref)
    np.savetxt(tmp_path / "sci.xyz", pts * 1e-3)           # scientific notation, 18 digits
    assert np.array_equal(tu.loadtxt_f32(tmp_path / "sci.xyz"), np.loadtxt(tmp_path / "sci.xyz").astype("float32"))


def test_loadtxt_exact_parser_cases(tmp_path):
    """loadtxt without strtod on the common path (csrc/textio.cu: 64-bit digit gathering, one exact double operation for
    <= 15 digits, x87 extended precision with a midpoint guard for numpy's own 19-digit lines, strtod for the rest): every
    digit count in both notations, float32 rounding midpoints and their double neighbours written with 17-19 digits
    (where a double off by one ulp would flip the float), decimal strings AT midpoints of two doubles, the syntax numpy
    accepts (comments, blank lines, tabs / commas / CRLF, no trailing newline, signs, nan / inf, subnormals, overlong
    mantissas), the multi-threaded path (> 1 MiB), and the errors."""
    import numpy as np
    from decimal import Decimal, getcontext
    from deepfit_b200 import _lib
    from deepfit_b200 import tutorial_utils as tu
    rng = np.random.default_rng(3)

    def same(path):
        ref = np.loadtxt(path, ndmin=2).astype("float32")
        got = tu.loadtxt_f32(path)
        return got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    for nd in list(range(1, 22)) + [25]:
        vals = rng.standard_normal((600, 3)) * 10.0 ** rng.integers(-12, 12, size=(600, 3))
        np.savetxt(tmp_path / "e.txt", vals, fmt="%%.%de" % (nd - 1))
        assert same(tmp_path / "e.txt"), "%d significant digits, scientific" % nd
    for nd in (0, 1, 3, 6, 9, 15, 17, 20):
        vals = rng.standard_normal((600, 3)) * 10.0 ** rng.integers(-3, 6, size=(600, 3))
        np.savetxt(tmp_path / "f.txt", vals, fmt="%%.%df" % nd)
        assert same(tmp_path / "f.txt"), "%d decimals, fixed" % nd
    for dtype, span in ((np.float32, 8), (np.float64, 30)):                      # numpy's own format, threaded path
        a = (rng.standard_normal((30000, 3)) * 10.0 ** rng.integers(-span, span, size=(30000, 3))).astype(dtype)
        np.savetxt(tmp_path / "np.txt", a)
        assert os.path.getsize(tmp_path / "np.txt") > (1 << 20) and same(tmp_path / "np.txt")
    f = rng.integers(0x30000000, 0x50000000, size=4000, dtype=np.uint32).view(np.float32)
    mid = (f.astype(np.float64) + np.nextafter(f, np.float32(np.inf)).astype(np.float64)) / 2
    for fmt in ("%.16e", "%.17e", "%.18e"):
        np.savetxt(tmp_path / "m.txt", np.stack([mid, np.nextafter(mid, np.inf), np.nextafter(mid, -np.inf)], 1), fmt=fmt)
        assert same(tmp_path / "m.txt"), fmt
    getcontext().prec = 60
    rows = []
    for x in rng.standard_normal(1500) * 10.0 ** rng.integers(-6, 6, size=1500):
        x = float(x)
        m = (Decimal(x) + Decimal(float(np.nextafter(x, np.inf)))) / 2
        rows.append(" ".join((format(m, ".18e"), format(m, ".17e"), format(m + Decimal(10) ** (m.adjusted() - 18), ".18e"))))
    open(tmp_path / "d.txt", "w").write("\n".join(rows) + "\n")
    assert same(tmp_path / "d.txt")
    txt = ("# header\n\n  1.5\t-2.25 ,+3e2 # tail\n\r\n0.000 -0.0 00012.50\r\n1e-320 1e30 1.7976931348623157e308\nnan inf -inf\n"
           "4.9e-324 2.2250738585072014e-308 123456789012345678901234567890e-20\n.5 5. 1.e3")
    open(tmp_path / "s.txt", "w").write(txt)
    want = np.array([[float(t) for t in ln.split("#")[0].replace(",", " ").split()]
                     for ln in txt.replace("\r", "").split("\n") if ln.split("#")[0].strip()], dtype=np.float64)
    with np.errstate(over="ignore"):
        want = want.astype(np.float32)
    got = tu.loadtxt_f32(tmp_path / "s.txt")
    assert got.shape == want.shape == (6, 3) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    open(tmp_path / "r.txt", "w").write("1 2 3\n4 5\n")
    with pytest.raises(_lib.DeepFitB200Error):
        tu.loadtxt_f32(tmp_path / "r.txt")
    for bad in ("1 2 x\n", "1 2 3\n4 5 6 7\n", "1 2 3\n4 y 6\n"):
        open(tmp_path / "x.txt", "w").write(bad)
        with pytest.raises(_lib.DeepFitB200Error):
            tu.loadtxt_f32(tmp_path / "x.txt")
    with pytest.raises(_lib.DeepFitB200Error):
        tu.loadtxt_f32(tmp_path / "missing.txt")
    open(tmp_path / "c.txt", "w").write("# nothing\n")
    assert tu.loadtxt_f32(tmp_path / "c.txt").shape == (0, 0)


def test_savetxt_exact_formatter_rounding_cases(tmp_path):
    """'%.18e' without printf (csrc/textio.cu: exact 128-bit integer arithmetic on float32-origin values and on true
    doubles): random bit patterns over the whole fast ranges, every neighbour of a power of ten, exponents where the
    rounding carries, exact ties, signed zeros, and the fallbacks (subnormals, huge / tiny values, nan / inf, three-digit
    exponents)."""
    import numpy as np
    from deepfit_b200 import tutorial_utils as tu
    rng = np.random.default_rng(7)

    def same(a):
        tu.savetxt(tmp_path / "ours.txt", a)
        np.savetxt(tmp_path / "numpy.txt", a)
        return open(tmp_path / "ours.txt", "rb").read() == open(tmp_path / "numpy.txt", "rb").read()

    bits = rng.integers(0x2c800000, 0x4c000000, size=600_000, dtype=np.uint32)
    bits[::2] |= 0x80000000
    assert same(bits.view(np.float32).reshape(-1, 3))
    near = []
    for kexp in range(-12, 9):
        v = np.float32(10.0 ** kexp)
        lo = hi = v
        near.append(v)
        for _ in range(5):
            lo = np.nextafter(lo, np.float32(0)); hi = np.nextafter(hi, np.float32(np.inf))
            near += [lo, hi]
    near = np.array(near, dtype=np.float32)
    assert same(np.resize(near, (len(near) // 3) * 3).reshape(-1, 3))
    for e in (113, 127, 149):       # a stride through every mantissa of three binades
        b = (np.uint32(e) << np.uint32(23)) | np.arange(0, 1 << 23, 37, dtype=np.uint32)
        a = b.view(np.float32)
        assert same(np.resize(a, (len(a) // 3) * 3).reshape(-1, 3))
    special = np.array([[0.0, -0.0, 1.0], [0.1, 0.5, -2.0], [1e-38, 3e38, 1.0 / 3], [16777215, 16777216, 8388608],
                        [9.9999999e5, 0.99999994, 1e-7], [2.0 ** -37, 2.0 ** -38, 2.0 ** 24], [1e-45, -1e-45, 65504.0]], dtype=np.float32)
    assert same(special)
    assert same(special.astype(np.float64))                                   # doubles that are exact float32 values
    assert same(np.array([[np.nan, np.inf, -np.inf], [1e-300, -1e300, 5e-324], [1e-99, 9.999e-100, 1e100], [0.1, 1.0 / 3, 1e99]]))
    # true doubles (curv / rad is one): the same integer arithmetic on a 53-bit mantissa, 1e-13 <= |x| < 1e19
    ex = rng.integers(1023 - 60, 1023 + 70, size=(150_000, 2), dtype=np.uint64)
    man = rng.integers(0, 1 << 52, size=(150_000, 2), dtype=np.uint64)
    sgn = rng.integers(0, 2, size=(150_000, 2), dtype=np.uint64)
    assert same(((sgn << np.uint64(63)) | (ex << np.uint64(52)) | man).view(np.float64))
    ex = rng.integers(1, 2046, size=(20_000, 2), dtype=np.uint64)              # every normal binade (fallbacks included)
    assert same(((ex << np.uint64(52)) | man[:20_000]).view(np.float64))
    p10 = np.array([10.0 ** k_ for k_ in range(-14, 20)])
    assert same(np.stack([np.nextafter(p10, 0), p10, np.nextafter(p10, np.inf)], 1))
    assert same(np.array([[m_ * 2.0 ** e for m_ in (1, 3, 5, 7, 9, 11, 4503599627370497.0)] for e in range(-45, 64)]))   # exact ties
    assert same(np.array([[float("9.9999999999999995e%d" % k_), float("1.00000000000000005e%d" % k_)] for k_ in range(-13, 19)]))
    assert same(np.array([[1e-13, 9.999999999999999e-14], [1e19, 9.999999999999998e18], [2.2e-308, 1.7976931348623157e308]]))
