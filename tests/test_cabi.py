"""The C-ABI library loads without a GPU and exports every symbol include/bpcuda.h declares; codegen + NVRTC work
on the CPU box; evaluation entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import bpinference_b200 as bp
from bpinference_b200 import _ffi, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bpcuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 35
    L = ctypes.CDLL(_ffi.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(_ffi.SYMBOLS) == names          # the ctypes mirror binds exactly the declared surface
    assert _ffi.lib().bpc_abi_version() == 2


def test_codegen_and_nvrtc_without_gpu():
    cfg = syn.make_config("cfg3", chains=2)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=True)
    src = eng.cuda_source
    assert "#define BP_KIND 2" in src and "#define BP_XDEP 1" in src and "exp(" in src and "bp_rates_vjp" in src
    cub = eng.cubin()
    assert cub[:4] == b"\x7fELF" and len(cub) > 10000          # sm_100a machine code from NVRTC
    assert eng.dim == 5


def test_stan_integer_division_quirk_is_reproduced_and_reported():
    """`1/2` pasted into Stan is integer division (= 0); the front end follows Stan and warns."""
    mod = bp.bp_model([[2.0], [0.0]], [1, 1], ["1/2*c[1]", "c[2]*(3/2.)*1.5"], 2, 0)
    eng = bp.stan_model(mod, [bp.prior_dist("uniform", [0, 1], [0, 2])] * 2)
    assert "integer division" in eng.warnings
    assert "0.0 * c[0]" in eng.cuda_source          # 1/2 -> 0
    assert "c[1] * 1.0" in eng.cuda_source          # R deparses `2.` as 2, so 3/2. is int division too -> 1
    assert "* 1.5" in eng.cuda_source


def test_no_cpu_fallback():
    if _ffi.lib().bpc_device_count() > 0:
        pytest.skip("a GPU is present")
    cfg = syn.make_config("cfg1", chains=2)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=True)
    with pytest.raises(_ffi.BpcError) as e:
        eng.log_prob_grad(cfg["data"], cfg["upars"])
    assert e.value.code == _ffi.E_CUDA and "no CPU fallback" in str(e.value)


def test_constrain_roundtrip_host():
    cfg = syn.make_config("cfg4", chains=3)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], noise_model=True)
    th = eng.constrain_pars(cfg["upars"])
    assert th.shape == (3, 4) and (th[:, :3] > 0).all() and (th[:, :3] < 2).all() and (th[:, 3] > 0).all()
    np.testing.assert_allclose(eng.unconstrain_pars(th), cfg["upars"], rtol=1e-12)


@pytest.mark.parametrize("name,expect", [("cfg5U", ("bp_otable", "bp_fusedo", "bp_toep", "bp_fusedc", "bp_ctable", "bp_fusedw", "bp_wpost", "bp_fused", "bp_grouped")),
                                         ("cfg2", ("bp_otable", "bp_fusedo", "bp_toep", "bp_grouped", "bp_fused")),
                                         ("cfg4", ("bp_fusedo", "bp_grouped")), ("cfg1", ("bp_onetype",))])
def test_every_module_compiles_for_sm100a_with_the_planned_resources(name, expect, tmp_path):
    """NVRTC output of each template family (no GPU needed): the kernels of every path are in the sm_100a cubin, and the
    headline kernel keeps what its launch geometry counts on -- 128 registers (four CTAs per SM), no spills to local memory."""
    import shutil
    import subprocess
    cfg = syn.make_config(name, chains=2, nobs=64)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod, cfg["priors"], simple_bd=cfg["kind"] == "simple_bd", noise_model=cfg["kind"] == "multitype_noise")
    cub = eng.cubin()
    assert cub[:4] == b"\x7fELF"
    for k in expect + ("bp_prologue", "bp_epilogue", "bp_loglik"):
        assert k.encode() in cub, k
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    path = tmp_path / "m.cubin"
    path.write_bytes(cub)
    head = subprocess.run(["cuobjdump", "-sass", "-fun", "bp_prologue", str(path)], capture_output=True, text=True).stdout
    assert "code for sm_100a" in head
    out = subprocess.run(["cuobjdump", "-res-usage", str(path)], capture_output=True, text=True).stdout
    res = dict(re.findall(r"Function (\w+):\s*\n\s*(REG:\d+ STACK:\d+)", out))
    if "bp_fusedo" in expect:
        regs, stack = (int(x.split(":")[1]) for x in res["bp_fusedo"].split())
        assert regs <= 128 and stack == 0, res["bp_fusedo"]


def test_module_cache_and_alternative_kernels_on_request(monkeypatch):
    """the same model description built twice in one process is compiled once (the cache key is the generated source, macros
    included); the measured alternatives of the octet passes (bp_fusedo2 / bp_fusedo9) are in a module only when BPC_O2 / BPC_O9 asks."""
    import time
    cfg = syn.make_config("cfg5U", chains=2, nobs=64)
    mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    monkeypatch.delenv("BPC_O2", raising=False)
    monkeypatch.delenv("BPC_O9", raising=False)
    monkeypatch.delenv("BPC_ALT_KERNELS", raising=False)
    a = bp.stan_model(mod, cfg["priors"])
    t0 = time.perf_counter()
    b = bp.stan_model(mod, cfg["priors"])
    assert time.perf_counter() - t0 < 1.0
    assert a.cubin() == b.cubin() and a.cuda_source == b.cuda_source
    assert b"bp_fusedo2" not in a.cubin() and b"bp_fusedo9" not in a.cubin() and b"bp_fusedo" in a.cubin()
    monkeypatch.setenv("BPC_O2", "1")
    c = bp.stan_model(mod, cfg["priors"])
    assert c.cuda_source != a.cuda_source and b"bp_fusedo2" in c.cubin() and b"bp_fusedo9" in c.cubin()


def test_r_shim_compiles_against_the_header():
    """r-package/src/bpcuda_rcall.c (the .Call binding INTEGRATION.md describes) cannot be built here -- R is not in the image -- but it
    can be type-checked: against include/bpcuda.h (every bpc_* call with the header's argument list) and against declarations of the
    part of R's C API it uses (tests/r_stub, written for this test).  Also: every registered routine exists with the registered
    number of SEXP arguments, and the R wrappers call only registered routines."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "r-package", "src", "bpcuda_rcall.c")
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                        "-Werror=int-conversion", "-Werror=missing-field-initializers", "-I" + os.path.join(root, "tests", "r_stub"),
                        "-I" + os.path.join(root, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    reg = dict((m.group(1), int(m.group(2))) for m in re.finditer(r'\{"(bpc_R_\w+)", \(DL_FUNC\)&\1, (\d+)\}', text))
    assert len(reg) >= 10
    for name, nargs in reg.items():
        m = re.search(r"SEXP %s\(([^)]*)\)\s*\{" % name, text)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == nargs and all(p.strip().startswith("SEXP ") for p in params), (name, nargs, m.group(1))
    rtext = "\n".join(ln.split("#")[0] for ln in open(os.path.join(root, "r-package", "R", "bpcuda.R")).read().splitlines())
    calls = []
    for m in re.finditer(r"\.Call\(", rtext):          # every .Call(...) with its top-level arguments
        depth, args, cur, i = 1, [], "", m.end()
        while depth:
            ch = rtext[i]
            depth += ch in "([{"
            depth -= ch in ")]}"
            if depth == 1 and ch == ",":
                args.append(cur)
                cur = ""
            elif depth:
                cur += ch
            i += 1
        args.append(cur)
        calls.append((args[0].strip().strip('"'), len(args) - 1))
    assert len(calls) >= 10
    for name, nargs in calls:
        assert reg.get(name) == nargs, (name, nargs, reg.get(name))


def test_non_finite_bounds_parameters_and_data_are_rejected():
    """(0, Inf) bounds used to produce Theta = Inf / NaN and silently rejected chains; NaN durations broke the sort in
    bpc_data_create.  Both are refused at the boundary now, before any device work."""
    L = _ffi.lib()
    for lo, hi in ((0.0, np.inf), (-np.inf, 1.0)):
        pr = _ffi.Prior(b"normal", 2, (ctypes.c_double * 3)(0.0, 1.0, 0.0), 1, lo, hi)
        assert L.bpc_check_prior(ctypes.byref(pr)) == _ffi.E_PRIOR and b"finite" in L.bpc_last_error()
    pr = _ffi.Prior(b"normal", 2, (ctypes.c_double * 3)(np.nan, 1.0, 0.0), 0, 0.0, 0.0)
    assert L.bpc_check_prior(ctypes.byref(pr)) == _ffi.E_PRIOR
    mod = bp.bp_model([[2.0], [0.0]], [1, 1], ["c[1]", "c[2]"], 2, 0)
    with pytest.raises(ValueError, match="finite"):
        bp.stan_model(mod, [dict(name="normal", params=[0.0, 1.0], bounds=[0.0, float("inf")])] * 2)
    cfg = syn.make_config("cfg2", chains=2, nobs=20)
    mod2 = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
    eng = bp.stan_model(mod2, cfg["priors"])
    for key, bad, msg in (("times", np.array([np.nan]), "times"), ("times", np.array([-1.0]), "times"), ("init_pop", None, "finite")):
        d = dict(cfg["data"])
        if bad is None:
            a = np.array(d[key], dtype=float)
            a[3, 1] = np.inf
            d[key] = a
        else:
            d[key] = bad
        with pytest.raises(_ffi.BpcError) as e:
            eng.data(d)
        assert e.value.code == _ffi.E_INVALID and msg in str(e.value)
