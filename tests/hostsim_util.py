"""TEST INFRASTRUCTURE: builds tests/hostsim/hostsim.cpp for one generated model and evaluates chains on the host.

This executes the BP_HD device functions of bp_kernels.cuh compiled by g++ — a checker for the kernel
arithmetic that runs without a GPU.  It is never imported by the product."""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(_HERE), "bpinference_b200", "csrc")
_CACHE = {}


def build(engine, counted=False):
    src = engine.cuda_source
    deps = open(os.path.join(_CSRC, "bp_kernels.cuh")).read() + open(os.path.join(_HERE, "hostsim", "hostsim.cpp")).read()
    key = hashlib.sha1((src + deps + str(counted)).encode()).hexdigest()[:16]
    if key in _CACHE:
        return _CACHE[key]
    d = os.path.join(tempfile.gettempdir(), "bp_hostsim_" + key)
    os.makedirs(d, exist_ok=True)
    msrc = os.path.join(d, "model.cu.h")
    with open(msrc, "w") as f:
        f.write(src)
    so = os.path.join(d, "hostsim.so")
    if not os.path.exists(so):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [cxx, "-O1" if counted else "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-DBP_HOSTSIM",
               '-DBP_MODEL_SRC="%s"' % msrc, "-I" + _CSRC, os.path.join(_HERE, "hostsim", "hostsim.cpp"), "-o", so]
        if counted:
            cmd.insert(1, "-DBP_COUNTED")
        subprocess.check_call(cmd)
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.hs_obs.argtypes = [dp, dp, dp, dp, C.c_double, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)]
    L.hs_obs.restype = C.c_int
    _CACHE[key] = L
    return L


def likelihood(engine, dat, theta, counted=False):
    """sum over observations of (lp, d lp/d theta [P], d lp/d sigma_obs), status, total applications, ops —
    for ONE chain with constrained parameters theta [dim]."""
    L = build(engine, counted)
    simple = engine.kind == 2
    N, n, P = int(dat["ndatapts"]), engine.ntypes, engine.nparams
    pop = np.asarray(dat["pop_vec"], dtype=float).reshape(N, -1)
    init = np.asarray(dat["init_pop"], dtype=float).reshape(N, -1)
    times = np.atleast_1d(np.asarray(dat["times"], dtype=float))
    fv = np.asarray(dat["function_var"], dtype=float).reshape(int(dat["ndep_levels"]), int(dat["ndep"]))
    kdep = max(engine.ndep, 1)
    th = np.ascontiguousarray(theta, dtype=float)
    dp = C.POINTER(C.c_double)
    lp_tot, cb_tot, sb_tot, st_tot, apps_tot = 0.0, np.zeros(P), 0.0, 0, 0
    ops = np.zeros(2, dtype=np.uint64)
    ops_tot = np.zeros(2, dtype=np.uint64)
    for k in range(N):
        t = times[k] if simple else times[int(dat["times_idx"][k]) - 1]
        x = np.zeros(kdep)
        x[:min(kdep, fv.shape[1])] = fv[int(dat["var_idx"][k]) - 1, :kdep]
        z0, xo = np.ascontiguousarray(init[k]), np.ascontiguousarray(pop[k])
        lp, sb, apps = C.c_double(), C.c_double(), C.c_int()
        cb = np.zeros(P)
        st = L.hs_obs(th.ctypes.data_as(dp), x.ctypes.data_as(dp), z0.ctypes.data_as(dp), xo.ctypes.data_as(dp), float(t), C.byref(lp),
                      cb.ctypes.data_as(dp), C.byref(sb), C.byref(apps), ops.ctypes.data_as(C.POINTER(C.c_ulonglong)))
        st_tot |= st
        if st == 0:
            lp_tot += lp.value
            cb_tot += cb
            sb_tot += sb.value
        apps_tot += apps.value
        ops_tot += ops
    return lp_tot, cb_tot, sb_tot, st_tot, apps_tot, ops_tot
