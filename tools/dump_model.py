"""Dumps the generated CUDA translation unit (and the NVRTC cubin) of a synthetic config: offline look at registers / SASS.
usage: python tools/dump_model.py cfg5U /tmp/cfg5   ->  /tmp/cfg5.cu, /tmp/cfg5.cubin"""
import ctypes as C, sys
import bpinference_b200 as bp
from bpinference_b200 import synthetic as syn, _ffi
name, out = sys.argv[1], sys.argv[2]
cfg = syn.make_config(name, chains=4, nobs=256)
mod = bp.bp_model(cfg["e_mat"], cfg["p_vec"], cfg["func_deps"], cfg["nparams"], cfg["ndep"])
eng = bp.stan_model(mod, cfg["priors"])
lib = _ffi.lib()
lib.bpc_model_cuda_source.restype = C.c_char_p
open(out + ".cu", "wb").write(lib.bpc_model_cuda_source(eng.handle))
data, size = C.c_void_p(), C.c_ulonglong()
if lib.bpc_model_cubin(eng.handle, C.byref(data), C.byref(size)) == 0 and size.value:
    open(out + ".cubin", "wb").write(C.string_at(data.value, size.value))
print("wrote", out + ".cu", size.value)
