// default_model.cu — bp_kernels.cuh instantiated by nvcc for the reference 3-type model of BASELINE.json's
// configs[4] (9 events: per type a division, a transition to the next type and a death; rates c[1..9]).
// Product models are compiled by NVRTC from the same header (codegen.cpp); this file exists so that build()
// compiles the hand-written kernels ahead of time and `nvcc -Xptxas -v` / cuobjdump show registers and SASS.
// The prelude below is what codegen.cpp generates for that model (tests/test_codegen.py checks they agree).
#include "default_model_prelude.inc"
