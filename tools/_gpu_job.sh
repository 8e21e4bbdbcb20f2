export PYTHONPATH=$PWD
BPC_SAMPLER_DEBUG=1 python tools/diag_three_type_grouped.py > gpurun_out/r2_diag3.log 2>&1
grep -v "^\[opt\]\|^\[hess\]" gpurun_out/r2_diag3.log | head -80; grep "^\[opt\]" gpurun_out/r2_diag3.log | awk '{print $7}' | sort -n | head -70 | tr '\n' ' '
