set -x
export PYTHONPATH=$PWD
BPC_SAMPLER_DEBUG=1 timeout 150 python tools/sample_cfg5.py cfg5G 20000 64 200 120 1 300 1 > gpurun_out/r2_dbg_g.log 2>&1
BPC_SAMPLER_DEBUG=1 timeout 150 python tools/sample_cfg5.py cfg5U 20000 64 200 120 1 300 1 > gpurun_out/r2_dbg_u.log 2>&1
timeout 150 python tools/sample_cfg5.py cfg5U 20000 64 200 120 1 300 0 > gpurun_out/r2_dbg_u_diag.log 2>&1
timeout 300 python tools/sample_cfg5.py cfg5U 100000 512 200 280 1 300 1 > gpurun_out/r2_samp_u_full.log 2>&1
grep -c rescue gpurun_out/r2_dbg_g.log gpurun_out/r2_dbg_u.log; grep "\[hess\]" gpurun_out/r2_dbg_u.log | head -5
grep -v "^\[" gpurun_out/r2_dbg_g.log | tail -9; grep -v "^\[" gpurun_out/r2_dbg_u.log | tail -9; tail -9 gpurun_out/r2_dbg_u_diag.log; tail -12 gpurun_out/r2_samp_u_full.log
