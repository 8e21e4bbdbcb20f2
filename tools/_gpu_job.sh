set -x
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -q -x > gpurun_out/r2_gputest2.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_gputest2.log
timeout 150 python tools/sample_cfg5.py cfg5G 20000 64 300 120 1 100 > gpurun_out/r2_samp_g20k_opt.log 2>&1
timeout 200 python tools/sample_cfg5.py cfg5U 20000 64 300 170 1 100 > gpurun_out/r2_samp_u20k_opt.log 2>&1
tail -5 gpurun_out/r2_gputest2.log; tail -14 gpurun_out/r2_samp_g20k_opt.log; tail -14 gpurun_out/r2_samp_u20k_opt.log
