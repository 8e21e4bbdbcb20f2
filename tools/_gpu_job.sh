set -x
export PYTHONPATH=$PWD
timeout 120 python tools/sample_cfg5.py cfg5U 100000 512 200 100 1 300 1 5000 > gpurun_out/r2_samp_u_full_2stage.log 2>&1
timeout 120 python tools/sample_cfg5.py cfg5U 100000 512 200 100 1 300 1 0 > gpurun_out/r2_samp_u_full_1stage.log 2>&1
python -m pytest tests/test_gpu_sampler_stats.py tests/test_gpu_sampler.py -q -x > gpurun_out/r2_gputest9.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_gputest9.log
tail -12 gpurun_out/r2_samp_u_full_2stage.log; tail -4 gpurun_out/r2_samp_u_full_1stage.log; tail -5 gpurun_out/r2_gputest9.log
