set -x
export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_parity.py tests/test_gpu_corners.py -q -x -k "octet or full_size or bench_shape or unbounded or deterministic" > gpurun_out/r2_gputest7.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_gputest7.log
B="python bench.py --steps 20 --warmup 5 --no-sampling --no-other-workloads --no-cpu-baseline"
BPC_O9=0 $B > gpurun_out/r2_bench_o9_off.json 2> gpurun_out/r2_bench_o9_off.err
$B > gpurun_out/r2_bench_o9_on.json 2> gpurun_out/r2_bench_o9_on.err
tail -5 gpurun_out/r2_gputest7.log; cut -c1-330 gpurun_out/r2_bench_o9_off.json; cut -c1-330 gpurun_out/r2_bench_o9_on.json; tail -3 gpurun_out/r2_bench_o9_on.err
