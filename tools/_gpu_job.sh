export PYTHONPATH=$PWD
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
tail -3 gpurun_out/r2_bench_d.err; cut -c1-200 gpurun_out/r2_bench_d.json
