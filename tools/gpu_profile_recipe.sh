# the profiling recipe of round 2 (run under gpurun): plain run first, ncu launch list, ncu --set full of the dominant kernels,
# compute-sanitizer attempts (closed on this pool: see profiles/r02_compute_sanitizer_closed_on_this_pool.txt)
set -x
export PYTHONPATH=$PWD
B="python bench.py --steps 2 --warmup 3 --no-sampling --no-other-workloads --no-cpu-baseline --no-parity"
$B > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg5U.csv $B > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bp_fusedo -s 3 -c 1 -f -o gpurun_out/r2_fusedo $B > gpurun_out/r2_ncu2.log 2>&1
python bench.py --workload cfg5K --steps 2 --warmup 3 --no-sampling --no-other-workloads --no-cpu-baseline --no-parity > gpurun_out/r2_prof_5k.json 2> gpurun_out/r2_prof_5k.err
ncu --set full --clock-control none --import-source on -k regex:bp_fused$ -s 3 -c 1 -f -o gpurun_out/r2_fused_cfg5K python bench.py --workload cfg5K --steps 2 --warmup 3 --no-sampling --no-other-workloads --no-cpu-baseline --no-parity > gpurun_out/r2_ncu3.log 2>&1
python bench.py --workload cfg5G --steps 2 --warmup 3 --no-sampling --no-other-workloads --no-cpu-baseline --no-parity > gpurun_out/r2_prof_5g.json 2> gpurun_out/r2_prof_5g.err
ncu --set full --clock-control none --import-source on -k regex:bp_grouped -s 3 -c 1 -f -o gpurun_out/r2_grouped_cfg5G python bench.py --workload cfg5G --steps 2 --warmup 3 --no-sampling --no-other-workloads --no-cpu-baseline --no-parity > gpurun_out/r2_ncu4.log 2>&1
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py cfg5U 19 3000 > gpurun_out/r2_sanitizer_${tool}_octet.log 2>&1
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py cfg2 8 250 > gpurun_out/r2_sanitizer_${tool}_cfg2.log 2>&1
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_sampler.py 16 40 > gpurun_out/r2_sanitizer_${tool}_sampler.log 2>&1
done
tail -3 gpurun_out/r2_sanitizer_*.log
ls -la gpurun_out/*.ncu-rep
