export PYTHONPATH=$PWD
B="python bench.py --workload cfg5G --steps 20 --warmup 5 --no-sampling --no-other-workloads --no-cpu-baseline --no-parity"
for t in 512 1024 2048 4096; do BPC_GROUPB_TILE=$t $B 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('tile $t', j['ms_per_step'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_cfg5G.csv python bench.py --workload cfg5G --steps 2 --warmup 3 --no-sampling --no-other-workloads --no-cpu-baseline --no-parity > /dev/null 2>&1
python tools/launch_times.py gpurun_out/r2_launches_cfg5G.csv | grep bp_
