export PYTHONPATH=$PWD
B="python bench.py --workload cfg5G --steps 20 --warmup 5 --no-sampling --no-other-workloads --no-cpu-baseline --no-parity"
for mb in 3 4; do for t in 128 256 512; do BPC_GC_MINBLOCKS=$mb BPC_GROUPC_TILE=$t $B 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('minblocks $mb tile $t', j['ms_per_step'])"; done; done
