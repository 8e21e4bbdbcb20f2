set -x
export PYTHONPATH=$PWD
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py tests/test_gpu_sampler_stats.py -q -k "multi or three_type" > gpurun_out/r2_gputest_n2.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_gputest_n2.log
NCCL_DEBUG=VERSION timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference --cpu-seconds 8 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err
tail -5 gpurun_out/r2_gputest_n2.log; tail -5 gpurun_out/r2_bench_n2.err; cut -c1-300 gpurun_out/r2_bench_n2.json; cut -c1-700 gpurun_out/r2_bench_ref_n2.json; cat gpurun_out/r2_sample_multi_2dev.json
