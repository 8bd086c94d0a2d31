set -x
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
timeout 300 $T bench.py --gpus 2 --steps 30 --warmup 5 --exchange fused > gpurun_out/bench_r2_n2_fused.json 2> gpurun_out/bench_r2_n2_fused.err; tail -c 1500 gpurun_out/bench_r2_n2_fused.json; tail -3 gpurun_out/bench_r2_n2_fused.err
timeout 300 $T bench.py --gpus 2 --steps 30 --warmup 5 --exchange nccl > gpurun_out/bench_r2_n2_nccl.json 2> gpurun_out/bench_r2_n2_nccl.err; tail -c 800 gpurun_out/bench_r2_n2_nccl.json; tail -3 gpurun_out/bench_r2_n2_nccl.err
timeout 400 $T bench.py --gpus 2 --steps 10 --warmup 3 --config cfg4 > gpurun_out/bench_r2_n2_cfg4.json 2> gpurun_out/bench_r2_n2_cfg4.err; tail -c 1500 gpurun_out/bench_r2_n2_cfg4.json; tail -3 gpurun_out/bench_r2_n2_cfg4.err
AUR_BENCH_ROWS=1250000 timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 3 --config cfg5 > gpurun_out/bench_r2_n2_cfg5.json 2> gpurun_out/bench_r2_n2_cfg5.err; tail -c 1500 gpurun_out/bench_r2_n2_cfg5.json; tail -3 gpurun_out/bench_r2_n2_cfg5.err
timeout 300 $T bench.py --gpus 2 --steps 8 --warmup 3 --config cfg3 > gpurun_out/bench_r2_n2_cfg3.json 2> gpurun_out/bench_r2_n2_cfg3.err; tail -c 1200 gpurun_out/bench_r2_n2_cfg3.json; tail -3 gpurun_out/bench_r2_n2_cfg3.err
