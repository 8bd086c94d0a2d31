set -x
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
O=gpurun_out/bench_r2_n$N
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/mgpu_test_n$N.log 2>&1; rc=$?; tail -5 gpurun_out/mgpu_test_n$N.log
if [ $rc -ne 0 ]; then echo "MULTI-GPU PARITY TEST FAILED: skipping the benches"; tail -40 gpurun_out/mgpu_test_n$N.log; exit 1; fi
timeout 300 $T bench.py --gpus $N --steps 50 --warmup 5 --exchange fused --graph > ${O}_fused_graph.json 2> ${O}_fused_graph.err; tail -c 300 ${O}_fused_graph.json; tail -3 ${O}_fused_graph.err
timeout 300 $T bench.py --gpus $N --steps 50 --warmup 5 --exchange fused > ${O}_fused.json 2> ${O}_fused.err; tail -c 300 ${O}_fused.json; tail -3 ${O}_fused.err
timeout 300 $T bench.py --gpus $N --steps 50 --warmup 5 --exchange nccl > ${O}_nccl.json 2> ${O}_nccl.err; tail -c 300 ${O}_nccl.json; tail -3 ${O}_nccl.err
timeout 400 $T bench.py --gpus $N --steps 10 --warmup 3 --config cfg4 --graph > ${O}_cfg4.json 2> ${O}_cfg4.err; tail -c 300 ${O}_cfg4.json; tail -3 ${O}_cfg4.err
AUR_BENCH_ROWS=${ROWS5:-1250000} timeout 600 $T bench.py --gpus $N --steps 20 --warmup 3 --config cfg5 > ${O}_cfg5.json 2> ${O}_cfg5.err; tail -c 300 ${O}_cfg5.json; tail -3 ${O}_cfg5.err
timeout 300 $T bench.py --gpus $N --steps 8 --warmup 3 --config cfg3 > ${O}_cfg3.json 2> ${O}_cfg3.err; tail -c 300 ${O}_cfg3.json; tail -3 ${O}_cfg3.err
