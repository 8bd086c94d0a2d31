T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_r2_n2_final.json 2> gpurun_out/bench_r2_n2_final.err; tail -c 300 gpurun_out/bench_r2_n2_final.json; tail -2 gpurun_out/bench_r2_n2_final.err
timeout 300 $T bench.py --gpus 2 --steps 5 --warmup 3 --config cfg3 > gpurun_out/bench_r2_n2_cfg3_final.json 2> gpurun_out/bench_r2_n2_cfg3_final.err; tail -c 300 gpurun_out/bench_r2_n2_cfg3_final.json; tail -2 gpurun_out/bench_r2_n2_cfg3_final.err
timeout 300 $T bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r2_n2_ref_final.json 2> gpurun_out/bench_r2_n2_ref_final.err; tail -c 200 gpurun_out/bench_r2_n2_ref_final.json
