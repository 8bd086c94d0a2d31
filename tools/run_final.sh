# Round-end style validation on one GPU: full GPU suite, smoke, default bench (both arms), ncu launch list.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
timeout 600 python bench.py 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json; tail -c 600 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/final_ref.err | tail -1 > gpurun_out/final_ref.json; tail -c 400 gpurun_out/final_ref.json
timeout 600 python bench.py --config cfg3 2>gpurun_out/final_cfg3.err | tail -1 > gpurun_out/final_cfg3.json; tail -c 500 gpurun_out/final_cfg3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-parity > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 600 python tools/daemon_load.py gpurun_out/daemon_load_r2.json 2>&1 | tail -4
