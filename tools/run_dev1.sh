timeout 1200 python -m pytest tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python bench.py --no-parity --no-encoder --steps 300 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f ms %.4f e2e %.0f kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']), d['roofline']['frac'])"; done
python tools/rows_sweep.py gpurun_out/rows_sweep_tail.json 2>&1 | tail -1
