timeout 1200 python -m pytest tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -2
python bench.py --no-parity --no-encoder --steps 300 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f ms %.4f e2e %.0f kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']), d.get('phases_ms'))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize --launch-skip 3 -c 1 -o gpurun_out/prof_fin_r2c -f python tools/profile_search.py tc2 1000000 256 32 6 2>&1 | tail -1
