timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_encoder.py -m gpu -x -q -k "pinned or many_units or tenant or scope or concurrent" 2>&1 | tail -4
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export AUR_NO_DIRECT_OUT=1; else unset AUR_NO_DIRECT_OUT; fi
  python bench.py --no-parity --steps 300 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no_direct=$v value %.0f e2e %.0f e2e_ms %.4f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done | tee gpurun_out/direct_out_ab_r2.txt
