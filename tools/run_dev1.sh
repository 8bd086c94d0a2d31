timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --config cfg3 2>gpurun_out/final_cfg3.err | tail -1 > gpurun_out/final_cfg3.json
python - <<'PY'
import json
c=json.load(open('gpurun_out/final_cfg3.json'))
print('cfg3 value', c['value'], 'e2e', c['e2e']['value'], 'from_text', c['from_text']['chunks_per_s'], c['from_text']['ms_per_batch'], 'tok', c['from_text']['tokenizer_only_chunks_per_s'])
PY
