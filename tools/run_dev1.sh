timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "bootstrap or chunker or retriever or embedding" 2>&1 | tail -2
timeout 600 python bench.py --config cfg3 2>gpurun_out/final_cfg3.err | tail -1 > gpurun_out/final_cfg3.json
python - <<'PY'
import json
c=json.load(open('gpurun_out/final_cfg3.json'))
print('cfg3 value', c['value'], 'from_text', c['from_text']['chunks_per_s'], c['from_text']['ms_per_batch'], 'tok', c['from_text']['tokenizer_only_chunks_per_s'], c['from_text']['tokenizer_only_chunks_per_s_1_thread'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize --launch-skip 3 -c 1 -o gpurun_out/prof_fin_r2c -f python tools/profile_search.py tc2 1000000 256 32 6 2>&1 | tail -1
