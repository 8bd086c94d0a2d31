python tools/decompose.py 125000 gpurun_out/decompose_125k_r2.json 2>&1 | tail -14
python tools/push_stats.py 125000 2>&1 | tail -12
