for n in 125000 1000000; do echo "== rows $n"; AURORA_B200_LIB=$PWD/aurora_b200/libaurora_b200_prof.so python tools/push_stats.py $n 2>&1 | tail -16; done | tee gpurun_out/push_stats_small_r2.txt
