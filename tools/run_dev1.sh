set -x
timeout 700 python -m pytest tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -4
python tools/ab.py libaurora_base.so libaurora_b200.so --rounds 2 2>&1 | tail -2 | tee gpurun_out/ab_r2_6.log
for cfg in "1250000 1024 100" "1250000 1024 32" "2000000 512 32"; do set -- $cfg; echo "== rows $1 nq $2 k $3 dim ${D:-}"; AUR_DIM=1024 python tools/profile_search.py tc2 $1 $2 $3 4 2>&1 | tail -2; AURORA_B200_AB_OLD_ABI=1 AURORA_B200_LIB=aurora_b200/libaurora_base.so AUR_DIM=1024 python tools/profile_search.py tc2 $1 $2 $3 4 2>&1 | tail -1 | sed 's/^/base: /'; done 2>&1 | tee gpurun_out/superblock_r2.log
AUR_DIM=768 python tools/profile_search.py tc2 4000000 512 32 4 2>&1 | tail -1 | sed 's/^/cfg5-like 4M x768 nq512: /' | tee -a gpurun_out/superblock_r2.log
AURORA_B200_AB_OLD_ABI=1 AURORA_B200_LIB=aurora_b200/libaurora_base.so AUR_DIM=768 python tools/profile_search.py tc2 4000000 512 32 4 2>&1 | tail -1 | sed 's/^/base cfg5-like: /' | tee -a gpurun_out/superblock_r2.log
python bench.py --steps 30 --warmup 5 --no-encoder --graph > gpurun_out/bench_r2_dev4g.json 2> gpurun_out/bench_r2_dev4g.err; tail -c 300 gpurun_out/bench_r2_dev4g.json; tail -3 gpurun_out/bench_r2_dev4g.err
python tools/rows_sweep.py gpurun_out/simtopk_rows_sweep_r2b.json > gpurun_out/sweep_r2b.log 2>&1; tail -2 gpurun_out/sweep_r2b.log
