set -x
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -5
for v in 1 0 1 0; do
  AUR_ATTN_V1=$v AUR_NSEQ=192 timeout 120 python tools/attn_prof.py 2>&1 | tail -1 | sed "s/^/V1=$v /"
done | tee gpurun_out/attn_only_ab_r2.log
AUR_NSEQ=192 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -c 1 -o gpurun_out/attn_tc2_r2 -f python tools/attn_prof.py 2>&1 | tail -3
