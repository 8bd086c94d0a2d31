AUR_ATTN=3 timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "attention" 2>&1 | tail -3
for rep in 1 2; do for v in 1 2 3; do
  AUR_ATTN=$v AUR_NSEQ=192 timeout 120 python tools/attn_prof.py 2>&1 | tail -1 | sed "s/^/mixed v$v /"
done; done | tee gpurun_out/attn_v3_ab.log
for L in 128 384 512; do for v in 2 3; do
  AUR_LEN=$L AUR_ATTN=$v AUR_NSEQ=192 timeout 120 python tools/attn_prof.py 2>&1 | tail -1 | sed "s/^/LEN=$L v$v /"
done; done | tee -a gpurun_out/attn_v3_ab.log
