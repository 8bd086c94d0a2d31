timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do
  AUR_ATTN_V1=1 AUR_NSEQ=192 timeout 120 python tools/attn_prof.py 2>&1 | tail -1 | sed "s/^/mixed v1 /"
  for var in 0 1 2 3; do
    AUR_ATTN_VAR=$var AUR_NSEQ=192 timeout 120 python tools/attn_prof.py 2>&1 | tail -1 | sed "s/^/mixed var=$var /"
  done
done | tee gpurun_out/attn_only_ab_r2.log
for L in 128 384 512; do
  AUR_LEN=$L AUR_ATTN_V1=1 AUR_NSEQ=192 timeout 120 python tools/attn_prof.py 2>&1 | tail -1 | sed "s/^/LEN=$L v1 /"
  for var in 0 1 2 3; do
  AUR_LEN=$L AUR_ATTN_VAR=$var AUR_NSEQ=192 timeout 120 python tools/attn_prof.py 2>&1 | tail -1 | sed "s/^/LEN=$L var=$var /"
done; done | tee -a gpurun_out/attn_only_ab_r2.log
