timeout 600 ncu --set full --clock-control none -k regex:"gemm_tc|attn_tc" --launch-skip 10 -c 5 -o gpurun_out/encoder_layer_r2 -f python tools/encoder_layer_prof.py 2>&1 | tail -3
