PROBE_BREAKDOWN=0 PROBE_CASES=48x1280 python scripts/vocoder_probe.py > gpurun_out/an_plain.log 2>&1 || exit 1
PROBE_BREAKDOWN=0 PROBE_CASES=48x1280 ncu --set full --clock-control none --import-source on -k regex:"head_act_split3|overlap_add|layernorm_split3|attn_softmax|transpose_v" --launch-skip 10 -c 5 -o gpurun_out/an_voc_head python scripts/vocoder_probe.py > gpurun_out/an_ncu.log 2>&1
ls -la gpurun_out/an_*
