PROBE_CASES=48x1280 python scripts/vocoder_probe.py > gpurun_out/w1_plain.log 2>&1 || exit 1
PROBE_BREAKDOWN=0 PROBE_CASES=48x1280 ncu --set full --clock-control none --import-source on -k regex:"dwconv_adaln_bulk_kernel|groupnorm|layernorm_split3|head_act_split3|attn_softmax|istft|overlap" --launch-skip 30 -c 12 -o gpurun_out/w1_voc_stream python scripts/vocoder_probe.py > gpurun_out/w1_ncu.log 2>&1
PROBE_BREAKDOWN=0 PROBE_CASES=48x1280 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 260 --csv --log-file gpurun_out/w1_voc_launches.csv python scripts/vocoder_probe.py > gpurun_out/w1_ncu2.log 2>&1
ls -la gpurun_out/w1_*
