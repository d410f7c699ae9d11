PROBE_BREAKDOWN=0 PROBE_CASES=48x1280 python scripts/vocoder_probe.py > gpurun_out/n1_plain.log 2>&1 || exit 1
PROBE_BREAKDOWN=0 PROBE_CASES=48x1280 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_persistent_kernel --launch-skip 40 -c 6 -o gpurun_out/n1_voc_gemm python scripts/vocoder_probe.py > gpurun_out/n1_ncu.log 2>&1
ls -la gpurun_out/n1_*
