PROBE_T0=40 PROBE_STEPS=3 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm" --launch-skip 700 -c 17 -o gpurun_out/aj_big_gemm python scripts/bigbatch_probe.py > gpurun_out/aj_ncu.log 2>&1
tail -3 gpurun_out/aj_ncu.log
ls -la gpurun_out/aj_*
