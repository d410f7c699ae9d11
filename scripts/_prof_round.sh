# capture sequence behind profiles/r02_*: plain run first (must exit 0), then the same command under ncu
set -x
python bench.py --short --steps 2 --warmup 3 > gpurun_out/q1_short_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/q1_launches.csv python bench.py --short --steps 2 --warmup 3 > gpurun_out/q1_short_ncu.log 2>&1
PROBE_BREAKDOWN=0 PROBE_CASES=48x1280 python scripts/vocoder_probe.py > gpurun_out/q1_voc_plain.log 2>&1 || exit 1
PROBE_BREAKDOWN=0 PROBE_CASES=48x1280 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_persistent_kernel --launch-skip 40 -c 6 -o gpurun_out/q1_voc_gemm python scripts/vocoder_probe.py > gpurun_out/q1_voc_ncu.log 2>&1
PROBE_BREAKDOWN=0 PROBE_CASES=48x1280 ncu --set full --clock-control none -k regex:"dwconv_adaln_tiled|groupnorm_stats|groupnorm_apply|transpose_v" --launch-skip 30 -c 6 -o gpurun_out/q1_voc_stream python scripts/vocoder_probe.py > gpurun_out/q1_voc_ncu2.log 2>&1
ls -la gpurun_out/q1_*
