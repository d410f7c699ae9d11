set -x
python bench.py --short --steps 2 --warmup 3 > gpurun_out/h1_short_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/h1_launches.csv python bench.py --short --steps 2 --warmup 3 > gpurun_out/h1_short_ncu.log 2>&1
TARGET_PRECISION=exact python scripts/ncu_step_target.py > gpurun_out/h1_target_plain.log 2>&1 || exit 1
TARGET_PRECISION=exact ncu --set full --clock-control none --import-source on -k regex:cluster_decode --launch-skip 4 -c 4 -o gpurun_out/h1_cluster_exact python scripts/ncu_step_target.py > gpurun_out/h1_ncu_exact.log 2>&1
PROBE_N=240 PROBE_CUT=8 python scripts/cluster_ncu_target.py > gpurun_out/h1_target8_plain.log 2>&1 || exit 1
PROBE_N=240 PROBE_CUT=8 ncu --set full --clock-control none --import-source on -k regex:cluster_decode --launch-skip 1 -c 1 -o gpurun_out/h1_cluster8 python scripts/cluster_ncu_target.py > gpurun_out/h1_ncu_8.log 2>&1
PROBE_N=64 PROBE_CUT=16 PROBE_PRECISION=exact ncu --set full --clock-control none --import-source on -k regex:cluster_decode --launch-skip 1 -c 1 -o gpurun_out/h1_cluster_exact_t110 python scripts/cluster_ncu_target.py > gpurun_out/h1_ncu_exact_t110.log 2>&1
PROBE_CHECK=0 PROBE_N=64 PROBE_PRECISION=exact PROBE_TRACE=1 python scripts/cluster_probe.py > gpurun_out/h1_probe_exact.log 2>&1
PROBE_CHECK=0 PROBE_N=64 PROBE_PRECISION=bf16 PROBE_TRACE=1 python scripts/cluster_probe.py > gpurun_out/h1_probe_bf16.log 2>&1
ls -la gpurun_out/h1_*
