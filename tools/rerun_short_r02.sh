set -x
SHORT="python bench.py --steps 2 --warmup 3 --frames 2048 --no-cpu-baseline --no-configs --e2e-steps 1 --e2e-frames 64 --e2e-rounds 1 --sustain-seconds 0"
$SHORT > gpurun_out/r02_short_plain.json 2> gpurun_out/r02_short_plain.err && \
 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv $SHORT > gpurun_out/r02_launches.log 2>&1 && \
 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused -s 3 -c 2 -f -o gpurun_out/r02_fused $SHORT > gpurun_out/r02_fused_ncu.log 2>&1
ls -la gpurun_out | grep -E "fused|launches|short"
