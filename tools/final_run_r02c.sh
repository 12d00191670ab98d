# One box session refreshing the round-2 evidence after the third pass (run under gpurun from the repo root):
#   tests, smoke, both bench arms, stage tables and probes, the launch list of a short bench run and the ncu --set full
#   captures of the persistent kernel, the warp-specialised energy + heat-map kernel and the packed mask kernels.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.txt
python __graft_entry__.py smoke > gpurun_out/r02_smoke.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err
python tools/bench_stages.py > gpurun_out/r02_stage_table.txt 2>&1
python tools/stage2_probe.py 8192 > gpurun_out/r02_stage2_probe.txt 2>&1
python tools/stage2_quick.py 8192 > gpurun_out/r02_stage2_quick.txt 2>&1
python tools/mask_probe.py 8192 > gpurun_out/r02_mask_probe.txt 2>&1
python tests/run_latency_probe.py > gpurun_out/r02_latency_probe.txt 2>&1
SHORT="python bench.py --steps 2 --warmup 3 --frames 2048 --no-cpu-baseline --no-configs --e2e-steps 1 --e2e-frames 64 --e2e-rounds 1 --sustain-seconds 0"
$SHORT > gpurun_out/r02_short_plain.json 2> gpurun_out/r02_short_plain.err && \
 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv $SHORT > gpurun_out/r02_launches.log 2>&1 && \
 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused -s 3 -c 2 -f -o gpurun_out/r02_fused $SHORT > gpurun_out/r02_fused_ncu.log 2>&1
python tools/ws_once.py > /dev/null 2>&1 && \
 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'energy_heat_ws|heat_stream' -f -o gpurun_out/r02_ws python tools/ws_once.py > gpurun_out/r02_ws_ncu.log 2>&1
python tools/mask_once.py > /dev/null 2>&1 && \
 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'packed' -f -o gpurun_out/r02_mask python tools/mask_once.py > gpurun_out/r02_mask_ncu.log 2>&1
cat gpurun_out/r02_pytest_gpu.txt gpurun_out/r02_smoke.txt gpurun_out/r02_mask_probe.txt
grep heatmap gpurun_out/r02_stage2_quick.txt
