# One box session refreshing the evidence after the last session of round 2 (run under gpurun from the repo root): tests,
# smoke, both bench arms, stage table and probes, the launch list of a short bench run, ncu --set full of the wide
# float64 kernels.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.txt
python __graft_entry__.py smoke > gpurun_out/r02_smoke.txt 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python tools/bench_stages.py > gpurun_out/r02_stage_table.txt 2>&1
python tools/stage2_quick.py 8192 > gpurun_out/r02_stage2_quick.txt 2>&1
python tools/overlay_probe.py > gpurun_out/r02_overlay_probe.txt 2>&1
python tests/run_latency_probe.py > gpurun_out/r02_latency_probe.txt 2>&1
SHORT="python bench.py --steps 2 --warmup 3 --frames 2048 --no-cpu-baseline --no-configs --e2e-steps 1 --e2e-frames 64 --e2e-rounds 1 --sustain-seconds 0"
$SHORT > gpurun_out/r02_short_plain.json 2> gpurun_out/r02_short_plain.err && \
 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv $SHORT > gpurun_out/r02_launches.log 2>&1
python tools/energy_once.py > /dev/null 2>&1 && \
 timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage2 -s 4 -f -o gpurun_out/r02_stage2_wide python tools/energy_once.py > gpurun_out/r02_wide_ncu.log 2>&1
cat gpurun_out/r02_pytest_gpu.txt gpurun_out/r02_smoke.txt gpurun_out/r02_stage2_quick.txt
cut -c1-400 gpurun_out/r02_bench_n1.json
