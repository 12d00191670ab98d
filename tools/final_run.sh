set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/bench_stages.py > gpurun_out/stage_table.txt 2>&1
python tools/secondary_once.py > gpurun_out/secondary_plain.log 2>&1 && \
 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'energy_kernel|heatmap|overlay|ciou|resize_mask|triplet|tile|normalize' -s 9 -c 12 -f -o gpurun_out/secondary python tools/secondary_once.py > gpurun_out/secondary_ncu.log 2>&1
SHORT="python bench.py --steps 2 --warmup 3 --frames 2048 --no-cpu-baseline --e2e-steps 1 --e2e-frames 64"
$SHORT > gpurun_out/short_plain.json 2> gpurun_out/short_plain.err && \
 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/launches.log 2>&1 && \
 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused -s 3 -c 2 -f -o gpurun_out/fused $SHORT > gpurun_out/fused_ncu.log 2>&1
ls -la gpurun_out | tail -20
cat gpurun_out/pytest_gpu.txt
