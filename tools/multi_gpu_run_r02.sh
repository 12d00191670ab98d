# One 8-GPU box session (run under `gpurun --gpus 8` from the repo root): the host <-> device fabric probe at N = 1, 2, 4, 8,
# the bench at N = 2, 4, 8 (one rank per GPU) and at N = 8 from ONE process, and the multi-GPU tests.
set -x
nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.txt 2>&1
lscpu | head -25 > gpurun_out/r02_lscpu_8gpu.txt 2>&1
python tools/h2d_scale_probe.py > gpurun_out/r02_h2d_scale_n1.txt 2>&1
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) tools/h2d_scale_probe.py > gpurun_out/r02_h2d_scale_n$N.txt 2>&1
done
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
done
python bench.py --gpus 8 --single-process --steps 20 --warmup 5 > gpurun_out/r02_bench_single_process_n8.json 2> gpurun_out/r02_bench_sp8.err
python -m pytest tests/test_gpu_group.py tests/test_gpu_multirank.py -q 2>&1 | tail -4 > gpurun_out/r02_pytest_multi_gpu.txt
grep -h -v "^\*\|OMP_NUM\|^$\|Warn" gpurun_out/r02_h2d_scale_n*.txt
tail -2 gpurun_out/r02_pytest_multi_gpu.txt
ls -la gpurun_out | grep "r02_bench_n[248]\|sp8\|single"
