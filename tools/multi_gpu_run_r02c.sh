# One 8-GPU box session with the final library of round 2 (run under `gpurun --gpus 8` from the repo root): the bench at N = 8
# and N = 4 (one rank per GPU), at N = 8 from ONE process, the reference arm under torchrun, and the multi-GPU tests.
set -x
for N in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02c_bench_n$N.json 2> gpurun_out/r02c_bench_n$N.err
done
python bench.py --gpus 8 --single-process --steps 20 --warmup 5 > gpurun_out/r02c_bench_single_process_n8.json 2> gpurun_out/r02c_bench_sp8.err
python -m pytest tests/test_gpu_group.py tests/test_gpu_multirank.py -q 2>&1 | tail -4 > gpurun_out/r02c_pytest_multi_gpu.txt
tail -2 gpurun_out/r02c_pytest_multi_gpu.txt
ls -la gpurun_out | grep "r02c_bench"
