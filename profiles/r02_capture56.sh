set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c56_2gpu.json 2> gpurun_out/c56.err
head -c 260 gpurun_out/r02_bench_c56_2gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 1 --warmup 3 --workload bzi77-full --no-cpu > gpurun_out/r02_bench_c56_2gpu_full.json 2>> gpurun_out/c56.err
head -c 260 gpurun_out/r02_bench_c56_2gpu_full.json; echo
