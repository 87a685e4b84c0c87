set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_c33_8gpu.json 2> gpurun_out/c33.err
head -c 300 gpurun_out/r02_bench_c33_8gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 3 --workload bzi77-full --no-cpu > gpurun_out/r02_bench_c33_8gpu_full.json 2>> gpurun_out/c33.err
head -c 300 gpurun_out/r02_bench_c33_8gpu_full.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 2 --warmup 1 --impl reference > gpurun_out/r02_bench_c33_8gpu_ref.json 2>> gpurun_out/c33.err
head -c 300 gpurun_out/r02_bench_c33_8gpu_ref.json; echo
tail -c 400 gpurun_out/c33.err
