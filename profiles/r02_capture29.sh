set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c29_2gpu.json 2> gpurun_out/c29.err
tail -c 600 gpurun_out/c29.err
head -c 300 gpurun_out/r02_bench_c29_2gpu.json
