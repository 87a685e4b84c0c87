set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02_bench_c48_4gpu.json 2> gpurun_out/c48.err
head -c 300 gpurun_out/r02_bench_c48_4gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 2 --warmup 1 --impl reference > gpurun_out/r02_bench_c48_4gpu_ref.json 2>> gpurun_out/c48.err
head -c 200 gpurun_out/r02_bench_c48_4gpu_ref.json; echo
