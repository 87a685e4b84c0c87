# Round 2, capture 10: two GPUs of one box: default weak-scaling bench, the whole configs[1] job (strong scaling), reference arm under torchrun, scalar loop after host trimming
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c10_2gpu.json 2> gpurun_out/c10_a.err; head -c 500 gpurun_out/r02_bench_c10_2gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload bzi77-full --no-cpu > gpurun_out/r02_bench_c10_2gpu_full.json 2> gpurun_out/c10_b.err; head -c 500 gpurun_out/r02_bench_c10_2gpu_full.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_c10_2gpu_ref.json 2> gpurun_out/c10_c.err; head -c 300 gpurun_out/r02_bench_c10_2gpu_ref.json; echo
python bench.py --workload bzi77-full --no-cpu > gpurun_out/r02_bench_c10_1gpu_full.json 2> gpurun_out/c10_d.err; head -c 300 gpurun_out/r02_bench_c10_1gpu_full.json; echo
python profiles/scalar_loop_profile.py 2>&1 | head -4
tail -3 gpurun_out/c10_*.err
