set -x
python bench.py --workload suh03 --no-cpu --no-extra > gpurun_out/r02_bench_c31_suh03.json 2> gpurun_out/c31.err
python bench.py --workload woodpile1111 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_bench_c31_woodpile.json 2>> gpurun_out/c31.err
