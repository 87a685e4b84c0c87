set -x
python bench.py --no-cpu > gpurun_out/r02_bench_c41_bzi77.json 2> gpurun_out/c41.err; head -c 200 gpurun_out/r02_bench_c41_bzi77.json; echo
python bench.py --workload woodpile1111 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_bench_c41_woodpile.json 2>> gpurun_out/c41.err; head -c 200 gpurun_out/r02_bench_c41_woodpile.json; echo
python bench.py --workload suh03 --no-cpu --no-extra > gpurun_out/r02_bench_c41_suh03.json 2>> gpurun_out/c41.err; head -c 200 gpurun_out/r02_bench_c41_suh03.json; echo
tail -3 gpurun_out/c41.err
