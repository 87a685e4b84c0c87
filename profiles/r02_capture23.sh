set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_c23.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c23.log
python bench.py --workload woodpile1111 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_bench_c23_woodpile.json 2> gpurun_out/c23.err; head -c 200 gpurun_out/r02_bench_c23_woodpile.json; echo
python bench.py --no-cpu > gpurun_out/r02_bench_c23_bzi77.json 2>> gpurun_out/c23.err; head -c 200 gpurun_out/r02_bench_c23_bzi77.json; echo
