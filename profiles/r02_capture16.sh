set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_c16.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c16.log
python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c16_bzi77.json 2> gpurun_out/bench_c16.err; head -c 260 gpurun_out/r02_bench_c16_bzi77.json; echo
