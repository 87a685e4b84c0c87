set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_c14.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c14.log
python bench.py --no-cpu > gpurun_out/r02_bench_c14_bzi77.json 2> gpurun_out/bench_c14.err; head -c 300 gpurun_out/r02_bench_c14_bzi77.json; echo
