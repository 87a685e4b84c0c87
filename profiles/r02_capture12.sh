set -x
python -m pytest tests -m gpu -q -x -k "zinv or suh03 or bzi or star" > gpurun_out/r02_pytest_gpu_c12.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c12.log
python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c12_bzi77.json 2> gpurun_out/bench_c12.err; head -c 300 gpurun_out/r02_bench_c12_bzi77.json; echo
