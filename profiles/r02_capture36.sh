set -x
timeout 600 python -m pytest tests -m gpu -q -x -k "zinv" > gpurun_out/r02_pytest_gpu_c36.log 2>&1; tail -15 gpurun_out/r02_pytest_gpu_c36.log
