# v43: kh_fields_fourier_batch (return_fourier=True). GPU parity suite, memcheck of the new path, default bench.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r01_pytest_gpu_v43.log 2>&1; tail -2 gpurun_out/r01_pytest_gpu_v43.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fields_twisted_sharding.py -q -m gpu -k "return_fourier" > gpurun_out/r01_memcheck_fourier_v43.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/r01_memcheck_fourier_v43.log
python bench.py > gpurun_out/r01_bench_v43_bzi77.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/r01_bench_v43_bzi77.json
