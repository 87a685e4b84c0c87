# v45: kh_idft_batch / fourier.idft. GPU parity suite, memcheck of the new operator, default bench (unchanged hot path).
set -x
python -m pytest tests -m gpu -q > gpurun_out/r01_pytest_gpu_v45.log 2>&1; tail -2 gpurun_out/r01_pytest_gpu_v45.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fields_twisted_sharding.py -q -m gpu -k "idft" > gpurun_out/r01_memcheck_idft_v45.log 2>&1; echo memcheck rc=$?; tail -3 gpurun_out/r01_memcheck_idft_v45.log
python bench.py > gpurun_out/r01_bench_v45_bzi77.json 2> gpurun_out/bench.err; tail -c 200 gpurun_out/r01_bench_v45_bzi77.json
