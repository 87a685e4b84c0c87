# Round 2, capture 7: accumulator-init addends, column-restricted bdmul, relative series tolerance / q <= 8: tests, accuracy, bench, launch list, zgemm DRAM traffic
set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_c7.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c7.log
python bench.py --no-cpu > gpurun_out/r02_bench_c7_bzi77.json 2> gpurun_out/bench_c7.err; head -c 400 gpurun_out/r02_bench_c7_bzi77.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c7.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_l.log 2>&1
