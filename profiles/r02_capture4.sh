# Round 2, capture 4: accumulator-init addends, column-restricted bdmul, relative series tolerance / q <= 8: tests, accuracy, bench, launch list, zgemm DRAM traffic
set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_c4.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c4.log
python profiles/accuracy_probe.py > gpurun_out/r02_accuracy_c4.jsonl 2> gpurun_out/acc_c4.err; cut -c 1-200 gpurun_out/r02_accuracy_c4.jsonl
python bench.py --no-cpu > gpurun_out/r02_bench_c4_bzi77.json 2> gpurun_out/bench_c4.err; head -c 400 gpurun_out/r02_bench_c4_bzi77.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_l.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:zgemm -c 80 --csv --log-file gpurun_out/r02_zgemm_dram_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_t.log 2>&1
