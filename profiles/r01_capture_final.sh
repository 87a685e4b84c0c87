# Round-1 closing capture (final code of the round: E = XB A^-1 schedule, flux-column forward chain): default bench + reference arm,
# launch list and GEMM DRAM traffic of the same command, the other BASELINE configs.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r01_pytest_gpu_v41.log 2>&1; tail -2 gpurun_out/r01_pytest_gpu_v41.log
python bench.py > gpurun_out/r01_bench_v41_bzi77.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/r01_bench_v41_bzi77.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_v41_reference_arm.json 2>&1
python bench.py --workload suh03 --no-cpu > gpurun_out/r01_bench_v41_suh03.json 2>&1
python bench.py --workload woodpile1111 --steps 2 --warmup 1 --no-cpu > gpurun_out/r01_bench_v41_woodpile1111.json 2>&1
python profiles/aux_bench.py > gpurun_out/r01_aux_bench_v41.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches_v41.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_l.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:zgemm --csv --log-file gpurun_out/r01_zgemm_dram_launches_v41.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_t.log 2>&1
head -c 700 gpurun_out/r01_aux_bench_v41.jsonl
