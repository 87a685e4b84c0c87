# Round-1 final capture (run on the GPU box): parity tests, default bench + reference arm, the other BASELINE configs,
# secondary benches, DRAM traffic of the GEMM launches.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r01_pytest_gpu_v38.log 2>&1; tail -2 gpurun_out/r01_pytest_gpu_v38.log
python bench.py > gpurun_out/r01_bench_v38_bzi77.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/r01_bench_v38_bzi77.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_v38_reference_arm.json 2>&1
python bench.py --workload suh03 --no-cpu > gpurun_out/r01_bench_v38_suh03.json 2>&1
python bench.py --workload woodpile1111 --steps 2 --warmup 1 --no-cpu > gpurun_out/r01_bench_v38_woodpile1111.json 2>&1
python profiles/aux_bench.py > gpurun_out/r01_aux_bench_v38.jsonl 2>&1
python profiles/fields_bench.py 17 > gpurun_out/r01_fields_bench_v38.jsonl 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:zgemm --csv --log-file gpurun_out/r01_zgemm_dram_launches_v38.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_t.log 2>&1
tail -n 3 gpurun_out/r01_aux_bench_v38.jsonl | cut -c1-300
