# Round-1 capture of HEAD (division-free Householder scalars / Wilkinson shift in zgeev): GPU parity tests, default bench,
# reference arm, the launch list of the same command, then the other BASELINE configs.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r01_pytest_gpu_v42.log 2>&1; tail -2 gpurun_out/r01_pytest_gpu_v42.log
python bench.py > gpurun_out/r01_bench_v42_bzi77.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/r01_bench_v42_bzi77.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_v42_reference_arm.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches_v42.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_l.log 2>&1
python bench.py --workload suh03 --no-cpu > gpurun_out/r01_bench_v42_suh03.json 2>&1
python bench.py --workload woodpile1111 --steps 2 --warmup 1 --no-cpu > gpurun_out/r01_bench_v42_woodpile1111.json 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r01_smoke_v42.log 2>&1; tail -1 gpurun_out/r01_smoke_v42.log
