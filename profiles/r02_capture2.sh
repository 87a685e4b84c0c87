# Round 2, capture 2: full GPU test-suite, default bench line (doubling method) with extras + reference arm, the whole configs[1]
# job, ncu launch list of one step, ncu --set full of one zgemm and one zinv launch of the same command.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_c2.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c2.log
python bench.py > gpurun_out/r02_bench_c2_bzi77.json 2> gpurun_out/bench_c2.err; head -c 600 gpurun_out/r02_bench_c2_bzi77.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_c2_reference_arm.json 2> gpurun_out/ref_c2.err; head -c 400 gpurun_out/r02_bench_c2_reference_arm.json; echo
python bench.py --workload bzi77-full --no-cpu > gpurun_out/r02_bench_c2_bzi77_full.json 2> gpurun_out/full_c2.err; head -c 600 gpurun_out/r02_bench_c2_bzi77_full.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:zgemm56u3 -s 40 -c 1 -o gpurun_out/r02_zgemm_c2 python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:zinv_dmma -s 6 -c 1 -o gpurun_out/r02_zinv_c2 python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_i.log 2>&1
ls -la gpurun_out | tail -12
