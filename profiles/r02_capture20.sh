set -x
python bench.py --workload woodpile1111 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_bench_c20_woodpile.json 2> gpurun_out/c20.err; head -c 200 gpurun_out/r02_bench_c20_woodpile.json; echo
python bench.py --workload suh03 --no-cpu --no-extra > gpurun_out/r02_bench_c20_suh03.json 2>> gpurun_out/c20.err; head -c 200 gpurun_out/r02_bench_c20_suh03.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_c20_woodpile.csv python bench.py --workload woodpile1111 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_w.log 2>&1
