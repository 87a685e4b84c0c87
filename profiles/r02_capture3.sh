# Round 2, capture 3: even/odd half-slice conversion, collapsed runs, fused Horner epilogue, flux-column gemv: tests, accuracy probe, bench.
set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_c3.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c3.log
python profiles/accuracy_probe.py > gpurun_out/r02_accuracy_c3.jsonl 2> gpurun_out/acc_c3.err; cut -c 1-330 gpurun_out/r02_accuracy_c3.jsonl
python bench.py --no-cpu > gpurun_out/r02_bench_c3_bzi77.json 2> gpurun_out/bench_c3.err; head -c 400 gpurun_out/r02_bench_c3_bzi77.json; echo
for th in 6 8 12; do KHEPRI_B200_THETA=$th python bench.py --no-cpu --no-extra --steps 3 > gpurun_out/r02_bench_c3_bzi77_th$th.json 2>> gpurun_out/bench_c3.err; head -c 300 gpurun_out/r02_bench_c3_bzi77_th$th.json; echo; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_l.log 2>&1
