# Round 2, capture 1: doubling method (slice series + self star products) against the eigensolver method on bzi77, GPU parity tests.
set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_c1.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c1.log
for m in eig doubling; do
  KHEPRI_B200_METHOD=$m python bench.py --no-cpu > gpurun_out/r02_bench_c1_bzi77_$m.json 2> gpurun_out/bench_$m.err; tail -c 1500 gpurun_out/r02_bench_c1_bzi77_$m.json
done
for th in 4 6 12 16; do
  KHEPRI_B200_METHOD=doubling KHEPRI_B200_THETA=$th python bench.py --no-cpu --steps 3 > gpurun_out/r02_bench_c1_bzi77_doubling_th$th.json 2>> gpurun_out/bench_th.err; head -c 330 gpurun_out/r02_bench_c1_bzi77_doubling_th$th.json; echo
done
KHEPRI_B200_METHOD=doubling python bench.py --workload suh03 --no-cpu > gpurun_out/r02_bench_c1_suh03_doubling.json 2>&1; head -c 330 gpurun_out/r02_bench_c1_suh03_doubling.json; echo
KHEPRI_B200_METHOD=eig python bench.py --workload suh03 --no-cpu > gpurun_out/r02_bench_c1_suh03_eig.json 2>&1; head -c 330 gpurun_out/r02_bench_c1_suh03_eig.json; echo
KHEPRI_B200_METHOD=doubling python bench.py --workload woodpile1111 --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_c1_woodpile_doubling.json 2>&1; head -c 330 gpurun_out/r02_bench_c1_woodpile_doubling.json; echo
KHEPRI_B200_METHOD=eig python bench.py --workload woodpile1111 --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_c1_woodpile_eig.json 2>&1; head -c 330 gpurun_out/r02_bench_c1_woodpile_eig.json; echo
