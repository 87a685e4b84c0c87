set -x
for v in 0 2 3; do KH_ZINV_L2_SMALL=$v python bench.py --no-cpu --no-extra --steps 3 > gpurun_out/r02_bench_c15_l2small$v.json 2>> gpurun_out/bench_c15.err; head -c 260 gpurun_out/r02_bench_c15_l2small$v.json; echo; done
