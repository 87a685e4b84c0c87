set -x
KH_ZINV_LA=1 python -m pytest tests -m gpu -q -k "zinv or golden or star" > gpurun_out/r02_pytest_gpu_c27.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c27.log
python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c27_base.json 2> gpurun_out/c27.err; head -c 200 gpurun_out/r02_bench_c27_base.json; echo
KH_ZINV_LA=1 python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c27_la.json 2>> gpurun_out/c27.err; head -c 200 gpurun_out/r02_bench_c27_la.json; echo
python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c27_base2.json 2>> gpurun_out/c27.err; head -c 200 gpurun_out/r02_bench_c27_base2.json; echo
KH_ZINV_LA=1 python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c27_la2.json 2>> gpurun_out/c27.err; head -c 200 gpurun_out/r02_bench_c27_la2.json; echo
