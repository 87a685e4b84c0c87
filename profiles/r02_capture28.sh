set -x
KH_ZINV_LA=2 python -m pytest tests -m gpu -q -k "zinv or star" > gpurun_out/r02_pytest_gpu_c28.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c28.log
KH_ZINV_LA=3 python -m pytest tests -m gpu -q -k "zinv or star" > gpurun_out/r02_pytest_gpu_c28b.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c28b.log
for m in 0 2 3 0 2 3; do
KH_ZINV_LA=$m python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c28_$m.json 2>> gpurun_out/c28.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_c28_$m.json').read().strip().splitlines()[-1]);print($m, d['value'], d['roofline']['kernels']['zinv']['avg_launch_ms'])"
done
