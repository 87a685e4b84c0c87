set -x
KH_ZGEMM_TMA=1 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_c30.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c30.log
for m in 0 1 0 1; do
KH_ZGEMM_TMA=$m timeout 300 python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c30_$m.json 2>> gpurun_out/c30.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_c30_$m.json').read().strip().splitlines()[-1]);print($m, d['value'], d['roofline']['kernels']['zgemm'])"
done
