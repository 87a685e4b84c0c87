set -x
KH_ZINV_NB8=1 python -m pytest tests -m gpu -q -k "zinv or suh03 or oblique or analytical or random_structures" > gpurun_out/r02_pytest_gpu_c21.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c21.log
for v in 0 1 0 1; do KH_ZINV_NB8=$v python bench.py --workload suh03 --no-cpu --no-extra > gpurun_out/r02_bench_c21_suh03_nb$v.json 2>> gpurun_out/c21.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_c21_suh03_nb$v.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("NB8=$v","%.0f"%d["value"],{k:(vv['share_of_kernel_time'],round(vv.get('achieved_tflops',0),2)) for k,vv in r["kernels"].items() if k in('zgemm','zinv')})
PY
done
