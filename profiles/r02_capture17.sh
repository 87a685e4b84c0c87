set -x
KH_ZGEMM_NC10=1 python -m pytest tests -m gpu -q -k "zgemm or suh03 or bzi" > gpurun_out/r02_pytest_gpu_c17.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c17.log
for v in 0 1 0 1; do KH_ZGEMM_NC10=$v python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c17_nc$v.json 2>> gpurun_out/bench_c17.err; head -c 230 gpurun_out/r02_bench_c17_nc$v.json; echo; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_c17_nc$v.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("NC10=$v","%.0f"%d["value"],{k:(vv['share_of_kernel_time'],round(vv.get('achieved_tflops',0),2)) for k,vv in r["kernels"].items() if k in('zgemm','zinv')})
PY
done
