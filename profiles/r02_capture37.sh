set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_c37.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c37.log
for nf in 1 17 51; do for m in 0 296; do
KH_ZINV_CLUSTER_MAXCTAS=$m python profiles/fields_bench.py $nf > gpurun_out/r02_fields_c37_${nf}_$m.jsonl 2>> gpurun_out/fields_c37.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_fields_c37_${nf}_$m.jsonl').read());print($nf, $m, round(d['ms_solve'],2), round(d['ms_fields'],2), d['field_kernels_ms']['zinv'], d['solve_kernels_ms']['zinv'])"
done; done
