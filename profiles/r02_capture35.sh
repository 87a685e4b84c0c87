set -x
for f in 0 1 2 3; do
KH_ZINV_L2_FLAGS=$f python profiles/fields_bench.py 17 > gpurun_out/r02_fields_c35_$f.jsonl 2>> gpurun_out/fields_c35.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_fields_c35_$f.jsonl').read());print($f, d['ms_solve'], d['ms_fields'], d['field_kernels_ms']['zinv'], d['solve_kernels_ms']['zinv'])"
done
