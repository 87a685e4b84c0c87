set -x
python -m pytest tests -m gpu -q -x -k "zgemm or golden or star" 2>&1 | tail -2
for i in 1 2; do python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c46_$i.json 2>> gpurun_out/c46.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_c46_$i.json').read().strip().splitlines()[-1]);print(d['value'], d['roofline']['kernels']['zgemm'])"; done
python bench.py --workload suh03 --no-cpu --no-extra > gpurun_out/r02_bench_c46_suh03.json 2>> gpurun_out/c46.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_c46_suh03.json').read().strip().splitlines()[-1]);print(d['value'], d['roofline']['kernels']['zgemm'])"
