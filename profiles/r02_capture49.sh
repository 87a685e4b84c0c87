set -x
python bench.py --no-cpu > gpurun_out/r02_bench_c49.json 2> gpurun_out/c49.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_c49.json').read().strip().splitlines()[-1]);print(d['value'], [round(e.get('solves_per_s',0),1) for e in d['extra'][:6]], [e.get('eig_fallbacks') for e in d['extra'][:5]])"
KH_FUZZ_TRIALS=150 timeout 1500 python -m pytest tests/test_fuzz_parity.py -m gpu -q 2>&1 | tail -3
python profiles/guard_probe.py 2>/dev/null | cut -c 1-200
python -m pytest tests -m gpu -q -k "not fuzz" 2>&1 | tail -2
