set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2
KH_FUZZ_TRIALS=400 KH_FUZZ_SEED=7 KH_FUZZ_LOG=gpurun_out/r02_fuzz_seed7.jsonl timeout 2400 python -m pytest tests/test_fuzz_parity.py -m gpu -q -k random_structures > gpurun_out/r02_fuzz_seed7_pytest.log 2>&1; tail -3 gpurun_out/r02_fuzz_seed7_pytest.log | cut -c 1-400
tail -1 gpurun_out/r02_fuzz_seed7.jsonl
python profiles/fields_bench.py 17 > gpurun_out/r02_fields_c52_17.jsonl 2>/dev/null; cut -c 1-260 gpurun_out/r02_fields_c52_17.jsonl
python profiles/fields_bench.py 51 > gpurun_out/r02_fields_c52_51.jsonl 2>/dev/null; cut -c 1-260 gpurun_out/r02_fields_c52_51.jsonl
