set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/c57_tests.txt; cat gpurun_out/c57_tests.txt
KH_FUZZ_TRIALS=372 KH_FUZZ_SEED=7 KH_FUZZ_LOG=gpurun_out/r02_fuzz_seed7.jsonl timeout 2400 python -m pytest tests/test_fuzz_parity.py -m gpu -q -k random_structures > gpurun_out/r02_fuzz_seed7_pytest.log 2>&1; tail -3 gpurun_out/r02_fuzz_seed7_pytest.log | cut -c 1-400
tail -1 gpurun_out/r02_fuzz_seed7.jsonl
python profiles/fields_bench.py 17 > gpurun_out/r02_fields_c57_17.jsonl 2>/dev/null; cut -c 1-200 gpurun_out/r02_fields_c57_17.jsonl
python profiles/fields_bench.py 51 > gpurun_out/r02_fields_c57_51.jsonl 2>/dev/null; cut -c 1-200 gpurun_out/r02_fields_c57_51.jsonl
