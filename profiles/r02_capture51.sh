set -x
KH_FUZZ_TRIALS=600 KH_FUZZ_SEED=7 KH_FUZZ_LOG=gpurun_out/r02_fuzz_seed7.jsonl timeout 2400 python -m pytest tests/test_fuzz_parity.py -m gpu -q -k random_structures > gpurun_out/r02_fuzz_seed7_pytest.log 2>&1; tail -3 gpurun_out/r02_fuzz_seed7_pytest.log | cut -c 1-400
tail -1 gpurun_out/r02_fuzz_seed7.jsonl
KH_FUZZ_TRIALS=300 KH_FUZZ_LOG=gpurun_out/r02_fuzz.jsonl timeout 2400 python -m pytest tests/test_fuzz_parity.py -m gpu -q -k random_structures > gpurun_out/r02_fuzz_pytest.log 2>&1; tail -3 gpurun_out/r02_fuzz_pytest.log | cut -c 1-400
tail -1 gpurun_out/r02_fuzz.jsonl
