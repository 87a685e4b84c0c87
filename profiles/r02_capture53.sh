set -x
KH_FUZZ_TRIALS=360 KH_FUZZ_SEED=7 KH_FUZZ_LOG=gpurun_out/r02_fuzz_seed7.jsonl timeout 2400 python -m pytest tests/test_fuzz_parity.py -m gpu -q -k random_structures > gpurun_out/r02_fuzz_seed7_pytest.log 2>&1; tail -6 gpurun_out/r02_fuzz_seed7_pytest.log | cut -c 1-900
tail -1 gpurun_out/r02_fuzz_seed7.jsonl
