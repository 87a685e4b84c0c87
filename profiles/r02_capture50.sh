set -x
KH_FUZZ_TRIALS=600 KH_FUZZ_SEED=7 KH_FUZZ_LOG=gpurun_out/r02_fuzz_seed7.jsonl timeout 2400 python -m pytest tests/test_fuzz_parity.py -m gpu -q > gpurun_out/r02_fuzz_seed7_pytest.log 2>&1; tail -6 gpurun_out/r02_fuzz_seed7_pytest.log | cut -c 1-700
tail -1 gpurun_out/r02_fuzz_seed7.jsonl; tail -1 gpurun_out/r02_fuzz_seed7.jsonl.special; cat gpurun_out/r02_fuzz_seed7.jsonl.twisted gpurun_out/r02_fuzz_seed7.jsonl.analytical
