set -x
KH_FUZZ_TRIALS=160 KH_FUZZ_LOG=gpurun_out/r02_fuzz2.jsonl timeout 1500 python -m pytest tests/test_fuzz_parity.py -m gpu -q -x -k special > gpurun_out/r02_fuzz2_pytest.log 2>&1; tail -12 gpurun_out/r02_fuzz2_pytest.log | cut -c 1-600
tail -1 gpurun_out/r02_fuzz2.jsonl.special
