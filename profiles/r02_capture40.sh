set -x
KH_FUZZ_TRIALS=264 KH_FUZZ_LOG=gpurun_out/r02_fuzz.jsonl timeout 1500 python -m pytest tests/test_fuzz_parity.py -m gpu -q -x > gpurun_out/r02_fuzz_pytest.log 2>&1; tail -5 gpurun_out/r02_fuzz_pytest.log
tail -1 gpurun_out/r02_fuzz.jsonl
python -m pytest tests -m gpu -q 2>&1 | tail -4
