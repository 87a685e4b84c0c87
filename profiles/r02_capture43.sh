set -x
KH_FUZZ_TRIALS=264 KH_FUZZ_LOG=gpurun_out/r02_fuzz.jsonl timeout 1500 python -m pytest tests/test_fuzz_parity.py -m gpu -q > gpurun_out/r02_fuzz_pytest.log 2>&1; tail -4 gpurun_out/r02_fuzz_pytest.log | cut -c 1-400
tail -1 gpurun_out/r02_fuzz.jsonl; tail -1 gpurun_out/r02_fuzz.jsonl.special
python profiles/guard_probe.py > gpurun_out/r02_guard_probe.jsonl 2> gpurun_out/guard.err; cat gpurun_out/r02_guard_probe.jsonl | cut -c 1-300
