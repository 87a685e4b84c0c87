set -x
KH_FUZZ_TRIALS=120 KH_FUZZ_LOG=gpurun_out/r02_fuzz3.jsonl timeout 1500 python -m pytest tests/test_fuzz_parity.py -m gpu -q -x -k "twisted or analytical" > gpurun_out/r02_fuzz3_pytest.log 2>&1; tail -12 gpurun_out/r02_fuzz3_pytest.log | cut -c 1-500
cat gpurun_out/r02_fuzz3.jsonl.twisted gpurun_out/r02_fuzz3.jsonl.analytical
