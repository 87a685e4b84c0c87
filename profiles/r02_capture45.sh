set -x
python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c45.json 2> gpurun_out/c45.err; head -c 250 gpurun_out/r02_bench_c45.json; echo
KH_FUZZ_TRIALS=300 KH_FUZZ_LOG=gpurun_out/r02_fuzz.jsonl timeout 1500 python -m pytest tests/test_fuzz_parity.py -m gpu -q > gpurun_out/r02_fuzz_pytest.log 2>&1; tail -4 gpurun_out/r02_fuzz_pytest.log | cut -c 1-400
tail -1 gpurun_out/r02_fuzz.jsonl; tail -1 gpurun_out/r02_fuzz.jsonl.special
python -m pytest tests -m gpu -q 2>&1 | tail -3
