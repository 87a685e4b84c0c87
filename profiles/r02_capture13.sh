set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_c13.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c13.log
python profiles/fields_bench.py 51 > gpurun_out/r02_fields_c13_51.jsonl 2> gpurun_out/fields_c13.err; cut -c 1-300 gpurun_out/r02_fields_c13_51.jsonl
KH_ZINV_L2_MAXBATCH=0 python profiles/fields_bench.py 51 > gpurun_out/r02_fields_c13_51_blocked.jsonl 2>> gpurun_out/fields_c13.err; cut -c 1-300 gpurun_out/r02_fields_c13_51_blocked.jsonl
python profiles/fields_bench.py 17 > gpurun_out/r02_fields_c13_17.jsonl 2>> gpurun_out/fields_c13.err; cut -c 1-300 gpurun_out/r02_fields_c13_17.jsonl
