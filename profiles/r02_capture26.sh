set -x
python -m pytest tests -m gpu -q -k "field or twisted" > gpurun_out/r02_pytest_gpu_c26.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c26.log
python profiles/fields_bench.py 17 > gpurun_out/r02_fields_c26_17.jsonl 2> gpurun_out/fields_c26.err; cut -c 1-300 gpurun_out/r02_fields_c26_17.jsonl
python profiles/fields_bench.py 51 > gpurun_out/r02_fields_c26_51.jsonl 2>> gpurun_out/fields_c26.err; cut -c 1-300 gpurun_out/r02_fields_c26_51.jsonl
