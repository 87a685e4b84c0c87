set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_c34.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c34.log
python profiles/fields_bench.py 17 > gpurun_out/r02_fields_c34_17.jsonl 2> gpurun_out/fields_c34.err; cut -c 1-300 gpurun_out/r02_fields_c34_17.jsonl
python profiles/fields_bench.py 51 > gpurun_out/r02_fields_c34_51.jsonl 2>> gpurun_out/fields_c34.err; cut -c 1-300 gpurun_out/r02_fields_c34_51.jsonl
python profiles/fields_bench.py 1 > gpurun_out/r02_fields_c34_1.jsonl 2>> gpurun_out/fields_c34.err; cut -c 1-300 gpurun_out/r02_fields_c34_1.jsonl
