set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2
python profiles/fields_bench.py 17 > gpurun_out/r02_fields_c54_17.jsonl 2>/dev/null; cut -c 1-200 gpurun_out/r02_fields_c54_17.jsonl
python profiles/fields_bench.py 51 > gpurun_out/r02_fields_c54_51.jsonl 2>/dev/null; cut -c 1-200 gpurun_out/r02_fields_c54_51.jsonl
python profiles/fields_bench.py 1 > gpurun_out/r02_fields_c54_1.jsonl 2>/dev/null; cut -c 1-200 gpurun_out/r02_fields_c54_1.jsonl
