set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_c24.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c24.log
python profiles/fields_bench.py 17 > gpurun_out/r02_fields_c24_17.jsonl 2> gpurun_out/fields_c24.err; cut -c 1-400 gpurun_out/r02_fields_c24_17.jsonl
python profiles/fields_bench.py 51 > gpurun_out/r02_fields_c24_51.jsonl 2>> gpurun_out/fields_c24.err; cut -c 1-400 gpurun_out/r02_fields_c24_51.jsonl
python profiles/fields_bench.py 1 > gpurun_out/r02_fields_c24_1.jsonl 2>> gpurun_out/fields_c24.err; cut -c 1-400 gpurun_out/r02_fields_c24_1.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fld_grid_body -c 1 -o gpurun_out/r02_fld_grid_c24 python profiles/fields_bench.py 17 > /dev/null 2>> gpurun_out/fields_c24.err
ls -la gpurun_out/*.ncu-rep
