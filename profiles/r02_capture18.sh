set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_c18.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c18.log
python profiles/worst_energy_probe.py > gpurun_out/r02_worst_energy.jsonl 2> gpurun_out/worst.err; cut -c 1-300 gpurun_out/r02_worst_energy.jsonl
KHEPRI_B200_METHOD=eig python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c18_bzi77_eig.json 2> gpurun_out/bench_c18.err; head -c 260 gpurun_out/r02_bench_c18_bzi77_eig.json; echo
python profiles/fields_bench.py 51 > gpurun_out/r02_fields_final_51.jsonl 2> gpurun_out/fields_c18.err; cut -c 1-260 gpurun_out/r02_fields_final_51.jsonl
