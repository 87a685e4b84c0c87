# Round 2, capture 8: TMA-staged inverse / Hessenberg / replay, acquire loads + stress tests of the QR relay, C5 field maps at 17 and 51 frequencies
set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_c8.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_c8.log
python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c8_bzi77.json 2> gpurun_out/bench_c8.err; head -c 300 gpurun_out/r02_bench_c8_bzi77.json; echo
KHEPRI_B200_METHOD=eig python bench.py --no-cpu --no-extra > gpurun_out/r02_bench_c8_bzi77_eig.json 2>> gpurun_out/bench_c8.err; head -c 300 gpurun_out/r02_bench_c8_bzi77_eig.json; echo
python profiles/fields_bench.py 17 > gpurun_out/r02_fields_c8_17.jsonl 2> gpurun_out/fields_c8.err; cut -c 1-400 gpurun_out/r02_fields_c8_17.jsonl
python profiles/fields_bench.py 51 > gpurun_out/r02_fields_c8_51.jsonl 2>> gpurun_out/fields_c8.err; cut -c 1-400 gpurun_out/r02_fields_c8_51.jsonl
