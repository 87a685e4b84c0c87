# compute-sanitizer on the round-2 kernels at HEAD (small shapes): incl. the cluster inverse, the 32-pivot block column, the batched field pipeline, the side stream
set -x
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitizer_r02.py > gpurun_out/r02_memcheck.log 2>&1; tail -8 gpurun_out/r02_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python profiles/racecheck_r02_dense.py > gpurun_out/r02_racecheck_dense.log 2>&1; tail -8 gpurun_out/r02_racecheck_dense.log
python -m pytest tests -m gpu -q -k "zinv" 2>&1 | tail -2
