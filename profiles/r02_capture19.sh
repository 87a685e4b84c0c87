# compute-sanitizer on the round-2 kernels (small shapes)
set -x
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitizer_r02.py > gpurun_out/r02_memcheck.log 2>&1; tail -6 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python profiles/racecheck_r02_dense.py > gpurun_out/r02_racecheck_dense.log 2>&1; tail -6 gpurun_out/r02_racecheck_dense.log
