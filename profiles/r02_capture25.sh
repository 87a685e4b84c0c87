set -x
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:fld_grid_body -c 1 -o gpurun_out/r02_fld_grid_c25 python profiles/fields_bench.py 17 > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log
ls -la gpurun_out/*.ncu-rep
