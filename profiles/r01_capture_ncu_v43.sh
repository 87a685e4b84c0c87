# ncu --set full of the three eigensolver kernels that lead the step by time (HEAD, bzi77 default bench), one launch each.
set -x
for k in zqr_packed zhessz zrot_apply; do
  timeout 170 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -c 1 -f -o gpurun_out/r01_${k}_v43 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
