# Round-1 (second capture, after the DMMA inverse / unit-balanced GEMM): launch list of the default bench command and
# ncu --set full of the kernels that changed or lead the time shares.  Run on the GPU box: sh profiles/r01_capture2.sh
set -x
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
BZ="python bench.py --steps 1 --warmup 1 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches_v36.csv $BZ > gpurun_out/ncu_l.log 2>&1
$NCU -k regex:zinv_dmma_body -s 2 -c 1 -o gpurun_out/r01_zinv_dmma $BZ > gpurun_out/ncu_1.log 2>&1
$NCU -k regex:zgemm56u3 -s 20 -c 1 -o gpurun_out/r01_zgemm56u3 $BZ > gpurun_out/ncu_2.log 2>&1
$NCU -k regex:zrot_apply -c 1 -o gpurun_out/r01_zrot_smem $BZ > gpurun_out/ncu_3.log 2>&1
$NCU -k regex:zhessz -c 1 -o gpurun_out/r01_zhessz $BZ > gpurun_out/ncu_4.log 2>&1
ls -la gpurun_out
