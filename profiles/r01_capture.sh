# Round-1 evidence capture (run on the GPU box): launch list, ncu --set full of the tiled large-n kernels and the field-map kernel, secondary benches.
set -x
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
WP="python bench.py --workload woodpile1111 --kpoints 1 --steps 1 --warmup 1 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_launches_woodpile.csv $WP > gpurun_out/ncu_l.log 2>&1
$NCU -k regex:zhb_panel -s 3 -c 1 -o gpurun_out/r01_zhb_panel $WP > gpurun_out/ncu_1.log 2>&1
$NCU -k regex:zqr_global -c 1 -o gpurun_out/r01_zqr_tiled $WP > gpurun_out/ncu_2.log 2>&1
$NCU -k regex:zinvb_panel -s 3 -c 1 -o gpurun_out/r01_zinvb_panel $WP > gpurun_out/ncu_3.log 2>&1
$NCU -k regex:zrot_apply -c 1 -o gpurun_out/r01_zrot_strip $WP > gpurun_out/ncu_4.log 2>&1
$NCU -k regex:fld_grid_body -c 1 -o gpurun_out/r01_fld_grid python profiles/fields_bench.py 3 > gpurun_out/ncu_5.log 2>&1
python profiles/fields_bench.py 17 > gpurun_out/fields27.jsonl 2>&1
python profiles/aux_bench.py > gpurun_out/aux27.jsonl 2>&1
python bench.py --workload woodpile1111 --kpoints 4 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench27_wp.json 2>&1
python bench.py > gpurun_out/bench27.json 2>&1
