# Round 2 closing capture (HEAD): GPU tests, smoke, default bench line (with CPU baseline = unmodified reference, extras), reference arm,
# whole configs[1] job, ncu launch list of one step, ncu --set full of one zgemm and one zinv launch of the same command.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
python profiles/accuracy_probe.py > gpurun_out/r02_accuracy_final.jsonl 2> gpurun_out/acc_final.err; cut -c 1-160 gpurun_out/r02_accuracy_final.jsonl
python profiles/fields_bench.py 51 > gpurun_out/r02_fields_final_51.jsonl 2> gpurun_out/fields_final.err; cut -c 1-300 gpurun_out/r02_fields_final_51.jsonl
python profiles/fields_bench.py 17 > gpurun_out/r02_fields_final_17.jsonl 2>> gpurun_out/fields_final.err; cut -c 1-300 gpurun_out/r02_fields_final_17.jsonl
python bench.py > gpurun_out/r02_bench_final_bzi77.json 2> gpurun_out/bench_final.err; head -c 300 gpurun_out/r02_bench_final_bzi77.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_final_reference_arm.json 2> gpurun_out/ref_final.err; head -c 300 gpurun_out/r02_bench_final_reference_arm.json; echo
python bench.py --workload bzi77-full --no-cpu > gpurun_out/r02_bench_final_bzi77_full.json 2> gpurun_out/full_final.err; head -c 300 gpurun_out/r02_bench_final_bzi77_full.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:zgemm56u3 -s 30 -c 1 -o gpurun_out/r02_zgemm_final python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:zinv_dmma -s 4 -c 1 -o gpurun_out/r02_zinv_final python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_i.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:zgemm -c 60 --csv --log-file gpurun_out/r02_zgemm_dram_final.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_t.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:fld_grid9_body -c 1 -o gpurun_out/r02_fld_grid_final python profiles/fields_bench.py 17 > gpurun_out/ncu_f.log 2>&1
python bench.py --workload woodpile1111 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_bench_final_woodpile.json 2> gpurun_out/wood_final.err; head -c 200 gpurun_out/r02_bench_final_woodpile.json; echo
python bench.py --workload suh03 --no-cpu --no-extra > gpurun_out/r02_bench_final_suh03.json 2> gpurun_out/suh_final.err; head -c 200 gpurun_out/r02_bench_final_suh03.json; echo
