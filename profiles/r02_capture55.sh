# ncu --set full of the two inverse kernels added late in round 2: the cluster / DSMEM inverse (C5 solve at one frequency) and the
# 32-pivot block column of the blocked inverse (woodpile 11x11 step)
set -x
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:zinv_cluster4_body -s 3 -c 1 -o gpurun_out/r02_zinv_cluster_final python profiles/fields_bench.py 1 > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:zinvb_panel32a_body -s 3 -c 1 -o gpurun_out/r02_zinvb_panel32_final python bench.py --workload woodpile1111 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_p.log 2>&1; tail -2 gpurun_out/ncu_p.log
ls -la gpurun_out/r02_zinv_cluster_final.ncu-rep gpurun_out/r02_zinvb_panel32_final.ncu-rep
