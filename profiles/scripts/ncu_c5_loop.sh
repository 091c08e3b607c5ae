set -e
TB200_C5_BLOCKS=1 timeout 300 python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/c5_plain.log 2>&1
TB200_C5_BLOCKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'etsi_loop_kernel|etsi_mf_kernel' --launch-skip 4 -c 2 -o gpurun_out/r02_c5_loop -f python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/ncu_c5_loop.log 2>&1
tail -2 gpurun_out/ncu_c5_loop.log
