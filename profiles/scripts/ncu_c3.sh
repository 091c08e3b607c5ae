set -e
timeout 300 python profiles/scripts/c3_once.py 3 > gpurun_out/c3_once.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'burst_batch' --launch-skip 2 -c 1 -o gpurun_out/r02_c3 -f python profiles/scripts/c3_once.py 3 > gpurun_out/ncu_c3.log 2>&1
tail -1 gpurun_out/c3_once.log
