set -e
timeout 300 python profiles/scripts/c4_once.py 2 > gpurun_out/c4_once.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'polybank800|refchan_' --launch-skip 4 -c 4 -o gpurun_out/r02_c4 -f python profiles/scripts/c4_once.py 2 > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/c4_once.log
