set -e
timeout 300 python profiles/scripts/n3n4_times.py > gpurun_out/n3n4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:voice --launch-skip 3 -c 1 -o gpurun_out/r02_n3n4 -f python profiles/scripts/n3n4_times.py > gpurun_out/ncu_n3n4.log 2>&1
cat gpurun_out/n3n4.log | tail -2
