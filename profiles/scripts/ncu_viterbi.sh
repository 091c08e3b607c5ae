set -e
timeout 300 python profiles/scripts/etsi_times.py > gpurun_out/etsi_times.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'etsi_decode_kernel' --launch-skip 2 -c 1 -o gpurun_out/r02_viterbi -f python profiles/scripts/etsi_times.py > gpurun_out/ncu_vit.log 2>&1
cat gpurun_out/etsi_times.log | tail -3
