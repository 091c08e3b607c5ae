set -e
timeout 300 python profiles/scripts/one_step.py fast fast32 > gpurun_out/one_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'chan_fast' --launch-skip 1 -c 2 -o gpurun_out/r02_chan -f python profiles/scripts/one_step.py fast fast32 > gpurun_out/ncu_chan.log 2>&1
tail -3 gpurun_out/one_step.log
