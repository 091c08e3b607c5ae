# Round-2 profile set: every command first runs plain (exit 0), then under ncu.  Outputs in gpurun_out/.
set -e
timeout 300 python profiles/scripts/one_step.py fast fast32 > gpurun_out/one_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'decim_fast_kernel|chan_fast|symbols_kernel|decode_fused|tail_fused' --launch-skip 4 -c 8 -o gpurun_out/r02_c1 -f python profiles/scripts/one_step.py fast fast32 > gpurun_out/ncu_c1.log 2>&1
timeout 300 python profiles/scripts/c4_once.py 2 > gpurun_out/c4_once.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'polybank800|refchan_edge' --launch-skip 2 -c 2 -o gpurun_out/r02_c4 -f python profiles/scripts/c4_once.py 2 > gpurun_out/ncu_c4.log 2>&1
TB200_C5_BLOCKS=1 timeout 300 python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/c5_plain.log 2>&1
TB200_C5_BLOCKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'etsi_demod_kernel|etsi_find_bursts|etsi_stitch|etsi_extract' --launch-skip 8 -c 4 -o gpurun_out/r02_c5 -f python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/ncu_c5f.log 2>&1
# launch lists (share of the step per kernel)
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/b_plain.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fast.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_l1.log 2>&1
timeout 300 python bench.py --workload c4 --steps 2 --warmup 3 > gpurun_out/c4_plain.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 > gpurun_out/ncu_l4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/ncu_l5.log 2>&1
echo profiles done
