set -e
timeout 500 python bench.py --steps 2 --warmup 3 > gpurun_out/b_extras.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'burst_batch_kernel|voice|detect_|sync_corr_kernel|decode_kernel|psd_ema' -c 14 -o gpurun_out/r02_other -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_other.log 2>&1
echo done
