set -e
# general-rate Viterbi (TCH/4.8 is its first launch after two warm-ups) and the AACH ML decoder
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'etsi_decode_gen_kernel|etsi_aach_kernel' --launch-skip 2 -c 1 -o gpurun_out/r02_codec_gen -f python profiles/scripts/etsi_times.py > gpurun_out/ncu_codec_gen.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'etsi_aach_kernel' --launch-skip 2 -c 1 -o gpurun_out/r02_aach -f python profiles/scripts/etsi_times.py > gpurun_out/ncu_aach.log 2>&1
tail -2 gpurun_out/ncu_aach.log
