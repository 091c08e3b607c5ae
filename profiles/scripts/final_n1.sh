timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 200 gpurun_out/bench_default.err
timeout 400 python bench.py --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 200 gpurun_out/bench_c4.err
timeout 400 python bench.py --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 200 gpurun_out/bench_c5.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 200 gpurun_out/bench_reference.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
