"""C1 step replayed as CUDA graphs: two captured pipelines alternating on two streams."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tetraear_b200 import pipeline, synth

n_chunks, CH = 183, 131072
x = synth.make_capture(CH * n_chunks, 130, seed=1234).reshape(n_chunks, CH)
dev = torch.from_numpy(x.view(np.float32).reshape(n_chunks, CH, 2)).cuda()
for lanes in (2, 3, 4, 6):
    pipes = [pipeline.ReceivePipeline(2.4e6, CH, max_chunks=n_chunks, mode="fast32", n_lanes=lanes) for _ in range(2)]
    graphs = [p.capture(dev) for p in pipes]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for p in pipes:
        p.stats.zero_()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 200
    main = torch.cuda.current_stream()
    a.record(main)
    for s in streams:
        s.wait_stream(main)
    t0 = time.perf_counter()
    for i in range(K):
        with torch.cuda.stream(streams[i & 1]):
            graphs[i & 1].replay()
    t1 = time.perf_counter()
    for s in streams:
        main.wait_stream(s)
    b.record(main)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    crc = int(pipes[0].stats[1].item() + pipes[1].stats[1].item())
    print(f"lanes {lanes}: {ms:.4f} ms/step  {n_chunks * CH / ms / 1e3:.0f} Msamples/s  host {1e3 * (t1 - t0) / K:.4f} ms/step  crc_pass/step {crc / K:.1f}")
