#!/usr/bin/env python
"""bench.py — IQ Msamples/s (and CRC-pass bursts/s) of the IQ->bursts receive path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c1|c3|c2]
                    [--mode exact|fast|fast32|fast-tc|fast32-tc|fast-tc4|fast32-tc4]

Workload c1 (default, BASELINE.json configs[0]): one synthetic TETRA-like downlink carrier, 10 s of
complex64 IQ at 2.4 MS/s = 183 whole chunks of 131072 samples (the ragged 0.1-chunk tail is
dropped, like a capture loop that stops on a chunk boundary), processed exactly as the
reference's receive loop does per chunk: SignalProcessor.process -> TetraDecoder.decode.
A "step" is one pass of the hot path over that 192 MB batch (larger than the 126 MB L2, so
no L2 flush is needed between steps).

One JSON line on stdout (rank 0).  `value` = device-resident throughput, `e2e` = the same
through the host-buffer API with H2D/D2H inside the timed region, `roofline` = the dominant
stage (decimator) against the measured HBM peak, `cpu_baseline` = the oracle port on the
box's host cores over a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHUNK = 131072
FS = 2.4e6
C1_CHUNKS = 183          # floor(24e6 / 131072)
DECIM_TRAFFIC_BYTES = 207979008   # dram__bytes_read.sum + dram__bytes_write.sum of decim_fast_kernel<10>, one 183-chunk launch,
                                   # ncu --set full (profiles/r02_ncu_c1.md: 191.96 MB + 16.02 MB); the -tc modes report null


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------ workload
def make_c1(seed=1234, n_chunks=C1_CHUNKS):
    from tetraear_b200 import synth
    iq = synth.make_capture(n_chunks * CHUNK, 130, seed=seed)
    return iq.reshape(n_chunks, CHUNK)


# ------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """process + decode of a list of chunks with the oracle port; returns (samples, crc_pass, seconds)."""
    from oracle import oracle
    chunks, fs = args
    t0 = time.perf_counter()
    crc = 0
    n = 0
    for ch in chunks:
        d, _, _ = oracle.process(ch, fs, 0.0)
        _, st = oracle.decode(d)
        crc += st["crc_pass"]
        n += len(ch)
    return n, crc, time.perf_counter() - t0


_REF = None


def _ref_worker(args):
    """process + decode of a list of chunks with the UNMODIFIED reference classes (oracle/_ref, staged by
    oracle/make_ref.py); returns (samples, crc_pass, seconds)."""
    global _REF
    import logging
    logging.disable(logging.CRITICAL)
    if _REF is None:
        from oracle import ref_bridge
        _REF = ref_bridge.load()
    SP, TD, _ = _REF
    chunks, fs = args
    t0 = time.perf_counter()
    crc = 0
    n = 0
    for ch in chunks:
        sp = SP(fs)
        d = sp.process(ch, 0)
        dec = TD(auto_decrypt=False)
        if len(d) >= 255:                      # the receive loop's rule, ui/modern.py:2074-2077
            dec.decode(d)
        crc += dec.protocol_parser.stats["crc_pass"]
        n += len(ch)
    return n, crc, time.perf_counter() - t0


def _ref_burst_worker(rows):
    """TetraProtocolParser.parse_burst of the unmodified reference over a block of 255-dibit rows."""
    global _REF
    import logging
    logging.disable(logging.CRITICAL)
    if _REF is None:
        from oracle import ref_bridge
        _REF = ref_bridge.load()
    p = _REF[2]()
    t0 = time.perf_counter()
    ok = 0
    for r in rows:
        b = p.parse_burst(r)
        ok += bool(b is not None and b.crc_ok)
    return len(rows), ok, time.perf_counter() - t0


def _psd_worker(args):
    """The reference's per-frame spectrum block (ui/modern.py:1928-1941 restated in oracle.psd_frame: Hann window, fft,
    fftshift, dB) + the GUI's dB-domain EMA (:929-935) over dense frames."""
    from oracle import oracle
    x, n_fft = args
    t0 = time.perf_counter()
    S = None
    for k in range(len(x) // n_fft):
        out = oracle.psd_frame(x[k * n_fft:(k + 1) * n_fft], FS, n_fft)
        P = out[1] if isinstance(out, tuple) else out
        S = P if S is None else S * 0.85 + P * 0.15
    return len(x) // n_fft * n_fft, time.perf_counter() - t0


def reference_available():
    from oracle import ref_bridge
    return ref_bridge.available()


def cpu_run(x, cores, reps=1, kind="port"):
    """The CPU implementation of the path over the chunks of x with `cores` worker processes: kind "reference" = the
    unmodified reference classes, "port" = the oracle's C + numpy restatement.  Returns (Msamples/s, bursts/s, wall s)."""
    import multiprocessing as mp
    worker = _cpu_worker
    if kind == "reference":
        worker = _ref_worker
    else:
        from oracle import oracle
        oracle.lib()                          # compile the C part before forking
    parts = [list(x[i::cores]) for i in range(cores)]
    parts = [p for p in parts if p]
    best = None
    ctx = mp.get_context("fork")
    with ctx.Pool(len(parts)) as pool:
        pool.map(worker, [([p[0][:4096]], FS) for p in parts])      # warm the workers
        for _ in range(reps):
            t0 = time.perf_counter()
            res = pool.map(worker, [(p, FS) for p in parts])
            wall = time.perf_counter() - t0
            n = sum(r[0] for r in res)
            crc = sum(r[1] for r in res)
            if best is None or wall < best[2]:
                best = (n / wall / 1e6, crc / wall, wall)
    return best


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            # "under load": the upper half of the samples (idle samples before/after are excluded)
            top = sorted(sm)[len(sm) // 2:]
            out = {"sm_mhz": float(np.median(top)), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------ GPU arm
def _cpulist(text):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def bind_to_gpu_numa(torch, local_rank):
    """Pin this rank to the CPUs of its GPU's NUMA node, so that the pinned staging buffer it allocates next is placed
    on that node (matters for the end-to-end figure when several ranks copy at once).  The node comes from sysfs
    (/sys/bus/pci/devices/<gpu>/numa_node + /sys/devices/system/node/nodeN/cpulist, intersected with the CPUs this
    process may use); NVML's affinity mask, which inside a container is often just the container's CPU set, is the
    fallback.  Returns (number of CPUs bound to, numa node or -1)."""
    allowed = set(os.sched_getaffinity(0))
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = sorted(set(_cpulist(f.read())) & allowed)
            if cpus:
                os.sched_setaffinity(0, cpus)
                return len(cpus), node
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        n = os.cpu_count()
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = sorted({i for i in range(n) if (mask[i // 64] >> (i % 64)) & 1} & allowed)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus), -1
    except Exception:
        return 0, -1


def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    torch.cuda.set_device(local_rank)
    from tetraear_b200 import pipeline, _lib
    all_cpus = os.sched_getaffinity(0)
    numa_cpus, numa_node = bind_to_gpu_numa(torch, local_rank) if world > 1 else (0, -1)

    # weak scaling over capture segments: the job's capture is world x 10 s, rank r owns segment r (183 chunks); chunks
    # are stateless (ui/modern.py:1915-1919), so the segments need no exchange — the only communication is the gather of
    # the frame records on rank 0 (NCCL, inside the timed region, every step)
    x = make_c1(seed=1234 + rank)
    n_chunks = x.shape[0]
    samples_per_step = n_chunks * CHUNK
    pipe = pipeline.ReceivePipeline(FS, CHUNK, max_chunks=n_chunks, mode=args.mode, device=local_rank, n_lanes=args.lanes)
    host = torch.from_numpy(x.view(np.float32).reshape(n_chunks, CHUNK, 2)).pin_memory()
    dev = host.to(f"cuda:{local_rank}")
    stream = torch.cuda.current_stream()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident: warm-up, then K timed steps
    for _ in range(3):
        pipe.run_device(dev)
    sync_all()
    l0 = pipe.launches
    pipe.run_device(dev)
    launches_per_step = pipe.launches - l0
    graphs, graph_error = None, None
    if args.graph:
        # one step = one CUDA graph launch; two captured pipelines (own workspaces / outputs) alternate on two
        # streams, so consecutive batches overlap exactly like the lanes of one batch do
        try:
            pipe_b = pipeline.ReceivePipeline(FS, CHUNK, max_chunks=n_chunks, mode=args.mode, device=local_rank, n_lanes=args.lanes)
            graphs = [pipe.capture(dev), pipe_b.capture(dev)]
            gstreams = [torch.cuda.Stream(device=local_rank), torch.cuda.Stream(device=local_rank)]
        except Exception as e:                   # capture unsupported in this environment: drive the steps from Python
            print(f"bench: CUDA graph capture failed ({e!r}); falling back to per-step launches", file=sys.stderr)
            graphs, graph_error = None, repr(e)[:200]
            torch.cuda.synchronize()

    gathered = None
    if world > 1:
        # final gather of every step: frame records + counts of all segments, to every rank's device (rank 0 reads them)
        gathered = [torch.empty((world, pipe.records.numel()), dtype=torch.uint8, device=dev.device) for _ in range(2)]
        gather_streams = [torch.cuda.Stream(device=local_rank), torch.cuda.Stream(device=local_rank)]
        # the records leave the pipeline through a small staging copy: the pipeline's next step then waits for that copy
        # (microseconds), not for the collective — otherwise each step's low-occupancy tail runs alone on the GPU
        # while its own pipeline sits in NCCL (measured: 0.178 instead of 0.145 ms per step at N = 2)
        staged = [torch.empty_like(pipe.records) for _ in range(2)]
        copied = [torch.cuda.Event(), torch.cuda.Event()]

    def gather_step(which, p, after):
        """ONE all-gather of pipeline p's records (frames + counts) once its step on stream `after` is complete (NCCL, on the
        pipeline's own gather stream, so that the other pipeline's next step is not held up)."""
        if os.environ.get("TB200_BENCH_NO_GATHER"):      # experiment switch: what the step costs without the collective
            return
        gs_ = gather_streams[which]
        gs_.wait_stream(after)
        with torch.cuda.stream(gs_):
            staged[which].copy_(p.records, non_blocking=True)
            copied[which].record(gs_)
            dist.all_gather_into_tensor(gathered[which], staged[which])

    def run_steps(k):
        if graphs is None:
            for _ in range(k):
                pipe.run_device(dev, join=False)     # batches pipeline through the lane streams, each lane in order
                if world > 1:
                    pipe.join()
                    gather_step(0, pipe, stream)
                    stream.wait_stream(gather_streams[0])
            pipe.join()                              # ... and all of them are complete before the stop event
            return
        for gs in gstreams:
            gs.wait_stream(stream)
        for i in range(k):
            with torch.cuda.stream(gstreams[i & 1]):
                if world > 1:
                    gstreams[i & 1].wait_event(copied[i & 1])              # this pipeline's previous records are staged
                graphs[i & 1].replay()
            if world > 1:
                gather_step(i & 1, pipe if (i & 1) == 0 else pipe_b, gstreams[i & 1])
        for gs in gstreams:
            stream.wait_stream(gs)
        if world > 1:
            for gs in gather_streams:
                stream.wait_stream(gs)

    run_steps(args.warmup)
    # everything rank 0 does alone (starting the clock sampler) happens BEFORE the barrier: with a collective inside the
    # timed region a late rank 0 is waited for by all the others in their first gather (measured: 0.8 ms per run)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pipe.stats.zero_()
    if graphs is not None:
        pipe_b.stats.zero_()
    sync_all()
    e0.record(stream)
    run_steps(args.steps)
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps
    stats = pipe.stats.cpu().numpy().copy()
    # ---- untimed check of the gather: the merged record list must equal what ONE GPU finds in the same segments
    shard_check = None
    if world > 1:
        last = 0 if graphs is None else (args.steps - 1) & 1
        if rank == 0:
            g_all = gathered[last].cpu().numpy()
            fbytes = int(np.prod(pipe.frames.shape))
            g_frames = [g_all[r, :fbytes] for r in range(world)]
            g_counts = np.stack([g_all[r, fbytes:].view(np.int32) for r in range(world)])
            from tetraear_b200 import _lib as _l, shard as _shard
            checked, n_merged = [], 0
            merged = []
            for r in range(world):
                rec = np.ascontiguousarray(g_frames[r]).view(_l.FRAME_DTYPE).reshape(n_chunks, -1)
                for c in range(n_chunks):
                    for f in rec[c, :int(g_counts[r, c])]:
                        if f["kept"]:
                            # key = absolute bit position of the frame in the job's capture (segments and chunks are disjoint)
                            merged.append((0, (r * n_chunks + c) * 4096 + int(f["start_pos"]), (r, c, int(f["number"]), int(f["type"]),
                                                                                                bool(f["crc_ok"]), bytes(f["bits"]))))
            merged = _shard.dedup_frames(sorted(merged, key=lambda t: (t[0], t[1])))
            n_merged = len(merged)
            solo = pipeline.ReceivePipeline(FS, CHUNK, max_chunks=n_chunks, mode=args.mode, device=local_rank, n_lanes=args.lanes)
            ok_all, detail = True, None
            for r in sorted({0, world - 1}):
                xr = x if r == rank else make_c1(seed=1234 + r)
                nfr1, fr1, _ = solo.run_host(xr)
                want = [(r, c, int(f["number"]), int(f["type"]), bool(f["crc_ok"]), bytes(f["bits"]))
                        for c in range(n_chunks) for f in fr1[c, :int(nfr1[c])] if f["kept"]]
                got = [t[2] for t in merged if t[2][0] == r]
                if got != want:
                    ok_all = False
                    bad = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), min(len(got), len(want)))
                    detail = f"segment {r}: {len(got)} merged vs {len(want)} single-GPU frames, first difference at frame {bad}"
                    print(f"bench: sharded gather differs from the single-GPU result ({detail})", file=sys.stderr)
                checked.append(r)
            del solo
            shard_check = {"segments": world, "merged_frames": n_merged, "verified_against_single_gpu": checked, "ok": ok_all}
            if detail:
                shard_check["detail"] = detail

    used_graphs = graphs is not None
    if used_graphs:
        stats = stats + pipe_b.stats.cpu().numpy()
        del graphs, pipe_b
        torch.cuda.empty_cache()

    # ---- end to end: pinned host IQ -> H2D -> kernels -> D2H of frame records, every step
    for _ in range(max(1, min(args.warmup, 2))):
        pipe.run_host(host)
    sync_all()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        nfr, frames, st_e2e = pipe.run_host(host)
    t1.record(stream)
    sync_all()
    ms_e2e = t0.elapsed_time(t1)
    clk = clocks.stop() if clocks else None
    # same end-to-end loop fed with the capture device's native int16 IQ (SC16Q11, 4 B/sample over PCIe): extra figure
    raw16 = torch.from_numpy(np.clip(np.rint(x.view(np.float32) * 32768.0), -32768, 32767).astype(np.int16)
                             .reshape(n_chunks, CHUNK, 2)).pin_memory()
    pipe.run_host(raw16)
    sync_all()
    t2, t3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t2.record(stream)
    for _ in range(args.steps):
        pipe.run_host(raw16)
    t3.record(stream)
    sync_all()
    ms_e2e16 = t2.elapsed_time(t3)

    # ---- the same device-resident loop in the other time-parallel arithmetic (float32 channel filter when the headline
    # runs the float64 one, and vice versa), timed the same way (CUDA-graph replay on two streams)
    alt = None
    alt_mode = {"fast": "fast32", "fast32": "fast", "fast-tc": "fast32-tc", "fast32-tc": "fast-tc"}.get(args.mode)
    if alt_mode and world == 1 and not args.no_extras:
        pa = pipeline.ReceivePipeline(FS, CHUNK, max_chunks=n_chunks, mode=alt_mode, device=local_rank, n_lanes=args.lanes)
        pb = pipeline.ReceivePipeline(FS, CHUNK, max_chunks=n_chunks, mode=alt_mode, device=local_rank, n_lanes=args.lanes)
        k_alt = min(args.steps, 200)
        try:
            ga = [pa.capture(dev), pb.capture(dev)]
            gst = [torch.cuda.Stream(device=local_rank), torch.cuda.Stream(device=local_rank)]

            def alt_steps(k):
                for gs_ in gst:
                    gs_.wait_stream(stream)
                for i in range(k):
                    with torch.cuda.stream(gst[i & 1]):
                        ga[i & 1].replay()
                for gs_ in gst:
                    stream.wait_stream(gs_)
        except Exception:
            torch.cuda.synchronize()

            def alt_steps(k):
                for _ in range(k):
                    pa.run_device(dev, join=False)
                pa.join()
        alt_steps(5)
        sync_all()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        alt_steps(k_alt)
        a1.record(stream)
        sync_all()
        ms_alt = a0.elapsed_time(a1) / k_alt
        alt = {"mode": alt_mode, "value": round(samples_per_step / (ms_alt / 1e3) / 1e6, 3), "unit": "Msamples/s",
               "ms_per_step": round(ms_alt, 4),
               "path_frac": round((samples_per_step * 8.0 + n_chunks * pipe.plan.max_symbols) / (ms_alt / 1e3) / 1e9 / measured_peaks()[0], 5),
               "note": "fast32 = channel filter recurrences in float32 (3e-6 from scipy's filtfilt) instead of float64 (1e-11); same "
                       "float32 time-parallel decimator; dibits / frames / CRC identical on all 183 chunks (tests/test_gpu_fullsize.py)"}
        del pa, pb
        torch.cuda.empty_cache()

    # ---- end to end through the mirrored CLASS API, one chunk per call like CaptureThread.run (ui/modern.py:1919-2077):
    # SignalProcessor.process(numpy chunk) -> TetraDecoder.decode(dibits), host numpy in and out, H2D/D2H inside
    class_api = None
    if world == 1 and not args.no_extras:
        from tetraear_b200.signal import SignalProcessor
        from tetraear_b200.core import TetraDecoder
        sp, dec = SignalProcessor(FS, mode=args.mode, device=local_rank), TetraDecoder(auto_decrypt=False, device=local_rank)
        for c in range(3):
            dec.decode(sp.process(x[c], 0.0))
        torch.cuda.synchronize()
        t_a = time.perf_counter()
        nfr_class = 0
        for c in range(n_chunks):
            nfr_class += len(dec.decode(sp.process(x[c], 0.0)))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t_a
        class_api = {"value": round(samples_per_step / dt / 1e6, 3), "unit": "Msamples/s", "ms_per_chunk": round(dt / n_chunks * 1e3, 4),
                     "frames": nfr_class, "h2d_bytes_per_call": CHUNK * 8,
                     "note": "per-chunk SignalProcessor.process + TetraDecoder.decode (the reference's own call pattern), wall clock, "
                             "incl. the host MAC header walk per frame"}
        del sp, dec

    # ---- per-stage device time for the roofline (separate pass so the events do not sit in `value`)
    # (single lane: the per-stage events of concurrent lanes would overlap and double count)
    pipe1 = pipe if args.lanes == 1 else pipeline.ReceivePipeline(FS, CHUNK, max_chunks=n_chunks, mode=args.mode, device=local_rank)
    pipe1.profile(True)
    for _ in range(min(args.steps, 20)):
        pipe1.run_device(dev)
    torch.cuda.synchronize()
    prof = pipe1.profile_read()
    pipe1.profile(False)

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_e2e16], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e16 = float(t[0]), float(t[1]), float(t[2])
        s = torch.tensor(stats, dtype=torch.int64, device=f"cuda:{local_rank}")
        dist.all_reduce(s)
        stats = s.cpu().numpy()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_samples = samples_per_step * world * args.steps
    value = total_samples / (ms / 1e3) / 1e6
    e2e = total_samples / (ms_e2e / 1e3) / 1e6
    bursts_per_s = float(stats[1]) / (ms / 1e3)

    # roofline of the dominant stage: algorithmic bytes = 8 B per input complex64 sample + 8/q per decimated output
    peak, peak_src = measured_peaks()
    dec_ms, dec_calls = prof["decimator"]
    q = pipe.plan.q
    alg_bytes = samples_per_step * 8.0 + n_chunks * pipe.plan.n_dec * 8.0
    dec_launch_ms = dec_ms / max(dec_calls, 1)
    achieved = alg_bytes / (dec_launch_ms / 1e3) / 1e9
    stage_ms = {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]}
    # whole process() path against the same roofline: 8 B in + 1 B/dibit out per sample
    path_bytes = samples_per_step * 8.0 + n_chunks * (pipe.plan.max_symbols) * 1.0
    step_ms = ms / args.steps

    # ---- CPU baseline: the unmodified reference (oracle/_ref) on a bounded sample, rank 0 only at N=1; the oracle's
    # C + numpy port alongside
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = host_cores()
        port_one = cpu_run(x[:8], 1)
        port_many = cpu_run(x, cores)
        if reference_available():
            one = cpu_run(x[:4], 1, kind="reference")
            many = cpu_run(x, cores, kind="reference")
            cpu = {"value": round(many[0], 3), "unit": "Msamples/s", "cores": cores, "kind": "reference",
                   "sample": f"all {n_chunks} chunks (131072 samples each) of the same capture, once, through the unmodified "
                             f"tetraear SignalProcessor.process + TetraDecoder.decode (oracle/_ref), {cores} worker processes "
                             f"({many[2]:.1f} s wall); 1-core: {one[0]:.3f} Msamples/s over 4 chunks",
                   "one_core_value": round(one[0], 3), "bursts_per_s": round(many[1], 1),
                   "port_value": round(port_many[0], 3), "port_one_core_value": round(port_one[0], 3),
                   "port_note": "oracle.process + oracle.decode (C + numpy restatement), same chunks and worker count"}
        else:
            cpu = {"value": round(port_many[0], 3), "unit": "Msamples/s", "cores": cores, "kind": "port",
                   "sample": f"all {n_chunks} chunks through oracle.process + oracle.decode, {cores} worker processes "
                             "(oracle/_ref is absent: run oracle/make_ref.py where /root/reference exists)",
                   "one_core_value": round(port_one[0], 3), "bursts_per_s": round(port_many[1], 1)}

    line = {
        "metric": "iq_msamples_per_s", "value": round(value, 3), "unit": "Msamples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(step_ms, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.mode.startswith("fast32") else "f32 decimator / f64 channel filter",
        "data": "synthetic",
        "config": {"workload": "c1: single synthetic TETRA downlink carrier, 10 s complex64 IQ @2.4 MS/s, "
                               f"{n_chunks} chunks x 131072 samples per GPU, process+decode per chunk (reference mode, "
                               f"arithmetic '{args.mode}': " + {
                                   "exact": "op-order-exact float32 decimator recurrence + float64 channel filter, bit-identical to the reference",
                                   "fast": "time-parallel float32 decimator + float64 channel filter, the reference's own precision per stage; "
                                           "dibits identical, soft symbols <= 1e-4 of full scale",
                                   "fast32": "time-parallel float32 decimator and float32 channel filter; dibits identical on the parity cases, "
                                             "soft symbols <= 1e-4 of full scale (tests/test_gpu_parity.py::test_fast_mode_within_tolerance)",
                                   "fast-tc": "as 'fast' with the decimator's linear part on the tensor cores (split-TF32 tcgen05 MMAs)",
                                   "fast32-tc": "as 'fast32' with the decimator's linear part on the tensor cores (split-TF32 tcgen05 MMAs)",
                                   }[args.mode] + ")",
                   "l2": "input 192 MB per step > 126 MB L2, no flush needed", "parallelism": f"capture segment per GPU x{world}",
                   "stream_lanes": args.lanes, "cuda_graph": bool(args.graph) and graph_error is None,
                   "cpus_bound_per_rank": numa_cpus, "numa_node_rank0": numa_node,
                   "gather": ("frame records of every segment all-gathered over NCCL inside the timed region, every step"
                              if world > 1 else "single segment")},
        "crc_pass_bursts_per_s": round(bursts_per_s, 1), "crc_pass_per_step": int(stats[1] // max(args.steps, 1)),
        "gpu_launches": int(launches),
        "e2e": {"value": round(e2e, 3), "unit": "Msamples/s", "h2d_bytes_per_step": pipe.h2d_bytes(n_chunks),
                "d2h_bytes_per_step": pipe.d2h_bytes(n_chunks), "ms_per_step": round(ms_e2e / args.steps, 4)},
        "e2e_sc16": {"value": round(total_samples / (ms_e2e16 / 1e3) / 1e6, 3), "unit": "Msamples/s",
                     "h2d_bytes_per_step": pipe.h2d_bytes(n_chunks, sc16=True), "ms_per_step": round(ms_e2e16 / args.steps, 4),
                     "note": "same loop fed with int16 SC16Q11 IQ (the capture device's format, converted on the GPU); "
                             "input quantised to int16, not part of the headline"},
        "roofline": {"bound": "hbm", "kernel": "decimator stage (cheby1 zero-phase IIR + ::q)", "achieved": round(achieved, 2),
                     "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
                     # dram__bytes_read.sum + dram__bytes_write.sum of the decimator kernel, one launch, ncu --set full
                     # (profiles/r02_ncu_c1.md); valid for the FFMA2 decimator of the fast modes on this workload only
                     "traffic": DECIM_TRAFFIC_BYTES if args.mode in ("fast", "fast32") else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "ms_per_launch": round(dec_launch_ms, 4),
                     "path_achieved_gbs": round(path_bytes / (step_ms / 1e3) / 1e9, 2),
                     "path_frac": round(path_bytes / (step_ms / 1e3) / 1e9 / peak, 5), "stage_ms": stage_ms},
        "clocks": clk,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if alt:
        line["value_other_arithmetic"] = alt
    if class_api:
        line["e2e_class_api"] = class_api
    if shard_check:
        line["shard_gather"] = shard_check
    if world == 1 and not args.no_extras:
        try:
            del dev, host
            torch.cuda.empty_cache()
            line["other_workloads"] = other_workloads(local_rank, peak)
        except Exception as e:          # the headline line must survive a failure of the extras
            line["other_workloads"] = {"error": repr(e)[:300]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ other configs (kernel figures)
def other_workloads(device, peak):
    """Short device-resident timings of the other BASELINE configs' kernels (not the headline):
    c2 PSD 4096-pt + EMA over 60 s @2.4 MS/s, c3 1M-burst parse+CRC (reference mode) and 1M SCH/HD
    blocks through the ETSI decoder, c4 800-channel channelizer + etsi receiver on 0.2 s @20 MS/s."""
    import torch
    from tetraear_b200 import _lib, synth, etsi
    from tetraear_b200.signal import SpectrumAnalyzer
    from tetraear_b200.core import TetraProtocolParser
    dev = torch.device(f"cuda:{device}")
    out = {}

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # ---- drop-in per-call latency: one 131072-sample chunk through the mirrored Python API, host numpy in/out
    from tetraear_b200.signal import SignalProcessor
    from tetraear_b200.core import TetraDecoder
    chunk = synth.make_capture(CHUNK, 130, seed=1234)
    sp, dec = SignalProcessor(FS, mode="fast"), TetraDecoder(auto_decrypt=False)
    for _ in range(3):
        frames = dec.decode(sp.process(chunk, 0.0))
    t0 = time.perf_counter()
    for _ in range(20):
        frames = dec.decode(sp.process(chunk, 0.0))
    lat = (time.perf_counter() - t0) / 20
    out["per_call_latency_one_chunk"] = {"ms": round(lat * 1e3, 3), "frames": len(frames),
                                         "note": "SignalProcessor.process + TetraDecoder.decode, wall clock incl. H2D/D2H and the "
                                                 "host MAC header walk; the chunk lasts 54.6 ms at 2.4 MS/s (reference: ~60 ms per chunk)"}
    del sp, dec

    # ---- c2: 60 s @2.4 MS/s, N = 4096 dense frames, power + EMA-smoothed output (float32)
    n = 144_000_000
    nfr = n // 4096
    g = torch.Generator(device=dev); g.manual_seed(1)
    iq = (0.1 * torch.randn((n, 2), generator=g, device=dev, dtype=torch.float32))
    an = SpectrumAnalyzer(FS, 4096, device=device, precision="f32")
    power = torch.empty((nfr, 4096), dtype=torch.float32, device=dev)
    smooth = torch.empty((nfr, 4096), dtype=torch.float32, device=dev)
    stats = torch.empty((nfr, 8), dtype=torch.float64, device=dev)
    ms = timed(lambda: an.run_device(iq, nfr, 4096, denoise=True, power=power, smoothed=smooth, stats=stats), reps=3, warm=1)
    byts = n * 8 + nfr * 4096 * 4 * 2          # ALGORITHMIC bytes: IQ read, power written, smoothed written;
    #                                            the EMA pass's re-read of the power array is traffic, not useful bytes
    out["c2_psd_4096_ema_60s_f32"] = {"msamples_per_s": round(n / ms / 1e3, 1), "ms": round(ms, 3), "frames": nfr,
                                  "achieved_gbs": round(byts / ms / 1e6, 1), "frac_of_hbm_peak": round(byts / ms / 1e6 / peak, 4)}
    # same 60 s with linear (Welch) averaging of 8 frames before the dB conversion ("w/ averaging"): 1/8 of the output
    an8 = SpectrumAnalyzer(FS, 4096, device=device, precision="f32", average=8)
    p8 = torch.empty((nfr // 8, 4096), dtype=torch.float32, device=dev)
    s8 = torch.empty((nfr // 8, 4096), dtype=torch.float32, device=dev)
    st8 = torch.empty((nfr // 8, 8), dtype=torch.float64, device=dev)
    ms8 = timed(lambda: an8.run_device(iq, nfr, 4096, denoise=True, power=p8, smoothed=s8, stats=st8), reps=3, warm=1)
    byts8 = n * 8 + (nfr // 8) * 4096 * 4 * 2
    out["c2_psd_4096_avg8_ema_60s_f32"] = {"msamples_per_s": round(n / ms8 / 1e3, 1), "ms": round(ms8, 3), "frames_out": nfr // 8,
                                       "achieved_gbs": round(byts8 / ms8 / 1e6, 1),
                                       "frac_of_hbm_peak": round(byts8 / ms8 / 1e6 / peak, 4)}
    del an8, p8, s8, st8
    # CPU side: the reference's per-frame block + EMA in numpy on the host cores, bounded sample
    try:
        import multiprocessing as mp
        cores = host_cores()
        xs = (0.1 * np.random.default_rng(1).standard_normal((cores, 2000 * 4096, 2))).astype(np.float32).view(np.complex64)[..., 0]
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_psd_worker, [(xs[i][:4096 * 8], 4096) for i in range(cores)])
            w0 = time.perf_counter()
            res = pool.map(_psd_worker, [(xs[i], 4096) for i in range(cores)])
            wall = time.perf_counter() - w0
        out["c2_psd_4096_ema_60s_f32"]["cpu_baseline"] = {
            "value": round(sum(r[0] for r in res) / wall / 1e6, 2), "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": f"{cores} x 2000 dense 4096-point frames (numpy restatement of ui/modern.py:1928-1941 + the EMA of :929-935), one worker per core"}
        del xs
    except Exception as e:
        out["c2_psd_4096_ema_60s_f32"]["cpu_baseline"] = {"error": repr(e)[:200]}
    del iq, power, smooth, stats, an
    torch.cuda.empty_cache()

    # ---- c3 (reference mode): 1M bursts x 255 dibits -> type / data bits / CRC / MAC validity
    nb = 1_000_000
    rows = torch.from_numpy(synth.make_burst_batch(nb, seed=99)).to(dev)
    rec = torch.empty((nb, 32), dtype=torch.uint8, device=dev)
    st = torch.zeros((3,), dtype=torch.int64, device=dev)
    p = TetraProtocolParser(device=device)
    h = p._handle()
    stream = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: _lib.check(h.lib.tb200_burst_batch(h.h, rows.data_ptr(), nb, rec.data_ptr(), None, st.data_ptr(), stream)))
    byts = nb * (255 + 32)
    crc_pass = int(rec[:, 28].sum().item())
    out["c3_burst_batch_1M"] = {"bursts_per_s": round(nb / ms * 1e3, 0), "ms": round(ms, 4), "crc_pass": crc_pass,
                                "achieved_gbs": round(byts / ms / 1e6, 1), "frac_of_hbm_peak": round(byts / ms / 1e6 / peak, 4)}
    if reference_available():
        try:
            import multiprocessing as mp
            cores = host_cores()
            host_rows = rows[:cores * 1500].cpu().numpy()
            with mp.get_context("fork").Pool(cores) as pool:
                pool.map(_ref_burst_worker, [host_rows[:8]] * cores)
                w0 = time.perf_counter()
                res = pool.map(_ref_burst_worker, [host_rows[i::cores] for i in range(cores)])
                wall = time.perf_counter() - w0
            out["c3_burst_batch_1M"]["cpu_baseline"] = {
                "value": round(sum(r[0] for r in res) / wall, 1), "unit": "bursts/s", "cores": cores, "kind": "reference",
                "sample": f"first {len(host_rows)} of the 1M bursts through the unmodified TetraProtocolParser.parse_burst, {cores} workers"}
        except Exception as e:
            out["c3_burst_batch_1M"]["cpu_baseline"] = {"error": repr(e)[:200]}
    del rows, rec

    # ---- c3 (etsi mode): 1M SCH/HD blocks: descramble + de-interleave + RCPC Viterbi + CRC
    rng = np.random.default_rng(7)
    pool = 4096
    t1 = rng.integers(0, 2, (pool, 124), dtype=np.uint8)
    init = etsi.scrambling_init(262, 42, 17)
    soft = etsi.to_soft(etsi.encode(t1, "SCH/HD", init), 0.5, rng)
    soft_d = torch.from_numpy(soft).to(dev).repeat(nb // pool + 1, 1)[:nb].contiguous()
    dec = etsi.ChannelDecoder("SCH/HD", device=device)
    bits = torch.empty((nb, 40), dtype=torch.uint8, device=dev)
    ok = torch.empty((nb,), dtype=torch.uint8, device=dev)
    met = torch.empty((nb,), dtype=torch.int32, device=dev)
    ms = timed(lambda: dec.decode_device(soft_d, init, bits, ok, met))
    out["c3_etsi_viterbi_1M"] = {"blocks_per_s": round(nb / ms * 1e3, 0), "ms": round(ms, 4), "crc_pass": int(ok.sum().item()),
                                 "achieved_gbs": round(nb * (216 + 45) / ms / 1e6, 1), "bound": "integer-ALU issue (one thread per block, 16 register ACS per step), not HBM"}
    del soft_d, bits, ok, met

    # ---- c4: 20 MS/s -> 800 x 25 kHz: channelizer + per-channel etsi receiver, 0.2 s of capture
    M, fs = 800, 20e6
    n = 4_000_000
    wide = (0.02 * torch.randn((n, 2), generator=g, device=dev, dtype=torch.float32))
    bits_s, _, init = etsi.make_etsi_stream(14, seed=3)
    base = etsi.modulate_dqpsk(bits_s, 50000.0)                    # one channel at the channelizer's output rate
    # place it on four carriers by up-sampling x400 with the channelizer prototype would be slow in numpy;
    # instead add it directly to four channelizer OUTPUT rows after the first pass (the receiver load is what matters)
    ch = etsi.Channelizer(M, device=device)
    nf = ch.n_frames(n)
    y = torch.empty((M, nf, 2), dtype=torch.float32, device=dev)
    rx = etsi.EtsiReceiver(50000.0, max_streams=M, max_samples=nf, device=device)
    ms_ch = timed(lambda: ch.run_device(wide, n, y), reps=3, warm=1)
    inj = torch.from_numpy(base.view(np.float32).reshape(-1, 2)).to(dev)
    L = min(inj.shape[0], nf)
    for k in (3, 100, 401, 777):
        y[k, :L] += inj[:L]
    ms_rx = timed(lambda: rx.run_device(y, nf, init), reps=3, warm=1)
    found = int(rx.n_bursts.sum().item())
    okb = int(rx.crc_ok.view(M, rx.max_bursts, 2)[[3, 100, 401, 777]].sum().item())
    # the per-channel timing loop is sequential: ArchiveReplay cuts the frame axis into overlapping time blocks that
    # are received concurrently (what the c5 workload does); same channelizer output, 4 blocks
    from tetraear_b200.replay import ArchiveReplay
    rp = ArchiveReplay(M, n, device=device, n_blocks=4)
    rp.rows.copy_(y)
    ms_rx4 = timed(lambda: rp.receive_rows(nf, init), reps=3, warm=1)
    recs = rp.collect(0)
    ok4 = len({(c, slot) for c, slot, _ in recs if c in (3, 100, 401, 777)})
    out["c4_800ch_20MSps_0.2s"] = {"channelizer_msamples_per_s": round(n / ms_ch / 1e3, 1), "channelizer_ms": round(ms_ch, 3),
                                   "receiver_ms": round(ms_rx, 3), "total_msamples_per_s": round(n / (ms_ch + ms_rx) / 1e3, 1),
                                   "receiver_ms_4_time_blocks": round(ms_rx4, 3),
                                   "total_msamples_per_s_4_time_blocks": round(n / (ms_ch + ms_rx4) / 1e3, 1),
                                   "bursts_found": found, "crc_ok_blocks_on_4_carriers": okb,
                                   "distinct_bursts_on_4_carriers_4_time_blocks": ok4,
                                   "channelizer_frac_of_hbm_peak": round((n * 8 + M * nf * 8) / ms_ch / 1e6 / peak, 4)}
    del rp
    del wide, y, rx, ch
    torch.cuda.empty_cache()

    # ---- N3: scanner heuristics, 64 dwells x 262144 samples (one per scanned frequency), device-resident
    from tetraear_b200.signal import TetraSignalDetector
    nd, nsd = 64, 262144
    dw = 0.05 * torch.randn((nd, nsd, 2), generator=g, device=dev, dtype=torch.float32)
    det = TetraSignalDetector(FS, device=device)
    det.stats_batch(dw, 5)
    hd = det._handle(nd * nsd)
    rec = torch.empty((nd, _lib.DETECT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    ms = timed(lambda: _lib.check(hd.lib.tb200_detect_batch(hd.h, dw.data_ptr(), _lib.DTYPE_C64, nd, nsd, nsd, FS, 5,
                                                            rec.data_ptr(), stream)))
    out["n3_detector_64x262144"] = {"msamples_per_s": round(nd * nsd / ms / 1e3, 1), "ms": round(ms, 4),
                                    "achieved_gbs": round(nd * nsd * 16 / ms / 1e6, 1),
                                    "frac_of_hbm_peak": round(nd * nsd * 16 / ms / 1e6 / peak, 4),
                                    "note": "power/windows + modulation steps + sync-bit search; the dwell is read twice (max|x| first)"}
    del dw

    # ---- N4: codec frames, 1M frames gathered from bit rows
    from tetraear_b200.voice import VoiceBlockBuilder
    nv = 1_000_000
    vb = VoiceBlockBuilder(device=device)
    vrows = torch.randint(0, 2, (nv, 432), generator=g, device=dev, dtype=torch.uint8)
    blocks = torch.empty((nv, 690), dtype=torch.int16, device=dev)
    vld = torch.empty((nv,), dtype=torch.uint8, device=dev)
    ms = timed(lambda: _lib.check(vb._h.lib.tb200_voice_blocks(vb._h.h, _lib.VOICE_BITS, vrows.data_ptr(), 432, None, None, None,
                                                                nv, 127.0, blocks.data_ptr(), vld.data_ptr(), stream)))
    out["n4_voice_blocks_1M"] = {"frames_per_s": round(nv / ms * 1e3, 0), "ms": round(ms, 4),
                                 "achieved_gbs": round(nv * (432 + 1380) / ms / 1e6, 1),
                                 "frac_of_hbm_peak": round(nv * (432 + 1380) / ms / 1e6 / peak, 4)}
    return out


# ------------------------------------------------------------------------------ c5: multichannel archive replay
def run_c5(args):
    """BASELINE config 5: 1 h multichannel archive replay sharded over the GPUs of one box, realised as a looped
    device-resident 0.8 s segment of 20 MS/s wideband IQ per rank (800 x 25 kHz, `etsi` mode: channelizer ->
    RRC/timing/soft bits -> TS search -> Viterbi/CRC).  Work unit = one segment; rank r owns segments r, r+N, ...
    (weak scaling: every rank processes `steps` segments).  No collective on the data path."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    torch.cuda.set_device(local_rank)
    from tetraear_b200 import etsi
    from tetraear_b200.replay import ArchiveReplay
    M, n = 800, 16_000_000
    dev = torch.device(f"cuda:{local_rank}")
    g = torch.Generator(device=dev); g.manual_seed(10 + rank)
    wide = 0.02 * torch.randn((n, 2), generator=g, device=dev, dtype=torch.float32)
    n_blocks = int(os.environ.get("TB200_C5_BLOCKS", "2"))
    rp = ArchiveReplay(M, n, device=local_rank, n_blocks=n_blocks)
    bits_s, _, init = etsi.make_etsi_stream(56, seed=3)
    inj = torch.from_numpy(etsi.modulate_dqpsk(bits_s, 50000.0).view(np.float32).reshape(-1, 2)).to(dev)
    L = min(inj.shape[0], rp.nf)
    carriers = list(range(5, M, 50))                     # 16 active carriers, injected at the channelizer output

    def step():
        rp.chan.run_device(wide, n, rp.rows)
        rp.rows[carriers, :L] += inj[:L]
        rp.receive_rows(rp.nf, init)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ok_blocks = sum(int(r.crc_ok.sum().item()) for r in rp.rxs)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
        c = torch.tensor([ok_blocks], dtype=torch.int64, device=dev); dist.all_reduce(c); ok_blocks = int(c[0])
    clk = clocks.stop() if clocks else None
    if rank == 0:
        peak, _ = measured_peaks()
        total = n * world * args.steps
        print(json.dumps({
            "metric": "iq_msamples_per_s", "value": round(total / ms / 1e3, 1), "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c5: multichannel archive replay, 20 MS/s -> 800 x 25 kHz etsi receive, looped 0.8 s "
                                   f"device-resident segment per rank, {rp.n_blocks} concurrent time block(s) x time-parallel demodulator segments, 16 carriers injected at the channelizer output",
                       "parallelism": f"segment per GPU x{world}", "l2": "input 128 MB per step > 126 MB L2"},
            "crc_pass_blocks_per_s": round(ok_blocks * world / (ms / args.steps) * 1e3 / max(world, 1), 1),
            "realtime_factor": round(total / ms / 1e3 / 20.0, 1),
            "roofline": {"bound": "hbm", "kernel": "channelizer + etsi receiver", "achieved": round(n * 8 * world * args.steps / ms / 1e6, 2),
                         "peak": peak, "unit": "GB/s", "frac": round(n * 8 * world * args.steps / ms / 1e6 / peak / world, 5), "traffic": None},
            "clocks": clk}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ c4: wideband, reference mode
C4_FS, C4_M, C4_CHUNK, C4_RING = 20e6, 800, 1 << 20, 16
C4_CARRIERS = [0, 1, -1, 7, -37, 120, 399, -400, 250, -333, 64, -199, 3, 311, -272, 158]


def _ref_c4_worker(args):
    """The reference's only way to a channel of a wideband chunk: frequency_shift + process + decode, per channel."""
    global _REF
    import logging
    logging.disable(logging.CRITICAL)
    if _REF is None:
        from oracle import ref_bridge
        _REF = ref_bridge.load()
    SP, TD, _ = _REF
    x, channels = args
    t0 = time.perf_counter()
    crc = 0
    for ch in channels:
        sp = SP(C4_FS)
        d = sp.process(sp.frequency_shift(x, ch * C4_FS / C4_M), 0)
        dec = TD(auto_decrypt=False)
        if len(d) >= 255:
            dec.decode(d)
        crc += dec.protocol_parser.stats["crc_pass"]
    return len(channels), crc, time.perf_counter() - t0


def run_c4(args):
    """BASELINE config 4 in REFERENCE mode: 20 MS/s wideband chunks of 2^20 samples, all 800 x 25 kHz channels of every
    chunk demodulated to dibits and decoded exactly as process(frequency_shift(x, f_k)) + decode would per channel
    (tests/test_gpu_multichannel.py).  A step = a ring of 16 different chunks (128 MB > L2).  Rank r owns its own ring
    (weak scaling over capture segments, no collective)."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    torch.cuda.set_device(local_rank)
    from tetraear_b200 import synth
    from tetraear_b200.multichannel import WidebandReceiver
    dev = torch.device(f"cuda:{local_rank}")
    base = synth.make_wideband(C4_CHUNK, C4_CARRIERS, C4_FS, C4_M, 1079, seed=5 + rank)
    host = torch.from_numpy(base.view(np.float32).reshape(-1, 2)).pin_memory()
    ring = torch.empty((C4_RING, C4_CHUNK, 2), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(40 + rank)
    for i in range(C4_RING):                       # same carriers, fresh wideband noise per ring slot
        ring[i].copy_(host, non_blocking=True)
        ring[i] += 0.005 * torch.randn((C4_CHUNK, 2), generator=g, device=dev, dtype=torch.float32)
    # two receivers (own workspaces) on two streams take alternate chunks: the latency-bound tail of one chunk (timing pick,
    # slicer, sync search: one small CTA per channel) overlaps the channel bank of the next
    rx = WidebandReceiver(C4_FS, C4_M, C4_CHUNK, mode="fast32", device=local_rank, keep_symbols=False)
    rxs = [rx, WidebandReceiver(C4_FS, C4_M, C4_CHUNK, mode="fast32", device=local_rank, keep_symbols=False)]
    side = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    stream = torch.cuda.current_stream()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    def step():
        fork = torch.cuda.Event()
        fork.record(stream)
        for s_ in side:
            s_.wait_event(fork)
        for i in range(C4_RING):
            with torch.cuda.stream(side[i & 1]):
                rxs[i & 1].run_device(ring[i])
        for s_ in side:
            ev = torch.cuda.Event()
            ev.record(s_)
            stream.wait_event(ev)

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    l0 = rxs[0].launches + rxs[1].launches
    step()
    launches_per_step = rxs[0].launches + rxs[1].launches - l0
    sync_all()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    for r_ in rxs:
        r_.stats.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    stats = (rxs[0].stats + rxs[1].stats).cpu().numpy().copy()
    # end to end: pinned host chunk -> H2D -> kernels -> D2H of the per-channel frame counts and records
    stage = [torch.empty((C4_CHUNK, 2), dtype=torch.float32, device=dev) for _ in range(2)]
    h_nfr = [torch.empty((C4_M,), dtype=torch.int32).pin_memory() for _ in range(2)]
    h_frames = [torch.empty(rx.frames.shape, dtype=torch.uint8).pin_memory() for _ in range(2)]

    def step_host():
        fork = torch.cuda.Event()
        fork.record(stream)
        for s_ in side:
            s_.wait_event(fork)
        for i in range(C4_RING):                  # upload / kernels / download of chunk i on stream i & 1
            k = i & 1
            with torch.cuda.stream(side[k]):
                stage[k].copy_(host, non_blocking=True)
                rxs[k].run_device(stage[k])
                h_nfr[k].copy_(rxs[k].n_frames, non_blocking=True)
                h_frames[k].copy_(rxs[k].frames, non_blocking=True)
        for s_ in side:
            ev = torch.cuda.Event()
            ev.record(s_)
            stream.wait_event(ev)
        stream.synchronize()

    step_host()
    sync_all()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_e2e = max(1, min(args.steps, 10))
    t0.record(stream)
    for _ in range(k_e2e):
        step_host()
    t1.record(stream)
    sync_all()
    ms_e2e = t0.elapsed_time(t1) / k_e2e
    clk = clocks.stop() if clocks else None
    # per-stage device time
    for hh in (rx.bank, rx.rx):
        hh.profile(True)
    for i in range(C4_RING):                      # one receiver, one stream: the stage events must not overlap
        rx.run_device(ring[i])
    torch.cuda.synchronize()
    prof = {}
    for hh in (rx.bank, rx.rx):
        for k, (tms, calls) in hh.profile_read().items():
            if calls:
                prof[k] = round(prof.get(k, 0.0) + tms / C4_RING, 4)
        hh.profile(False)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        c = torch.tensor(stats, dtype=torch.int64, device=dev); dist.all_reduce(c); stats = c.cpu().numpy()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    samples_per_step = C4_RING * C4_CHUNK
    total = samples_per_step * world * args.steps
    value = total / (ms / 1e3) / 1e6
    step_ms = ms / args.steps
    cpu = None
    if world == 1 and not args.no_cpu and reference_available():
        import multiprocessing as mp
        cores = host_cores()
        chans = C4_CARRIERS[:8]
        parts = [chans[i::cores] for i in range(min(cores, len(chans)))]
        with mp.get_context("fork").Pool(len(parts)) as pool:
            pool.map(_ref_c4_worker, [(base[:65536], p[:1]) for p in parts])
            w0 = time.perf_counter()
            res = pool.map(_ref_c4_worker, [(base, p) for p in parts])
            wall = time.perf_counter() - w0
        per_channel_s = wall * len(parts) / len(chans)        # CPU seconds per (channel, chunk)
        cpu = {"value": round(C4_CHUNK / (per_channel_s * C4_M / cores) / 1e6, 4), "unit": "Msamples/s", "cores": cores,
               "kind": "reference",
               "sample": f"8 of 800 channels of one 2^20-sample chunk through the unmodified frequency_shift + process + decode "
                         f"({wall:.2f} s wall on {len(parts)} workers), scaled x100 to all 800 channels on {cores} cores",
               "crc_pass_8_channels": int(sum(r[1] for r in res))}
    alg = samples_per_step * 8.0 + C4_RING * C4_M * rx.stride * 1.0
    line = {
        "metric": "iq_msamples_per_s", "value": round(value, 3), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c4: wideband 20 MS/s capture, 800 x 25 kHz channels, reference mode (every channel == "
                               "process(frequency_shift(x, f_k)) + decode), 16 chunks x 2^20 samples per step per GPU, 16 active carriers",
                   "l2": "ring of 16 chunks = 128 MB per step > 126 MB L2",
                   "parallelism": f"capture segment per GPU x{world}; two receivers on two streams take alternate chunks"},
        "channel_msamples_per_s": round(value * C4_M, 1), "realtime_factor": round(value / 20.0, 1),
        "crc_pass_bursts_per_s": round(float(stats[1]) / (ms / 1e3), 1), "gpu_launches": int(launches_per_step * args.steps),
        "e2e": {"value": round(samples_per_step * world / (ms_e2e / 1e3) / 1e6, 3), "unit": "Msamples/s",
                "h2d_bytes_per_step": samples_per_step * 8, "d2h_bytes_per_step": C4_RING * (C4_M * 4 + int(np.prod(rx.frames.shape))),
                "ms_per_step": round(ms_e2e, 4)},
        "roofline": {"bound": "hbm", "kernel": "whole c4 step (channel bank + channel filter + slicer + decode)",
                     "achieved": round(alg / (step_ms / 1e3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                     "frac": round(alg / (step_ms / 1e3) / 1e9 / peak, 5), "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg, "stage_ms_per_chunk": prof,
                     "note": "8 B per wideband sample read once + 1 B per dibit (SURVEY 8(d)); the channel bank is a direct two-stage "
                             "DFT (FP32-issue bound), not HBM bound"},
        "clocks": clk,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port: the reference is pure Python and
    /root/reference does not exist on the GPU box) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as entry
    from oracle import build as ob
    ob.build()
    cores = host_cores()
    kind = "reference" if reference_available() else "port"
    x = make_c1(n_chunks=C1_CHUNKS)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_run(x[:cores], cores, kind=kind)
    # bounded sample: the whole capture per step when K steps of it fit in ~90 s of wall clock, else its first chunks
    _, _, t_full = cpu_run(x, cores, kind=kind)
    sample_chunks = C1_CHUNKS
    if t_full * args.steps > 90.0:
        sample_chunks = int(max(2 * cores, min(C1_CHUNKS, C1_CHUNKS * 90.0 / (t_full * args.steps))))
    x = x[:sample_chunks]
    walls, crcs = [], 0.0
    for _ in range(args.steps):
        v, b, wall = cpu_run(x, cores, kind=kind)
        walls.append(wall)
        crcs += b * wall
    total = sample_chunks * CHUNK * args.steps
    value = total / sum(walls) / 1e6
    line = {
        "impl": "reference", "metric": "iq_msamples_per_s", "value": round(value, 3), "unit": "Msamples/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * sum(walls) / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 decimator / f64 channel filter", "data": "synthetic",
        "config": {"workload": "c1: single synthetic TETRA downlink carrier @2.4 MS/s, process+decode per 131072-sample chunk",
                   "sample": f"{sample_chunks} of {C1_CHUNKS} chunks per step"},
        "crc_pass_bursts_per_s": round(crcs / sum(walls), 1),
        "cpu_baseline": {"value": round(value, 3), "unit": "Msamples/s", "cores": cores, "kind": kind,
                         "sample": f"{sample_chunks} chunks x 131072 samples per step, {cores} worker processes, " +
                                   ("the unmodified tetraear SignalProcessor.process + TetraDecoder.decode staged under oracle/_ref"
                                    if kind == "reference" else
                                    "oracle.process + oracle.decode (C + numpy restatement; oracle/_ref absent)")},
        "e2e": {"value": round(value, 3), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="enqueue every step from Python instead of replaying the captured CUDA graphs")
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("TB200_BENCH_LANES", "2")),
                    help="sub-batches processed on concurrent streams (overlaps the latency-bound tail kernels)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=os.environ.get("TB200_BENCH_MODE", "fast"), choices=["exact", "fast", "fast32", "fast-tc", "fast32-tc", "fast-tc4", "fast32-tc4"])
    ap.add_argument("--workload", default="c1", choices=["c1", "c4", "c5"],
                    help="c1 = headline; c4 = wideband 800-channel chunk in reference mode; c5 = multichannel archive replay (etsi mode)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the other-workload kernel figures")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c4":
        run_c4(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
