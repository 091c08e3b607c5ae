# small end-to-end run of every entry point for compute-sanitizer memcheck
import sys, numpy as np
sys.path.insert(0, '.')  # run from the repo root
import __graft_entry__ as g; g.build()
from tetraear_b200 import synth, etsi, pipeline
from tetraear_b200.signal import SignalProcessor, SpectrumAnalyzer
from tetraear_b200.core import TetraDecoder, TetraProtocolParser
iq = synth.make_capture(131072, 130, seed=1)
for mode in ("exact", "fast"):
    sp = SignalProcessor(2.4e6, mode=mode)
    for n in (131072, 40001, 500, 29):
        d = sp.process(iq[:n], 0.0); d = sp.process(iq[:n], 321.0)
    sp.decimate(iq[:5000], 10); sp.filter_signal(iq[:3000]); sp.frequency_shift(iq[:100], 5.0)
dec = TetraDecoder(auto_decrypt=False)
d = SignalProcessor(2.4e6).process(iq, 0.0)
print(len(dec.decode(d)), len(dec.decode(np.tile(d, 8))))
p = TetraProtocolParser(); r = p.parse_burst_batch(synth.make_burst_batch(1001, seed=2)); print(r["crc_ok"].sum())
an = SpectrumAnalyzer(2.4e6, 4096); print(an.run(iq[:4096*5], denoise=True)["power"].shape)
pipe = pipeline.ReceivePipeline(2.4e6, 131072, max_chunks=3, mode="fast", n_lanes=2)
x = synth.make_capture(3*131072, 130, seed=3).reshape(3, -1)
print(pipe.run_host(x)[2])
bits, t1, init = etsi.make_etsi_stream(6, seed=1)
xe = etsi.modulate_dqpsk(bits, 50000.0)
rx = etsi.EtsiReceiver(50000.0, 2, len(xe)); print(len(rx.run(np.stack([xe, xe]), init)[0]))
ch = etsi.Channelizer(16); print(ch.run(iq[:5000]).shape)
cd = etsi.ChannelDecoder("SCH/F"); print(cd.decode(etsi.to_soft(etsi.encode(np.zeros((3,268),np.uint8),"SCH/F",3)),3)[1])
print("sanitize run ok")
