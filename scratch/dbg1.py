import numpy as np, sys
sys.path.insert(0,'.')
from oracle import oracle
from tetraear_b200 import synth, pipeline
from tetraear_b200.signal import SignalProcessor
n_chunks, L = 6, 131072
cap = synth.make_capture(n_chunks * L, 130, seed=4321, mix=(0.6, 0.2, 0.2))
x = cap.reshape(n_chunks, L).copy()
sp = SignalProcessor(2.4e6)
for c in range(n_chunks):
    d = sp.process(x[c], 0.0)
    od, osym, info = oracle.process(x[c], 2.4e6, 0.0)
    e = np.abs(sp.symbols - osym)
    print('single', c, np.array_equal(d, od), e.max(), np.argmax(e), len(osym), info['best_phase'])
    dec = sp.decimate(x[c], 10); odec = oracle.decimate(x[c], 10)
    ed = np.abs(dec-odec); print('   decim', ed.max(), np.argmax(ed))
    fl = sp.filter_signal(odec, 25000, 240000.0); ofl = oracle.filter_signal(odec, 25000, 240000.0)
    ef = np.abs(fl-ofl); print('   filt', ef.max(), np.argmax(ef))
for nb in (1,2,6):
    pipe = pipeline.ReceivePipeline(2.4e6, L, max_chunks=8, keep_symbols=True)
    pipe.run_host(x[:nb])
    sym = pipe.symbols.cpu().numpy().view(np.complex128)[..., 0]
    for c in range(nb):
        od, osym, info = oracle.process(x[c], 2.4e6, 0.0)
        e = np.abs(sym[c,:len(osym)]-osym); print('batch',nb,c, e.max(), np.argmax(e))
