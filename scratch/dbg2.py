import numpy as np, sys
sys.path.insert(0,'.')
import __graft_entry__ as g; g.build()
from oracle import oracle
from tetraear_b200 import synth, pipeline, _lib
n_chunks, L = 2, 131072
cap = synth.make_capture(n_chunks * L, 130, seed=4321, mix=(0.6, 0.2, 0.2))
x = cap.reshape(n_chunks, L).copy()
pipe = pipeline.ReceivePipeline(2.4e6, L, max_chunks=8, keep_symbols=True)
pipe.run_host(x)
dec = np.empty((2, 13108), np.complex64)
_lib.check(pipe.h.lib.tb200_debug_read(pipe.h.h, 0, dec.ctypes.data, dec.nbytes))
flt = np.empty((2, 13108), np.complex128)
_lib.check(pipe.h.lib.tb200_debug_read(pipe.h.h, 1, flt.ctypes.data, flt.nbytes))
for c in range(2):
    od = oracle.decimate(x[c], 10)
    e = np.abs(dec[c]-od); print('dec', c, e.max(), np.argmax(e), e[-5:])
    of = oracle.filter_signal(od, 25000, 240000.0)
    e = np.abs(flt[c]-of); print('flt', c, e.max(), np.argmax(e), e[-5:])
