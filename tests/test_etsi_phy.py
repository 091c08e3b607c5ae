"""`etsi` mode physical layer on the GPU: modulator -> RRC/timing/soft demod -> TS search -> channel
decoder.  No reference implementation exists (parity unpinned): truth = what was transmitted."""
import numpy as np
import pytest

from tetraear_b200 import etsi


def test_modulator_ideal_demod_cpu():
    bits, t1, init = etsi.make_etsi_stream(3, seed=1)
    x = etsi.modulate_dqpsk(bits, 72000.0)
    sps, span = 4, 6
    h = etsi.rrc_pulse(np.arange(-span * sps, span * sps + 1) / sps)
    y = np.convolve(x, h) / sps
    s = y[(np.arange(len(bits) // 2 + 1) + span) * sps + span * sps]
    d = s[1:] * np.conj(s[:-1])
    rx = np.stack([(d.imag < 0), (d.real < 0)], 1).reshape(-1).astype(np.uint8)
    assert np.array_equal(rx, bits)
    assert bits[244:266].tolist() == etsi.TS_N.tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("fs,t_off,sigma", [(72000.0, 0.0, 0.0), (72000.0, 0.37, 0.05), (50000.0, 0.21, 0.05),
                                           (36000.0, 0.6, 0.03)])
def test_receiver_round_trip(fs, t_off, sigma):
    import __graft_entry__ as entry
    entry.build()
    n_bursts = 24
    streams, truth = [], []
    for s in range(3):
        bits, t1, init = etsi.make_etsi_stream(n_bursts, seed=10 + s)
        streams.append(etsi.modulate_dqpsk(bits, fs, sigma=sigma, t_offset=t_off + 0.1 * s, seed=s))
        truth.append(t1)
    n = min(len(v) for v in streams)
    x = np.stack([v[:n] for v in streams])
    rx = etsi.EtsiReceiver(fs, max_streams=3, max_samples=n)
    res = rx.run(x, init)
    for s in range(3):
        found = res[s]
        assert len(found) >= n_bursts - 4, (s, len(found))          # the loop needs a burst or so to lock
        good = 0
        for b in found:
            # identify the transmitted burst by its position (bursts are 510 bits back to back)
            idx = int(round(b["start"] / 510.0))
            assert abs(b["start"] - 510 * idx) <= 4
            if b["crc_ok"].all() and 0 <= idx < n_bursts:
                assert np.array_equal(b["blocks"], truth[s][idx]), (s, idx)
                good += 1
        assert good >= n_bursts - 5, (s, good)
