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


def test_bsch_fields_round_trip_cpu():
    f = dict(system_code=1, colour_code=17, timeslot=2, frame=11, multiframe=42, sharing_mode=1, ts_reserved_frames=3,
             u_plane_dtx=1, frame18_extension=0, reserved=0, mcc=262, mnc=42, neighbour_cell_broadcast=2,
             cell_service_level=1, late_entry=1)
    b = etsi.bsch_pack(**f)
    assert len(b) == 60 and etsi.bsch_unpack(b) == f
    bits, t1, init = etsi.make_etsi_stream(9, seed=2, sync_every=4)
    assert bits[214:252].tolist() == etsi.TS_Y.tolist() and bits[14:94].tolist() == etsi.FREQ_CORRECTION.tolist()
    assert bits[510 + 244:510 + 266].tolist() == etsi.TS_N.tolist()
    assert init == etsi.scrambling_init(262, 42, 17)


@pytest.mark.gpu
@pytest.mark.parametrize("fs,sigma", [(72000.0, 0.0), (50000.0, 0.05)])
def test_sync_burst_acquisition_bootstraps_the_scrambling_code(fs, sigma):
    """No network code given: find the synchronisation bursts (38-bit sequence y), decode BSCH (init 3), read
    MCC / MNC / colour code, derive the scrambling init and decode the normal bursts with it."""
    import __graft_entry__ as entry
    entry.build()
    n_bursts, k = 26, 6
    nets = [(262, 42, 17), (901, 16383, 63), (1, 0, 0)]
    streams, truth = [], []
    for s, net in enumerate(nets):
        bits, t1, init = etsi.make_etsi_stream(n_bursts, seed=40 + s, sync_every=k, network=net)
        streams.append(etsi.modulate_dqpsk(bits, fs, sigma=sigma, t_offset=0.15 * s, seed=s))
        truth.append((t1, init))
    n = min(len(v) for v in streams)
    x = np.stack([v[:n] for v in streams])
    rx = etsi.EtsiReceiver(fs, max_streams=3, max_samples=n)
    acq = rx.acquire(x)
    for s, net in enumerate(nets):
        assert len(acq[s]) >= n_bursts // k - 1, (s, len(acq[s]))          # the first burst may precede timing lock
        for a in acq[s]:
            idx = int(round(a["start"] / 510.0))
            assert abs(a["start"] - 510 * idx) <= 4 and idx % k == 0
            f = a["fields"]
            assert (f["mcc"], f["mnc"], f["colour_code"]) == net and f["timeslot"] == idx % 4
            assert a["scrambling_init"] == truth[s][1]
    # the derived code decodes the traffic of its own stream
    for s in range(3):
        res = rx.run(x[s:s + 1], acq[s][0]["scrambling_init"])[0]
        good = 0
        for b in res:
            idx = int(round(b["start"] / 510.0))
            if idx % k and b["crc_ok"].all() and 0 <= idx < n_bursts:
                assert np.array_equal(b["blocks"], truth[s][0][idx])
                good += 1
        assert good >= (n_bursts - n_bursts // k) - 6, (s, good)


@pytest.mark.gpu
def test_carrier_offset_correction():
    """Feed-forward AFC (tb200_etsi_afc): carrier offsets up to +-2 kHz are estimated to a few Hz from sum d^4 and
    removed before the searches."""
    import __graft_entry__ as entry
    entry.build()
    fs, n_bursts = 50000.0, 20
    offs = [-2000.0, -900.0, 0.0, 650.0, 1800.0]
    streams, truth = [], []
    for s, f in enumerate(offs):
        bits, t1, init = etsi.make_etsi_stream(n_bursts, seed=70 + s)
        streams.append(etsi.modulate_dqpsk(bits, fs, sigma=0.04, freq=f, t_offset=0.1 * s, seed=s))
        truth.append(t1)
    n = min(len(v) for v in streams)
    x = np.stack([v[:n] for v in streams])
    rx = etsi.EtsiReceiver(fs, max_streams=len(offs), max_samples=n)
    res = rx.run(x, init)
    est = rx.carrier_offsets_hz(len(offs))
    with_afc = []
    for s, f in enumerate(offs):
        assert abs(est[s] - f) < 15.0, (f, est[s])
        good = sum(1 for b in res[s] if b["crc_ok"].all() and np.array_equal(b["blocks"], truth[s][int(round(b["start"] / 510.0))]))
        assert good >= n_bursts - 5, (f, good)
        with_afc.append(good)
    # the same capture without the correction: the rotated decisions cost bursts at +-2 kHz (40 degrees of the 45 available)
    rx0 = etsi.EtsiReceiver(fs, max_streams=len(offs), max_samples=n, afc=False)
    res0 = rx0.run(x, init)
    without = [sum(1 for b in res0[s] if b["crc_ok"].all()) for s in range(len(offs))]
    assert all(w <= a for w, a in zip(without, with_afc)), (without, with_afc)
    assert without[0] < with_afc[0] or without[4] < with_afc[4], (without, with_afc)


@pytest.mark.gpu
@pytest.mark.parametrize("fs,sigma,rate_ppm", [(50000.0, 0.0, 0.0), (50000.0, 0.05, 0.0), (72000.0, 0.05, 60.0), (36000.0, 0.03, -40.0)])
def test_segmented_demodulator_matches_the_sequential_loop(fs, sigma, rate_ppm):
    """tb200_etsi_demod_segmented cuts every stream into concurrently running segments (warm-up, then symbol-exact
    stitching): same number of symbols and the same decisions as the one sequential loop, also when the transmitter's
    symbol clock is off (the segment boundaries are nominal instants, the loops track the real ones)."""
    import __graft_entry__ as entry
    entry.build()
    n_bursts = 40                                                     # 10 200 symbols: 16 segments of 640
    streams = []
    for s in range(4):
        bits, t1, init = etsi.make_etsi_stream(n_bursts, seed=90 + s)
        streams.append(etsi.modulate_dqpsk(bits, fs, symbol_rate=18000.0 * (1 + rate_ppm * 1e-6), sigma=sigma, t_offset=0.13 * s + 0.2,
                                           seed=s))
    n = min(len(v) for v in streams)
    x = np.stack([v[:n] for v in streams])
    out = {}
    for seg in (False, True, 8):                                       # sequential, two-stage (default), eight-lane
        rx = etsi.EtsiReceiver(fs, max_streams=4, max_samples=n, segmented=bool(seg), afc=False, seg_lanes=8 if seg == 8 else 0)
        res = rx.run(x, init)
        out[seg] = (rx.n_soft.cpu().numpy().copy(), rx.soft.cpu().numpy().copy(), res)
        assert (rx._scratch is not None) == bool(seg)                  # the segmented forms did run
    ns0, soft0, res0 = out[False]
    for form in (True, 8):
      ns1, soft1, res1 = out[form]
      for s in range(4):
          assert abs(int(ns0[s]) - int(ns1[s])) <= 2 * 2, (s, ns0[s], ns1[s])      # the last symbols may fall either side of the end
          m = min(int(ns0[s]), int(ns1[s]))
          assert m > 2 * 255 * (n_bursts - 1)
          a, b = soft0[s, :m], soft1[s, :m]
          # skip the first burst (both loops are still pulling in there, identically, but decisions on it are not the point)
          agree = np.mean((a[510:] > 0) == (b[510:] > 0))
          assert agree >= (1.0 if sigma == 0.0 else 0.9995), (s, agree)
          assert np.max(np.abs(a[510:] - b[510:])) < (0.25 if sigma == 0.0 else 0.6)    # timing jitter of two loops, same decisions
          good0 = sum(bool(v["crc_ok"].all()) for v in res0[s])
          good1 = sum(bool(v["crc_ok"].all()) for v in res1[s])
          assert good1 >= good0 - 1 and good1 >= n_bursts - 5, (s, good0, good1)


@pytest.mark.gpu
@pytest.mark.parametrize("fs,sigma", [(72000.0, 0.0), (50000.0, 0.08)])
def test_training_sequence_search_all_sequences(fs, sigma):
    """tb200_etsi_find_sequence: every training sequence of cl. 9.4.4.3 is found where the modulator put it — the tail
    sequence q across every slot boundary (q(1..10) of one burst + q(11..22) of the next), normal n / p in the middle of
    their bursts (bit 244), the sync sequence y (bit 214) and an extended sequence x planted in one block."""
    import __graft_entry__ as entry
    entry.build()
    n_bursts = 20
    bits, t1, init = etsi.make_etsi_stream(n_bursts, seed=5, sync_every=7)
    bits_p, _, _ = etsi.make_etsi_stream(n_bursts, seed=6, ts=etsi.TS_P)
    x_at = 510 * 9 + 60                                              # inside BKN1 of burst 9, on a symbol boundary
    bits_p = bits_p.copy()
    bits_p[x_at:x_at + 30] = etsi.TS_X
    xs = [etsi.modulate_dqpsk(b, fs, sigma=sigma, t_offset=0.2 * i, seed=i) for i, b in enumerate((bits, bits_p))]
    n = min(len(v) for v in xs)
    rx = etsi.EtsiReceiver(fs, max_streams=2, max_samples=n)
    rx.run(np.stack([v[:n] for v in xs]), init)                     # demodulates; soft bits stay on the device
    lock = 2 * 510                                                   # the timing loop may need a burst or two

    def near(hits, want, tol=4):
        got = np.array([p for p, c in hits])
        return [w for w in want if len(got) and np.abs(got - w).min() <= tol]

    q_hits = rx.find_sequence(2, "q", threshold=0.8)
    for s in range(2):
        want = [510 * k - 10 for k in range(1, n_bursts) if 510 * k > lock]
        assert len(near(q_hits[s], want)) >= len(want) - 1, (s, q_hits[s])
        assert all(c <= 1.0 + 1e-5 for p, c in q_hits[s])
    n_hits = rx.find_sequence(2, "n", threshold=0.8)
    want_n = [510 * k + 244 for k in range(n_bursts) if k % 7 and 510 * k > lock]
    assert len(near(n_hits[0], want_n)) >= len(want_n) - 1
    assert not near(n_hits[1], [510 * k + 244 for k in range(n_bursts)])          # stream 1 carries p, not n
    p_hits = rx.find_sequence(2, "p", threshold=0.8)
    want_p = [510 * k + 244 for k in range(n_bursts) if 510 * k > lock]
    assert len(near(p_hits[1], want_p)) >= len(want_p) - 1 and not near(p_hits[0], want_n)
    y_hits = rx.find_sequence(2, "y", threshold=0.8)
    want_y = [510 * k + 214 for k in range(0, n_bursts, 7) if 510 * k > lock]
    assert near(y_hits[0], want_y) == want_y and len(y_hits[1]) == 0
    x_hits = rx.find_sequence(2, "x", threshold=0.85, min_gap_symbols=20)
    assert near(x_hits[1], [x_at]) == [x_at]
    # x = n(12..22) n(1..19): it also lights up at every normal sequence n, 11 bits early, with partial correlation only
    assert all(c < 0.85 or abs(p - x_at) <= 4 for p, c in x_hits[1])


@pytest.mark.gpu
@pytest.mark.parametrize("fs,sigma", [(72000.0, 0.0), (50000.0, 0.1)])
def test_aach_through_the_receive_chain(fs, sigma):
    """Broadcast block of every normal burst: modulator -> demodulator -> burst search -> tb200_etsi_extract_bits ->
    (30, 14) soft ML decoder == the access-assignment word that was sent."""
    import __graft_entry__ as entry
    entry.build()
    n_bursts = 24
    bits, t1, init, words = etsi.make_etsi_stream(n_bursts, seed=21, aach=True)
    x = etsi.modulate_dqpsk(bits, fs, sigma=sigma, t_offset=0.3, seed=2)
    rx = etsi.EtsiReceiver(fs, max_streams=1, max_samples=len(x))
    res = rx.run(x[None, :], init, aach=True)[0]
    assert len(res) >= n_bursts - 4
    good = 0
    for b in res:
        idx = int(round(b["start"] / 510.0))
        if 0 <= idx < n_bursts and b["crc_ok"].all():
            assert np.array_equal(b["aach"], words[idx]) and b["aach_margin"] > 0, idx
            good += 1
    assert good >= n_bursts - 5
