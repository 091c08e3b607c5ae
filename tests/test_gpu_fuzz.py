"""Randomised parity of the CUDA path against the oracle (which is pinned to the reference goldens):
many small seeded cases covering ragged lengths, corrupted training sequences, 8-PSK symbol input,
random frequency offsets and sample rates.  Needs a B200."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import __graft_entry__ as entry
    entry.build()
    from tetraear_b200 import signal, core, synth
    return signal, core, synth


def test_decode_fuzz(mods):
    signal, core, synth = mods
    rng = np.random.default_rng(2024)
    n_frames_total = 0
    for case in range(120):
        kind = case % 4
        n_b = int(rng.integers(1, 14))
        bits = np.concatenate([synth.make_burst(rng, ("pass", "near", "random")[int(rng.integers(0, 3))],
                                                sync_type=bool(rng.random() < 0.1)) for _ in range(n_b)])
        # corrupt some training sequences by 0..6 bits, shift by a random bit offset, cut raggedly
        for k in range(n_b):
            ne = int(rng.integers(0, 7))
            if ne:
                bits[510 * k + 216 + rng.choice(22, ne, replace=False)] ^= 1
        shift = int(rng.integers(0, 300))
        bits = np.concatenate([rng.integers(0, 2, shift, dtype=np.uint8), bits])
        bits = bits[: len(bits) - int(rng.integers(0, 200))]
        bits = bits[: len(bits) // 2 * 2]
        d = synth.bits_to_dibits(bits)
        if kind == 3:
            d = rng.integers(0, 8, len(d), dtype=np.uint8)             # 8-PSK style input -> mapping path
        dec = core.TetraDecoder(auto_decrypt=False)
        recs = dec.decode_records(d)
        kept = [r for r in recs if r["kept"]]
        ofr, ost = oracle.decode(d)
        assert [(int(r["sync_pos"]), int(r["number"]), int(r["type"]), int(r["enc_mode"]), int(r["burst_type"]),
                 bool(r["crc_ok"]), bool(r["mac_ok"])) for r in kept] == \
               [(f["sync_pos"], f["number"], f["type"], f["enc_mode"], f["burst_type"], f["burst_crc"], f["mac_ok"])
                for f in ofr], case
        host = core.TetraProtocolParser()
        for r, f in zip(kept, ofr):
            assert np.array_equal(np.unpackbits(r["bits"])[:510], f["bits"]), case
            # N1: the fixed-offset MAC header fields on the record equal the host header walk
            nd = 510 if int(r["burst_type"]) == 5 else 216
            pdu = host.parse_mac_pdu(np.unpackbits(r["data"])[:nd])
            if pdu is not None:
                # ... and the payload statistics (length, distinct bytes: the entropy heuristic; first bytes: SDS pre-filters)
                assert int(r["mac_data_bytes"]) == len(pdu.data) and int(r["mac_unique"]) == len(set(pdu.data)), case
                assert int(r["mac_b0"]) == (pdu.data[0] if len(pdu.data) > 0 else 0), case
                assert int(r["mac_b1"]) == (pdu.data[1] if len(pdu.data) > 1 else 0), case
            else:
                assert int(r["mac_data_bytes"]) == 0 and int(r["mac_unique"]) == 0, case
            if pdu is not None and pdu.pdu_type.name in ("MAC_RESOURCE", "MAC_END"):
                assert int(r["mac_length"]) == pdu.length and int(r["mac_fill"]) == pdu.fill_bits, case
                if pdu.pdu_type.name == "MAC_RESOURCE":
                    assert int(r["mac_address"]) == pdu.address, case
        st = dec.protocol_parser.stats
        assert (st["total_bursts"], st["crc_pass"], st["crc_fail"]) == (ost["total_bursts"], ost["crc_pass"], ost["crc_fail"]), case
        n_frames_total += len(ofr)
    assert n_frames_total > 100


def test_burst_batch_fuzz(mods):
    signal, core, synth = mods
    rng = np.random.default_rng(7)
    d = synth.make_burst_batch(20000, seed=123)
    d[:500] = rng.integers(0, 4, (500, 255), dtype=np.uint8)
    d[500:520] = 0
    d[520:540, :] = 3
    p = core.TetraProtocolParser()
    r = p.parse_burst_batch(d)
    o = oracle.parse_bursts(d)
    assert np.array_equal(r["burst_type"], o["burst_type"])
    assert np.array_equal(r["crc_ok"], o["crc_ok"])
    assert np.array_equal(r["data_bits"], o["data_bits"])
    mac = np.array([oracle.mac_pdu_valid(o["data_bits"][i, :o["n_data"][i]]) for i in range(0, len(d), 7)], np.uint8)
    assert np.array_equal(r["mac_ok"][::7], mac)
    assert p.stats["crc_pass"] == int(o["crc_ok"].sum())


@pytest.mark.parametrize("mode", ["exact", "fast", "fast32"])
def test_process_fuzz(mods, mode):
    signal, core, synth = mods
    rng = np.random.default_rng(99)
    for case in range(15):
        # q = 10, 10, 7, none, 8 (all three FAST instantiations + the undecimated path)
        fs, sps_total = [(2.4e6, 130), (2.4e6, 130), (1.8e6, 98), (240000.0, 13), (2.0e6, 111)][case % 5]
        n = int(rng.integers(3000, 70000))
        f = float(rng.choice([0.0, 0.0, rng.uniform(-3000, 3000)]))
        iq = synth.make_capture(n, sps_total, seed=500 + case, sigma=0.03)
        if f:
            iq = (iq * np.exp(2j * np.pi * f * np.arange(n) / fs)).astype(np.complex64)
        sp = signal.SignalProcessor(fs, mode=mode)
        d = sp.process(iq, f)
        od, osym, _ = oracle.process(iq, fs, f)
        assert np.array_equal(d, od), (case, mode, n, fs, f)
        tol = 1e-9 if mode == "exact" else 1e-4
        assert len(sp.symbols) == len(osym) and np.abs(sp.symbols - osym).max() <= tol * max(1.0, np.abs(osym).max()), (case, mode)
