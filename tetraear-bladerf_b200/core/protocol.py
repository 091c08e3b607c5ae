"""``TetraProtocolParser`` — drop-in for the PHY part of
``tetraear.core.protocol.TetraProtocolParser`` (/root/reference/tetraear/core/protocol.py:142-347).

``parse_burst`` and its helpers (burst typing, data-bit extraction, the CRC-16 "soft" check,
``_calculate_crc16``) run on the GPU through ``tb200_burst_batch`` / ``tb200_crc16_rows``.
``parse_mac_pdu`` is the stateful header walk of protocol.py:349-596 kept on the host: it
touches a few dozen bits of the handful of frames a chunk yields and owns the fragment buffer
("MAC parsing stays sequential on host in burst order").  SDS / LIP / call
metadata text parsing is out of scope and stays with the reference's parser.
"""
import logging
from dataclasses import dataclass
from enum import Enum
from typing import Optional

import numpy as np

from .. import _lib

logger = logging.getLogger(__name__)


class BurstType(Enum):
    NormalUplink = 1
    NormalDownlink = 2
    ControlUplink = 3
    ControlDownlink = 4
    Synchronization = 5
    Linearization = 6


class PDUType(Enum):
    MAC_RESOURCE = 0
    MAC_FRAG = 1
    MAC_END = 2
    MAC_BROADCAST = 3
    MAC_SUPPL = 4
    MAC_U_SIGNAL = 5
    MAC_DATA = 6
    MAC_U_BLK = 7


@dataclass
class TetraBurst:
    """A TETRA burst (255 symbols) — same fields as protocol.py:66-76."""
    burst_type: BurstType
    slot_number: int
    frame_number: int
    training_sequence: np.ndarray
    data_bits: np.ndarray
    crc_ok: bool
    scrambling_code: int = 0
    colour_code: int = 0


@dataclass
class MacPDU:
    """MAC layer PDU — same fields as protocol.py:112-122."""
    pdu_type: PDUType
    encrypted: bool
    address: Optional[int]
    length: int
    data: bytes
    fill_bits: int = 0
    encryption_mode: int = 0
    reassembled_data: Optional[bytes] = None


def _bits_value(bits):
    v = 0
    for b in bits:
        v = (v << 1) | int(b)
    return v


def _bits_bytes(bits):
    """BitArray(bits).tobytes(): MSB-first, zero padded on the right."""
    b = np.asarray(bits, dtype=np.uint8)
    return np.packbits(b).tobytes() if len(b) else b""


class TetraProtocolParser:
    SYMBOLS_PER_SLOT = 255
    SLOTS_PER_FRAME = 4
    FRAMES_PER_MULTIFRAME = 18
    MULTIFRAMES_PER_HYPERFRAME = 60
    SYNC_CONTINUOUS_DOWNLINK = [1, 1, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 1, 0, 0]
    SYNC_DISCONTINUOUS_DOWNLINK = [0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 0, 0, 1, 1]

    def __init__(self, device=0):
        self.current_frame_number = 0
        self.current_multiframe = 0
        self.current_hyperframe = 0
        self.mcc = None
        self.mnc = None
        self.la = None
        self.colour_code = None
        self.stats = {
            'total_bursts': 0, 'crc_pass': 0, 'crc_fail': 0, 'clear_mode_frames': 0, 'encrypted_frames': 0,
            'decrypted_frames': 0, 'voice_calls': 0, 'data_messages': 0, 'control_messages': 0,
        }
        self.fragment_buffer = bytearray()
        self.fragment_metadata = {}
        self.device = device
        self._h = None

    # ------------------------------------------------------------------ GPU plumbing
    def _handle(self):
        if self._h is None:
            self._h = _lib.Handle(max_chunks=0, max_chunk_len=0, device=self.device, max_bursts=4096)
        return self._h

    def parse_burst_batch(self, dibits):
        """[n, 255] uint8 -> dict(burst_type, crc_ok, mac_ok, hdr, data_bits [n,510], n_data); updates stats.
        The batched form of parse_burst (C3)."""
        d = np.ascontiguousarray(dibits, dtype=np.uint8).reshape(-1, 255)
        n = d.shape[0]
        h = self._handle()
        rec = np.empty((n, 32), np.uint8)
        full = np.empty((n, 64), np.uint8)
        stats = np.zeros(3, np.int64)
        step = 4096
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            _lib.check(h.lib.tb200_burst_batch_host(h.h, d[lo:hi].ctypes.data, hi - lo, rec[lo:hi].ctypes.data,
                                                    full[lo:hi].ctypes.data, stats.ctypes.data, None))
        self.stats['total_bursts'] += int(stats[0])
        self.stats['crc_pass'] += int(stats[1])
        self.stats['crc_fail'] += int(stats[2])
        return unpack_burst_records(rec, full)

    # ------------------------------------------------------------------ reference API
    def parse_burst(self, symbols, slot_number: int = 0) -> Optional[TetraBurst]:
        """protocol.py:192-244."""
        if len(symbols) < self.SYMBOLS_PER_SLOT:
            logger.warning(f"Insufficient symbols for burst: {len(symbols)} < {self.SYMBOLS_PER_SLOT}")
            return None
        row = (np.asarray(symbols[:self.SYMBOLS_PER_SLOT]).astype(np.int64) & 3).astype(np.uint8)
        r = self.parse_burst_batch(row[None, :])
        is_sync = int(r["burst_type"][0]) == BurstType.Synchronization.value
        bits = r["all_bits"][0]
        nd = int(r["n_data"][0])
        return TetraBurst(
            burst_type=BurstType(int(r["burst_type"][0])), slot_number=slot_number,
            frame_number=self.current_frame_number,
            training_sequence=bits[108:130].astype(np.int64) if is_sync else bits[108:122].astype(np.int64),
            data_bits=r["data_bits"][0, :nd].astype(np.int64), crc_ok=bool(r["crc_ok"][0]),
            colour_code=self.colour_code or 0)

    def _burst_bits_to_row(self, bits):
        b = np.zeros(510, np.uint8)
        v = np.asarray(bits).astype(np.uint8)[:510] & 1
        b[:len(v)] = v
        return ((b[0::2] << 1) | b[1::2])[None, :]

    def _detect_burst_type(self, bits) -> BurstType:
        """protocol.py:246-254 (expects the 510 burst bits)."""
        sync_pos = len(bits) // 2
        return BurstType.Synchronization if self._check_sync_pattern(bits[sync_pos:sync_pos + 22]) else BurstType.NormalDownlink

    def _check_sync_pattern(self, bits) -> bool:
        """protocol.py:256-265 — evaluated by the burst kernel on a burst carrying `bits` at 255..276."""
        if len(bits) < 22:
            return False
        burst = np.zeros(510, np.uint8)
        burst[255:277] = np.asarray(bits[:22]).astype(np.uint8) & 1
        saved = dict(self.stats)
        r = self.parse_burst_batch(self._burst_bits_to_row(burst))
        self.stats = saved
        return int(r["burst_type"][0]) == BurstType.Synchronization.value

    def _extract_training_sequence(self, bits, burst_type):
        return bits[108:130] if burst_type == BurstType.Synchronization else bits[108:122]

    def _extract_data_bits(self, bits, burst_type):
        """protocol.py:277-290 (pure slicing, no arithmetic)."""
        if burst_type in (BurstType.NormalDownlink, BurstType.NormalUplink):
            return np.concatenate([bits[0:108], bits[122:230]])
        return bits

    def _crc16_rows(self, rows):
        import torch
        from .. import _dev
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        h = self._handle()
        d = _dev.to_dev(rows, self.device)
        out = _dev.empty((rows.shape[0],), torch.int16, self.device)
        _lib.check(h.lib.tb200_crc16_rows(h.h, d.data_ptr(), rows.shape[0], rows.shape[1], rows.shape[1],
                                          out.data_ptr(), _dev.stream_ptr(self.device)))
        return out.cpu().numpy().view(np.uint16)

    def _calculate_crc16(self, bits) -> np.ndarray:
        """protocol.py:331-347 -> array of 16 bits, MSB first."""
        b = (np.asarray(bits).astype(np.int64) & 1).astype(np.uint8)
        crc = 0xFFFF if len(b) == 0 else int(self._crc16_rows(b[None, :])[0])
        return np.array([(crc >> i) & 1 for i in range(15, -1, -1)])

    def _check_crc(self, bits) -> bool:
        """protocol.py:292-329 for a bit array of any length (two CRC rows on the GPU)."""
        bits = np.asarray(bits)
        if len(bits) < 16:
            return False
        b = (bits.astype(np.int64) & 1).astype(np.uint8)
        ones = int(b.sum())
        if ones == 0 or ones == len(b):
            return False
        payload, rx = b[:-16], b[-16:]
        if len(payload) == 0:
            crcs = np.array([0xFFFF, 0xFFFF], np.uint16)
        else:
            crcs = self._crc16_rows(np.stack([payload, payload[::-1]]))
        rxv = _bits_value(rx)
        return bool(bin(int(crcs[0]) ^ rxv).count("1") <= 2 or bin(int(crcs[1]) ^ rxv).count("1") <= 2)

    # ------------------------------------------------------------------ MAC header walk (host, stateful)
    def parse_mac_pdu(self, bits) -> Optional[MacPDU]:
        """protocol.py:349-596: type / encryption mode / address / length fields, fragment buffer."""
        n = len(bits)
        if n < 8:
            return None
        b = np.asarray(bits).astype(np.int64) & 1
        kind = int(b[0]) * 2 + int(b[1])
        pdu_type = (PDUType.MAC_RESOURCE, PDUType.MAC_FRAG, PDUType.MAC_BROADCAST, PDUType.MAC_END)[kind]
        mode = int(b[2]) * 2 + int(b[3])
        encrypted = mode > 0
        address, length, data, fill = None, 0, b"", 0

        def sized_payload(pos, length):
            nbits = length * 8
            if nbits > n - pos + 16:
                return None
            seg = b[pos:pos + nbits] if (nbits > 0 and n >= pos + nbits) else b[pos:]
            return _bits_bytes(seg)

        if pdu_type == PDUType.MAC_RESOURCE:
            fill = int(b[4])
            if n < 5 + 24 + 6:
                return None
            address = _bits_value(b[5:29])
            length = _bits_value(b[29:35])
            data = sized_payload(35, length)
            if data is None:
                return None
            self.fragment_buffer = bytearray(data)
            self.fragment_metadata = {'address': address, 'encrypted': encrypted, 'mode': mode}
        elif pdu_type == PDUType.MAC_FRAG:
            fill = int(b[4])
            data = _bits_bytes(b[5:])
            self.fragment_buffer.extend(data)
            if self.fragment_metadata:
                encrypted = self.fragment_metadata.get('encrypted', False)
                address = self.fragment_metadata.get('address')
        elif pdu_type == PDUType.MAC_BROADCAST:
            if mode == 0:                                   # SYSINFO
                if n < 4 + 30:
                    return None
                self.mcc = _bits_value(b[4:14])
                self.mnc = _bits_value(b[14:28])
                self.colour_code = _bits_value(b[28:34])
                if self.mcc < 200 or self.mcc > 799 or self.mnc > 999:
                    return None
            data = _bits_bytes(b[4:])
        else:                                               # MAC-END
            fill = int(b[4])
            if n < 5 + 6:
                return None
            length = _bits_value(b[5:11])
            data = sized_payload(11, length)
            if data is None:
                return None
            self.fragment_buffer.extend(data)
            if self.fragment_metadata:
                encrypted = self.fragment_metadata.get('encrypted', False)
                address = self.fragment_metadata.get('address')
        self.stats['encrypted_frames' if encrypted else 'clear_mode_frames'] += 1
        pdu = MacPDU(pdu_type=pdu_type, encrypted=encrypted, address=address, length=length, data=data,
                     fill_bits=fill, encryption_mode=mode)
        if pdu_type == PDUType.MAC_END:
            if self.fragment_buffer:
                pdu.reassembled_data = bytes(self.fragment_buffer)
                if self.fragment_metadata:
                    if not pdu.address:
                        pdu.address = self.fragment_metadata.get('address')
                    pdu.encrypted = self.fragment_metadata.get('encrypted', False)
                self.fragment_buffer = bytearray()
                self.fragment_metadata = {}
        elif pdu_type == PDUType.MAC_RESOURCE:
            pdu.reassembled_data = bytes(data)
        return pdu

    def get_statistics(self):
        """protocol.py:1261-1275."""
        total = self.stats['clear_mode_frames'] + self.stats['encrypted_frames']
        clear_pct = self.stats['clear_mode_frames'] / total * 100 if total > 0 else 0
        enc_pct = self.stats['encrypted_frames'] / total * 100 if total > 0 else 0
        return {**self.stats, 'clear_mode_percentage': clear_pct, 'encrypted_percentage': enc_pct,
                'crc_success_rate': (self.stats['crc_pass'] / max(1, self.stats['total_bursts'])) * 100}


def unpack_burst_records(rec, full):
    """tb200_burst_batch output ([n,32] records, [n,64] packed burst bits) -> numpy fields."""
    rec = np.asarray(rec)
    n = rec.shape[0]
    burst_type = rec[:, 27].copy()
    all_bits = np.unpackbits(np.asarray(full), axis=1)[:, :510]
    n_data = np.where(burst_type == 5, 510, 216).astype(np.int32)
    data_bits = np.zeros((n, 510), np.uint8)
    normal = np.unpackbits(rec[:, :27], axis=1)            # [n, 216]
    data_bits[:, :216] = normal
    sync_rows = burst_type == 5
    data_bits[sync_rows] = all_bits[sync_rows]
    return {"burst_type": burst_type, "crc_ok": rec[:, 28].copy(), "mac_ok": rec[:, 29].copy(), "hdr": rec[:, 30].copy(),
            "sync_match": rec[:, 31].copy(), "n_data": n_data, "data_bits": data_bits, "all_bits": all_bits}
