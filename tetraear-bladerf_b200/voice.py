"""Codec-frame builder on the output side of the path: 690 x int16 per frame for the
reference's ACELP decoder, built on the GPU by ``tb200_voice_blocks``.

Mirrors ``ModernTetraGUI._extract_voice_slot_from_symbols`` (/root/reference/tetraear/ui/modern.py:2309-2423),
the block builder of ``continuous_capture.py:100-124`` and the fallback packing of ``ui/modern.py:2143-2200``.
"""
import numpy as np

from . import _lib


class VoiceBlockBuilder:
    def __init__(self, device=0):
        self.device = device
        self._h = _lib.Handle(sample_rate=2.4e6, max_chunks=0, max_chunk_len=0, mode=_lib.MODE_EXACT, device=device)

    # ---- device-level: tensors in, tensors out -----------------------------------------------------------
    def blocks_device(self, mode, src, n_frames, n_src=None, frame_stream=None, frame_pos=None, scale=127.0):
        """src: CUDA tensor (uint8 rows, or float32 rows for VOICE_SOFT); returns (int16 [n_frames, 690], uint8 [n_frames])."""
        import torch
        from . import _dev
        blocks = _dev.empty((max(n_frames, 1), 690), torch.int16, self.device)
        valid = _dev.empty((max(n_frames, 1),), torch.uint8, self.device)
        ptr = lambda t: t.data_ptr() if t is not None else None
        _lib.check(self._h.lib.tb200_voice_blocks(self._h.h, mode, src.data_ptr(), src.stride(0) if src.dim() > 1 else src.shape[0],
                                                  ptr(n_src), ptr(frame_stream), ptr(frame_pos), n_frames, float(scale),
                                                  blocks.data_ptr(), valid.data_ptr(), _dev.stream_ptr(self.device)))
        return blocks[:n_frames], valid[:n_frames]

    # ---- reference-shaped: numpy in, bytes out -----------------------------------------------------------
    def _run(self, mode, rows, lens=None, pos=None, scale=127.0):
        import torch
        from . import _dev
        n = len(pos) if pos is not None else rows.shape[0]
        if n == 0:
            return np.zeros((0, 690), np.int16), np.zeros(0, np.uint8)
        src = _dev.to_dev(rows, self.device)
        d_len = _dev.to_dev(np.asarray(lens, np.int32), self.device) if lens is not None else None
        d_pos = _dev.to_dev(np.asarray(pos, np.int32), self.device) if pos is not None else None
        b, v = self.blocks_device(mode, src, n, n_src=d_len, frame_pos=d_pos, scale=scale)
        return b.cpu().numpy(), v.cpu().numpy()

    def from_symbols(self, frames, demodulated_symbols):
        """One codec frame (1380 bytes) or None per frame dict — modern.py:2309-2423 applied to a list of frames
        of ONE chunk's symbol stream."""
        d = np.ascontiguousarray(np.asarray(demodulated_symbols).astype(np.uint8))[None, :]
        pos = [(-1 if f.get("position") is None else int(f["position"])) for f in frames]
        if d.shape[1] == 0:
            return [None] * len(frames)
        b, v = self._run(_lib.VOICE_SYMBOLS, d, lens=[d.shape[1]], pos=pos)
        return [b[i].astype("<i2").tobytes() if v[i] else None for i in range(len(frames))]

    def _rows(self, bit_rows):
        lens = np.array([len(r) for r in bit_rows], np.int32)
        rows = np.zeros((len(bit_rows), max(432, int(lens.max()) if len(lens) else 432)), np.uint8)
        for i, r in enumerate(bit_rows):
            rows[i, :len(r)] = (np.asarray(r).astype(np.int64) != 0)
        return rows, lens

    def from_bits(self, bit_rows):
        """continuous_capture.py:100-124 for a list of per-frame bit arrays (None where fewer than 432 bits)."""
        if not len(bit_rows):
            return []
        rows, lens = self._rows(bit_rows)
        b, v = self._run(_lib.VOICE_BITS, rows, lens=lens)
        return [b[i].astype("<i2").tobytes() if v[i] else None for i in range(len(bit_rows))]

    def from_params(self, bit_rows):
        """ui/modern.py:2143-2200 (fallback parameter packing)."""
        if not len(bit_rows):
            return []
        rows, lens = self._rows(bit_rows)
        b, v = self._run(_lib.VOICE_PARAMS, rows, lens=lens)
        return [b[i].astype("<i2").tobytes() if v[i] else None for i in range(len(bit_rows))]

    def from_soft(self, soft_rows, scale=127.0):
        """etsi mode: float soft bits (+ = 1) [n, >=432] -> int16 [n, 690]."""
        s = np.ascontiguousarray(soft_rows, dtype=np.float32)
        b, _ = self._run(_lib.VOICE_SOFT, s, scale=scale)
        return b
