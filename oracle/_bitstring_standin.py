"""Minimal stand-in for ``bitstring.BitArray`` (TEST INFRASTRUCTURE ONLY).

The reference imports ``bitstring`` (core/decoder.py:6, core/protocol.py:25), which is
not installed in this image and cannot be installed (no network).  Only the surface the
reference actually touches is provided: construction from an
iterable of 0/1 or from ``bytes``, ``len``, slicing, ``.uint``, ``.int``, ``.bin``,
``.tobytes()``.  Used by ``oracle/ref_bridge.py`` when generating golden fixtures in the
build container; never imported by the product package.
"""


class BitArray:
    __slots__ = ("_b",)

    def __init__(self, auto=None, bytes=None, bin=None, length=None):  # noqa: A002
        if bytes is not None:
            auto = bytes
        if bin is not None:
            self._b = [1 if c == "1" else 0 for c in bin if c in "01"]
        elif auto is None:
            self._b = []
        elif isinstance(auto, (builtins_bytes, bytearray)):
            self._b = [(byte >> (7 - i)) & 1 for byte in auto for i in range(8)]
        elif isinstance(auto, BitArray):
            self._b = list(auto._b)
        else:
            self._b = [1 if bool(x) else 0 for x in auto]

    def __len__(self):
        return len(self._b)

    def __getitem__(self, key):
        if isinstance(key, slice):
            out = BitArray()
            out._b = self._b[key]
            return out
        return bool(self._b[key])

    def __iter__(self):
        return iter(bool(x) for x in self._b)

    @property
    def bin(self):
        return "".join("1" if x else "0" for x in self._b)

    @property
    def uint(self):
        if not self._b:
            raise ValueError("Cannot interpret a zero length bitstring as an integer.")
        v = 0
        for x in self._b:
            v = (v << 1) | x
        return v

    @property
    def int(self):
        u = self.uint
        n = len(self._b)
        return u - (1 << n) if self._b[0] else u

    @property
    def hex(self):
        return self.tobytes().hex()[: (len(self._b) + 3) // 4]

    def tobytes(self):
        pad = (-len(self._b)) % 8
        bits = self._b + [0] * pad
        out = bytearray()
        for i in range(0, len(bits), 8):
            v = 0
            for x in bits[i:i + 8]:
                v = (v << 1) | x
            out.append(v)
        return builtins_bytes(out)

    def __eq__(self, other):
        return isinstance(other, BitArray) and self._b == other._b

    def __repr__(self):
        return f"BitArray(bin='{self.bin}')"


import builtins as _b  # noqa: E402

builtins_bytes = _b.bytes
