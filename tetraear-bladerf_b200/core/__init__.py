"""Mirror of ``tetraear.core`` for the burst/frame decoder interfaces on the IQ->bursts path
(reference: tetraear/core/__init__.py:10-20).  Crypto, validator and location helpers are out
of scope and stay with the reference."""
from .protocol import TetraProtocolParser, TetraBurst, MacPDU, BurstType, PDUType
from .decoder import TetraDecoder

__all__ = ["TetraDecoder", "TetraProtocolParser", "TetraBurst", "MacPDU", "BurstType", "PDUType"]
