"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE.  The GPU box has no /root/reference, so nothing in the ``-m gpu``
tests, ``smoke()`` or ``bench.py`` may import this module; it exists for
``tests/golden/make_golden.py`` and for the container-only cross-checks in
``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TETRAEAR_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tetraear"))


def load():
    """Return (SignalProcessor, TetraDecoder, TetraProtocolParser) classes of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if "bitstring" not in sys.modules:
        from . import _bitstring_standin
        mod = types.ModuleType("bitstring")
        mod.BitArray = _bitstring_standin.BitArray
        sys.modules["bitstring"] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)      # at the END: the reference also has a top-level `tests` package
    from tetraear.signal.processor import SignalProcessor
    from tetraear.core.decoder import TetraDecoder
    from tetraear.core.protocol import TetraProtocolParser
    return SignalProcessor, TetraDecoder, TetraProtocolParser
