"""Import the UNMODIFIED reference: from /root/reference in the build container, else from oracle/_ref/.

TEST INFRASTRUCTURE.  The GPU box has no /root/reference; ``oracle/make_ref.py`` stages byte-identical copies of
the seven modules of the path under the git-ignored ``oracle/_ref/`` (it travels with the snapshot like the built
libraries), so that ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's own code there and the
cross-checks of ``tests/test_oracle_vs_reference.py`` run on both sides.  ``tests/golden/make_golden.py`` uses the
full tree (it also reads ui/modern.py source lines).  The product never imports this module.
"""
import os
import sys
import types

_FULL = os.environ.get("TETRAEAR_REFERENCE", "/root/reference")
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FULL_TREE = os.path.isdir(os.path.join(_FULL, "tetraear"))            # ui/modern.py etc. present (build container)
REFERENCE_ROOT = _FULL if FULL_TREE else _STAGED


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "tetraear", "signal", "processor.py"))


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
