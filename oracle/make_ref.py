"""Stage the UNMODIFIED reference modules of the IQ->bursts path under oracle/_ref/ (TEST INFRASTRUCTURE).

The reference is pure Python; the GPU box has no /root/reference, but oracle/_ref/ is git-ignored build output
that travels with the repo snapshot (like the built .so files).  This recipe copies, byte for byte, the seven
files that `tetraear.signal.processor.SignalProcessor` and `tetraear.core.decoder.TetraDecoder` need to import
(their transitive closure inside the package; third-party needs are numpy, scipy and `bitstring`, for which
oracle/_bitstring_standin.py stands in) and writes a SHA-256 manifest.  Consumers: `bench.py --impl reference` and
the `cpu_baseline` leg (the reference's own CPU implementation timed on the box's host cores) and the live
cross-checks of tests/test_oracle_vs_reference.py.  Nothing under oracle/_ref/ is ever imported by the product.

    python oracle/make_ref.py        # needs /root/reference (the build container)
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("TETRAEAR_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = [
    "tetraear/__init__.py",
    "tetraear/core/__init__.py",
    "tetraear/core/crypto.py",
    "tetraear/core/protocol.py",
    "tetraear/core/decoder.py",
    "tetraear/signal/__init__.py",
    "tetraear/signal/processor.py",
]


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, f)) for f in FILES)


def build(force: bool = False) -> bool:
    """Copy the files when the source tree is present; returns whether oracle/_ref is complete afterwards."""
    if not os.path.isdir(os.path.join(SRC_ROOT, "tetraear")):
        return available()
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(OUT, rel)
        with open(src, "rb") as f:
            data = f.read()
        manifest[rel] = hashlib.sha256(data).hexdigest()
        if force or not os.path.exists(dst) or open(dst, "rb").read() != data:
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            tmp = dst + f".{os.getpid()}.tmp"
            with open(tmp, "wb") as f:
                f.write(data)
            os.replace(tmp, dst)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC_ROOT, "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    print("oracle/_ref complete:", build(force=True))
