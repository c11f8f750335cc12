"""ORACLE — TEST INFRASTRUCTURE ONLY.  Recipe for `oracle/_ref/`.

The reference is three loose Python files (no package, no build system).  This recipe
places byte-identical copies of the two files the hot path lives in
(`/root/reference/nets.py`, `/root/reference/main.py`) into the git-ignored
`oracle/_ref/` so that they travel to the GPU box with the snapshot (like the built
`.so`), where `/root/reference` does not exist.  Nothing under `oracle/_ref/` is ever
committed or imported by the product package; `oracle/ref_shims.py` imports it for
 * `bench.py --impl reference` (the reference's own classes timed on the host cores),
 * the boundary test that runs the reference's unmodified `Handler` loops on the
   cgs_b200 classes (tests/test_gpu_boundary.py).

    python oracle/make_ref.py        # authoring container only; no-op without /root/reference
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CGS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("nets.py", "main.py")


def make(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "nets.py")):
        if verbose:
            print(f"oracle/make_ref: {SRC} not present; keeping {DST} as it is")
        return os.path.isfile(os.path.join(DST, "nets.py"))
    os.makedirs(DST, exist_ok=True)
    sums = []
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        sums.append(f"{hashlib.sha256(open(os.path.join(DST, f), 'rb').read()).hexdigest()}  {f}")
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(sums) + "\n")
    if verbose:
        print("oracle/make_ref: copied", ", ".join(FILES), "->", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
