"""Stages the reference's own Track-M modules for the `--impl reference` arm of bench.py — TEST INFRASTRUCTURE ONLY.

    python -m oracle.stage_reference        (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python and not an installable distribution (`pip install /root/reference` stops at setuptools'
flat-layout package discovery: eleven top-level modules, no build configuration), and /root/reference does not exist on
the GPU box.  This recipe copies the files of the benchmarked path, byte for byte, from where they lie under
/root/reference into oracle/_ref/reference/ — a git-ignored directory that travels to the GPU box with the snapshot
exactly like the built .so does — so that bench.py can time the UNMODIFIED reference (`kind: "reference"`) instead of the
oracle port.  Nothing under oracle/_ref/ is committed and nothing in the product imports it.
"""
from __future__ import annotations

import shutil
from pathlib import Path

REF = Path("/root/reference")
DST = Path(__file__).resolve().parent / "_ref" / "reference"
FILES = ["model_m.py", "trainer.py", "adapter/__init__.py", "adapter/clip_adapter.py", "adapter/peclip.py"]


def stage() -> bool:
    """True when oracle/_ref/reference holds the reference files (copied now or earlier)."""
    if not REF.exists():
        return all((DST / f).exists() for f in FILES)
    for f in FILES:
        (DST / f).parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(REF / f, DST / f)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference not available", DST)
