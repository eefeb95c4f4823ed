"""Recipe that stages the UNMODIFIED reference hot path for the GPU box (test / benchmark infrastructure only).

    python oracle/build_ref.py            # run in the build container; __graft_entry__.build() calls stage()

The reference is pure Python, so "building" it is copying the two modules of the hot path -- byte for byte, no edits --
from the read-only checkout into ``oracle/_ref/models/``.  ``oracle/_ref/`` is git-ignored (no reference source enters
the history) but travels with the working tree to the GPU box, exactly like the built ``libpangu_b200.so``.  There,
``oracle/ref_runner.py`` imports it through the 15-line ``timm`` shim (``oracle/timm_shim``: the only two names
``models/layers.py:9`` needs) so that ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs time the reference's
own CPU forward (``kind: "reference"``) instead of the oracle port.  The product never imports anything from here.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("models/layers.py", "models/pangu_model.py")       # the files the hot path lives in


def stage(reference: str | None = None) -> str | None:
    """Copy the hot-path modules of ``reference`` into oracle/_ref; returns the destination or None if absent."""
    ref = reference or os.environ.get("PANGU_REFERENCE", "/root/reference")
    if not all(os.path.isfile(os.path.join(ref, f)) for f in FILES):
        return None
    manifest = {}
    for f in FILES:
        dst = os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(ref, f), dst)
        with open(dst, "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    open(os.path.join(DEST, "models", "__init__.py"), "w").close()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref, "sha256": manifest, "note": "unmodified copies; see oracle/build_ref.py"}, fh, indent=1)
    return DEST


if __name__ == "__main__":
    print(stage() or "reference checkout not found: nothing staged")
