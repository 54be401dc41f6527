"""Build ``oracle/_ref``: the reference's own environment, byte-compiled, so that it can be TIMED on
the GPU box's host cores beside the CUDA path (bench.py ``--impl reference`` / ``cpu_baseline``,
``kind: "reference"``).  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            (runs in the build container; __graft_entry__.build() calls it)

``/root/reference`` does not exist on the GPU box and the reference is pure Python, so the recipe
compiles every module of ``/root/reference/backend`` where it lies (``py_compile``) and packs ONLY
the resulting ``.pyc`` files -- no source -- plus the reference's ``config.yaml`` (its ``Config`` reads
``<parent of backend>/config.yaml``, backend/config.py:8-9) into ONE binary,
``oracle/_ref/reference_backend.bin`` (a zip archive; loose ``.pyc`` files do not survive the
snapshot to the GPU box).  ``ref_loader`` unpacks it into a scratch directory at run time, where the
sourceless packages import like the originals.  ``oracle/_ref`` is listed in ``.gitignore`` (it never
enters history) and not in ``.gpurunignore`` (it travels with the snapshot, like the built ``.so``
files).  The image on the GPU box has the same interpreter (3.12), which is what a ``.pyc`` requires;
``ref_loader`` refuses an archive built by another version.
"""
from __future__ import annotations

import importlib.util
import io
import json
import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


ARCHIVE = os.path.join(REF_DIR, "reference_backend.bin")


def build(src_root: str = "/root/reference", quiet: bool = False) -> str | None:
    backend = os.path.join(src_root, "backend")
    if not os.path.isfile(os.path.join(backend, "envs", "lander.py")):
        if not quiet:
            print(f"make_ref: no reference checkout at {src_root}; keeping whatever {REF_DIR} holds")
        return ARCHIVE if os.path.isfile(ARCHIVE) else None
    shutil.rmtree(REF_DIR, ignore_errors=True)
    os.makedirs(REF_DIR)
    n = 0
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_DEFLATED) as z, tempfile.TemporaryDirectory() as tmp:
        for dirpath, dirnames, files in os.walk(backend):
            dirnames[:] = [d for d in dirnames if d != "__pycache__"]
            rel = os.path.relpath(dirpath, src_root)
            for f in sorted(files):
                if f.endswith(".py"):
                    out = os.path.join(tmp, "m.pyc")
                    py_compile.compile(os.path.join(dirpath, f), cfile=out,
                                       dfile=os.path.join("reference", rel, f), doraise=True)
                    z.write(out, os.path.join(rel, f + "c"))
                    n += 1
        z.write(os.path.join(src_root, "config.yaml"), "config.yaml")
        z.writestr("BUILD.json", json.dumps({"python": list(sys.version_info[:3]),
                                             "magic": importlib.util.MAGIC_NUMBER.hex(), "modules": n}))
    with open(ARCHIVE, "wb") as fh:
        fh.write(buf.getvalue())
    if not quiet:
        print(f"make_ref: {n} reference modules byte-compiled into {ARCHIVE} ({os.path.getsize(ARCHIVE)} bytes)")
    return ARCHIVE


if __name__ == "__main__":
    build(*(sys.argv[1:2]))
