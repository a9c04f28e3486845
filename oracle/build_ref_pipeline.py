#!/usr/bin/env python
"""oracle/build_ref_pipeline.py — build the UNMODIFIED reference pipeline into oracle/_ref_pipeline/ (test infrastructure).

    python oracle/build_ref_pipeline.py [--force]

The whole reference package (`/root/reference/indelpost`: ten Cython modules + ssw.c + the pure-Python modules) is copied
to a scratch directory under /tmp and built there, unmodified, with the reference's own compiler flags
(setup.py:37: `-Wno-unused-function`), against the stub `pysam` of oracle/pysam_stub (real pysam / htslib are absent and
cannot be installed offline).  Only build PRODUCTS are installed into oracle/_ref_pipeline/ (compiled extension modules,
plus the package's pure-Python modules exactly like a `pip install --target` would place them); the directory is
git-ignored and travels to the GPU box with the snapshot, like oracle/_ref.  Nothing from the reference is committed.

Consumers: tests/refpipe.py (pipeline-level parity tests), bench.py's `pipeline` extra (loci/s, reference SSW vs GPU SSW)
and tests/golden/make_pipeline_golden.py.  No product module imports it.
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref_pipeline")
STUB = os.path.join(HERE, "pysam_stub")
SHIM = os.path.join(HERE, "ref_pileup_shim.pyx")   # our own four-line accessor of the cdef make_pileup

SETUP = '''
from setuptools import setup, Extension
from Cython.Build import cythonize
import glob, os
exts = []
for p in sorted(glob.glob("pysam/*.pyx")):
    exts.append(Extension("pysam." + os.path.basename(p)[:-4], [p]))
for p in sorted(glob.glob("indelpost/*.pyx")):
    name = os.path.basename(p)[:-4]
    src = [p] + (["indelpost/ssw.c"] if name == "sswpy" else [])
    exts.append(Extension("indelpost." + name, src, include_dirs=["."], extra_compile_args=["-Wno-unused-function", "-w"]))
exts.append(Extension("refshim", ["refshim.pyx"], include_dirs=["."], extra_compile_args=["-w"]))
setup(ext_modules=cythonize(exts, language_level=3, include_path=["."], quiet=True))
'''


def source_stamp():
    h = hashlib.sha256()
    for root in (os.path.join(REFERENCE, "indelpost"), STUB):
        for p in sorted(glob.glob(os.path.join(root, "**", "*"), recursive=True)):
            if os.path.isfile(p) and not p.endswith((".pyc", ".so")):
                h.update(p.encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    with open(SHIM, "rb") as fh:
        h.update(fh.read())
    h.update(sys.version.encode())
    return h.hexdigest()


def build(force=False):
    if not os.path.isdir(os.path.join(REFERENCE, "indelpost")):
        print("reference tree absent: keeping prebuilt oracle/_ref_pipeline (if any)")
        return os.path.isdir(OUT)
    stamp = source_stamp()
    stamp_file = os.path.join(OUT, ".stamp")
    if not force and os.path.exists(stamp_file) and open(stamp_file).read().strip() == stamp:
        return True
    tmp = tempfile.mkdtemp(prefix="indelpost_ref_")
    try:
        shutil.copytree(os.path.join(REFERENCE, "indelpost"), os.path.join(tmp, "indelpost"))
        shutil.copytree(os.path.join(STUB, "pysam"), os.path.join(tmp, "pysam"))
        shutil.copy2(SHIM, os.path.join(tmp, "refshim.pyx"))
        with open(os.path.join(tmp, "setup.py"), "w") as fh:
            fh.write(SETUP)
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-3000:] + r.stderr[-6000:])
            raise SystemExit("reference pipeline build failed")
        if os.path.isdir(OUT):
            shutil.rmtree(OUT)
        for pkg in ("indelpost", "pysam"):
            os.makedirs(os.path.join(OUT, pkg))
            for p in glob.glob(os.path.join(tmp, pkg, "*")):
                # build products + what an installed package needs at run time; no C / Cython sources
                if p.endswith(".so") or p.endswith(".py"):
                    shutil.copy2(p, os.path.join(OUT, pkg, os.path.basename(p)))
        for p in glob.glob(os.path.join(tmp, "refshim*.so")):
            shutil.copy2(p, os.path.join(OUT, os.path.basename(p)))
        with open(stamp_file, "w") as fh:
            fh.write(stamp + "\n")
        print("built", OUT)
        return True
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    build(force="--force" in sys.argv)
