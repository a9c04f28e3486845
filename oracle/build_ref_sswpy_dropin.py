#!/usr/bin/env python
"""oracle/build_ref_sswpy_dropin.py — the link-level drop-in of INTEGRATION.md §1, built for the tests (test infrastructure).

    python oracle/build_ref_sswpy_dropin.py [--force]

The reference's UNMODIFIED `indelpost/sswpy.pyx` is cythonized in a scratch directory against `include/compat/ssw.h` and linked
against `indelpost_b200/libswb200.so` INSTEAD of the reference's `ssw.c` -- exactly the one-line setup.py change a maintainer
would make.  Only the build product (one extension module, `sswpy_dropin/sswpy.*.so`) is installed into oracle/_ref_sswpy/
(git-ignored; travels to the GPU box with the snapshot like oracle/_ref).  Its run path points at the in-tree libswb200.so.
tests/test_gpu_parity.py::test_link_level_dropin_of_reference_sswpy replays tests/golden/sswpy_api.json through it.
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get("REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref_sswpy")

SETUP = '''
from setuptools import setup, Extension
from Cython.Build import cythonize
ext = Extension("sswpy", ["sswpy.pyx"],                                  # the reference lists ["indelpost/sswpy.pyx", "indelpost/ssw.c"]
                include_dirs=[{inc!r}], libraries=["swb200"], library_dirs=[{lib!r}],
                runtime_library_dirs=["$ORIGIN/../../../indelpost_b200"], extra_compile_args=["-Wno-unused-function", "-w"])
setup(ext_modules=cythonize([ext], language_level=3, quiet=True))
'''


def stamp():
    h = hashlib.sha256()
    for p in (os.path.join(REFERENCE, "indelpost", "sswpy.pyx"), os.path.join(ROOT, "include", "swb200.h"), os.path.join(ROOT, "include", "compat", "ssw.h"), __file__):
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(sys.version.encode())
    return h.hexdigest()


def build(force=False):
    src = os.path.join(REFERENCE, "indelpost", "sswpy.pyx")
    if not os.path.isfile(src):
        print("reference tree absent: keeping prebuilt oracle/_ref_sswpy (if any)")
        return os.path.isdir(OUT)
    if not os.path.exists(os.path.join(ROOT, "indelpost_b200", "libswb200.so")):
        raise SystemExit("build indelpost_b200/libswb200.so first (make -C indelpost_b200/csrc)")
    st = stamp()
    sf = os.path.join(OUT, ".stamp")
    if not force and os.path.exists(sf) and open(sf).read().strip() == st and glob.glob(os.path.join(OUT, "sswpy_dropin", "sswpy*.so")):
        return True
    tmp = tempfile.mkdtemp(prefix="sswpy_dropin_")
    try:
        shutil.copy2(src, os.path.join(tmp, "sswpy.pyx"))
        with open(os.path.join(tmp, "setup.py"), "w") as fh:
            fh.write(SETUP.format(inc=os.path.join(ROOT, "include", "compat"), lib=os.path.join(ROOT, "indelpost_b200")))
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-3000:] + r.stderr[-6000:])
            raise SystemExit("drop-in build of the reference sswpy.pyx failed")
        if os.path.isdir(OUT):
            shutil.rmtree(OUT)
        os.makedirs(os.path.join(OUT, "sswpy_dropin"))
        for p in glob.glob(os.path.join(tmp, "sswpy*.so")):
            shutil.copy2(p, os.path.join(OUT, "sswpy_dropin", os.path.basename(p)))
        with open(sf, "w") as fh:
            fh.write(st + "\n")
        print("built", OUT)
        return True
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    build(force="--force" in sys.argv)
