"""Where does the UNMODIFIED reference spend a locus outside Smith-Waterman?  (analysis tool)

cProfile cannot see into the reference's compiled Cython and the image has neither perf nor py-spy, so this samples native
stacks with SIGPROF (tools/research/native_sampler.c) while the reference pipeline of oracle/_ref_pipeline runs synthetic cfg3
loci, and symbolises the addresses with dladdr + `nm`.  Reported per function: self samples (innermost frame) and inclusive
samples of the innermost frame that belongs to an indelpost module (so libc / CPython time is charged to the Cython function
that caused it).

    python tools/research/profile_reference_locus.py [--loci 6] [--config cfg3]
"""
from __future__ import annotations

import argparse
import bisect
import collections
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def build_sampler():
    so = "/tmp/native_sampler.so"
    src = os.path.join(ROOT, "tools", "research", "native_sampler.c")
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.sampler_frame.restype = C.c_void_p
    lib.sampler_frame.argtypes = [C.c_int, C.c_int]
    lib.sampler_module.restype = C.c_char_p
    lib.sampler_module.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    return lib


class Symbols:
    def __init__(self):
        self.tables = {}

    def table(self, path):
        t = self.tables.get(path)
        if t is None:
            addrs, names = [], []
            try:
                out = subprocess.run(["nm", "-n", "--defined-only", path], capture_output=True, text=True).stdout
                if not out.strip():
                    out = subprocess.run(["nm", "-n", "-D", "--defined-only", path], capture_output=True, text=True).stdout
                for line in out.splitlines():
                    p = line.split()
                    if len(p) >= 3 and p[1] in "tTwW":
                        addrs.append(int(p[0], 16)); names.append(p[2])
            except Exception:  # noqa: BLE001
                pass
            t = self.tables[path] = (addrs, names)
        return t

    def name(self, path, offset):
        addrs, names = self.table(path)
        i = bisect.bisect_right(addrs, offset) - 1
        return names[i] if i >= 0 else "?"


def pretty(sym):
    # __pyx_f_9indelpost_6pileup_make_pileup / __pyx_pw_9indelpost_6varaln_16VariantAlignment_1__cinit__ ...
    for pre in ("__pyx_f_9indelpost_", "__pyx_pf_9indelpost_", "__pyx_pw_9indelpost_", "__pyx_gb_9indelpost_", "__pyx_fuse_"):
        if sym.startswith(pre):
            return sym[len(pre):]
    return sym


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=6)
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--interval-us", type=int, default=500)
    a = ap.parse_args()
    import bench_pipeline as BP
    import loci
    import refpipe

    refpipe.load()
    lcs = [loci.make_locus(**sp) for sp in BP.make_specs(a.config, a.loci)]
    refpipe.run_locus(lcs[0])
    lib = build_sampler()
    lib.sampler_start(a.interval_us)
    for lc in lcs:
        refpipe.run_locus(lc)
    n = lib.sampler_stop()
    syms = Symbols()
    self_c, incl_c, mod_c = collections.Counter(), collections.Counter(), collections.Counter()
    cache = {}

    def resolve(addr):
        r = cache.get(addr)
        if r is None:
            base = C.c_size_t(0)
            m = lib.sampler_module(addr, C.byref(base))
            if not m:
                r = ("?", "?")
            else:
                path = m.decode()
                r = (os.path.basename(path), pretty(syms.name(path, addr - base.value)))
            cache[addr] = r
        return r

    for k in range(n):
        d = lib.sampler_depth(k)
        frames = [resolve(lib.sampler_frame(k, j)) for j in range(2, d)]       # skip the handler and the signal trampoline
        if not frames:
            continue
        self_c[frames[0]] += 1
        mod_c[frames[0][0]] += 1
        for f in frames:
            if f[0].startswith(("pileup.", "varaln.", "localn.", "gappedaln.", "softclip.", "utilities.", "variant.", "contig.", "sswpy.", "local_reference.")):
                incl_c[f] += 1
                break
        else:
            incl_c[("(no indelpost frame)", frames[0][1])] += 1
    print(f"{n} samples of {a.interval_us} us CPU time over {a.loci} {a.config} loci ({n * a.interval_us / 1e3 / a.loci:.1f} ms per locus)")
    print("\n-- by module of the innermost frame")
    for (m, c) in mod_c.most_common(12):
        print(f"{100 * c / n:6.1f} %  {m}")
    print("\n-- inclusive, by the innermost indelpost function on the stack")
    for (f, c) in incl_c.most_common(28):
        print(f"{100 * c / n:6.1f} %  {f[0]:28s} {f[1]}")
    print("\n-- self time, innermost frame")
    for (f, c) in self_c.most_common(22):
        print(f"{100 * c / n:6.1f} %  {f[0]:28s} {f[1]}")


if __name__ == "__main__":
    main()
