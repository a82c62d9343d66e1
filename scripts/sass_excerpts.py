#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel instruction-mix histograms of the hot kernels in seqlib_b200/libseqlib_b200.so and
the listing of the wavefront step loop (between two consecutive SHFL.UP).  usage: python scripts/sass_excerpts.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "seqlib_b200", "libseqlib_b200.so")
OUT = os.path.join(ROOT, "profiles")
KERNELS = {
    "seed2": r"k_seed2ILi16ELi6ELi1",
    "seed3": r"k_seed3",
    "chain_build": r"k_chain_build",
    "extend_wave": r"k_extend_waveILi(4|8)",
    "ext_wave_batch": r"k_ext_waveILi4",
    "extend_group": r"k_extend_groupILi8ELb1",
    "finalize_dp": r"k_finalize_dpILi8",
    "k_ec": r"k_ec",
    "seedtab": r"k_seedtab_level",
}


def main():
    names = subprocess.run(["cuobjdump", "-elf", SO], capture_output=True, text=True).stdout
    syms = sorted(set(re.findall(r"\.text\.(_Z\w+)", names)))
    for tag, pat in KERNELS.items():
        for sym in [s for s in syms if re.search(pat, s)][:2]:
            sass = subprocess.run(["cuobjdump", "-sass", "-fun", sym, SO], capture_output=True, text=True).stdout
            ins = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l)]
            ops = collections.Counter(re.sub(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?", "", l).split()[0].rstrip(";") for l in ins)
            path = os.path.join(OUT, "r02_sass_%s.txt" % tag)
            with open(path, "a" if os.path.exists(path) and sym != [s for s in syms if re.search(pat, s)][0] else "w") as f:
                f.write("== %s: %d SASS instructions\n" % (sym, len(ins)))
                for op, n in ops.most_common(40):
                    f.write("  %-28s %6d\n" % (op, n))
                if "wave" in tag:
                    idx = [i for i, l in enumerate(ins) if "SHFL.UP" in l]
                    if len(idx) >= 2:
                        f.write("\n-- one wavefront step (between two SHFL.UP): %d instructions for 2 cells\n" % (idx[1] - idx[0]))
                        for l in ins[idx[0]:idx[1]]:
                            f.write(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) + "\n")
                if tag == "seed2":
                    f.write("\n-- 256-bit gathers (Occ blocks / table sectors):\n")
                    for l in ins:
                        if "LDG.E" in l and "256" in l:
                            f.write(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) + "\n")
            print(path)


if __name__ == "__main__":
    main()
