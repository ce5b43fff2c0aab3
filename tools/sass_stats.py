#!/usr/bin/env python
"""Static look at a kernel's hot loop (no GPU needed): instruction mix of the SASS between two addresses and the
single-warp issue-model time of one trip through it (stall counts from the control bits, fixed latencies of the
ALU/FMA pipes; /opt/skills/guides/B300_MICROARCH.md "Single-warp issue model").

  tools/sass_stats.py <lib.so> <mangled kernel name> [--loops]            list backward branches (loops) with sizes
  tools/sass_stats.py <lib.so> <mangled kernel name> <lo> <hi>            mix + model for [lo, hi] (hex addresses)
"""
import collections
import re
import subprocess
import sys

ALU = ("VIADDMNMX", "VIMNMX", "VIMNMX3", "IADD3", "IADD", "LOP3", "SHF", "PRMT", "ISETP", "SEL", "R2P", "P2R", "MOV", "IMNMX",
       "LEA", "PLOP3", "IABS", "VIADD", "FMNMX", "SGXT", "BMSK", "FLO", "POPC", "UMOV", "CS2R", "S2R")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2")


def dump(lib, fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
    ins = []
    lines = out.splitlines()
    i = 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
            hi = int(m2.group(1), 16) if m2 else 0
            ins.append((int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), hi))
            i += 2
        else:
            i += 1
    return ins


def opname(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]


def main():
    lib, fun = sys.argv[1], sys.argv[2]
    ins = dump(lib, fun)
    if not ins:
        raise SystemExit("kernel not found")
    if len(sys.argv) < 5:
        for a, t, lo, hi in ins:
            if "BRA" in t and "BRA.DIV" not in t:
                m = re.search(r"0x([0-9a-f]+)", t)
                if m and int(m.group(1), 16) < a:
                    b = int(m.group(1), 16)
                    n = (a - b) // 16 + 1
                    if n > 1300:
                        continue
                    body = [x for x in ins if b <= x[0] <= a]
                    mix = collections.Counter(opname(tt) for _, tt, _, _ in body)
                    alu = sum(c for o, c in mix.items() if o in ALU)
                    fma = sum(c for o, c in mix.items() if o in FMA)
                    st = sum(max((h >> 41) & 0xf, 1) for _, _, _, h in body)
                    dpx = mix["VIADDMNMX"] + mix["VIMNMX3"] + mix["VIMNMX"]
                    print(f"loop {b:#x}..{a:#x} {n:5d} instr  alu {alu:4d} fma {fma:4d} dpx {dpx:4d} sel {mix['SEL']:3d} bra {mix['BRA']:2d} stallsum {st:5d}")
        print("total", len(ins))
        return
    lo_a, hi_a = int(sys.argv[3], 16), int(sys.argv[4], 16)
    body = [x for x in ins if lo_a <= x[0] <= hi_a]
    mix = collections.Counter(opname(t) for _, t, _, _ in body)
    n = len(body)
    alu = sum(c for o, c in mix.items() if o in ALU)
    fma = sum(c for o, c in mix.items() if o in FMA)
    print(f"{n} instructions: ALU-pipe {alu}, FMA-pipe {fma}, other {n - alu - fma}")
    for o, c in mix.most_common():
        print(f"  {o:12s} {c}")
    # issue model: stall field = bits 105..108 of the 128-bit word -> bits 41..44 of the high half
    t = 0
    for _, _, _, hi in body:
        stall = (hi >> 41) & 0xf
        t += max(stall, 1)
    print(f"sum of stall counts (lone warp, no scoreboard waits): {t} cycles; ALU pipe busy {2 * alu} cycles, FMA pipe {2 * fma}")


if __name__ == "__main__":
    main()
