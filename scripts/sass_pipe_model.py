#!/usr/bin/env python
"""Static pipe model of a kernel from its SASS (no GPU needed): per innermost loop body, the instruction mix and the cycles
each integer pipe needs per warp, using the issue costs measured on B200 by scripts/ubench/pipes.cu (profiles/r1_ubench.txt):
IMAD 2.1, IMAD.WIDE 4.6, IMAD.HI 5.1 cycles per warp instruction on the multiplier pipe ("fmaheavy"); 2.0 on the ALU pipe for
IADD3 / VIADD / VIADDMNMX / LOP3 / SHF / ISETP / SEL / LEA. IMAD.IADD / IMAD.MOV / IMAD.X / IMAD.SHL run on the multiplier pipe too.

  python scripts/sass_pipe_model.py plonky3-recursion_b200/libp3r_b200.so k_hash_rows KoalaBear

Prints one line per loop (backward branch) of every matching kernel: instructions, multiplier-pipe cycles, ALU cycles,
memory / shuffle / barrier counts. The larger of the two pipe numbers bounds the loop; compare with the issue slots (= the
instruction count) to see which one limits."""
import re
import subprocess
import sys
from collections import Counter

HEAVY = {"IMAD": 2.1, "IMAD.U32": 2.1, "IMAD.IADD": 2.1, "IMAD.MOV": 2.1, "IMAD.MOV.U32": 2.1, "IMAD.X": 2.1, "IMAD.SHL": 2.1,
         "IMAD.SHL.U32": 2.1, "IMAD.WIDE": 4.6, "IMAD.WIDE.U32": 4.6, "IMAD.HI": 5.1, "IMAD.HI.U32": 5.1}
ALU_PREFIX = ("IADD3", "VIADD", "VIMNMX", "LOP3", "SHF", "ISETP", "SEL", "LEA", "PRMT", "IABS", "MOV", "PLOP3", "FLO", "POPC")
MEM_PREFIX = ("LDG", "STG", "LDS", "STS", "LD.", "ST.", "LDC", "LDCU", "ATOM", "RED", "LDSM")


def kernels(lib, patterns):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            ins = re.sub(r"^@!?U?P[0-9T]+\s+", "", m.group(2))
            out[cur].append((int(m.group(1), 16), ins))
    return {k: v for k, v in out.items() if all(p in k for p in patterns)}


def model(ins):
    c = Counter(t.split()[0] for _, t in ins)
    heavy = sum(HEAVY.get(op, 0.0) * n for op, n in c.items())
    alu = sum(2.0 * n for op, n in c.items() if op.startswith(ALU_PREFIX) and op not in HEAVY)
    mem = sum(n for op, n in c.items() if op.startswith(MEM_PREFIX))
    shfl = sum(n for op, n in c.items() if op.startswith("SHFL"))
    bar = sum(n for op, n in c.items() if op.startswith(("BAR", "WARPSYNC", "BSYNC")))
    return len(ins), heavy, alu, mem, shfl, bar, c


def main():
    lib, patterns = sys.argv[1], sys.argv[2:]
    for name, ins in kernels(lib, patterns).items():
        n, heavy, alu, mem, shfl, bar, c = model(ins)
        print(f"{name}\n  whole kernel: {n} instr, multiplier pipe {heavy:.0f} cyc, ALU {alu:.0f} cyc, mem {mem}, shfl {shfl}, bar {bar}")
        loops = []
        for a, t in ins:
            if t.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", t)
                if m and int(m.group(1), 16) < a:
                    loops.append((int(m.group(1), 16), a))
        inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
        for lo, hi in inner:
            body = [(a, t) for a, t in ins if lo <= a <= hi]
            n, heavy, alu, mem, shfl, bar, c = model(body)
            top = ", ".join(f"{op} {k}" for op, k in c.most_common(6))
            print(f"  loop {lo:#06x}-{hi:#06x}: {n:5d} instr | multiplier {heavy:7.0f} | ALU {alu:7.0f} | mem {mem:3d} shfl {shfl:3d} bar {bar:2d} | {top}")


if __name__ == "__main__":
    main()
