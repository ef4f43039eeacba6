#!/usr/bin/env python
"""Static check of the DMMA schedule ptxas produced (no GPU needed).

    python tools/sass_dmma_chains.py <object-or-cubin> <kernel-name-substring> [--dump]

Walks the SASS of one kernel (cuobjdump -sass) and reports, per region between block barriers:
  * DMMA count, NOPs (ptxas pads a DMMA whose successor depends on it with a NOP: DMMA results have a fixed ~29-cycle
    latency, the stall field holds 15), LDS/STS/DFMA-class counts,
  * the histogram of dependency distances: how many DMMAs back the accumulator of each DMMA was last written
    (distance 1 = back-to-back dependent = the warp idles for the full DMMA latency; >= 2 covers it at 16 cycles/DMMA),
  * the histogram of LDS -> first-use distances in DMMAs issued in between (operand prefetch depth).
Round 2 found layer 1 of the fused kernel scheduled as serial chains (distance 1) this way: ncu's `wait` stalls on DMMA
and NOP were 28 % of phase A (profiles/r02_summary.md)."""
import collections
import re
import subprocess
import sys


def kernel_sass(path, name):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, on = [], False
    for line in txt.split("\n"):
        if "Function :" in line:
            on = name in line
            if on:
                out.append(("F", line.strip()))
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((m.group(1), m.group(2).strip()))
    return out


def regs(tok, width=2):
    m = re.match(r"R(\d+)", tok.strip())
    if not m:
        return []
    b = int(m.group(1))
    return list(range(b, b + width))


def analyse(ins, dump=False):
    regions, cur = [], []
    for a, s in ins:
        if a == "F":
            continue
        cur.append(s)
        if "BAR.SYNC" in s:
            regions.append(cur)
            cur = []
    regions.append(cur)
    for ri, reg in enumerate(regions):
        nd = sum("DMMA" in s for s in reg)
        if nd == 0:
            continue
        last, n = {}, 0
        dist = collections.Counter()
        pend = {}                      # register -> DMMA index at LDS time
        lds_use = collections.Counter()
        seq = []
        for s in reg:
            s2 = re.sub(r"^@!?U?P\d+\s+", "", s)
            op = s2.split()[0]
            if op.startswith("LDS"):
                m = re.match(r"LDS(\.\w+)*\s+(R\d+),", s2)
                w = 4 if ".128" in op else (2 if ".64" in op else 1)
                if m:
                    for r in regs(m.group(2), w):
                        pend[r] = n
            m = re.match(r"DMMA\.8x8x4\s+(R\d+),\s*(R\d+),\s*(R\d+),\s*(R\d+)", s2)
            if m:
                d, a_, b_, c_ = m.groups()
                dd = n - last[c_] if c_ in last else 99
                dist[min(dd, 9)] += 1
                seq.append(dd)
                for tok in (a_, b_):
                    for r in regs(tok):
                        if r in pend:
                            lds_use[min(n - pend.pop(r), 9)] += 1
                last[d] = n
                n += 1
        cnt = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0] for s in reg)
        print(f"region {ri}: {len(reg)} instr, DMMA {nd}, NOP {cnt['NOP']}, LDS {cnt['LDS']}, STS {cnt['STS']}, "
              f"DFMA/DMUL/DADD {cnt['DFMA'] + cnt['DMUL'] + cnt['DADD']}")
        print("   dependency distance (DMMAs since the accumulator was written; 9 = >= 9 / first use):",
              dict(sorted(dist.items())))
        print("   LDS -> first DMMA use, in DMMAs issued between:", dict(sorted(lds_use.items())))
        if dump:
            for k in range(0, len(seq), 32):
                print("     ", " ".join(str(min(x, 99)) for x in seq[k:k + 32]))


if __name__ == "__main__":
    ins = kernel_sass(sys.argv[1], sys.argv[2])
    if not ins:
        sys.exit(f"no kernel matching {sys.argv[2]!r}")
    print(ins[0][1])
    analyse(ins, "--dump" in sys.argv)
