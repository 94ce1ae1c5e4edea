#!/usr/bin/env python
"""Static SASS instruction count of a kernel's hot loop, by issue pipe (C4's INT32-issue fraction).

    python tools/sass_loop_count.py            # match_table_kernel<4> (k = 12) -> profiles/r02_match_table_sass_counts.json

The hot loop is taken to be the widest backward branch of the function that holds no CTA barrier (for match_table_kernel: one ROUND of a warp = 32 k-mers
against the 1024 constants of a group = 32 768 pair tests).  Pipes as measured in B300_MICROARCH.md ("Pipe rates"): LOP3 / SHF /
IADD3 / ISETP / LEA / PRMT / POPC... issue on the ALU pipe at one warp instruction per 2 cycles per SM sub-partition, IMAD on the
FMA pipe at the same rate, every instruction takes one issue slot (1 per cycle per sub-partition).  bench.py turns these counts,
the trip count of the workload and the measured time into `secondary.match_c4.int32`."""
import collections
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "kmer-extension_b200" / "libkmer_cuda.so"
ALU = {"LOP3", "SHF", "IADD3", "IADD", "ISETP", "LEA", "PRMT", "SEL", "VIADD", "IABS", "IMNMX", "VIMNMX", "BMSK", "SGXT", "PLOP3", "P2R", "R2P", "MOV"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2"}
LSU = {"LDS", "STS", "ATOMS", "LDG", "STG", "ATOMG", "RED", "LDC", "SHFL", "LDSM"}


def loop_counts(mangled: str):
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, str(LIB)], capture_output=True, text=True, check=True).stdout
    ins = []
    for ln in sass.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*)", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    bars = [addr for addr, op, _ in ins if op.startswith("BAR")]
    best = None
    for addr, op, rest in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if not t or int(t.group(1), 16) >= addr:
                continue
            lo_ = int(t.group(1), 16)
            if any(lo_ <= b <= addr for b in bars):       # a loop with a CTA barrier inside is not one warp's inner loop
                continue
            if best is None or addr - lo_ > best[1] - best[0]:
                best = (lo_, addr)
    lo, hi = best
    ops = collections.Counter(op.split(".")[0] for addr, op, _ in ins if lo <= addr <= hi)
    pipe = collections.Counter()
    for o, n in ops.items():
        pipe["alu" if o in ALU else "fma" if o in FMA else "lsu" if o in LSU else "other"] += n
    return {"function": mangled, "loop": [hex(lo), hex(hi)], "instructions": sum(ops.values()), "by_pipe": dict(pipe), "by_opcode": dict(ops.most_common()),
            "function_instructions": len(ins)}


def main():
    fn = "_ZN4kmer18match_table_kernelILi4EEEviPKmmiPKNS_10MatchConstEPKijPjmPy"
    r = loop_counts(fn)
    r["what"] = ("match_table_kernel<4> (k = 12: 4 chunks of 3 bases): one round of a warp = 32 k-mers x 1024 constants = 32768 pair tests; "
                 "static count of the round loop body from cuobjdump -sass of the shipped libkmer_cuda.so (predicated-off atomics included)")
    r["pair_tests_per_loop_trip"] = 32 * 1024
    out = ROOT / "profiles" / "r02_match_table_sass_counts.json"
    out.write_text(json.dumps(r, indent=1) + "\n")
    print(json.dumps({k: r[k] for k in ("loop", "instructions", "by_pipe")}))


if __name__ == "__main__":
    sys.exit(main())
