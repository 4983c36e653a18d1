"""Per-kernel SASS opcode counts of the in-tree library (profiles/sass_opcodes.txt).

    python scripts/sass_opcodes.py > profiles/sass_opcodes.txt

Evidence for the instruction selection DESIGN.md claims: DMMA (FP64 tensor pipe) in the Gram / Ritz-update / gradient kernels,
UBLKCP (1-D TMA bulk copies) + SYNCS (mbarrier) in the streaming Gram kernels, FFMA2 (packed fp32) in the FP32 SpMM and the
synthesis kernels, no UTC*MMA / LDTM (tcgen05 has no FP64 kind: nothing on this path is a low-precision GEMM).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "diffsound_b200", "libdiffsound_sm100.so")
OPS = ["DMMA", "DFMA", "FFMA2", "FFMA", "HMMA", "UBLKCP", "UBLKPF", "UTMA", "SYNCS", "LDG", "STG", "LDS", "STS", "SHFL", "ATOM", "ATOMS", "ATOMG", "RED",
       "UTCMMA", "UTCHMMA", "LDTM", "BAR", "MUFU", "CCTL", "ACQBULK", "UCGABAR"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    counts = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for o in OPS:
                if op == o or (o in ("UTMA", "UTCMMA", "UTCHMMA") and op.startswith(o)):
                    counts[cur][o] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)}  (cubin archs: {', '.join(arch)})")
    print("# one line per kernel: total SASS instructions, then the non-zero counts of the opcodes that carry a design claim")
    tot = collections.Counter()
    for (k, c), name in zip(counts.items(), demangle):
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("ds::", "")
        if "cub::" in name or "thrust::" in name or "CUB_" in name:
            name = "[CUB] " + name[:70]
        parts = " ".join(f"{o}={c[o]}" for o in OPS if c[o])
        print(f"{name[:64]:64s} n={c['_all']:6d}  {parts}")
        tot.update(c)
    print("# library totals: " + " ".join(f"{o}={tot[o]}" for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
