"""Per-kernel SASS evidence of the Blackwell-native paths: counts of the mnemonics that prove the fp64 tensor pipe (DMMA),
TMA (UTMALDG / UTMASTG), bulk copies (UBLKCP), mbarriers (SYNCS), cluster / DSMEM traffic (UCGABAR, MAPA, "shared::cluster"
lower to ST/LD with .CLUSTER? -> we count any opcode containing CGA) and plain fp64 FMA (DFMA) in the shipped libbbgpu.so.
Usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bboptpy_b200", "libbbgpu.so")
WATCH = ["DMMA", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UCGABAR", "MAPA", "DFMA", "DADD", "DMUL", "MUFU", "REDG", "ATOMG", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1).split(".")[0]
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or (w == "UCGABAR" and op.startswith("UCGABAR")):
                    kernels[cur][w] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): instruction counts per kernel")
    print(f"# {'kernel':90s} {'total':>7s} " + " ".join(f"{w:>7s}" for w in WATCH))
    for (name, c), dm in zip(kernels.items(), demangled):
        short = re.sub(r"\(.*", "", dm)
        short = short.replace("(anonymous namespace)::", "")
        print(f"{short[:92]:92s} {c['_total']:7d} " + " ".join(f"{c[w]:7d}" for w in WATCH))


if __name__ == "__main__":
    main()
