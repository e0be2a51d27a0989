"""Per-source-line instruction and stall-sample totals of one kernel in an .ncu-rep, joined to the line table of the
object file it was built from (ncu's CSV source page carries SASS only).
Usage: python tools/ncu_lines.py report.ncu-rep build/obj.o kernel_substring [top]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os


def main():
    rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if "Address" in r and "Source" in r)
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    sass = [r for r in rows if len(r) == len(hdr) and r[ia].startswith("0x")]
    base = int(sass[0][ia], 16)
    per_off = {int(r[ia], 16) - base: (int(r[ii]), int(r[isamp]), r[hdr.index("Source")].strip()) for r in sass}
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, stdout=subprocess.DEVNULL, check=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], stdout=subprocess.PIPE, text=True).stdout
    lines = dis.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
    cur = ("?", 0)
    agg = {}
    total_i = total_s = 0
    for l in lines[start + 1:]:
        if l.startswith("//-----") or l.startswith(".text."):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m:
            off = int(m.group(1), 16)
            if off in per_off:
                n, s, _ = per_off[off]
                a = agg.setdefault(cur, [0, 0])
                a[0] += n
                a[1] += s
                total_i += n
                total_s += s
    print(f"# {rep} :: {kern}: {total_i} warp instructions, {total_s} stall samples")
    src_cache = {}
    for (f, ln), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        text = ""
        for cand in (os.path.join("bboptpy_b200", "csrc", f),):
            if os.path.exists(cand):
                src_cache.setdefault(cand, open(cand).read().splitlines())
                if ln - 1 < len(src_cache[cand]):
                    text = src_cache[cand][ln - 1].strip()[:90]
        print(f"{100 * s / max(total_s, 1):5.1f}% smp {100 * n / max(total_i, 1):5.1f}% inst  {f}:{ln:<5d} {text}")


if __name__ == "__main__":
    main()
