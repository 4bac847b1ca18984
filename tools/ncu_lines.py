#!/usr/bin/env python
"""Attribute ncu per-instruction counters to CUDA source lines (needs -lineinfo).

    python tools/ncu_lines.py <report.ncu-rep> <kernel-mangled-substring> [top]

Joins `ncu --page source --csv` (SASS with "Instructions Executed") with `nvdisasm -g` of the
engine cubin by instruction offset, then prints the hottest source lines.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BYSTALL = "--by-stall" in sys.argv  # sort by stall samples instead of instructions
if BYSTALL:
    sys.argv.remove("--by-stall")
OUTER = "--outer" in sys.argv  # attribute inlined code to the kernel-level line that called it
if OUTER:
    sys.argv.remove("--outer")


SKIP = None  # --skip N: the N-th result of a report that holds several kernels
for _a in list(sys.argv):
    if _a.startswith("--skip="):
        SKIP = int(_a.split("=")[1])
        sys.argv.remove(_a)


def disasm_lines(func_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "rgpot_b200", "librgpot_b200.so")],
                   cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    # the engine is several translation units: disassemble every cubin, the function is in one of them
    out = ""
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        out += subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], check=True, capture_output=True,
                              text=True).stdout
    mapping = {}
    in_func, cur, block_open = False, None, False
    for line in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            in_func = func_sub in m.group(1)
            continue
        if not in_func:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            # a block of annotations: innermost first, outermost (kernel-level) last
            if OUTER or not block_open:
                cur = loc
            block_open = True
            continue
        block_open = False
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m:
            mapping[int(m.group(1), 16)] = (cur, m.group(2))
    return mapping


def main():
    rep, func = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    sel = ["--launch-skip", str(SKIP), "--launch-count", "1"] if SKIP is not None else []
    txt = subprocess.run(["ncu", "-i", rep, *sel, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    H = rows[hdr]
    ca, ci, ct = H.index("Address"), H.index("Instructions Executed"), H.index("Thread Instructions Executed")
    cs = H.index("Warp Stall Sampling (All Samples)")
    data = []
    for r in rows[hdr + 1:]:
        try:
            data.append((int(r[ca], 16), int(r[ci]), int(r[ct]), int(r[cs]), r[H.index("Source")]))
        except (ValueError, IndexError):
            pass
    base = data[0][0]
    mapping = disasm_lines(func)
    per = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for addr, n, t, s, src in data:
        loc = mapping.get(addr - base, (None, ""))[0] or ("?", 0)
        for k, v in enumerate((n, t, s)):
            per[loc][k] += v
            tot[k] += v
    print(f"total warp-instr {tot[0]:.3e}  thread-instr {tot[1]:.3e}  samples {tot[2]}")
    src_cache = {}
    for loc, (n, t, s) in sorted(per.items(), key=lambda kv: -kv[1][2 if BYSTALL else 0])[:top]:
        fname, ln = loc
        text = ""
        path = os.path.join(ROOT, "rgpot_b200", "csrc", fname)
        if os.path.exists(path):
            src_cache.setdefault(path, open(path).read().splitlines())
            if 0 < ln <= len(src_cache[path]):
                text = src_cache[path][ln - 1].strip()
        print(f"{100 * n / tot[0]:5.1f}% instr {100 * s / max(tot[2], 1):5.1f}% stall  lanes {t / max(n, 1):4.1f}  "
              f"{fname}:{ln}  {text[:90]}")


if __name__ == "__main__":
    main()
