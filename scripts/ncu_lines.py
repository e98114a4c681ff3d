#!/usr/bin/env python
"""Development aid: joins the SASS-level source page of an ncu report (`ncu -i X.ncu-rep --page source --csv`) with the
line table of the library (`nvdisasm -g -c` of the cubin inside coexpression_v2.so) and prints executed instructions and
stall samples per CUDA source line of one kernel.
    python scripts/ncu_lines.py gpurun_out/X.ncu-rep count_dedup_kernelILi0 [top] [regex of the demangled kernel name, for
    reports that hold several kernels]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("COEX_SO") or os.path.join(root, "coexpression_b200", "csrc", "coexpression_v2.so")   # (COEX_SO: the library of the commit the report was taken with)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
line_of = {}
inside, cur = False, ("?", 0)
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
flt = ["--kernel-name", "regex:" + sys.argv[4]] if len(sys.argv) > 4 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + flt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ia, ii, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
base = None
agg = {}
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) <= ii or not r[ii].isdigit():
        continue
    addr = int(r[ia], 16) if r[ia].startswith("0x") or re.match(r"^[0-9a-f]+$", r[ia]) else int(r[ia])
    if base is None:
        base = addr
    key = line_of.get(addr - base, ("?", -1))
    a = agg.setdefault(key, [0, 0])
    a[0] += int(r[ii]); a[1] += int(r[isamp] or 0)
    tot_i += int(r[ii]); tot_s += int(r[isamp] or 0)
src = {}
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
for key, (ni, ns) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, l = key
    if f not in src and f != "?":
        try:
            src[f] = open(os.path.join(os.path.dirname(so), f)).read().split("\n")     # the sources next to the library
        except OSError:
            src[f] = []
    text = src.get(f, [])[l - 1].strip()[:100] if f in src and 0 < l <= len(src[f]) else ""
    print(f"{100 * ni / tot_i:5.1f}% inst {100 * ns / max(tot_s, 1):5.1f}% stall  {f}:{l}  {text}")
