"""Attribute ncu per-SASS-instruction counters to CUDA source lines (tools, not product).

usage: ncu_lines.py <report.ncu-rep> <kernel-regex> <lib.so> [top]
Needs the library compiled with -lineinfo.  Joins `ncu --page source --csv` (SASS view) with
`nvdisasm -g` line annotations by instruction offset.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kre, so = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# first kernel instance only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
kname = rows[hdr_i - 1][1]
inst = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    inst.append(dict(zip(hdr, r)))
base = int(inst[0]["Address"], 16)
print("kernel:", kname[:120], "instructions:", len(inst))

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", kname.split("(")[0].split("::")[-1].split("<")[0])
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    secs = re.split(r"\n//-+ \.text\.", sass)
    for sec in secs[1:]:
        name = sec.split(" ", 1)[0]
        if mangled_hint not in name:
            continue
        # template args must match: compare instruction count
        cur, table, n = None, {}, 0
        for ln in sec.splitlines():
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
            if m:
                table[int(m.group(1), 16)] = cur
                n += 1
        if n == len(inst):
            line_of = table
            break
    if line_of:
        break
if not line_of:
    sys.exit("could not match the kernel in the library (instruction counts differ)")

agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for d in inst:
    off = int(d["Address"], 16) - base
    key = line_of.get(off)
    ni = int(d["Instructions Executed"]); ns = int(d["# Samples"])
    agg[key][0] += ni; agg[key][1] += ns
    tot_i += ni; tot_s += ns
src_cache = {}
def src(key):
    if key is None:
        return ""
    fn, ln = key
    for root in ("astrild_b200/csrc", "."):
        p = os.path.join(root, fn)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][ln - 1].strip()[:100]
    return ""
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for key, (ni, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*ni/tot_i:5.1f}% inst {100*ns/max(tot_s,1):5.1f}% smp  {key[0] if key else '?'}:{key[1] if key else 0:<5d} {src(key)}")
