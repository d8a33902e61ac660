#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction stall samples with nvdisasm line info and print the
hottest CUDA source lines.   usage: ncu_lines.py report.ncu-rep lib.so kernel_substring [launch_idx [mangled_substring]]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
sect = sys.argv[5] if len(sys.argv) > 5 else kern      # substring of the mangled name (.text section) when it differs from ncu's kernel name
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# offset -> (file, line) for the kernel
off2line, cur, inside = {}, None, False
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text."):
        inside = sect in ln
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name"')
want = int(sys.argv[4]) if len(sys.argv) > 4 else 0
blk = [b for b in blocks if kern in b.split("\n")[0]][want]
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
hdr = rows[1]
ia, isamp = hdr.index("Address"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "(" not in h]
base = None
per_line = collections.Counter()
per_line_stall = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    addr = int(r[ia], 16)
    base = addr if base is None else base
    s = int(r[isamp] or 0)
    key = off2line.get(addr - base, ((None, 0), "?"))[0]
    per_line[key] += s
    for i, h in stall_cols:
        if r[i] not in ("", "0"):
            per_line_stall[key][h] += int(r[i])
tot = sum(per_line.values())
print("kernel", kern, "total samples", tot)
src_cache = {}
for key, s in per_line.most_common(40):
    f, l = key if key else (None, 0)
    text = ""
    if f:
        path = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
        if os.path.exists(path):
            src_cache.setdefault(path, open(path).read().splitlines())
            text = src_cache[path][l - 1].strip()[:90] if l - 1 < len(src_cache[path]) else ""
    top = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in per_line_stall[key].most_common(3))
    print("%5.1f%%  %s:%s  %s   [%s]" % (100.0 * s / max(tot, 1), f, l, text, top))
