#!/usr/bin/env python
"""Print the SASS of one kernel for a source line range. usage: sass_lines.py <symbol> <file> <lo> <hi>"""
import os, re, subprocess, sys, tempfile
sym, f, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath("superterrainplus_b200/libshf_b200.so")], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, x) for x in os.listdir(tmp) if x.endswith(".cubin")][0]
text = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
on = False; line = 0; fname = ""; out = []
for ln in text.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        on = sym in ln; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if "inlined at" not in ln:
            line = int(m.group(2)); fname = m.group(1).split('/')[-1]
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: out.append((fname, line, m.group(1), m.group(2).strip()))
    elif re.match(r"\s*\.L_", ln): out.append(("", 0, "", ln.strip()))
sel = [i for i, o in enumerate(out) if o[0] == f and lo <= o[1] <= hi]
if sel:
    for o in out[min(sel):max(sel) + 1]: print(o[1], o[2], o[3])
print("total instructions in kernel:", sum(1 for o in out if o[2]))
