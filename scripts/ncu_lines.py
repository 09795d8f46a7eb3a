#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep (read here, no GPU needed).

ncu's CSV source page is per SASS instruction without line numbers, so the SASS of the very same build is disassembled
with `nvdisasm --print-line-info` and matched instruction by instruction.

usage: python scripts/ncu_lines.py gpurun_out/prof.ncu-rep superterrainplus_b200/libshf_b200.so emit_kernelILi2E [top] [kernel regex]
(the kernel regex selects the launch inside a report that holds several kernels; default = the name before "ILi")
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, symbol):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    out, on, line, fname = [], False, 0, ""
    for ln in text.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            on = symbol in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if "inlined at" not in ln:
                line = int(m.group(2))
                fname = os.path.basename(m.group(1))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append(((fname, line), m.group(2).strip()))
    return out


def main():
    rep, lib, symbol = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    kre = sys.argv[5] if len(sys.argv) > 5 else symbol.split("ILi")[0]
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre, "-c", "1"],
                                          capture_output=True, text=True).stdout.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    ii, si = h.index("Instructions Executed"), h.index("# Samples")
    body = rows[hdr + 1:]
    for i, r in enumerate(body):  # a report may repeat the kernel's section; keep the first one
        if r and r[0] == "Kernel Name":
            body = body[:i]
            break
    inst = [(int(r[ii]), int(r[si]), r[1]) for r in body if len(r) > ii and r[ii].isdigit()]
    sass = sass_lines(lib, symbol)
    if len(sass) != len(inst):
        print(f"warning: {len(sass)} disassembled instructions vs {len(inst)} profiled; is the .so the profiled build?")
    srcs = {}

    def text_of(key):
        f, line = key
        if f not in srcs:
            try:
                srcs[f] = open(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)).read().splitlines()
            except OSError:
                srcs[f] = []
        return srcs[f][line - 1].strip()[:100] if 0 < line <= len(srcs[f]) else ""
    per = {}
    for (n, s, _), (line, _) in zip(inst, sass):
        a = per.setdefault(line, [0, 0])
        a[0] += n
        a[1] += s
    tot_n = sum(a[0] for a in per.values()) or 1
    tot_s = sum(a[1] for a in per.values()) or 1
    print(f"total warp instructions {tot_n}, samples {tot_s}")
    for line, (n, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * n / tot_n:5.1f}% inst {100 * s / tot_s:5.1f}% samples  {line[0].replace('shf_', '').replace('.cuh', '')}:{line[1]}: {text_of(line)}")


if __name__ == "__main__":
    main()
