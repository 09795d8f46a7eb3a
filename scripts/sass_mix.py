#!/usr/bin/env python
"""Static SASS instruction mix of every kernel in the built library (no GPU needed): instruction count by class, the
memory instructions by kind, and the mnemonics that show what the kernel is made of (mbarrier SYNCS, cp.async LDGSTS,
VOTE / POPC / REDUX for the compaction, ATOMS, ...). usage: python scripts/sass_mix.py > profiles/r2_sass_mix.txt"""
import collections
import os
import re
import subprocess
import sys
import tempfile

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "superterrainplus_b200", "libshf_b200.so")
CLASSES = [
    ("global load", r"^LDG|^LD\.E|^LDC|^LDCU"), ("global store", r"^STG|^ST\.E"), ("shared load", r"^LDS"), ("shared store", r"^STS"),
    ("cp.async", r"^LDGSTS|^LDGDEPBAR|^DEPBAR"), ("mbarrier / barrier", r"^SYNCS|^BAR|^WARPSYNC|^MEMBAR|^FENCE"),
    ("atomics", r"^ATOM|^RED"), ("vote / popc / shuffle / redux", r"^VOTE|^POPC|^SHFL|^REDUX|^MATCH|^FLO|^BREV"),
    ("integer ALU", r"^IADD|^IMAD|^LOP|^SHF|^LEA|^ISETP|^SEL|^PRMT|^IABS|^VIADD|^VIMNMX|^IMNMX|^MOV|^CS2R|^S2R|^S2UR|^R2UR|^UIADD|^UMOV|^ULOP|^USHF|^ULEA|^UISETP|^UIMAD|^USEL|^PLOP|^UPLOP|^P2R|^R2P|^I2I|^SGXT|^BMSK|^UFLO|^UPRMT|^UP2UR|^UR2UP|^USGXT|^UBMSK|^UPOPC"),
    ("float / convert", r"^F|^I2F|^F2I|^MUFU|^D|^HADD|^HMUL|^HFMA|^I2FP|^F2FP"), ("branch / control", r"^BRA|^BSSY|^BSYNC|^EXIT|^CALL|^RET|^JMP|^NOP|^BREAK|^YIELD|^NANOSLEEP|^BPT|^WARPSYNC|^ERRBAR|^CCTL|^ACQBULK|^ENDCOLLECTIVE|^UCGABAR"),
]


def main():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", cubin], capture_output=True, text=True).stdout
    kernels, name = collections.OrderedDict(), None
    for ln in text.splitlines():
        m = re.match(r"//-+ \.text\.(\S+)", ln)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and name:
            kernels[name].append(m.group(1))
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# static SASS mix of {os.path.relpath(LIB)} (sm_100a), nvdisasm; counts are instructions in the binary, not executed")
    for (sym, ops), pretty in zip(kernels.items(), demangle):
        pretty = re.sub(r"\(.*", "", pretty)
        if "emit_kernel" in pretty and not re.search(r"<[12], 8>", pretty):
            continue   # the K = 4, 8 and 16-bit-ring instances differ only in widths
        if "events_kernel" in pretty and "<2, " not in pretty:
            continue
        if "vscan_kernel" in pretty and "<2, true>" not in pretty:
            continue
        if "presence_kernel<1>" in pretty or "remap_kernel<1>" in pretty:
            continue
        total = len(ops)
        by = collections.Counter()
        for o in ops:
            for cname, pat in CLASSES:
                if re.match(pat, o):
                    by[cname] += 1
                    break
            else:
                by["other"] += 1
        mn = collections.Counter(o.split(".")[0] for o in ops)
        print(f"\n== {pretty}: {total} instructions")
        print("   " + ", ".join(f"{k} {v} ({100 * v / total:.0f}%)" for k, v in by.most_common()))
        tell = ["SYNCS", "LDGSTS", "VOTE", "POPC", "REDUX", "SHFL", "ATOMS", "ATOMG", "RED", "LDS", "STS", "LDG", "STG", "BAR", "NANOSLEEP", "I2FP", "FMUL", "DADD", "DMUL", "UTMALDG", "UTCMMA"]
        print("   mnemonics: " + ", ".join(f"{t} {mn[t]}" for t in tell if mn[t]))


if __name__ == "__main__":
    main()
