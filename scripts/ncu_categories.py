#!/usr/bin/env python
"""Groups the per-source-line shares printed by scripts/ncu_lines.py (emit_kernel) into the parts of the kernel.
The line ranges below belong to shf_events.cuh / shf_kernels.cuh as committed with profiles/r2_emit_instruction_shares.txt.

usage: python scripts/ncu_lines.py <rep> superterrainplus_b200/libshf_b200.so emit_kernelILi2ELi8 600 > lines.txt
       python scripts/ncu_categories.py lines.txt
"""
import re,sys
cats=[("dense slide (load_counts / slide_cols / store_sums + segment loop)",[("events",427,521),("events",769,799)]),
("packed emission (lists <= 16 events)",[("events",861,909)]),
("unpacked emission + staged_count",[("events",909,950),("events",523,531)]),
("active-list rebuild",[("events",805,860)]),
("mbarrier polling / arrive",[("kernels",551,587),("events",694,694),("events",768,768),("events",961,963),("events",714,714)]),
("producers (ring fill)",[("events",619,720),("kernels",560,565)]),
("consumer batch bookkeeping / item setup",[("events",722,768),("events",800,804),("events",950,966),("events",532,618)]),
]
tot={k:0.0 for k,_ in cats}; samples={k:0.0 for k,_ in cats}; other=0.0; others=[]
for ln in open(sys.argv[1]):
    m=re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% samples\s+(\w+):(\d+):",ln)
    if not m: continue
    p,s,f,l=float(m.group(1)),float(m.group(2)),m.group(3),int(m.group(4))
    for k,rs in cats:
        if any(ff==f and a<=l<=b for ff,a,b in rs):
            tot[k]+=p; samples[k]+=s; break
    else:
        other+=p; others.append(ln.strip()[:90])
for k,_ in cats: print("%5.1f%% inst %5.1f%% samples  %s"%(tot[k],samples[k],k))
print("%5.1f%% other"%other); print("\n".join(others[:8]))
