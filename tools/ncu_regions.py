import csv, subprocess, sys, io
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
for r in rows:
    if r and r[0] == "Line No": hdr = r; break
print(hdr[:12])
cur=None; reg={}
import os, re
_SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "latticeqmc_b200", "csrc", "sweep_l2.cuh")
_ANCHORS = [("gemm", "struct L2PanelIssue"), ("flush_smem", "__device__ void l2_flush("), ("propose_smem", "__device__ void l2_propose_slice("),
            ("tmem helpers", "constexpr int L2_KDT"), ("flush_tmem1", "__device__ void l2_flush_tmem("), ("flush_tmem2", "__device__ void l2_flush_tmem2("),
            ("propose_tmem", "__device__ void l2_propose_slice_tmem("), ("tmemx", "constexpr int L2_TMEMX_COLS"),
            ("gj_inverse", "__device__ void l2_gj_inverse("), ("recompute", "__device__ void l2_recompute("), ("wrap", "__device__ void l2_wrap("),
            ("kernel", "struct L2Params")]
_lines = open(_SRC).read().split("\n")
_bounds = []
for name, pat in _ANCHORS:
    for i, l in enumerate(_lines):
        if pat in l:
            _bounds.append((i + 1, name)); break
_bounds.sort()
def region(f, ln):
    if f == "sweep_l2.cuh":
        cur = "sweep_l2.cuh (head)"
        for start, name in _bounds:
            if ln >= start - 3: cur = name
        return cur
    return f
tot=0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name"): continue
    if r[0] != "" and r[2] == "-":
        try: s = int(r[4]); ex = int(r[7])
        except ValueError: continue
        k = region(cur, int(r[0])); a = reg.get(k,(0,0)); reg[k]=(a[0]+s,a[1]+ex); tot+=s
for k,v in sorted(reg.items(), key=lambda kv:-kv[1][0]): print(f"{k:20s} samples {v[0]:9d} {100*v[0]/tot:5.1f}%  inst {v[1]:13d}")
