import csv, subprocess, sys, io
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
for r in rows:
    if r and r[0] == "Line No": hdr = r; break
print(hdr[:12])
cur=None; reg={}
def region(f, ln):
    if f == "sweep_l2.cuh":
        for name, lo, hi in (("gemm",160,345),("flush_smem",346,420),("propose_smem",421,600),("tmem helpers",601,640),("flush_tmem1",641,690),("flush_tmem2",691,750),("propose_tmem",751,905),("gj_inverse",906,1040),("recompute",1041,1085),("wrap",1086,1110),("kernel",1111,1200)):
            if lo <= ln <= hi: return name
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
