"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): opcode mix, stall reasons, divergence.
Usage: ncu -i rep --page source --csv -k regex:NAME | python scripts/ncu_sass_summary.py [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 12
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:90], "hdr": None, "data": []}
        sections.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections[:1]:
    hdr, data = sec["hdr"], sec["data"]
    ix = {h: i for i, h in enumerate(hdr)}
    num = lambda r, k: float(r[ix[k]].replace(",", "") or 0)
    tot_s = sum(num(r, "# Samples") for r in data) or 1
    tot_n = sum(num(r, "Instructions Executed") for r in data) or 1
    print(sec["name"])
    print(f"SASS lines {len(data)}  samples {tot_s:.0f}  warp-instr {tot_n:.0f}")
    op_n, op_s = collections.Counter(), collections.Counter()
    for r in data:
        t = r[ix["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        op_n[op] += num(r, "Instructions Executed"); op_s[op] += num(r, "# Samples")
    print("by executed instr :", ", ".join(f"{k} {100*v/tot_n:.1f}%" for k, v in op_n.most_common(top + 6)))
    print("by stall samples  :", ", ".join(f"{k} {100*v/tot_s:.1f}%" for k, v in op_s.most_common(top)))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    sv = sorted(((sum(num(r, h) for r in data), h) for h in stalls), reverse=True)
    print("stall reasons     :", ", ".join(f"{h[6:]} {100*v/tot_s:.1f}%" for v, h in sv[:8]))
    act = sum(num(r, "Avg. Threads Executed") * num(r, "Instructions Executed") for r in data) / tot_n
    print(f"avg active threads per executed instr: {act:.1f} / 32")
    print("hottest SASS by samples:")
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top]:
        print(f"  {100*num(r,'# Samples')/tot_s:5.1f}%  exec {num(r,'Instructions Executed'):9.0f}  thr {num(r,'Avg. Threads Executed'):4.1f}  {r[ix['Source']][:80]}")
