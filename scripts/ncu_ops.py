"""Opcode histogram (dynamic, from an `ncu --page source --csv` dump) per kernel section, plus the top source lines.
    python scripts/ncu_ops.py dump.csv [top]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "body": []}
        sections.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) > 6:
        cur["body"].append(r)
for s in sections:
    ix = {h: i for i, h in enumerate(s["hdr"])}
    ex, src, samp = ix["Instructions Executed"], ix["Source"], ix["# Samples"]
    ops, opss = collections.Counter(), collections.Counter()
    for r in s["body"]:
        t = r[src].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(r[ex] or 0)
        opss[op] += int(r[samp] or 0)
    tot, ts = sum(ops.values()), sum(opss.values())
    print("==", s["name"][:110])
    print("   static", len(s["body"]), "dynamic warp-inst", tot, "samples", ts)
    for k, v in ops.most_common(top):
        print(f"   {k:12s} {v:10d} {100 * v / tot:5.1f}%   samples {100 * opss[k] / max(ts, 1):5.1f}%")
    stall = collections.Counter()
    for i, h in enumerate(s["hdr"]):
        if h.startswith("stall_"):
            for r in s["body"]:
                try:
                    stall[h] += int(r[i] or 0)
                except (ValueError, IndexError):
                    pass
    print("   stalls:", {k[6:]: v for k, v in stall.most_common(10)})
