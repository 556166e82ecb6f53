"""Dynamic instruction counts per source line: joins an `ncu --page source --csv` dump (per-instruction executed counts,
by address) with `nvdisasm -g` line info of the same kernel from the built library.
    python scripts/ncu_lines.py dump.csv <kernel-substring> [--so PATH] [--top N] [--op OPCODE]"""
import argparse, collections, csv, os, re, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("csv"); ap.add_argument("kernel")
ap.add_argument("--so", default=os.path.join(ROOT, "xai-audio-deepfakes_b200", "libaddvisor_sm100.so"))
ap.add_argument("--top", type=int, default=40); ap.add_argument("--op", default=None)
a = ap.parse_args()
a.so = os.path.abspath(a.so)
rows = list(csv.reader(open(a.csv)))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows if len(r) > 6 and r[0] not in ("Address", "Kernel Name")]
seen, dyn = set(), []
for r in body:
    if r[0] in seen: continue
    seen.add(r[0]); dyn.append((int(r[0], 16) if r[0].startswith("0x") else int(r[0]), r[ix["Source"]], int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)))
base = min(d[0] for d in dyn)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", a.so], cwd=tmp, capture_output=True)
lines = {}
for cb in os.listdir(tmp):
    if not cb.endswith(".cubin"): continue
    syms = subprocess.run(["readelf", "-sW", os.path.join(tmp, cb)], capture_output=True, text=True).stdout
    for ln in syms.splitlines():
        f = ln.split()
        if len(f) >= 8 and f[3] == "FUNC" and a.kernel in f[-1]:
            out = subprocess.run(["nvdisasm", "-g", "-fun", f[0].rstrip(":"), os.path.join(tmp, cb)], capture_output=True, text=True).stdout
            cur = None
            for t in out.splitlines():
                m = re.search(r'//## File "([^"]+)", line (\d+)', t)
                if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
                m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", t)
                if m: lines[int(m.group(1), 16)] = cur
per, pers = collections.Counter(), collections.Counter()
tot = 0
for addr, src, n, smp in dyn:
    t = src.split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    if a.op and op != a.op: continue
    per[lines.get(addr - base)] += n; pers[lines.get(addr - base)] += smp; tot += n
print("total", tot)
for k, v in per.most_common(a.top): print(f"{v:10d} {100*v/max(tot,1):5.1f}%  samples {pers[k]:5d}  {k}")
