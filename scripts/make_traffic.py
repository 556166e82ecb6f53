"""profiles/traffic.json from an `ncu --page raw --csv` dump of scripts/prof_traffic.py (one row per captured launch).
    python scripts/make_traffic.py raw.csv --batch 1024 [--out profiles/traffic.json] [--source profiles/<csv name>]
Per kernel: DRAM bytes read + written and warp-instructions per launch, scaled to the 64-clip launch bench.py times
(the capture runs at `--batch` clips so that the outputs exceed L2 and the write-back is counted)."""
import argparse, csv, json, os
ap = argparse.ArgumentParser()
ap.add_argument("csv"); ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"))
ap.add_argument("--source", default=None)
a = ap.parse_args()
rows = [r for r in csv.reader(open(a.csv)) if r]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
units = rows[rows.index(hdr) + 1]
body = rows[rows.index(hdr) + 2:]


def val(r, name):
    x, u = float(r[ix[name]].replace(",", "")), units[ix[name]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "": 1.0, "inst": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}
    return x * scale.get(u, 1.0)


ALG = {  # algorithmic bytes per clip, cfg-2: N = 64000, F = 257, T = 401
    "explain4": 4 * 64000 + 4 * 257 * 401 + 8 * 64000,
    "stft3_X": 4 * 64000 + 8 * 257 * 401,
    "stft3_X_mag_phase": 4 * 64000 + 16 * 257 * 401,
    "istft4": 8 * 257 * 401 + 4 * 64000,
    "mel_fused": 4 * 64000 + 4 * 80 * 251,
}
out, seen_stft = {}, 0
for r in body:
    name = r[ix["Kernel Name"]]
    if "explain4_kernel" in name: key, b = "explain4", a.batch
    elif "stft3_kernel" in name:   # stft3_kernel<NF, MAG, PHASE, ...>: "<512, 0, 0" = spectrum only
        targs = [t.strip() for t in name.split("<", 1)[1].split(">")[0].split(",")] if "<" in name else []
        if len(targs) < 3:
            key, b = ("stft3_X", a.batch) if seen_stft % 2 == 0 else ("stft3_X_mag_phase", a.batch)
            seen_stft += 1
        else:
            key, b = ("stft3_X", a.batch) if targs[1] in ("0", "false") and targs[2] in ("0", "false") else ("stft3_X_mag_phase", a.batch)
    elif "istft4_kernel" in name: key, b = "istft4", a.batch
    elif "mel_fused_kernel" in name: key, b = "mel_fused", 64
    else: continue
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    out[key] = {"kernel": name.split("(")[0], "captured_clips": b, "per_64_clips": {
        "read_bytes": rd * 64 / b, "write_bytes": wr * 64 / b, "dram_bytes": (rd + wr) * 64 / b,
        "algorithmic_bytes": ALG[key] * 64, "warp_inst": val(r, "smsp__inst_executed.sum") * 64 / b},
        "dram_over_algorithmic": (rd + wr) / (ALG[key] * b), "duration_us_under_ncu": val(r, "gpu__time_duration.sum")}
e = out.get("explain4", {}).get("per_64_clips", {})
doc = {"explain_kernel_bytes_per_launch": e.get("dram_bytes"), "explain_kernel_warp_inst_per_launch": e.get("warp_inst"),
       "read_bytes": e.get("read_bytes"), "write_bytes": e.get("write_bytes"),
       "note": f"dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum per launch from ONE `ncu --set full "
               f"--clock-control none` capture of scripts/prof_traffic.py at {a.batch} clips per launch (outputs exceed the 126 MB "
               f"L2, so the write-back is counted), scaled to the 64-clip launch bench.py times; regenerate with scripts/make_traffic.py",
       "source": a.source, "kernels": out}
json.dump(doc, open(a.out, "w"), indent=1)
print(json.dumps({k: {"dram_over_algorithmic": round(v["dram_over_algorithmic"], 3), "MB_per_64": round(v["per_64_clips"]["dram_bytes"] / 1e6, 2),
                      "Minst_per_64": round(v["per_64_clips"]["warp_inst"] / 1e6, 2)} for k, v in out.items()}))
