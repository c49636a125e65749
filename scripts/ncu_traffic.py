"""sums dram__bytes_read/write over the kernels of ONE step (from the last walk_fused launch to the end)
of an `ncu --set full` raw CSV (ncu -i X.ncu-rep --page raw --csv) and writes profiles/r2_traffic.json
plus a per-kernel table.  usage: ncu_traffic.py raw.csv input_bytes [out.json]"""
import csv, json, sys
raw, n_bytes = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/r2_traffic.json"
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, k):
    v = float(r[ix[k]]); u = units[ix[k]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def ms(r):
    return float(r[ix["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "ns": 1e-6}[units[ix["gpu__time_duration.sum"]]]
body = rows[2:]
walks = [i for i, r in enumerate(body) if "walk_fused_kernel<" in r[ix["Kernel Name"]] and ", 0>" in r[ix["Kernel Name"]]]
first = walks[-1]
per = []
for r in body[first:]:
    per.append({"kernel": r[ix["Kernel Name"]].split("(")[0], "ms": ms(r),
                "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum")})
tot = sum(p["dram_read"] + p["dram_write"] for p in per)
agg = {}
for p in per:
    a = agg.setdefault(p["kernel"], {"launches": 0, "ms": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    a["launches"] += 1; a["ms"] += p["ms"]; a["dram_read"] += p["dram_read"]; a["dram_write"] += p["dram_write"]
json.dump({"input_bytes": n_bytes, "dram_bytes_per_step": tot, "kernels": agg}, open(out, "w"), indent=1)
for k, a in agg.items():
    print(f'{k[:60]:60s} x{a["launches"]:2d} {a["ms"]:8.3f} ms  read {a["dram_read"]/1e6:9.1f} MB  write {a["dram_write"]/1e6:9.1f} MB')
print("total", tot / 1e6, "MB =", tot / n_bytes, "bytes per input byte")
