"""sums dram__bytes_read/write over the kernels of ONE step from an `ncu --set full` raw CSV
(ncu -i X.ncu-rep --page raw --csv) and writes profiles/r1_traffic.json + a per-kernel table."""
import csv, json, sys
raw, n_bytes, first, count = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, k):
    v = float(r[ix[k]]); u = units[ix[k]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
per = []
for r in rows[2 + first: 2 + first + count]:
    per.append({"kernel": r[ix["Kernel Name"]].split("(")[0], "ms": float(r[ix["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "ns": 1e-6}[units[ix["gpu__time_duration.sum"]]],
                "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum")})
tot = sum(p["dram_read"] + p["dram_write"] for p in per)
json.dump({"input_bytes": n_bytes, "dram_bytes_per_step": tot, "kernels": per}, open("profiles/r1_traffic.json", "w"), indent=1)
for p in per: print(f'{p["kernel"]:28s} {p["ms"]:8.3f} ms  read {p["dram_read"]/1e6:9.1f} MB  write {p["dram_write"]/1e6:9.1f} MB')
print("total", tot / 1e6, "MB =", tot / n_bytes, "bytes per input byte")
