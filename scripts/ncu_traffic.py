"""Extracts per-launch DRAM traffic and duration of the captured kernels from an `ncu --set full` report:
   python scripts/ncu_traffic.py gpurun_out/prof_sweep.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_traffic.json
bench.py reads profiles/*_ncu_traffic.json (latest) for roofline.traffic."""
import csv, io, json, subprocess, sys

out = {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")}
    unit = {n: rows[1][i] for n, i in col.items()}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[unit["dram__bytes_read.sum"]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[unit["dram__bytes_write.sum"]]
        e = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "duration_us": 0.0})
        e["launches"] += 1
        e["dram_bytes"] += rd + wr
        e["duration_us"] += float(r[col["gpu__time_duration.sum"]])
for e in out.values():
    e["dram_bytes_per_launch"] = round(e.pop("dram_bytes") / e["launches"])
    e["duration_us_per_launch_under_ncu"] = round(e.pop("duration_us") / e["launches"], 2)
print(json.dumps({"source": sys.argv[1:], "kernels": out}, indent=1))
