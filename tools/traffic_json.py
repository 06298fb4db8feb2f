"""Turn an ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of the bench's
adain_tma_kernel launches) into profiles/adain_traffic.json (what bench.py reports as roofline.traffic).
    python tools/traffic_json.py gpurun_out/bench_traffic.csv [batch]"""
import csv, io, json, os, sys
src = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
per = {}
for r in rows:
    if "adain_tma_kernel" not in r["Kernel Name"]:
        continue
    d = per.setdefault(int(r["ID"]), {})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"].lower()
    name = r["Metric Name"]
    if name.startswith("dram__bytes"):
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        d[name] = v * mult
    else:
        mult = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6, "second": 1e3}[unit]
        d["ms"] = v * mult
launches = [per[k] for k in sorted(per)][-5:]          # the last step's five launches
levels = [256, 128, 64, 32, 16]
hw = 512 * 512
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over the 5 adain_tma_kernel "
                 f"launches of one bench step (batch {batch}); tools/traffic_json.py", "per_launch": []}
tot_d = tot_a = 0.0
for c, l in zip(levels, launches):
    alg = (3 if c == 256 else 4) * batch * c * hw * 4
    dram = l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]
    tot_d += dram; tot_a += alg
    out["per_launch"].append({"level_C": c, "algorithmic_bytes": alg, "dram_bytes": dram, "ratio": dram / alg,
                              "ncu_duration_ms": l["ms"], "ncu_GBs_algorithmic": alg / l["ms"] / 1e6})
out["dram_bytes_per_launch_avg"] = tot_d / 5
out["algorithmic_bytes_per_launch_avg"] = tot_a / 5
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "adain_traffic.json"), "w"), indent=1)
print(json.dumps({"dram_over_algorithmic": tot_d / tot_a, "dram_bytes_per_launch_avg": tot_d / 5}))
