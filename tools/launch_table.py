"""ncu --csv launch list (gpu__time_duration + tensor pipe + dram bytes) -> per-kernel table and the WHOLE-OP
tensor-pipe figure of a launch range: sum(time_i * pipe_i) / sum(time_i) (pack / row / finalize kernels in the denominator).
usage: python tools/launch_table.py file.csv [first_id last_id]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = {}
for r in rows[hi + 1:]:
    if len(r) > vi:
        d.setdefault(int(r[ii]), {"kernel": r[ki]})[r[mi]] = float(r[vi].replace(",", "")) if r[vi] else 0.0
ids = sorted(d)
lo, hi_ = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (ids[0], ids[-1])
tot = pipe = rd = wr = 0.0
out = []
for i in ids:
    if lo <= i <= hi_:
        e = d[i]
        t = e.get("gpu__time_duration.sum", 0.0) / 1e3
        p = e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        tot += t; pipe += t * p
        rd += e.get("dram__bytes_read.sum", 0.0); wr += e.get("dram__bytes_write.sum", 0.0)
        name = e["kernel"].split("(")[0].split("::")[-1]
        out.append({"id": i, "kernel": name, "us": round(t, 2), "tensor_pipe_pct": round(p, 2),
                    "dram_read_MB": round(e.get("dram__bytes_read.sum", 0.0) / 1e6, 2), "dram_write_MB": round(e.get("dram__bytes_write.sum", 0.0) / 1e6, 2)})
print(json.dumps({"launches": out, "total_us": round(tot, 2), "whole_op_tensor_pipe_pct": round(pipe / tot, 2) if tot else None,
                  "dram_MB": round((rd + wr) / 1e6, 2)}, indent=1))
