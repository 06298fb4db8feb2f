"""Summarise an .ncu-rep: key raw metrics + aggregated stall reasons + hottest SASS lines.
usage: python tools/ncu_summary.py file.ncu-rep [topN]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum"]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {units[i]}")
    print("---")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hdr_i:
    h = rows[hdr_i[0]]
    end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
    body = [r for r in rows[hdr_i[0] + 1:end] if len(r) == len(h)]
    si = h.index("# Samples")
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si]) for r in body if r[si].isdigit()) or 1
    agg = {s: sum(int(r[h.index(s)]) for r in body if r[h.index(s)].isdigit()) for s in stalls}
    print("total samples", tot)
    for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        print(f"  {s}: {100 * v / tot:.1f}%")
    top = sorted([r for r in body if r[si].isdigit()], key=lambda r: -int(r[si]))[:topn]
    for r in top:
        st = {s: int(r[h.index(s)]) for s in stalls if r[h.index(s)].isdigit() and int(r[h.index(s)]) > 0}
        print(r[si], r[1].strip()[:64], dict(sorted(st.items(), key=lambda x: -x[1])[:2]))
