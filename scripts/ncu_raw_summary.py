"""Pick the metrics the profiles/ summaries quote out of an `ncu --page raw --csv` export."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct"]

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
if len(rows) < 3:
    sys.exit(f"{sys.argv[1]}: no kernel was captured (empty ncu export)")
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print(f"## {r[col['Kernel Name']][:70]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
    for w in WANT:
        for h in hdr:
            if h == w or h.endswith(w):
                print(f"  {w:80s} {r[col[h]]:>18s} {units[col[h]]}")
                break
