"""Print a compact set of ncu raw metrics for the first kernel whose name matches argv[2] in report argv[1]."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__inst_executed_pipe_tc", "sm__pipe_tensor_subpipe", "sm__pipe_tc",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "issue_stalled", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct", "tensor", "tma", "smsp__average_warp_latency"]
for r in rows[2:]:
    if pat not in r[4]:
        continue
    print(r[4][:80])
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in KEYS) and ".min" not in h and ".max" not in h and "peak_sustained" not in h.split("pct_of_")[0]:
            try:
                if float(v.replace(",", "")) == 0:
                    continue
            except ValueError:
                pass
            if "per_second" in h or ".sum.pct" in h:
                continue
            print("  %-95s %-10s %s" % (h, u, v))
    break
