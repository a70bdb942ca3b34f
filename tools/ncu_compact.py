"""Compact per-kernel table from `ncu --set full` reports: python tools/ncu_compact.py rep1.ncu-rep [rep2 ...]"""
import csv, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
        ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum", "UTCHMMA bf16 ops"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__inst_executed_op_tma_ld.sum", "TMA load instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__shared_mem_per_block_static", "static smem/block")]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print("## %s: empty report\n" % rep)
        continue
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("## %s   [%s]" % (r[col["Kernel Name"]][:110], rep.split("/")[-1]))
        for k, label in KEYS:
            if k in col and r[col[k]] not in ("", "0"):
                print("  %-34s %14s %s" % (label, r[col[k]][:14], units[col[k]]))
        st = []
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  top stalls (warps per issue)      " + ", ".join("%s %.2f" % (n, v) for v, n in st[:4]))
        print()
