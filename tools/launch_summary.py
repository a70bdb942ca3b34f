"""Aggregate an `ncu --csv` launch list (gpu__time_duration.sum [+ dram__bytes_*]) per kernel name."""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
iname, imetric, ival, iunit, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
per_launch = {}
for r in rows[1:]:
    d = per_launch.setdefault(r[iid], {"name": r[iname]})
    v = float(r[ival].replace(",", ""))
    u = r[iunit]
    if r[imetric].startswith("gpu__time"):
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)       # -> ms
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r[imetric]] = v
agg = {}
for d in per_launch.values():
    name = re.sub(r"\(.*", "", d["name"])
    name = re.sub(r"void |pnx::|\(anonymous namespace\)::", "", name)[:100]
    a = agg.setdefault(name, [0.0, 0, 0.0, 0.0])
    a[0] += d.get("gpu__time_duration.sum", 0.0)
    a[1] += 1
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[0] for a in agg.values())
print("# %d launches, %.3f ms summed kernel time, DRAM read %.2f GB write %.2f GB (serialised, cold-cache per-launch numbers under ncu)" % (
    sum(a[1] for a in agg.values()), tot, sum(a[2] for a in agg.values()) / 1e9, sum(a[3] for a in agg.values()) / 1e9))
print("# ms_total  share  launches  dram_rd_MB  dram_wr_MB  GB/s   kernel")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%9.3f  %5.1f%%  %6d  %10.1f  %10.1f  %6.0f  %s" % (a[0], 100 * a[0] / tot, a[1], a[2] / 1e6, a[3] / 1e6,
                                                         (a[2] + a[3]) / max(a[0], 1e-9) / 1e6, name))
