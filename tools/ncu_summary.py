"""Summarise an `ncu --page raw --csv` dump: one line per launch with the metrics the roofline needs."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
names = sys.argv[2].split(",") if len(sys.argv) > 2 else None
hdr, units, data = rows[0], rows[1], rows[2:]
M = [("t_us", "gpu__time_duration.sum"), ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum"),
     ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
     ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("l2sect_M", "lts__t_sectors.sum"), ("l2hit%", "lts__t_sector_hit_rate.pct"),
     ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("grid", "launch__grid_size"), ("smemKB", "launch__shared_mem_per_block_dynamic"),
     ("ctas/sm(smem)", "launch__occupancy_limit_shared_mem"), ("regs", "launch__registers_per_thread"),
     ("bankconf_M", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")]
idx = {k: hdr.index(m) for k, m in M if m in hdr}
def conv(k, v, u):
    v = float(v.replace(",", ""))
    if k == "t_us": return v / 1e3 if u == "ns" else (v if u in ("us", "usecond") else v * 1e3 if u == "ms" else v)
    if k.startswith("dram") and k.endswith("MB"): return v / 1e6 if u == "byte" else v / 1e3 if u == "Kbyte" else v if u == "Mbyte" else v * 1e3
    if k in ("l2sect_M", "bankconf_M"): return v / 1e6
    if k == "smemKB": return v / 1024 if u == "byte" else v if u == "Kbyte" else v
    return v
print("layer".ljust(10) + "".join(k.rjust(14) for k in idx))
ki = hdr.index("Kernel Name")
for i, r in enumerate(data):
    nm = names[i] if names and i < len(names) else r[ki][:9]
    print(nm.ljust(10) + "".join(f"{conv(k, r[j], units[j]):14.2f}" for k, j in idx.items()))
