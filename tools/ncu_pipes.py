"""Pipe-level view of an ncu report (which unit limits a kernel): python tools/ncu_pipes.py <report.ncu-rep> [kernel-regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_tex_mem_texture.sum",
        "l1tex__t_sectors_pipe_tex_mem_texture_lookup_miss.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size"]
seen = set()
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    short = name.split("(")[0].replace("void ", "")
    if (pat and not pat.search(name)) or short in seen:
        continue
    seen.add(short)
    print("=====", short)
    for k in WANT:
        if k in hdr:
            print("  %-82s %18s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    # warp stall reasons (cycles a warp waits per issued instruction), largest first
    stalls = []
    for i, k in enumerate(hdr):
        m = re.match(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active", k)
        if m and r[i] not in ("", "n/a"):
            try:
                stalls.append((float(r[i].replace(",", "")), m.group(1)))
            except ValueError:
                pass
    for v, n in sorted(stalls, reverse=True)[:8]:
        print("  stall %-40s %8.2f warps per issue" % (n, v))
