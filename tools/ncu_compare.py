"""Side-by-side pipe / stall view of every launch in an ncu report: python tools/ncu_compare.py <report.ncu-rep>"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum", "l1tex__t_sectors_pipe_tex_mem_texture.sum",
        "l1tex__t_sectors_pipe_tex_mem_texture_lookup_miss.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
cols = []
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    cols.append((re.sub(r"\(.*", "", name).replace("void gdt::", ""), r))
print("%-66s" % "kernel", *["%16s" % c[0][:16] for c in cols])
print("%-66s" % "template args", *["%16s" % re.sub(r".*<", "<", c[1][hdr.index("Kernel Name")].split("(")[0])[-16:] for c in cols])
for k in WANT:
    if k in hdr:
        print("%-66s" % k[:66], *["%16s" % c[1][hdr.index(k)][:16] for c in cols])
for i, k in enumerate(hdr):
    m = re.match(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active", k)
    if m:
        print("%-66s" % ("stall " + m.group(1)), *["%16s" % c[1][i][:16] for c in cols])
