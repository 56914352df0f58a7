"""Text summary of an ncu report for profiles/:  python tools/ncu_summary.py <report.ncu-rep> > profiles/<name>.txt
One block per distinct (kernel, grid): duration, DRAM bytes, throughput percentages, occupancy, registers, pipe usage."""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [("duration", "gpu__time_duration.sum"), ("dram_read", "dram__bytes_read.sum"), ("dram_write", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("registers", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("dyn_smem", "launch__shared_mem_per_block_dynamic"), ("sm_clock", "sm__cycles_elapsed.avg.per_second"),
        ("warp_inst", "smsp__inst_executed.sum")]
seen = set()
print("# ncu --set full --clock-control none summary of %s (first launch of each kernel/grid)" % rep.split("/")[-1])
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    key = (name, r[idx["launch__grid_size"]])
    if key in seen:
        continue
    seen.add(key)
    print("\n" + name)
    for label, col in WANT:
        if col in idx:
            print("    %-18s %s %s" % (label, r[idx[col]], units[idx[col]]))
