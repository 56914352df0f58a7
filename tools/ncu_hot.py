"""Per-source-line hot spots from an ncu report:  python tools/ncu_hot.py <report.ncu-rep> <kernel-regex> [top]
Reads `ncu --page source --print-source cuda,sass --csv` and prints, per CUDA source line, the share of stall samples and
executed warp instructions, the dominant stall reasons and excess shared-memory wavefronts."""
import csv, subprocess, sys, collections, io

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kre], capture_output=True, text=True).stdout
cur_file, hdr, table, seen_fn = None, None, collections.OrderedDict(), []
def num(x):
    try: return int(float(x))
    except Exception: return 0
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if r[1] not in seen_fn: seen_fn.append(r[1])
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr[4:], r[4:]))
    e = table.setdefault((cur_file, r[0]), {"src": r[1], "samples": 0, "inst": 0, "stalls": collections.Counter(), "excess": 0})
    e["samples"] += num(d.get("# Samples")); e["inst"] += num(d.get("Instructions Executed"))
    e["excess"] += num(d.get("L1 Wavefronts Shared Excessive"))
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and num(v):
            e["stalls"][k[6:]] += num(v)
tot_s = sum(e["samples"] for e in table.values()) or 1
tot_i = sum(e["inst"] for e in table.values()) or 1
print("functions:", [f[:50] for f in seen_fn], " total samples:", tot_s, " warp-inst:", tot_i)
allst = collections.Counter()
for e in table.values(): allst.update(e["stalls"])
print("stall mix:", ", ".join("%s %.0f%%" % (k, 100.0 * v / max(sum(allst.values()), 1)) for k, v in allst.most_common(8)))
for (f, line), e in sorted(table.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = ",".join("%s:%d" % kv for kv in e["stalls"].most_common(3))
    print("%5.1f%% smp %5.1f%% inst  %-16s:%-4s %-64s [%s]%s" % (100.0 * e["samples"] / tot_s, 100.0 * e["inst"] / tot_i, f[:16], line,
          e["src"].strip()[:64], st, (" smem_excess=%d" % e["excess"]) if e["excess"] else ""))
