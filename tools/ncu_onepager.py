"""One-page markdown summary of an .ncu-rep (run here, no GPU needed): launch shape, time, DRAM traffic, pipe
utilisation, stall reasons, and the source lines with the most stall samples / instructions.
    python tools/ncu_onepager.py gpurun_out/x.ncu-rep [--top 25] [--traffic-json profiles/traffic.json --workload "..."]
With --traffic-json the DRAM bytes of the captured launch are written as the per-launch traffic bench.py reports."""
import argparse, csv, io, json, os, subprocess, sys, collections

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("--top", type=int, default=25)
ap.add_argument("--traffic-json"); ap.add_argument("--workload", default="")
a = ap.parse_args()

def page(name):
    out = subprocess.run(["ncu", "-i", a.rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def g(k, fmt="%s"):
    if k not in m: return "n/a"
    v, u = m[k]
    try: v = fmt % float(v.replace(",", ""))
    except ValueError: pass
    return ("%s %s" % (v, u)).strip()
def byt(k):
    v, u = m[k]; v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]

print("# ncu --set full: `%s`" % m.get("Kernel Name", ("?",))[0])
print("source: `%s`\n" % os.path.basename(a.rep))
rows = [("gpu__time_duration.sum", "%.3f"), ("launch__grid_size", "%d"), ("launch__block_size", "%d"),
        ("launch__registers_per_thread", "%d"), ("launch__shared_mem_per_block_static", "%.0f"),
        ("launch__shared_mem_per_block_dynamic", "%.0f"), ("launch__occupancy_limit_registers", "%d"),
        ("launch__occupancy_limit_shared_mem", "%d"), ("launch__occupancy_limit_warps", "%d"),
        ("dram__bytes_read.sum", "%.3f"), ("dram__bytes_write.sum", "%.3f"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%.1f"), ("lts__t_sector_hit_rate.pct", "%.1f"),
        ("l1tex__t_sector_hit_rate.pct", "%.1f"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "%.1f"), ("sm__inst_executed.avg.per_cycle_elapsed", "%.3f"),
        ("smsp__inst_executed.sum", "%.4g"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "%.2f"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "%.1f"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "%.1f"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "%.1f"),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "%.1f"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "%.1f"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "%.1f"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "%.1f")]
print("| metric | value |\n|---|---|")
for k, f in rows:
    if k in m: print("| %s | %s |" % (k, g(k, f)))
tr = byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum")
print("| dram bytes per launch (read + write) | %.4g B |" % tr)
print("\nstall reasons (warps per issue):\n")
st = sorted(((float(v[0]), k.split("issue_stalled_")[1].split("_per_issue")[0]) for k, v in m.items()
             if "issue_stalled" in k and k.endswith("per_issue_active.ratio")), reverse=True)
print(", ".join("%s %.2f" % (n, v) for v, n in st if v >= 0.05))

if a.traffic_json:
    with open(a.traffic_json, "w") as f:
        json.dump({"dram_bytes_per_launch": tr, "dram_bytes_read": byt("dram__bytes_read.sum"),
                   "dram_bytes_write": byt("dram__bytes_write.sum"),
                   "gpu_time_ms_under_ncu": float(m["gpu__time_duration.sum"][0].replace(",", "")),
                   "kernel": m.get("Kernel Name", ("?",))[0], "workload": a.workload,
                   "source": os.path.basename(a.rep), "tool": "tools/ncu_onepager.py"}, f, indent=1)

out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = {}
tot_s = tot_i = 0
fname, ci = "?", None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": ci = {n: k for k, n in enumerate(r) if n not in ("Source",)}; continue
    if ci is None or len(r) < 8 or r[2] != "-": continue          # CUDA-line aggregates carry "-" as address
    try:
        smp = float(r[ci["# Samples"]] or 0); ins = float(r[ci["Instructions Executed"]] or 0); thr = float(r[ci["Thread Instructions Executed"]] or 0)
    except (ValueError, KeyError):
        continue
    e = agg.setdefault((fname, r[0], r[1].strip()), [0, 0, 0]); e[0] += smp; e[1] += ins; e[2] += thr
    tot_s += smp; tot_i += ins
if agg:
    print("\ntotal warp instructions %.4g, stall samples %d" % (tot_i, tot_s))
    byfile = collections.Counter(); byfile_i = collections.Counter()
    for (f, _, _), (smp, ins, _) in agg.items(): byfile[f] += smp; byfile_i[f] += ins
    for f, v in byfile.most_common(): print("  %-28s samples %5.1f%%  instructions %5.1f%%" % (f, 100 * v / max(tot_s, 1), 100 * byfile_i[f] / max(tot_i, 1)))
    print("\ntop source lines by stall samples (share of samples, share of warp instructions, active lanes):\n")
    print("```")
    for (f, ln, txt), (smp, ins, thr) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print("%-20s %4s smp %5.1f%% inst %5.1f%% lanes %4.1f | %s" % (f, ln, 100 * smp / max(tot_s, 1), 100 * ins / max(tot_i, 1), thr / ins if ins else 0, txt[:120]))
    print("```")
