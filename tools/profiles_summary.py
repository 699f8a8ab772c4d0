"""tools/profiles_summary.py -- turn the ncu artefacts in gpurun_out/ into the tracked summaries
under profiles/ (launch shares, per-kernel metrics, DRAM traffic of the dominant kernel)."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

def raw(rep):
    """rep: a .ncu-rep, or the CSV of its raw page (``ncu -i x.ncu-rep --page raw --csv > x.raw.csv``)."""
    if rep.endswith(".ncu-rep") and not os.path.isfile(rep):
        rep = rep[:-8] + ".raw.csv"
    if rep.endswith(".csv"):
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    return r[0], r[1], r[2:]

# ---- launch list -> shares of one evaluation
lines = [l for l in open(os.path.join(G, "%s_launches.csv" % tag)) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = {}
order = []
for row in rows:
    k = row["Kernel Name"].split("(")[0]
    per.setdefault(k, []).append(float(row["Metric Value"]))
    if k not in order: order.append(k)
evalk = [k for k in order if any(s in k for s in ("blu_phi", "blu_grad", "blu_hess"))]
tot = sum(sum(per[k]) / len(per[k]) for k in evalk)
out = ["# %s: ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 2`" % tag,
       "(`--metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and serialised: compare SHARES)", "",
       "| kernel | launches | mean us | share of one evaluation |", "|---|---|---|---|"]
for k in order:
    mean = sum(per[k]) / len(per[k]) / 1e3
    share = ("%.1f%%" % (100 * mean * 1e3 / tot)) if k in evalk else "(setup)"
    out.append("| `%s` | %d | %.1f | %s |" % (k, len(per[k]), mean, share))
open(os.path.join(P, "%s_launches.md" % tag), "w").write("\n".join(out) + "\n")
os.replace(os.path.join(G, "%s_launches.csv" % tag), os.path.join(P, "%s_launches.csv" % tag)) if False else None
import shutil; shutil.copy(os.path.join(G, "%s_launches.csv" % tag), os.path.join(P, "%s_launches.csv" % tag))

want = ['launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
def summarize(rep, title, fname):
    hdr, units, rws = raw(rep)
    ki = hdr.index("Kernel Name")
    o = ["# " + title, "", "| metric | unit | " + " | ".join("`%s`" % r[ki].split("(")[0][:40] for r in rws) + " |", "|---|---|" + "---|" * len(rws)]
    for i, h in enumerate(hdr):
        if h in want:
            o.append("| %s | %s | %s |" % (h, units[i], " | ".join(r[i] for r in rws)))
    st = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            st.append((h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), [r[i] for r in rws]))
    st.sort(key=lambda x: -float(x[1][0] or 0))
    o += ["", "Top warp stall reasons (warps per issue-active cycle):", ""]
    for h, v in st[:6]:
        o.append("* %s: %s" % (h, ", ".join(v)))
    open(os.path.join(P, fname), "w").write("\n".join(o) + "\n")
    return hdr, units, rws

hdr, units, rws = summarize(os.path.join(G, "%s_hess.ncu-rep" % tag), "%s: ncu --set full, blu_hess_kernel<4,true> (N=15, L=32767), 2 launches" % tag, "%s_hess_kernel.md" % tag)
def col(name):
    i = hdr.index(name); return units[i], [float(r[i]) for r in rws]
ur, rd = col("dram__bytes_read.sum"); uw, wr = col("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = sum(rd) / len(rd) * scale[ur] + sum(wr) / len(wr) * scale[uw]
json.dump({"kernel": "blu_hess_kernel<4,true>", "dram_bytes_per_launch": traffic, "dram_read_bytes": sum(rd) / len(rd) * scale[ur],
           "dram_write_bytes": sum(wr) / len(wr) * scale[uw], "algorithmic_bytes_per_launch": 8.0 * 32767 * 32767,
           "source": "profiles/%s_hess_kernel.md (ncu --set full --clock-control none, 2 launches)" % tag},
          open(os.path.join(P, "hess_kernel_traffic.json"), "w"), indent=1)
if os.path.isfile(os.path.join(G, "%s_small.ncu-rep" % tag)) or os.path.isfile(os.path.join(G, "%s_small.raw.csv" % tag)):
    summarize(os.path.join(G, "%s_small.ncu-rep" % tag), "%s: ncu --set full, setup + streaming kernels (N=15)" % tag, "%s_stream_kernels.md" % tag)
if os.path.isfile(os.path.join(G, "%s_stream20.ncu-rep" % tag)) or os.path.isfile(os.path.join(G, "%s_stream20.raw.csv" % tag)):
    summarize(os.path.join(G, "%s_stream20.ncu-rep" % tag), "%s: ncu --set full, Phi / gradient streaming kernels at N=20 (L=1048575)" % tag, "%s_stream_kernels_N20.md" % tag)
print(open(os.path.join(P, "%s_launches.md" % tag)).read())
print(open(os.path.join(P, "hess_kernel_traffic.json")).read())
