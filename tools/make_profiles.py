"""Turn gpurun_out/ ncu artefacts into the tracked, judge-readable summaries under profiles/.

    python tools/make_profiles.py r01

 - <round>_launches.csv           the ncu launch list of `bench.py` (--metrics gpu__time_duration.sum)
 - <round>_launch_shares.md       per-kernel count / total time / share of the step
 - <round>_kernels.json           selected `ncu --set full` metrics per profiled kernel
 - traffic.json                   dram bytes per launch of k_onesweep at 2^26 pairs (read by bench.py)
"""
import csv, json, os, re, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)

def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("usrt::<unnamed>::", "").replace("unnamed>::", "").replace("usrt::", "")
    return name.strip()

# ---- launch list ------------------------------------------------------------------------------
src = os.path.join(G, rnd + "_launches.csv")
if os.path.exists(src):
    lines = [l for l in open(src) if l.startswith('"')]
    open(os.path.join(P, rnd + "_launches.csv"), "w").writelines(lines)
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Block Size"], r["Grid Size"]])
        a[0] += 1; a[1] += float(r["Metric Value"])
    tot = sum(a[1] for a in agg.values())
    ours = {k: a for k, a in agg.items() if k.startswith("k_")}
    tot_ours = sum(a[1] for a in ours.values())
    with open(os.path.join(P, rnd + "_launch_shares.md"), "w") as f:
        f.write("# %s: ncu launch list of `python bench.py --steps 2 --warmup 3` (cold-cache, serialised: compare shares)\n\n" % rnd)
        f.write("| kernel | launches | total us | share of all | share of this repo's kernels | block | grid (last) |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% | %s | %s | %s |\n" % (k[:70], a[0], a[1] / 1e3, 100 * a[1] / tot,
                    ("%.1f%%" % (100 * a[1] / tot_ours)) if k in ours else "-", a[2], a[3]))
    print("launch shares written;", len(rows), "launches")

# ---- full-set captures ---------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
out = {}
for rep, label in ((rnd + "_prof_sort.ncu-rep", "sort 2^26 pairs (tools/sort_bench.py --n 26)"),
                   (rnd + "_prof_build_trace.ncu-rep", "bench.py step (1,048,576 tris, 1920x1080)")):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        continue
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        k = short(r[idx["Kernel Name"]])
        key = "%s @ %s" % (k, label)
        if key in out:
            continue
        out[key] = {w: (r[idx[w]] + " " + units[idx[w]]).strip() for w in WANT if w in idx}
json.dump(out, open(os.path.join(P, rnd + "_kernels.json"), "w"), indent=1)
for key, m in out.items():
    if key.startswith("k_onesweep") and "dram__bytes_read.sum" in m:
        def mb(s):
            v, u = s.split()[0], s.split()[1]
            return float(v) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]
        t = mb(m["dram__bytes_read.sum"]) + mb(m["dram__bytes_write.sum"])
        json.dump({"k_onesweep_2p26_bytes_per_launch": t, "source": "ncu --set full, " + key,
                   "algorithmic_bytes_per_launch": 16 * (1 << 26)}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
        print("traffic", t)
print("kernels:", list(out))
