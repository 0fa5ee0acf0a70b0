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
for rep, label in ((rnd + "_prof_sort.ncu-rep", "sort 2^26 pairs (tools/micro/sort_lab.cu; r01: tools/sort_bench.py --n 26)"),
                   (rnd + "_prof_build_trace.ncu-rep", "bench.py step (1,048,576 tris, 1920x1080)"),
                   (rnd + "_prof_trace_c1.ncu-rep", "configs[0]: 65,536-tri soup, 512x512 primary rays (tools/ncu_trace.py c1)"),
                   (rnd + "_prof_trace_c4.ncu-rep", "configs[3]: 16,777,216-tri soup, 2,073,600 incoherent rays (tools/ncu_trace.py c4)")):
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
        n_same = sum(1 for kk in out if kk.startswith(key))
        if n_same and not k.startswith("k_onesweep"):
            continue
        if n_same:
            key += " #%d" % (n_same + 1)
        out[key] = {w: (r[idx[w]] + " " + units[idx[w]]).strip() for w in WANT if w in idx}
json.dump(out, open(os.path.join(P, rnd + "_kernels.json"), "w"), indent=1)
for key, m in out.items():
    if key.startswith("k_onesweep") and "2^26" in key and "dram__bytes_read.sum" in m and " #" not in key:
        def mb(s):
            v, u = s.split()[0], s.split()[1]
            return float(v) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]
        t = mb(m["dram__bytes_read.sum"]) + mb(m["dram__bytes_write.sum"])
        json.dump({"k_onesweep_2p26_bytes_per_launch": t, "source": "ncu --set full, " + key,
                   "algorithmic_bytes_per_launch": 16 * (1 << 26)}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
        print("traffic", t)
print("kernels:", list(out))

# ---- SASS evidence: what the shipped .so really contains ------------------------------------------------------------
so = os.path.join(ROOT, "unitysimpleraytracing_b200", "libusrt_b200.so")
if os.path.exists(so):
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    lines = ["# %s: SASS evidence from `cuobjdump -sass unitysimpleraytracing_b200/libusrt_b200.so` (sm_100a)\n" % rnd,
             "Per kernel: instruction count and how often the mnemonics the design relies on occur. No TMA / tensor-core",
             "mnemonics are expected: nothing on this path is a dense contraction or a tile copy (DESIGN.md section 4).\n",
             "| kernel | SASS instrs | LDG.E.128 | LDG.E.256 (ld.global.nc.v8) | LDG.E.64 | STG.E.64 | ATOMS.ADD | ATOMG.E.EXCH.128 | VOTE | SHFL | LDL/STL (spills, stack) | UTMALDG/UBLKCP/UTCMMA |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    want = ("k_onesweep", "k_histogram", "k_morton", "k_distribute_keys", "k_construct_tree", "k_construct_bvh", "k_trace_primary",
            "k_trace_rays", "k_peer_scatter_plan", "k_shade", "k_diffuse_rays")
    excerpts = []
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        k = short(dem.replace("(anonymous namespace)::", ""))
        if not any(w in k for w in want):
            continue
        body = [l for l in f.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        cnt = lambda pat: sum(1 for l in body if re.search(pat, l))
        lines.append("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |" % (
            k[:90], len(body), cnt(r"LDG\.E\.128"), cnt(r"LDG\.E(\.\w+)*\.256|LDG\.E\.ENL2\.256"), cnt(r"LDG\.E\.64"), cnt(r"STG\.E\.64"),
            cnt(r"ATOMS\.ADD"), cnt(r"ATOMG\.E\.EXCH(\.\w+)*\.128|EXCH.*128"), cnt(r"VOTE"), cnt(r"SHFL"), cnt(r"\b(LDL|STL)"),
            cnt(r"UTMALDG|UBLKCP|UTCMMA|UTCHMMA")))
        if "k_onesweep" in k and "256, 24, 3>, unsigned int, true, false, unsigned int, 2>" in k:
            first = next((i for i, l in enumerate(body) if "ATOMS.ADD" in l and "R" in l), 0)
            excerpts.append(("k_onesweep (interleaved records in and out): the ranking loop -- one returning ATOMS.ADD per key", body[max(first - 2, 0):first + 14]))
        if "k_trace_primary<false>" in k:
            first = next((i for i, l in enumerate(body) if ".256" in l), 0)
            excerpts.append(("k_trace_primary<strict>: the packed 64-byte node as two 256-bit loads", body[max(first - 1, 0):first + 4]))
        if "k_construct_bvh" in k:
            first = next((i for i, l in enumerate(body) if "EXCH" in l), 0)
            excerpts.append(("k_construct_bvh: cross-block merge through 128-bit atomic exchanges", body[max(first - 1, 0):first + 4]))
    lines.append("")
    for title, ex in excerpts:
        lines.append("## " + title + "\n\n```")
        lines += [l.rstrip()[:110] for l in ex]
        lines.append("```\n")
    open(os.path.join(P, rnd + "_sass.md"), "w").write("\n".join(lines) + "\n")
    print("sass summary written")
