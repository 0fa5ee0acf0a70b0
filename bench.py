#!/usr/bin/env python
"""bench.py -- headline benchmark of the LBVH build + ray-cast path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): 1,048,576-triangle scene (tessellated sphere + height field),
one STEP = full rebuild (Morton -> radix sort -> DistributeKeys -> tree -> refit) followed by a
1920x1080 primary-ray cast. Metric = Mrays/s = rays per step / device time per step.

Prints ONE JSON line (rank 0). The timed region is a K-step block (barrier + synchronize on both sides, CUDA
events, max over ranks) REPEATED until at least 0.5 s of device time has been measured; `ms_per_step` is the median
block, so a short --steps still gives a reproducible number. After the timed region the frame the LAST timed step
left on the device is checked: build buffers and (rank 0's camera) the frame against the digests generated from the
reference's own code (tests/golden/ref_digests.json), against the oracle's frame of the cpu_baseline leg, and at
N > 1 every rank re-traces its peers' cameras and bit-compares the records they stored into its frame over NVLink
(`verified`). Extra keys beside the contract: `sequential`, `stages_ms`, `sort` (2^26-pair key/value sort),
`roofline` (one onesweep radix pass, timed live with CUDA events), `rooflines`, `cpu_baseline`, and at N > 1
`strong` (the fixed 1080p frame sharded over the ranks) and `sort_dist` (2^28 / 2^30 pairs through the fused
NVLink bucket exchange).

N > 1: rays are sharded (BASELINE configs[4] shape, north_star (a)): every rank holds a replica of the BVH
(rebuilt each step, deterministic), traces its own 1080p sample of an N-spp frame, and the trace kernel stores
every hit record into every rank's frame over NVLink peer memory (USRT_BENCH_EXCHANGE=nccl: an NCCL all-gather
instead). Steps alternate over USRT_BENCH_CONTEXTS (default 3) contexts per GPU.
"""
import argparse
import hashlib
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 1920, 1080
SORT_LOG2 = 26
MIN_TIMED_MS = 500.0                 # the timed region is repeated until this much device time has been measured
# algorithmic bytes per unit (DESIGN.md "Kernels and rooflines")
BYTES = dict(morton=88, sort_pair=68, sort_pass_pair=16, distribute=8, tree=36,
             bvh=140)                # the reference refit's 140 B/tri; the packed traversal arrays K5 also
                                     # writes (64 + 96 B/tri) are reported as `extra_bytes_per_unit`, not counted
BVH_EXTRA_BYTES = 64 + 96
METRIC = "Mrays/s (1080p, 1M tris; step = full LBVH rebuild incl. sort + primary-ray cast)"
WORKLOAD = "configs[1]: 1,048,576-tri sphere+height-field, full rebuild + 1920x1080 primary rays"


def common_config(n_tris):
    """Identical in both arms (the driver compares them key by key)."""
    return {"workload": WORKLOAD, "triangles": int(n_tris), "frame": "%dx%d" % (W, H), "rays_per_step_per_gpu": W * H,
            "trace_mode": "strict (reference visiting order, no culling)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_config1():
    """Digests written by tests/golden/make_ref_golden.py from the REFERENCE'S OWN CODE (oracle/_ref)."""
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "ref_digests.json")))["config1_scene_1048576"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            # under load = samples in the upper half of the observed range
            hi = [x for x in sm if x >= 0.5 * max(sm)]
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def build_scene():
    from unitysimpleraytracing_b200 import meshes
    tris = meshes.scene_c2()
    return tris, meshes.SCENE_C2_CAMERA


def camera_for_rank(cam, rank):
    """Rank r traces sample r of the frame: the same camera nudged by a sub-pixel offset in x/y."""
    m = np.array(cam["cam_to_world"], np.float32).copy()
    m[0, 3] += np.float32(0.0131 * rank)
    m[1, 3] += np.float32(0.0071 * rank)
    return m


# =================================================================================================
# reference arm: the reference's algorithm (oracle C++ twin, pinned to oracle/_ref) on the host cores
# =================================================================================================
def cpu_step(ref_mod, tris, cam, rows, threads, keep_frame=False):
    """One CPU step: full single-threaded rebuild (the reference's CPU stages are serial loops; its GPU stages have no
    CPU form, the oracle runs them as scalar loops) + `rows` rows of the 1080p frame traced on `threads` host
    threads (rows == H: the whole frame, nothing extrapolated)."""
    tm = {}
    t0 = time.perf_counter()
    scene = ref_mod.Scene(tris, timings=tm)
    t_build = time.perf_counter() - t0
    y0 = (H - rows) // 2
    t1 = time.perf_counter()
    frame = scene.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], y0=y0, y1=y0 + rows, threads=threads,
                                counters=keep_frame)
    counts = None
    if keep_frame:                                  # oracle-counted work per ray, reported beside rays/s
        frame, counts = frame
    t_trace_rows = time.perf_counter() - t1
    t_frame = t_trace_rows * (H / rows)
    out = dict(step_s=t_build + t_frame, build_s=t_build, trace_rows_s=t_trace_rows, trace_frame_s=t_frame, stages=tm,
               rows=rows, y0=y0)
    if keep_frame:
        out["frame"] = frame
        out["scene"] = scene
        nr = float(rows * W)
        out["work_per_ray"] = {"node_box_tests": float(counts[0]) / nr, "triangle_box_tests": float(counts[1]) / nr,
                               "triangle_tests": float(counts[2]) / nr, "max_stack_depth": int(counts[3]),
                               "note": "counted by the CPU oracle on the same frame (strict = the reference's walk: no culling)"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                     # rank 0 alone runs the CPU arm
    from oracle import usrt_oracle as O
    O.build()
    tris, cam = build_scene()
    threads = os.cpu_count() or 1
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    budget_s = float(os.environ.get("USRT_BENCH_CPU_BUDGET_S", "150"))
    # first step (a warm-up step when W > 0) is always a whole frame: it also sizes the per-step sample
    first = cpu_step(O, tris, cam, H, threads)
    remaining = steps + max(warmup - 1, 0)
    rows = H
    if remaining * first["step_s"] > budget_s:     # bounded sample: fewer rows per step, every step still a full rebuild
        per_step = budget_s / remaining
        rows = int(H * max(per_step - first["build_s"], 0.02) / first["trace_frame_s"])
        rows = max(40, min(H, rows // 8 * 8))
    for _ in range(max(warmup - 1, 0)):
        cpu_step(O, tris, cam, rows, threads)
    samples = [cpu_step(O, tris, cam, rows, threads) for _ in range(steps)] if warmup > 0 else \
        [first] + [cpu_step(O, tris, cam, rows, threads) for _ in range(steps - 1)]
    step_s = statistics.mean(s["step_s"] for s in samples)
    value = W * H / step_s / 1e6
    build_ms = statistics.mean(s["build_s"] for s in samples) * 1e3
    sample = ("per step: full 1M-tri rebuild on 1 thread (%.0f ms) + %s traced on %d threads" %
              (build_ms, "the whole 1920x1080 frame" if rows == H else
               "%d of %d rows (frame time extrapolated by rows: %d steps of whole frames would exceed the %.0f s budget)"
               % (rows, H, steps + warmup, budget_s), threads))
    line = {
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32",
        "data": "synthetic",
        "config": common_config(len(tris)),
        "parallelism": "host threads (rank 0 only)",
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stages_ms": {"build": build_ms, "trace_frame": statistics.mean(s["trace_frame_s"] for s in samples) * 1e3},
        "note": "reference is GPU-only HLSL + Unity C#; this arm times its algorithm restated in C++ (oracle/usrt_oracle.cpp, "
                "pinned bit-for-bit to the reference's own text compiled by oracle/build_ref.sh), see DESIGN.md",
    }
    print(json.dumps(line), flush=True)


# =================================================================================================
# this repo's arm
# =================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    from unitysimpleraytracing_b200 import _lib, host

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # the hit-record all-gather runs beside the next step's kernels: fewer NCCL channels leave more SM
        # slots to the traversal (measured at 8 GPUs: 12 channels 11.1 Grays/s, NCCL default 10.7, 8 channels 10.1)
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "12")
        dist.init_process_group("nccl", device_id=dev)
    peak_gbs, peak_src = measured_peaks()
    K, Wm = max(args.steps, 1), max(args.warmup, 3)

    tris, cam = build_scene()
    n = len(tris)
    rays = W * H
    m = camera_for_rank(cam, rank)
    hit_dtype = np.dtype([("distance", "<f4"), ("triangleIndex", "<u4"), ("uv", "<f4", 2)])

    # D independent contexts (own scene buffers and stream) on this GPU, used round-robin: step i+1's rebuild --
    # a chain of short, latency-bound kernels -- runs beside step i's traversal. Every step still does all of
    # its work (full rebuild, full 1080p cast) on its own context; D=1 gives the strictly sequential form,
    # which is also measured below (`sequential`) because the per-stage times add up to that one.
    D = max(1, int(os.environ.get("USRT_BENCH_CONTEXTS", "3")))
    ctxs = [host.Context(n, device=local_rank) for _ in range(D)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(D)]
    torch.cuda.set_stream(streams[0])
    for c, st in zip(ctxs, streams):
        assert st.cuda_stream != 0
        c.set_stream(st.cuda_stream)       # the library's kernels and the torch events that time them share it
        c.upload_triangles(tris)
    ctx, stream = ctxs[0], streams[0]      # the single-context legs (stages, trace-only, sort) run on context 0

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    # N > 1: every rank must end each step holding all N frames. exchange "peer" (default): the trace kernel
    # stores each hit record into every rank's frame buffer over NVLink peer memory (usrt_set_hit_mirrors),
    # a one-element all-reduce is the "frame complete" fence; "nccl": trace, then all-gather.
    exchange = os.environ.get("USRT_BENCH_EXCHANGE", "peer") if world > 1 else None
    peers, gathered = None, None
    if exchange == "peer":
        from unitysimpleraytracing_b200 import dist as udist
        # CUDA IPC needs peer access between the GPUs and a shared PID/IPC namespace; if any rank cannot map its
        # peers, every rank switches to the NCCL all-gather (still this repo's kernels; only the exchange differs)
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            peers = [udist.PeerFrameExchange(c, rays, buffers=1) for c in ctxs]
        except Exception as e:                                          # noqa: BLE001
            sys.stderr.write("bench.py: peer-memory frame exchange unavailable on rank %d (%s)\n" % (rank, e))
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            for c in ctxs:
                c.set_hit_mirrors([])
            peers, exchange = None, "nccl"
        else:
            for px in peers:
                px.select(0)
    if exchange == "nccl":
        gathered = [torch.empty(world * rays * 4, dtype=torch.float32, device=dev) for _ in range(D)]

    # A block starts from an idle GPU (barrier + synchronize). Left alone, the D streams then start in lockstep -- D
    # rebuilds side by side, then D casts side by side, ... -- and may or may not drift apart within a short block (K = 20:
    # 0.920 or 0.955 ms/step from one block to the next). So the first rebuild of stream j waits for the first rebuild of
    # stream j-1: the streams enter the block staggered, the way a running pipeline is, and stay that way. Same work.
    stagger = os.environ.get("USRT_BENCH_STAGGER", "1") != "0" and D > 1
    first_rebuild_done = [None] * D

    def step(i):
        j = i % D
        c = ctxs[j]
        with torch.cuda.stream(streams[j]):
            if stagger and 0 < i < D and first_rebuild_done[i - 1] is not None:
                streams[j].wait_event(first_rebuild_done[i - 1])
            c.rebuild()
            if stagger and i < D - 1:
                first_rebuild_done[i] = torch.cuda.Event()
                first_rebuild_done[i].record(streams[j])
            c.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=False)
            if peers is not None:
                peers[j].fence()           # stream j resumes (step i+D) once every rank's frame i has landed everywhere
            elif world > 1:
                ptr, cnt = c.hits_device()
                dist.all_gather_into_tensor(gathered[j], _as_tensor(torch, ptr, cnt * 4, dev))

    def run_steps(k):
        """k steps; device time from before the first kernel to after the last one, over all D streams."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(streams[0])
        for st in streams[1:]:
            st.wait_event(ev0)
        for i in range(k):
            step(i)
        for st in streams[1:]:
            streams[0].wait_stream(st)
        ev1.record(streams[0])
        return ev0, ev1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_block():
        """EXACTLY K steps between barrier + synchronize on both sides; device time, max over ranks."""
        barrier()
        ev0, ev1 = run_steps(K)
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1))

    run_steps(max(Wm, D))
    barrier()

    # ---- timed region -------------------------------------------------------------------------
    launches0 = sum(c.kernel_launches for c in ctxs)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_wall0 = time.perf_counter()
    blocks = [timed_block()]
    # repeat the K-step block until >= MIN_TIMED_MS of device time is on record (every rank takes the same decision:
    # block times are already the max over ranks)
    reps = int(min(400, max(3, math.ceil(MIN_TIMED_MS / max(blocks[0], 1e-3)))))
    for _ in range(reps - 1):
        blocks.append(timed_block())
    t_wall = time.perf_counter() - t_wall0
    launches = sum(c.kernel_launches for c in ctxs) - launches0
    total_ms = statistics.median(blocks)
    ms_per_step = total_ms / K
    value = world * rays / (ms_per_step * 1e-3) / 1e6
    clocks = sampler.stop() if sampler else None

    # ---- verification of what the timed steps left on the device ------------------------------------------
    golden = golden_config1()
    last_j = (K - 1) % D
    checks = {}
    lc = ctxs[last_j]
    ptr, cnt = lc.hits_device()
    gpu_frame = _as_tensor(torch, ptr, cnt * 4, dev).cpu().numpy().view(hit_dtype).copy()      # the last timed frame of this rank
    if golden is not None:
        got = dict(sortedMortonCodes=lc.download(_lib.BUF_KEYS), sortedTriangleIndices=lc.download(_lib.BUF_TRIANGLE_INDEX),
                   triangleAABB=lc.download(_lib.BUF_TRIANGLE_AABB), internalNodes=lc.download(_lib.BUF_INTERNAL_NODES, n - 1),
                   leafNodes=lc.download(_lib.BUF_LEAF_NODES), bvhData=lc.download(_lib.BUF_BVH_DATA, n - 1))
        checks["build_buffers_vs_reference_digests"] = bool(golden.get("triangles") == sha(tris) and
                                                            all(sha(v) == golden[k] for k, v in got.items()))
        del got
        if rank == 0:
            checks["frame_vs_reference_digest"] = bool(sha(gpu_frame) == golden["primary_%dx%d" % (W, H)])
    if world > 1:
        # every rank re-traces every rank's camera on its own replica and bit-compares what the peers' kernels
        # stored into ITS frame buffers over NVLink (or what the all-gather delivered) during the last timed steps.
        # (timed_block ended with a barrier: nobody is still storing; the re-traces must not store to the peers either)
        for c in ctxs:
            c.set_hit_mirrors([])
        local = [ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], camera_for_rank(cam, r), download=True).tobytes()
                 for r in range(world)]
        ok_slots = True
        for j in range(D):
            buf = (peers[j].frame(0) if peers is not None else gathered[j]).cpu().numpy().view(hit_dtype).reshape(world, rays)
            for r in range(world):
                ok_slots = ok_slots and (buf[r].tobytes() == local[r])
        checks["peer_written_frames_bit_equal_local_retrace"] = bool(ok_slots)
        del local
    if peers is not None:
        barrier()
        for px in peers:
            px.close()                     # clears the mirrors: the legs below are single-GPU
        barrier()

    # ---- the same step strictly sequential on ONE context, L2 flushed between steps, one event pair per step
    def seq_step(i, e0=None, e1=None):
        flush.fill_(i & 0xFF)                                           # L2 flush, outside the event pair
        if e0 is not None:
            e0.record(stream)
        ctx.rebuild()
        ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=False)
        if e1 is not None:
            e1.record(stream)

    seq_steps = min(max(K, 50), 100)
    seq_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(seq_steps)]
    for i in range(3):
        seq_step(i)
    for i in range(seq_steps):
        seq_step(i, *seq_ev[i])
    torch.cuda.synchronize()
    seq_ms = statistics.median(a.elapsed_time(b) for a, b in seq_ev)
    # per-stage device times (events recorded inside the library on the same stream), on extra
    # steps of the same workload so that the queries' host syncs stay out of the timed regions
    stage_acc, sort_acc = {}, {}
    ctx.enable_stage_timing(True)       # (timed rebuilds are enqueued launch by launch, untimed ones replay a CUDA graph)
    for i in range(7):
        seq_step(i)
        if i < 2:
            continue
        for k, v in ctx.last_rebuild_ms().items():
            stage_acc.setdefault(k, []).append(v)
        for k, v in ctx.last_sort_ms().items():
            sort_acc.setdefault(k, []).append(v)
    ctx.enable_stage_timing(False)
    stages = {k: statistics.mean(v) for k, v in stage_acc.items()}
    stages["sort_kernels"] = {k: statistics.mean(v) for k, v in sort_acc.items()}
    # rebuild alone as the product runs it (CUDA-graph replay), L2 flushed
    rb_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
    for i, (a, b) in enumerate(rb_ev):
        flush.fill_(i & 0xFF); a.record(stream); ctx.rebuild(); b.record(stream)
    torch.cuda.synchronize()
    rebuild_graph_ms = statistics.median(a.elapsed_time(b) for a, b in rb_ev)
    # trace-only time, measured directly
    tr_steps = 50
    tr_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(tr_steps)]
    for i in range(tr_steps):
        flush.fill_(i & 0xFF)
        tr_ev[i][0].record(stream)
        ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=False)
        tr_ev[i][1].record(stream)
    torch.cuda.synchronize()
    trace_ms = statistics.median(a.elapsed_time(b) for a, b in tr_ev)
    stages["trace"] = trace_ms
    # the distance-culled traversal (usrt_set_trace_mode(1)): NOT the parity mode, reported separately -- it skips boxes that
    # start beyond the current closest hit, which cannot change a result in exact arithmetic but is not the reference's walk
    culled = None
    if rank == 0:
        strict_frame = ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=True)
        ctx.set_trace_mode(1)
        culled_frame = ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=True)
        cu_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
        for i, (a, b) in enumerate(cu_ev):
            flush.fill_(i & 0xFF); a.record(stream)
            ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=False)
            b.record(stream)
        torch.cuda.synchronize()
        cu_ms = statistics.median(a.elapsed_time(b) for a, b in cu_ev)
        differ = int((strict_frame.view(np.uint32).reshape(-1, 4) != culled_frame.view(np.uint32).reshape(-1, 4)).any(1).sum())
        culled = {"ms": cu_ms, "mrays_s": rays / (cu_ms * 1e-3) / 1e6, "records_differing_from_strict": differ,
                  "note": "non-parity mode, not part of `value`"}
        # mode 2: culled + the nearer of two hit internal children first, also non-parity
        ctx.set_trace_mode(2)
        near_frame = ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=True)
        nf_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
        for i, (a, b) in enumerate(nf_ev):
            flush.fill_(i & 0xFF); a.record(stream)
            ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=False)
            b.record(stream)
        torch.cuda.synchronize()
        ctx.set_trace_mode(0)
        nf_ms = statistics.median(a.elapsed_time(b) for a, b in nf_ev)
        culled["near_first"] = {"ms": nf_ms, "mrays_s": rays / (nf_ms * 1e-3) / 1e6, "records_differing_from_strict":
                                int((strict_frame.view(np.uint32).reshape(-1, 4) != near_frame.view(np.uint32).reshape(-1, 4)).any(1).sum())}
        del strict_frame, culled_frame, near_frame
    barrier()

    # ---- e2e: the same step through the C ABI with HOST buffers, on EVERY rank -------------------------------
    # Frames are pipelined over two contexts: frame i+1's upload (copy engine, PCIe) overlaps frame i's
    # kernels; every frame still pays its own 128 MiB upload and its own 33 MB of hit records, which the
    # trace kernel writes straight into the page-locked frame. The strictly synchronous form follows (rank 0).
    # Host buffers come from the library (usrt_host_alloc = cudaHostAlloc): what a host without a CUDA binding would use.
    # (A bare 128 MiB copy from such memory runs at 55.5 GB/s against 51.7 GB/s from a block pinned after the fact,
    # tools/h2d_probe.py; the pipelined frame time is 2.55 ms = 52.6 GB/s either way.)
    E = ctxs[:2] if D >= 2 else [ctxs[0], host.Context(n, device=local_rank)]
    pinned_tris = ctx.host_alloc(tris.nbytes)
    pinned_tris[:] = tris.view(np.uint8).reshape(-1)
    tris_h = pinned_tris.view(tris.dtype)
    pinned_hits = [ctx.host_alloc(rays * 16) for _ in E]
    hits_h = [p.view(hit_dtype) for p in pinned_hits]
    e2e_steps = max(8, min(K, 40))

    def e2e_run(k):
        for i in range(k):
            c = E[i % 2]
            c.sync()                                  # frame i-2 is complete: hits_h[i % 2] has been handed over
            c.upload_triangles_async(tris_h)          # H2D of the step's input
            c.rebuild()
            c.trace_primary_async(W, H, cam["near"], cam["tan_half_fov"], m, hits_h[i % 2])   # D2H of the result
        for c in E:
            c.sync()

    e2e_run(4)
    e2e_blocks = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        torch.cuda.synchronize()
        e2e_blocks.append(max_over_ranks((time.perf_counter() - t0) / e2e_steps * 1e3))
    e2e_ms = statistics.median(e2e_blocks)
    checks["e2e_frame_equals_device_frame"] = bool(hits_h[1].tobytes() == gpu_frame.tobytes())

    # the same pipeline with the ADDITIVE positions-only upload (usrt_upload_positions_async: 48 of the 128 bytes of every
    # Triangle, all the build and the traversal read) -- reported beside the full-struct number, not instead of it
    pinned_pos = ctx.host_alloc(n * 48)
    pos_h = pinned_pos.view(np.float32).reshape(n, 12)
    pos_h[:] = tris.view(np.float32).reshape(n, 32)[:, :12]

    def e2e_pos_run(k):
        for i in range(k):
            c = E[i % 2]
            c.sync()
            c.upload_positions(pos_h, pinned=True)
            c.rebuild()
            c.trace_primary_async(W, H, cam["near"], cam["tan_half_fov"], m, hits_h[i % 2])
        for c in E:
            c.sync()

    e2e_pos_run(4)
    e2e_pos_blocks = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        e2e_pos_run(e2e_steps)
        torch.cuda.synchronize()
        e2e_pos_blocks.append(max_over_ranks((time.perf_counter() - t0) / e2e_steps * 1e3))
    e2e_pos_ms = statistics.median(e2e_pos_blocks)
    if os.environ.get("USRT_BENCH_DEBUG"):
        sys.stderr.write("e2e blocks %s pos blocks %s\n" % (e2e_blocks, e2e_pos_blocks))
    checks["e2e_positions_only_frame_equals_device_frame"] = bool(hits_h[1].tobytes() == gpu_frame.tobytes() and
                                                                   hits_h[0].tobytes() == gpu_frame.tobytes())
    for c in E:
        c.upload_triangles(tris_h)         # back to the full-struct source for the legs below
    ctx.host_free(pinned_pos)
    del pinned_pos, pos_h

    line = None
    e2e_sync_ms = None
    if rank == 0:
        def e2e_sync_step():
            ctx.upload_triangles(tris_h)                  # synchronous H2D
            ctx.rebuild()
            ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=True, out=hits_h[0])

        for _ in range(2):
            e2e_sync_step()
        checks["e2e_synchronous_equals_pipelined"] = bool(hits_h[1].tobytes() == hits_h[0].tobytes())
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_sync_step()
        torch.cuda.synchronize()
        e2e_sync_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    if D < 2:
        E[1].close()
    del tris_h, hits_h
    for p in [pinned_tris] + pinned_hits:
        ctx.host_free(p)
    del pinned_tris, pinned_hits
    barrier()

    # ---- N > 1 extras: strong scaling of the fixed frame, multi-GPU sort -------------------------------------
    strong = sort_dist = None
    if world > 1:
        strong = strong_scaling_leg(torch, dist, ctx, stream, cam, world, rank, dev, K, barrier, max_over_ranks, hit_dtype)
        sort_dist = dist_sort_leg(torch, dist, ctx, world, rank, dev, barrier, max_over_ranks)

    if rank == 0:
        # ---- sort leg: 2^26 (key, value) pairs resident in HBM ------------------------------------
        ns = 1 << SORT_LOG2
        g = torch.Generator(device=dev); g.manual_seed(0x5EED)
        keys0 = torch.randint(-2 ** 31, 2 ** 31 - 1, (ns,), dtype=torch.int32, device=dev, generator=g)
        vals0 = torch.arange(ns, dtype=torch.int32, device=dev)
        keys = torch.empty_like(keys0); vals = torch.empty_like(vals0)
        sort_steps = 10
        sort_ms, pass_ms, hist_ms = [], [], []
        ctx.enable_stage_timing(True)       # per-kernel events inside the library (same stream)
        for i in range(3 + sort_steps):
            keys.copy_(keys0); vals.copy_(vals0)
            ctx.sort_pairs_device(keys.data_ptr(), vals.data_ptr(), ns)
            if i >= 3:
                t = ctx.last_sort_ms()
                sort_ms.append(t["total"]); hist_ms.append(t["histogram"])
                pass_ms.append(statistics.mean([t["pass0"], t["pass8"], t["pass16"], t["pass24"]]))
        torch.cuda.synchronize()
        ctx.enable_stage_timing(False)
        # verified on the device: ascending keys, stable (values ascend inside runs of equal keys), a permutation
        ku = keys.to(torch.int64) & 0xFFFFFFFF
        asc = bool((ku[1:] >= ku[:-1]).all().item())
        stable = bool(((ku[1:] != ku[:-1]) | (vals[1:] > vals[:-1])).all().item())
        perm = bool((keys0[vals.to(torch.int64)] == keys).all().item())
        checks["sort_2p26_sorted_stable_permutation"] = asc and stable and perm
        del keys0, vals0, keys, vals, ku
        s_ms, p_ms = statistics.median(sort_ms), statistics.median(pass_ms)
        sort_gbs = BYTES["sort_pair"] * ns / (s_ms * 1e-3) / 1e9
        pass_gbs = BYTES["sort_pass_pair"] * ns / (p_ms * 1e-3) / 1e9

        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("k_onesweep_2p26_bytes_per_launch")
            except Exception:
                traffic = None

        def roof(name, bytes_per_unit, units, ms, **extra):
            gbs = bytes_per_unit * units / (ms * 1e-3) / 1e9
            d = {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s",
                 "frac": gbs / peak_gbs, "bytes_per_unit": bytes_per_unit, "units": units, "ms": ms}
            d.update(extra)
            return d

        rooflines = [
            roof("k_morton (K1, 1M tris)", BYTES["morton"], n, stages["morton"]),
            roof("sort total (K2: histogram + 4 passes, 1M pairs)", BYTES["sort_pair"], n, stages["sort"]),
            roof("k_distribute_keys (K3, 1M keys)", BYTES["distribute"], n, stages["distribute"]),
            roof("k_construct_tree (K4, 1M tris)", BYTES["tree"], n, stages["tree"]),
            roof("k_construct_bvh (K5, 1M tris)", BYTES["bvh"], n, stages["bvh"], extra_bytes_per_unit=BVH_EXTRA_BYTES,
                 note="140 B/tri = the reference refit (SURVEY 8d); K5 also writes the packed traversal arrays (extra)"),
            roof("full rebuild (K1..K5, 1M tris, CUDA-graph replay)", 340, n, rebuild_graph_ms),
            roof("sort total (K2, 2^26 pairs)", BYTES["sort_pair"], ns, s_ms),
        ]

        # ---- CPU baseline: the oracle on a bounded sample (3 whole steps), host cores of this box ----------
        from oracle import usrt_oracle as O
        O.build()
        threads = os.cpu_count() or 1
        cs = [cpu_step(O, tris, cam, H, threads, keep_frame=(i == 0)) for i in range(3)]
        cpu_value = rays / statistics.mean(s["step_s"] for s in cs) / 1e6
        # free check: the oracle's whole frame (rank 0's camera) against the frame the timed run left on the GPU
        checks["frame_vs_oracle_whole_frame"] = bool(cs[0]["frame"].tobytes() == gpu_frame.tobytes())
        work_per_ray = cs[0]["work_per_ray"]
        sc = cs[0]["scene"]
        checks["build_buffers_vs_oracle"] = bool(
            np.array_equal(ctx.download(_lib.BUF_KEYS), sc.sortedMortonCodes) and
            ctx.download(_lib.BUF_BVH_DATA, n - 1).tobytes() == sc.bvhData[:n - 1].tobytes() and
            ctx.download(_lib.BUF_INTERNAL_NODES, n - 1).tobytes() == sc.internalNodes[:n - 1].tobytes())
        del sc

    if world > 1:
        flag = torch.tensor([1 if all(checks.values()) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verified = bool(flag.item())
    else:
        verified = all(checks.values())

    if rank == 0:
        seq_value = rays / (seq_ms * 1e-3) / 1e6
        line = {
            "metric": METRIC + "; value = %d contexts per GPU pipelined, sequential.value = one context, L2 flushed" % D,
            "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+u32", "data": "synthetic",
            "config": common_config(n),
            "parallelism": "1 GPU" if world == 1 else "ray-sharded x%d, replicated BVH, %s" % (
                world, "hit records stored to every rank by the trace kernel over NVLink peer memory" if exchange == "peer"
                else "all-gather of hit records overlapped"),
            "run": {"contexts": D,
                    "l2": ("no flush in the timed region: the %d contexts used round-robin hold 357 MB of scene buffers each, "
                           "far beyond the 126 MB L2" % D) if D > 1 else
                          "not flushed; the scene buffers (357 MB) exceed the 126 MB L2 (see `sequential` for the flushed form)",
                    "timed_blocks": len(blocks), "block_ms_min": min(blocks), "block_ms_median": total_ms, "block_ms_max": max(blocks),
                    "timed_device_ms_total": sum(blocks),
                    "note": "each block = exactly `steps` steps between barrier+synchronize, CUDA events, max over ranks; "
                            "ms_per_step = median block / steps"},
            "verified": verified, "checks": checks,
            "sequential": {"ms_per_step": seq_ms, "value": seq_value, "steps": seq_steps,
                           "note": "same step on ONE context, L2 flushed (512 MiB write) between steps, one CUDA event pair per "
                                   "step on this rank (median); stages_ms add up to this"},
            "stages_ms": stages,
            "trace_mrays_s": rays / (trace_ms * 1e-3) / 1e6, "trace_work_per_ray": work_per_ray, "trace_culled": culled,
            "build_ms": rebuild_graph_ms, "build_ms_launch_by_launch": stages["total"],
            "sort": {"pairs": ns, "ms": s_ms, "mkeys_s": ns / (s_ms * 1e-3) / 1e6, "achieved_gbs": sort_gbs,
                     "frac_of_measured_peak": sort_gbs / peak_gbs, "frac_of_8tbs": sort_gbs / 8000.0,
                     "histogram_ms": statistics.median(hist_ms), "pass_ms": p_ms,
                     "cub_calibration": "cub::DeviceRadixSort::SortPairs on the same box, 2^26 pairs: 1.950 ms "
                                        "(tools/micro/cub_sort_calib.cu, profiles/r02_cub_calibration.txt)"},
            "roofline": {"kernel": "k_onesweep (one 8-bit radix pass over 2^26 key/value pairs)", "bound": "hbm",
                         "achieved": pass_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": pass_gbs / peak_gbs,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES["sort_pass_pair"] * ns, "ms_per_launch": p_ms},
            "rooflines": rooflines,
            "cpu_baseline": {"value": cpu_value, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": "3 whole steps: full 1M-tri rebuild on 1 thread (%.0f ms) + the whole 1920x1080 frame traced on "
                                       "%d threads (%.0f ms); nothing extrapolated" % (
                                           statistics.mean(s["build_s"] for s in cs) * 1e3, threads,
                                           statistics.mean(s["trace_frame_s"] for s in cs) * 1e3)},
            "e2e": {"value": world * rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(n * 128), "d2h_bytes_per_step": int(rays * 16), "ranks": world,
                    "synchronous_ms_per_step": e2e_sync_ms, "synchronous_value": rays / (e2e_sync_ms * 1e-3) / 1e6,
                    "positions_only": {"value": world * rays / (e2e_pos_ms * 1e-3) / 1e6, "ms_per_step": e2e_pos_ms,
                                       "h2d_bytes_per_step": int(n * 48), "d2h_bytes_per_step": int(rays * 16),
                                       "note": "additive API usrt_upload_positions_async: the 48 bytes per triangle the build and "
                                               "the traversal read; same frames, bit for bit"},
                    "note": "per frame and per rank: usrt_upload_triangles_async(pinned host) + usrt_rebuild + "
                            "usrt_trace_primary_async(pinned host frame), frames alternate over two contexts so the next upload "
                            "overlaps the kernels; every rank runs it at once, value = all ranks' rays / the slowest rank's wall "
                            "time (median of 3 blocks); bytes are per rank; synchronous_* = rank 0 alone, one context, every call blocking"},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / (len(blocks) * K),
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
        }
        if strong is not None:
            line["strong"] = strong
        if sort_dist is not None:
            line["sort_dist"] = sort_dist
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for c in ctxs:
        c.close()
    if line is not None:
        print(json.dumps(line), flush=True)


def strong_scaling_leg(torch, dist, ctx, stream, cam, world, rank, dev, K, barrier, max_over_ranks, hit_dtype):
    """STRONG scaling (north_star (a) as stated): ONE fixed 1920x1080 frame, its row blocks interleaved over the ranks,
    every rank rebuilds its replica and traces its share; the trace kernel stores the records into every rank's frame
    over NVLink and a one-element all-reduce fences the frame. Also times that fence alone."""
    from unitysimpleraytracing_b200 import dist as udist, meshes  # noqa: F401
    block_rows = 8
    per, local_rows = udist.shard_layout(H, world, block_rows)
    slot = local_rows * W
    m0 = camera_for_rank(cam, 0)
    out = {"frame": "%dx%d" % (W, H), "block_rows": block_rows}
    try:
        px = udist.PeerFrameExchange(ctx, slot, buffers=1)
    except Exception as e:                                              # noqa: BLE001
        return {"unavailable": "peer memory: %s" % e}
    px.select(0, include_own=False)

    def one():
        ctx.rebuild()
        ctx.trace_primary_sharded(W, H, cam["near"], cam["tan_half_fov"], m0, block_rows, rank, world, dev_out=px.slot_ptr(0))
        px.fence()

    with torch.cuda.stream(stream):
        for _ in range(5):
            one()
        blocks = []
        for _ in range(5):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(K):
                one()
            e1.record(stream)
            barrier()
            blocks.append(max_over_ranks(e0.elapsed_time(e1)) / K)
        # the fence alone
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(50):
            px.fence()
        e1.record(stream)
        barrier()
        fence_ms = max_over_ranks(e0.elapsed_time(e1)) / 50
        # correctness: the assembled frame on this rank == a local trace of the whole frame
        g = px.frame(0).cpu().numpy().view(hit_dtype).reshape(world, slot)
        frame = udist.assemble_frame(g, W, H, world, block_rows)
        ctx.set_hit_mirrors([])
        want = ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m0, download=True)
        ok = torch.tensor([1 if frame.tobytes() == want.tobytes() else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    px.close()
    ms = statistics.median(blocks)
    out.update({"ms_per_step": ms, "value": W * H / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "scaling": "strong",
                "fence_all_reduce_ms": fence_ms, "verified": bool(ok.item()),
                "note": "step = full rebuild on every rank (replicas; the build does not shard) + this rank's 1/N of the frame + "
                        "fence; the rebuild (~0.2 ms) is the serial part"})
    return out


def dist_sort_leg(torch, dist, ctx, world, rank, dev, barrier, max_over_ranks):
    """BASELINE configs[2] beyond one GPU: 2^28 and 2^30 uint32 pairs split evenly over the ranks, sorted through
    PeerSortExchange (top-byte buckets scattered straight into the owners' buffers over NVLink by the partition kernel,
    then a local 4-pass sort). Verified on the device: globally ascending, stable, and the same multiset."""
    from unitysimpleraytracing_b200 import dist as udist
    res = {}
    ctx.use_torch_stream()
    for lg in (28, 30):
        total = 1 << lg
        n = total // world
        try:
            g = torch.Generator(device=dev); g.manual_seed(0xD157 + rank)
            keys = torch.randint(-2 ** 31, 2 ** 31 - 1, (n,), dtype=torch.int32, device=dev, generator=g)
            vals = torch.arange(rank * n, (rank + 1) * n, dtype=torch.int64, device=dev).to(torch.int32)   # global index (wraps at 2^31: compared as uint32)
            cap = int(n * 1.05) + 4096
            px = udist.PeerSortExchange(ctx, cap)
            in_sum = (keys.to(torch.int64) & 0xFFFFFFFF).sum() * 1 + (vals.to(torch.int64) & 0xFFFFFFFF).sum() * 3
            times = []
            for it in range(4):
                barrier()
                t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
                t0.record()
                rk, rv = px.sort(keys, vals)
                t1.record()
                torch.cuda.synchronize()
                if it >= 1:
                    times.append(max_over_ranks(t0.elapsed_time(t1)))
            ku = rk.to(torch.int64) & 0xFFFFFFFF
            vu = rv.to(torch.int64) & 0xFFFFFFFF
            ok = bool((ku[1:] >= ku[:-1]).all().item()) and bool(((ku[1:] != ku[:-1]) | (vu[1:] > vu[:-1])).all().item())
            # chunk boundaries: my first key >= the previous rank's last key (stability across ranks follows from the
            # source-rank-major landing order, which the value check inside equal-key runs covers within a chunk)
            edge = torch.tensor([int(ku[0].item()) if len(ku) else 0, int(ku[-1].item()) if len(ku) else 0], dtype=torch.int64, device=dev)
            edges = torch.empty(2 * world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(edges, edge)
            e = edges.cpu().numpy().reshape(world, 2)
            ok = ok and all(e[r][0] >= e[r - 1][1] for r in range(1, world))
            out_sum = ku.sum() * 1 + vu.sum() * 3
            sums = torch.stack([in_sum, out_sum, torch.tensor(n, device=dev), torch.tensor(len(ku), device=dev)]).to(torch.int64)
            dist.all_reduce(sums)
            ok = ok and int(sums[0].item()) == int(sums[1].item()) and int(sums[2].item()) == int(sums[3].item()) == total
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ms = statistics.median(times)
            res["2^%d" % lg] = {"pairs": total, "ms": ms, "gpairs_s": total / (ms * 1e-3) / 1e9,
                                "nvlink_bytes_total": int(8 * total * (world - 1) / world),
                                "verified": bool(flag.item())}
            del keys, vals, rk, rv, ku, vu
            px.close()
            torch.cuda.empty_cache()
        except Exception as e:                                          # noqa: BLE001
            res["2^%d" % lg] = {"unavailable": str(e)[:200]}
            barrier()
    res["note"] = ("pairs split evenly over the ranks; ms = device time of the whole distributed sort (histogram, counter all-gather, "
                   "device-side plan, fused scatter over NVLink, fence, local 4-pass sort), max over ranks, median of 3; "
                   "nvlink_bytes_total = 8 B x pairs x (N-1)/N expected for uniform keys")
    return res


def _as_tensor(torch, ptr, numel_f32, dev):
    """Zero-copy float32 view of library-owned device memory via __cuda_array_interface__."""
    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (int(numel_f32),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(w, device=dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
