#!/usr/bin/env python
"""bench.py -- throughput of the polar front end (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

A "step" is one pass of the hot path over one batch of synthetic frames per GPU:
Cartesian sweeps -> (rho, phi, z) transform -> polar hard voxelization -> mean VFE -> (pillar
grids) scatter into the dense polar BEV canvas.  Default workload = BASELINE.json configs[1]:
nuScenes 10-sweep frames, NUSC-PILLAR grid, batch 8 per GPU.  Frames are independent, so N GPUs
each take their own batch (weak scaling, no collective on the path); value = all points / max-rank
time.  Prints ONE JSON line on rank 0 (schema in the task contract), with
  value     device-resident inputs, kernels only (CUDA events, max over ranks)
  e2e       same metric through PolarFrontEnd.forward_host: pinned host buffers, H2D + all D2H
            inside the timed region
  roofline  dominant kernel: algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the oracle C port of the reference's CPU path on this box's host cores
`--impl reference` times that CPU port alone (all host threads) on the same config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from partner_b200 import synth  # noqa: E402

WORKLOADS = {
    # name: (grid tag, frame kind, generator kwargs, frames per GPU, config id, has canvas)
    "nusc_pillar_mean_canvas_b8": ("NUSC-PILLAR", "nusc", {}, 8, 2, True),
    "nusc_pillar_mean_canvas_b2": ("NUSC-PILLAR", "nusc", {}, 2, 2, True),      # experiments: L2-sized batches
    "nusc_pillar_mean_canvas_b4": ("NUSC-PILLAR", "nusc", {}, 4, 2, True),
    "waymo_partner_mean_b16": ("WAYMO-PARTNER", "waymo", dict(nsweeps=1, time_column=True), 16, 4, False),
    "waymo3_partner_mean_b8": ("WAYMO-PARTNER", "waymo", dict(nsweeps=3, time_column=True), 8, 5, False),
}
METRIC = "polar_voxelize_vfe_scatter_throughput"
UNIT = "Mpoints/s"
# stage names of pv_profile_mean_canvas per pipeline (pv_profile_pipeline: 1 list-based, 2 list-free)
STAGES_BY_PIPELINE = {1: ["bin_insert", "cell_flags", "scan", "place", "emit"],
                      2: ["insert", "cells", "scan", "finalize", "heavy"]}
STAGES = STAGES_BY_PIPELINE[2]
N_SETS = 4          # rotating input sets so a step never finds its inputs in the 126 MB L2


STAGE_KERNELS = {"insert": ("kf_insert",), "cells": ("kf_cells",), "scan": ("kf_scan",),
                 "heavy": ("kf_heavy_points", "kf_heavy_cells"), "finalize": ("kf_finalize",),
                 "bin_insert": ("k_bin_insert",), "cell_flags": ("k_cell_flags",), "place": ("k_place",),
                 "emit": ("k_emit",)}
LAUNCHES_PER_STEP = {1: 7, 2: 5}      # kernels per step: list-based (incl. the scan-tile layout), list-free


def ncu_traffic(stage):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the stage's kernels from
    the newest committed ncu --set full capture under profiles/, or None."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_kernels.json")))
    if not files:
        return None, None
    try:
        with open(files[-1]) as f:
            prof = json.load(f)
        tot = 0.0
        for k in prof["kernels"]:
            if any(name in k["kernel"] for name in STAGE_KERNELS[stage]):
                tot += k.get("dram_traffic_bytes", 0.0)
        return (tot or None), os.path.relpath(files[-1], ROOT)
    except Exception:
        return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_pass(frames, grid, threads):
    """The reference's CPU path (oracle C port): per frame transform_points + points_to_voxel
    (dense map included) + mean VFE + scatter, one frame per worker like its DataLoader workers."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    g = synth.GRIDS[grid]
    ref = oracle.VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    gs = ref.grid_size
    pillar = int(gs[2]) == 1

    def one(f):
        polar = oracle.transform_points(f)
        vox, coor, num, _, _ = ref.generate(polar)
        feats = oracle.vfe_mean(vox, num)
        if pillar:
            c4 = np.pad(coor, ((0, 0), (1, 0)))
            oracle.scatter(feats, c4, 1, [int(gs[0]), int(gs[1]), 1])
        return vox.shape[0]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:      # ctypes releases the GIL
        list(ex.map(one, frames))
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    grid, kind, kw, per_gpu, cfg_id, _ = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    frames = synth.make_batch(kind, cfg_id, per_gpu, **kw)
    npts = sum(f.shape[0] for f in frames)
    import oracle
    oracle.lib()
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_pass(frames[:max(1, min(len(frames), threads))], grid, threads)
    times = [cpu_reference_pass(frames, grid, threads) for _ in range(args.steps)]
    dt = sum(times)
    val = npts * args.steps / dt / 1e6
    sample = "%d steps x %d frames (%d points), one frame per thread" % (args.steps, len(frames), npts)
    _emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_s": len(frames) * args.steps / dt,
        "config": {"workload": args.workload, "grid": grid, "frames_per_step": len(frames),
                   "points_per_step": npts, "note": "reference CPU path (oracle C port of numba/numpy/torch code)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


_JSON_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner when NCCL_DEBUG is set on the box), so the real stdout is kept for the result line and
    file descriptor 1 is pointed at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(result):
    _JSON_OUT.write(json.dumps(result) + "\n")
    _JSON_OUT.flush()


def main():
    global N_SETS
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="nusc_pillar_mean_canvas_b8", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--streams", type=int, default=4,
                    help="independent batches in flight on separate CUDA streams (1 = strictly serial steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sets", type=int, default=4, help="rotating input sets (each with its own workspace/graph)")
    ap.add_argument("--launch", default="graph", choices=["graph", "python"],
                    help="graph: multi-step CUDA graphs (no host call between steps); python: one replay per step from Python")
    ap.add_argument("--graph-steps", type=int, default=20, help="steps captured per CUDA graph")
    ap.add_argument("--ws-per-stream", type=int, default=0, help="1: one workspace per stream instead of per input set")
    ap.add_argument("--lib", default=None, help="development aid: time another build of the C-ABI library (A/B runs)")
    args = ap.parse_args()
    N_SETS = max(1, args.sets)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 20          # bounded sample: a CPU step takes ~0.1-1 s
        return run_reference(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    from partner_b200 import PolarFrontEnd, _lib
    if args.lib:
        _lib.SO_PATH = os.path.abspath(args.lib)
    from partner_b200 import functional as F
    from partner_b200._lib import ptr, current_stream
    import ctypes
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    grid, kind, kw, per_gpu, cfg_id, has_canvas = WORKLOADS[args.workload]
    g = synth.GRIDS[grid]
    fe = PolarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"], cartesian=True, device=dev)

    # ---- synthetic inputs: N_SETS different batches per rank, resident in HBM --------------
    sets = []
    for s in range(N_SETS):
        frames = synth.make_batch(kind, cfg_id, per_gpu, first_frame=(rank * N_SETS + s) * per_gpu, **kw)
        sizes = [f.shape[0] for f in frames]
        off = np.zeros(per_gpu + 1, np.int32)
        np.cumsum(sizes, out=off[1:])
        sets.append(dict(frames=frames, sizes=sizes, n=int(off[-1]), cap=max(sizes),
                         h_points=torch.from_numpy(np.concatenate(frames)).pin_memory(),
                         h_off=torch.from_numpy(off).pin_memory()))
    c_in = sets[0]["frames"][0].shape[1]
    for s in sets:
        s["d_points"] = s["h_points"].to(dev)
        s["d_off"] = s["h_off"].to(dev)
    cap_all = max(s["cap"] for s in sets)
    in_bytes = sum(s["n"] for s in sets) * c_in * 4

    # ---- device-resident path: one CUDA graph per input set ------------------------------------
    n_streams = max(1, min(args.streams, N_SETS))
    runners = []
    fes = []
    for k, s in enumerate(sets):
        # one workspace per STREAM, not per input set: a batch in flight owns its workspace, and a workspace
        # that comes round again every n_streams steps stays in the 126 MB L2 (inputs and outputs still rotate)
        fe_s = PolarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"], cartesian=True,
                             device=dev, workspace_tag=k % n_streams if args.ws_per_stream else k)
        fes.append(fe_s)
        if args.no_graph:
            out = fe_s.forward_device(s["d_points"], s["d_off"], per_gpu, cap_all)
            runners.append((lambda s=s, out=out, f=fe_s: f.forward_device(s["d_points"], s["d_off"], per_gpu, cap_all, out=out), out))
        else:
            out = fe_s.capture(s["d_points"], s["d_off"], per_gpu, cap_all)
            runners.append((fe_s.replay, out))
    torch.cuda.synchronize()
    counts = [int(out.voxel_counts.sum().item()) for _, out in runners]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, n_streams):
        """K steps round-robin over n_streams CUDA streams (independent batches overlap), timed with
        CUDA events on the current stream, which every side stream forks from and joins into."""
        cur = torch.cuda.current_stream(dev)
        streams = [torch.cuda.Stream(dev) for _ in range(n_streams)] if n_streams > 1 else [cur]
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for st in streams:
            if st is not cur:
                st.wait_event(e0)
        for k in range(steps):
            with torch.cuda.stream(streams[k % n_streams]):
                fn(k)
        for st in streams:
            if st is not cur:
                ev = torch.cuda.Event()
                ev.record(st)
                cur.wait_event(ev)
        e1.record(cur)
        barrier()
        return e0.elapsed_time(e1)

    step = lambda k: runners[k % N_SETS][0]()        # noqa: E731

    def capture_rounds(rounds, branches_n):
        """ONE CUDA graph holding rounds * N_SETS steps: set k runs on branch k % branches_n (fork / join
        inside the capture), so a replay puts that many independent batches in flight without a
        single host call in between -- a Python `with stream: graph.replay()` per step costs more
        host time (~70 us) than the step takes on the device."""
        gr = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(dev)
        side = [torch.cuda.Stream(dev) for _ in range(branches_n - 1)]
        branches = [cap] + side
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.graph(gr, stream=cap):
            fork = torch.cuda.Event()
            fork.record(cap)
            for st in side:
                st.wait_event(fork)
            for _ in range(rounds):
                for k, s_ in enumerate(sets):
                    with torch.cuda.stream(branches[k % branches_n]):
                        fes[k].forward_device(s_["d_points"], s_["d_off"], per_gpu, cap_all, out=runners[k][1])
            for st in side:
                ev = torch.cuda.Event()
                ev.record(st)
                cap.wait_event(ev)
        return gr

    def timed_graphs(gr, per_replay, steps):
        """steps = q * per_replay + r: q replays of the multi-step graph, then r single-step graphs."""
        cur = torch.cuda.current_stream(dev)
        q, r = divmod(steps, per_replay)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for _ in range(q):
            gr.replay()
        for k in range(r):
            step(k)
        e1.record(cur)
        barrier()
        return e0.elapsed_time(e1)

    if args.no_graph or args.launch == "python":
        timed(step, args.warmup, n_streams)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms_total = timed(step, args.steps, n_streams)
        ms_single = timed(step, args.steps, 1) if n_streams > 1 else ms_total
        steps_per_graph = 1
    else:
        rounds = max(1, args.graph_steps // N_SETS)
        steps_per_graph = rounds * N_SETS
        g_multi = capture_rounds(rounds, n_streams)
        g_serial = capture_rounds(rounds, 1) if n_streams > 1 else g_multi
        timed_graphs(g_multi, steps_per_graph, max(args.warmup, steps_per_graph))
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms_total = timed_graphs(g_multi, steps_per_graph, args.steps)
        ms_single = timed_graphs(g_serial, steps_per_graph, args.steps) if n_streams > 1 else ms_total
    pts_done = sum(sets[k % N_SETS]["n"] for k in range(args.steps))

    # ---- end-to-end path: pinned host buffers, H2D + D2H inside the timed region ----------
    ios = []
    for s, f in zip(sets, fes):
        io = f.make_host_io(s["n"], per_gpu, c_in)
        io["h_points"].copy_(s["h_points"])
        io["h_offsets"].copy_(s["h_off"])
        ios.append(io)
    e2e_steps = max(4, min(args.steps, 48))
    traffic = [0, 0]

    def e2e_step(k):
        a, b = fes[k % N_SETS].forward_host_async(ios[k % N_SETS], cap_all)
        traffic[0] += a
        traffic[1] += b
    timed(e2e_step, 4, n_streams)
    traffic = [0, 0]
    t0 = time.perf_counter()
    e2e_ms = timed(e2e_step, e2e_steps, n_streams)
    e2e_wall = (time.perf_counter() - t0) * 1e3
    h2d, d2h = traffic
    # fully synchronous variant (one step at a time, slices copied after reading the counts)
    t0 = time.perf_counter()
    for k in range(8):
        fes[k % N_SETS].forward_host(ios[k % N_SETS], cap_all)
    e2e_sync_ms = (time.perf_counter() - t0) * 1e3 / 8
    e2e_pts = sum(sets[k % N_SETS]["n"] for k in range(e2e_steps))
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-stage CUDA-event times (same inputs, rotating) -> dominant kernel roofline ----
    lib = _lib.load()
    pipeline = lib.pv_profile_pipeline(fe.cfg)
    STAGES = STAGES_BY_PIPELINE[pipeline]
    stage = np.zeros(len(STAGES), np.float64)
    reps = 5
    for k in range(N_SETS):
        s, (_, out) = sets[k], runners[k]
        ms = (ctypes.c_float * len(STAGES))()
        ws = out.ws
        F.check(lib.pv_profile_mean_canvas(fe.cfg, ptr(s["d_points"]), ptr(s["d_off"]), per_gpu, s["n"], c_in, 1,
                                           out.n_cap, out.f_cap, ptr(ws), ws.numel(), ptr(out.coors), ptr(out.num_points),
                                           ptr(out.voxel_counts), ptr(out.mean_feats), ptr(out.canvas),
                                           current_stream(dev), reps, ms), "pv_profile_mean_canvas")
        stage += np.array(list(ms), np.float64) / N_SETS
    torch.cuda.synchronize()

    # ---- reduce over ranks: max time, sum of work ------------------------------------------
    vec = torch.tensor([ms_total, e2e_ms, float(pts_done), float(e2e_pts), ms_single], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = vec.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vec.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_ms, ms_single = float(mx[0]), float(mx[1]), float(mx[4])
        pts_all, e2e_pts_all = float(sm[2]), float(sm[3])
    else:
        pts_all, e2e_pts_all = float(pts_done), float(e2e_pts)

    if rank == 0:
        peak, peak_src = peaks()
        C = c_in + 2
        n_avg = float(np.mean([s["n"] for s in sets]))
        m_avg = float(np.mean(counts))
        cells = int(fe.grid_size[0]) * int(fe.grid_size[1])
        # algorithmic bytes per step on one GPU: points read once, every required
        # output written once; map / lists / workspace traffic is NOT counted.
        # insert reads every point row once; finalize writes every output once (the canvas in cell order,
        # zeros included)
        out_bytes = (16 + 4 + 4 * C) * m_avg + (4.0 * C * cells * per_gpu if has_canvas else 0.0)
        alg = ({"insert": 4.0 * n_avg * c_in, "finalize": out_bytes} if pipeline == 2 else
               {"bin_insert": 4.0 * n_avg * c_in, "emit": out_bytes})
        path_bytes = sum(alg.values())
        step_ms = ms_total / args.steps
        live = {STAGES[i]: float(stage[i]) for i in range(len(STAGES)) if stage[i] > 0}
        dom = max(STAGES, key=lambda k: live.get(k, 0.0))
        dom_ms = live[dom]
        dom_bytes = alg.get(dom, 0.0)
        ach = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_bytes > 0 else None
        traffic, traffic_src = ncu_traffic(dom)
        result = {
            "metric": METRIC, "value": pts_all / (ms_total * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "frames_per_s": per_gpu * world * args.steps / (ms_total * 1e-3),
            "config": {"workload": args.workload, "grid": grid, "frames_per_gpu_per_step": per_gpu,
                       "points_per_gpu_per_step": n_avg, "voxels_per_gpu_per_step": m_avg, "c_in": c_in,
                       "max_points": g["max_points"], "max_voxels": g["max_voxels"],
                       "parallelism": "frame-sharded x%d, no collective" % world,
                       "l2": "%d rotating input sets (%.0f MB) + outputs exceed the 126 MB L2" % (N_SETS, in_bytes / 1e6),
                       "launch": ("eager" if args.no_graph else "cuda-graph replay, one per step issued from Python"
                                  if args.launch == "python" else
                                  "cuda-graph replay, %d steps per graph on %d branches (no host call between steps)"
                                  % (steps_per_graph, n_streams)),
                       "pipeline": {1: "list-based (voxelize.cu)", 2: "list-free (fused.cu)"}[pipeline],
                       "streams": n_streams,
                       "note": "steps are independent batches; with streams > 1 consecutive steps overlap on "
                               "separate CUDA streams (each with its own workspace); single_stream = strictly serial"},
            "single_stream": {"value": pts_all / (ms_single * 1e-3) / 1e6, "unit": UNIT,
                              "ms_per_step": ms_single / args.steps},
            "e2e": {"value": e2e_pts_all / (e2e_ms * 1e-3) / 1e6, "unit": UNIT,
                    "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps, "wall_ms_per_step": e2e_wall / e2e_steps, "steps": e2e_steps,
                    "streams": n_streams, "synchronous_ms_per_step": e2e_sync_ms},
            "gpu_launches": LAUNCHES_PER_STEP[pipeline] * args.steps,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": (ach / peak) if ach else None, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": dom_ms,
                         "stage_ms": live, "stage_ms_sum": float(sum(live.values())),
                         "note": "HBM fraction as the contract asks; the path is bound by scattered L2 requests, "
                                 "not bytes (DESIGN.md 3.1)"},
            "roofline_path": {"algorithmic_bytes_per_step": path_bytes, "achieved": path_bytes / (step_ms * 1e-3) / 1e9,
                              "peak": peak, "unit": "GB/s", "frac": path_bytes / (step_ms * 1e-3) / 1e9 / peak,
                              "frac_of_8000": path_bytes / (step_ms * 1e-3) / 1e9 / 8000.0},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            import oracle
            oracle.lib()
            frames = sets[0]["frames"]
            cpu_reference_pass(frames[:min(len(frames), threads)], grid, threads)          # warm-up
            reps_cpu = 3
            dt = sum(cpu_reference_pass(frames, grid, threads) for _ in range(reps_cpu))
            result["cpu_baseline"] = {
                "value": sets[0]["n"] * reps_cpu / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "%d passes over %d frames (%d points), one frame per thread, oracle C port of the "
                          "reference's numba/numpy/torch CPU path" % (reps_cpu, len(frames), sets[0]["n"]),
                "frames_per_s": len(frames) * reps_cpu / dt}
        _emit(result)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
