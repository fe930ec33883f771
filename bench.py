#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: path-bounces/s and s/frame.

  python bench.py --gpus N --steps K --warmup W [--config c2]     our arm (one process per GPU)
  python bench.py --impl reference --gpus N ... [--config c2]     the reference's own CPU renderer

--config names one of BASELINE.json's configs (default c2, the one the metric is quoted on):
  c1        Weekend random-spheres scene 400x225, 10 spp (the reference's own CPU-runnable case)
  c2        Weekend final scene (487 spheres) 1200x800, 500 spp, depth 50
  c3        triangles/cuda OBJ mesh in the lit room (968-triangle stand-in for Suzanne), 800x800, 1500 spp
  c4        rt_next_week BVH scene (moving spheres, checker, dielectric) 1200x800, 1000 spp
  c5        Weekend final scene 3840x2160, 5000 spp (the multi-GPU config)
  nw_final  rt_next_week final scene (boxes, 1000-sphere cluster, media, perlin, image) 800x800, 1000 spp
  c3_instanced, nw_final_instanced   the same with translate / rotate_y as instances of the two-level BVH

A "step" is one complete frame. N > 1 splits the SAMPLES of the frame over the ranks (strong scaling of one
frame, as BASELINE.json's metric asks), each rank accumulating into its own frame, combined by the library's
rt_reduce (ncclReduce of the R,G,B lanes to rank 0). Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "path_bounces_per_second"
UNIT = "Mpath-bounces/s"

# scene function of a_dive_into_ray_tracing_b200.scenes, its keyword arguments, frame, profile
CONFIGS = {
    "c1": dict(scene="weekend", kw={}, W=400, H=225, spp=10,
               workload="weekend_random_spheres_487_400x225_10spp_depth50"),
    "c2": dict(scene="weekend", kw={}, W=1200, H=800, spp=500,
               workload="weekend_final_scene_487_spheres_1200x800_500spp_depth50"),
    "c3": dict(scene="obj_room", kw={"mesh": "blob968"}, W=800, H=800, spp=1500,
               workload="triangles_cuda_obj_room_968_triangle_mesh_800x800_1500spp_depth50"),
    "c4": dict(scene="next_week", kw={}, W=1200, H=800, spp=1000,
               workload="rt_next_week_bvh_scene_moving_spheres_1200x800_1000spp_depth50"),
    "c5": dict(scene="weekend", kw={}, W=3840, H=2160, spp=5000,
               workload="weekend_final_scene_487_spheres_3840x2160_5000spp_depth50"),
    "nw_final": dict(scene="next_week_final", kw={}, W=800, H=800, spp=1000,
                     workload="rt_next_week_final_scene_800x800_1000spp_depth50"),
    # the same two scenes with translate / rotate_y as instances of the two-level BVH (rt_group / rt_instance)
    "c3_instanced": dict(scene="obj_room", kw={"mesh": "blob968", "instanced": True}, W=800, H=800, spp=1500,
                         workload="triangles_cuda_obj_room_968_triangle_mesh_as_instance_800x800_1500spp_depth50"),
    "nw_final_instanced": dict(scene="next_week_final", kw={"instanced": True}, W=800, H=800, spp=1000,
                               workload="rt_next_week_final_scene_cluster_as_instance_800x800_1000spp_depth50"),
}
DATA = {"weekend": "synthetic (the reference's random_scene() under glibc's default seed, float-rounded fixture)",
        "obj_room": "synthetic (procedural 968-triangle mesh in the room of obj_render.cu:384-524)",
        "next_week": "synthetic (rt_next_week random_scene() distribution, numpy Philox seed 1984)",
        "next_week_final": "synthetic (rt_next_week final scene distribution, procedural earth map)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-variants", action="store_true",
                    help="also time the as-is (-O0) / -O2 one-thread / -O2 x nproc-process builds of the CPU reference")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the bounded baseline sample")
    ap.add_argument("--no-multi-gpu-check", action="store_true")
    ap.add_argument("--cpu-worker", default="", help=argparse.SUPPRESS)  # internal: one process of the nproc-process variant
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    cfg["name"] = a.config
    cfg["W"] = a.width or cfg["W"]
    cfg["H"] = a.height or cfg["H"]
    cfg["spp"] = a.spp or cfg["spp"]
    a.cfg = cfg
    return a


def config_dict(cfg):
    """identical in both arms (the driver compares them): what the workload IS, nothing about how it ran"""
    return {"workload": cfg["workload"], "config": cfg["name"], "width": cfg["W"], "height": cfg["H"], "spp": cfg["spp"],
            "max_depth": 50}


def make_scene(cfg):
    from a_dive_into_ray_tracing_b200 import scenes
    fn = getattr(scenes, cfg["scene"])
    return fn(width=cfg["W"], height=cfg["H"], **cfg["kw"])


# ------------------------------------------------------------------ CPU reference arm
def _sample_rows(H):
    return list(range(8, H, 16)) or [H // 2]


def cpu_reference_sample(cfg, target_seconds, variant="threads", threads=None):
    """The reference's own CPU implementation of the path on a bounded sample of the SAME workload: the
    config's view, every 16th row, spp chosen so that the sample takes about target_seconds.
    Weekend configs: oracle/_ref/libref_l0.so = the unmodified rt_in_one_weekend sources (its worker() + std::thread
    split, main.cpp:267-334), kind "reference". variant: "threads" (-O2, std::thread split over all cores: the
    reference as it runs, rand() lock included), "asis" (the reference Makefile's flags: no -O), "o2_1thread",
    "o2_procs" (-O2, one PROCESS per core on row slabs: no shared rand() lock = the fair all-core figure).
    CUDA-tree configs (c3, c4, nw_final) have no CPU implementation in the reference: the plain-C restatement of
    their device code (oracle L1-32, pinned to the reference's CUDA code run on a B200) on all cores, kind "port"."""
    from oracle import pyoracle
    W, H = cfg["W"], cfg["H"]
    ncores = os.cpu_count() or 1
    rows = _sample_rows(H)
    weekend = cfg["scene"] == "weekend"
    l0 = None
    if weekend:
        try:
            l0 = pyoracle.L0(asis=(variant == "asis"))
        except Exception:
            l0 = None
    if l0 is not None and variant == "o2_procs":
        return _cpu_procs_sample(cfg, target_seconds, ncores)
    if l0 is not None:
        threads = 1 if variant == "o2_1thread" else (threads or ncores)
        kind = "reference"
        cam13 = pyoracle.WEEKEND_CAM13(W / H)
        secs, seg = 0.0, 0
        for j in rows[:2]:  # calibrate on two rows at 2 spp
            s_, g_, _ = l0.worker_timed(W, H, 2, cam13, j * W, (j + 1) * W, threads, seed=1)
            secs += s_
        per_row_spp = secs / (2 * 2)
        spp = int(max(1, min(cfg["spp"], round(target_seconds / (per_row_spp * len(rows))))))
        secs = 0.0
        for j in rows:
            s_, g_, _ = l0.worker_timed(W, H, spp, cam13, j * W, (j + 1) * W, threads, seed=1)
            secs += s_
            seg += g_
    else:
        kind = "port"
        threads = threads or ncores
        sc = make_scene(cfg)
        orc = pyoracle.L1(64 if weekend else 32)
        from concurrent.futures import ThreadPoolExecutor

        def run(spp_):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL: one row per task
                segs = list(ex.map(lambda j: orc.render(sc, sc.profile, W, H, spp_, seed=1, rows=(j, j + 1),
                                                        want_sumsq=False)[2], rows))
            return time.perf_counter() - t0, int(sum(segs))

        s1, _ = run(1)
        spp = int(max(1, min(cfg["spp"], round(target_seconds / max(s1, 1e-3)))))
        secs, seg = run(spp)
    paths = len(rows) * W * spp
    return {"value": seg / secs / 1e6, "unit": UNIT, "cores": threads, "kind": kind, "variant": variant,
            "sample": "%dx%d view, every 16th row (%d rows), %d spp, depth 50: %d paths, %d bounces in %.2f s"
                      % (W, H, len(rows), spp, paths, seg, secs),
            "seconds": secs, "bounces": seg, "paths": paths,
            "extrapolated_s_per_frame": secs * (H / len(rows)) * (cfg["spp"] / spp)}


def _cpu_procs_sample(cfg, target_seconds, nproc):
    """-O2 reference, nproc independent PROCESSES (each its own glibc rand() state, no lock contention), the sample
    rows dealt round-robin; wall clock around the whole group."""
    W, H = cfg["W"], cfg["H"]
    rows = _sample_rows(H)

    def launch(spp):
        ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--config", cfg["name"], "--width", str(W),
                                "--height", str(H), "--cpu-worker", "%d,%d,%d" % (k, nproc, spp)],
                               stdout=subprocess.PIPE, text=True) for k in range(nproc)]
        outs = [p.communicate()[0].strip().splitlines()[-1].split() for p in ps]
        # the processes run side by side: the group's time is the slowest process's RENDER time (interpreter start-up
        # and library loading are not part of the reference's work and are left out)
        return max(float(o[1]) for o in outs), sum(int(o[0]) for o in outs)

    s1, _ = launch(1)
    spp = int(max(1, min(cfg["spp"], round(target_seconds / max(s1, 1e-3)))))
    secs, seg = launch(spp)
    paths = len(rows) * W * spp
    return {"value": seg / secs / 1e6, "unit": UNIT, "cores": nproc, "kind": "reference", "variant": "o2_procs",
            "sample": "%dx%d view, every 16th row (%d rows), %d spp, depth 50, %d processes side by side: %d paths, %d bounces, "
                      "slowest process %.2f s" % (W, H, len(rows), spp, nproc, paths, seg, secs),
            "seconds": secs, "bounces": seg, "paths": paths,
            "extrapolated_s_per_frame": secs * (H / len(rows)) * (cfg["spp"] / spp)}


def cpu_worker(args):
    """one process of the o2_procs variant: rows k, k+n, ... of the sample through the unmodified reference"""
    from oracle import pyoracle
    k, n, spp = (int(x) for x in args.cpu_worker.split(","))
    W, H = args.cfg["W"], args.cfg["H"]
    l0 = pyoracle.L0()
    cam13 = pyoracle.WEEKEND_CAM13(W / H)
    seg = 0
    t0 = time.perf_counter()
    for j in _sample_rows(H)[k::n]:
        seg += l0.render(W, H, spp, cam13, seed=1 + k, rows=(j, j + 1), want_sumsq=False)[2]
    print(seg, time.perf_counter() - t0, flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.cfg
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_sample(cfg, 1.0)
    vals, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_reference_sample(cfg, args.cpu_seconds)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    cpu = {"value": v, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
           "note": "each step = a bounded sample of the workload on the host CPU, extrapolated to the frame"}
    if args.cpu_variants and cfg["scene"] == "weekend":
        cpu["variants"] = cpu_variants(cfg, args.cpu_seconds)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64" if cfg["scene"] == "weekend" else "f32",
           "data": DATA[cfg["scene"]], "config": config_dict(cfg), "cpu_baseline": cpu,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "extrapolated_s_per_frame": last["extrapolated_s_per_frame"]}
    print(json.dumps(out), flush=True)


def cpu_variants(cfg, seconds):
    """BASELINE.md section 3 / SURVEY.md 8d: the three other ways to time the CPU reference, a few seconds each"""
    out = {}
    for name in ("asis", "o2_1thread", "o2_procs"):
        try:
            c = cpu_reference_sample(cfg, max(3.0, seconds / 3.0), variant=name)
            out[name] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated_s_per_frame")}
        except Exception as e:  # pragma: no cover
            out[name] = {"value": None, "error": str(e)}
    out["what"] = {"asis": "the reference Makefile's flags (no -O), its std::thread split over all cores",
                   "o2_1thread": "-O2, one thread", "o2_procs": "-O2, one process per core (no shared rand() lock): "
                   "the fair all-core CPU figure"}
    return out


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                pw.append(float(c[3]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if c[5 + k].lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic_bytes(cfg_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of k_render per launch, from the newest committed
    `ncu --set full` capture of this config's kernel (profiles/r*_<cfg>_k_render_raw.csv; config 2 captures carry
    no config tag). The captures are taken at reduced spp: the read side (scene staging) does not depend on spp,
    the write side (per-chunk partial frames) is scaled to the frame's chunk count by the caller. None if absent."""
    import csv
    import glob
    try:
        tag = "" if cfg_name in ("c2", "c1", "c5") else {"c3": "config3_", "nw_final": "final_scene_"}.get(cfg_name, cfg_name + "_")
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9]*_%sk_render_raw.csv" % tag)),
                       key=lambda f: os.path.basename(f))
        if not tag:
            files = [f for f in files if "config" not in f and "final_scene" not in f]
        rows = list(csv.reader(open(files[-1])))
        d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return (sum(float(d[k]) * mult[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")),
                os.path.basename(files[-1]))
    except Exception:
        return None, None


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from a_dive_into_ray_tracing_b200 import capi, ctypes_defs as D
    from a_dive_into_ray_tracing_b200.dist import init_comm, render_frame, sample_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    cfg = args.cfg
    W, H, spp, K, Wm = cfg["W"], cfg["H"], cfg["spp"], args.steps, max(args.warmup, 3)

    sc = make_scene(cfg)
    profile = sc.profile
    ctx = capi.Context(profile=profile, device=local, seed=1984)
    ctx.upload(sc).build_accel(1)
    if world > 1:
        init_comm(ctx, rank, world)  # the library's own NCCL communicator (rt_comm_init); torch only ships the id
    accum = torch.zeros(H, W, 4, device=dev, dtype=torch.float32)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)  # > 126 MB L2
    rgb_host = ctx.pinned_array((H, W, 3), np.uint8)  # rt_host_alloc: the 8-bit frame lands in page-locked memory
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()

    def frame():
        accum.zero_()
        render_frame(ctx, W, H, spp, accum, rank, world)

    for _ in range(Wm):
        frame()
    torch.cuda.synchronize()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.stats_reset()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        frame()
        ev[k][1].record()
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    st = ctx.stats()
    t = torch.tensor([sum(ms_steps), float(st["segments"]), float(st["paths"]), float(st["kernel_launches"])],
                     device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_ms = float(tmax[0])
    else:
        total_ms = float(t[0])
    segments, paths, launches = float(t[1]), float(t[2]), float(t[3])
    ms_per_step = total_ms / K
    value = segments / K / (ms_per_step * 1e-3) / 1e6

    # ---- per-phase breakdown (CUDA events, a separate pass of 3 frames outside the timed region; max over ranks)
    def phase_pass():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ctx.stats_reset()
        barrier()
        torch.cuda.synchronize()
        e[0].record()
        accum.zero_()
        e[1].record()
        begin, count = sample_range(spp, rank, world)
        ctx.render_device(W, H, count, begin, accum.data_ptr(), stream)
        e[2].record()
        if world > 1:
            ctx.reduce(W, H, accum.data_ptr(), root=0, uniform_count=True, stream_ptr=stream)
        e[3].record()
        if rank == 0:
            ctx.resolve_device(W, H, accum.data_ptr(), want_linear=False, want_rgb8=True, stream_ptr=stream,
                               out_rgb8=rgb_host)
        e[4].record()
        torch.cuda.synchronize()
        s = ctx.stats()
        return [e[0].elapsed_time(e[1]), s["ms_k_render"], s["ms_k_combine"], e[2].elapsed_time(e[3]),
                e[3].elapsed_time(e[4]), e[0].elapsed_time(e[4])]

    ph = np.median(np.array([phase_pass() for _ in range(3)]), axis=0)
    ph_t = torch.tensor(ph, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ph_t, op=dist.ReduceOp.MAX)
    phases = dict(zip(["zero_ms", "k_render_ms", "k_combine_ms", "reduce_ms", "resolve_d2h_ms", "total_ms"],
                      [float(x) for x in ph_t]))
    phases["note"] = ("median of 3 frames outside the timed region, max over ranks; reduce_ms on a rank includes waiting "
                      "for the slowest rank's render")

    # ---- end to end through the public API with HOST buffers: scene in host memory ->
    # upload (H2D) -> BVH build -> render -> reduce -> resolve -> 8-bit image in (pinned) host memory
    h2d = int(sc.spheres.nbytes + sc.materials.nbytes + sc.triangles.nbytes + sc.quads.nbytes + 96 + 32)
    d2h = W * H * 3
    e2e_ms, e2e_parts = [], []
    for k in range(max(3, min(K, 5)) + 1):
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.upload(sc)
        t1 = time.perf_counter()
        ctx.build_accel(1)
        t2 = time.perf_counter()
        frame()
        if rank == 0:
            ctx.resolve_device(W, H, accum.data_ptr(), want_linear=False, want_rgb8=True, stream_ptr=stream,
                               out_rgb8=rgb_host)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        barrier()
        if k > 0:
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
            e2e_parts.append([1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)])
    e2e_t = torch.tensor([float(np.mean(e2e_ms))] + list(np.mean(np.array(e2e_parts), axis=0)), device=dev,
                         dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms_mean = float(e2e_t[0])
    e2e_value = segments / K / (e2e_ms_mean * 1e-3) / 1e6

    # ---- multi-GPU image check (SURVEY.md 8c protocol item 3): the reduced frame of the N-rank sample split
    # against a full single-GPU render of rank 0, outside every timed region
    mg_check = None
    if world > 1 and not args.no_multi_gpu_check:
        frame()
        torch.cuda.synchronize()
        barrier()
        if rank == 0:
            full = torch.zeros_like(accum)
            ctx.render_device(W, H, spp, 0, full.data_ptr(), stream)
            torch.cuda.synchronize()
            a = (accum[..., :3] / accum[..., 3:4]).double()
            b = (full[..., :3] / full[..., 3:4]).double()
            rel = ((a - b).abs() / b.abs().clamp_min(1e-3)).max().item()
            mg_check = {"max_rel_diff_vs_single_gpu": rel, "counts_equal": bool(torch.equal(accum[..., 3], full[..., 3])),
                        "what": "reduced %d-rank frame vs rank 0 rendering all %d spp alone; per-pixel mean radiance, "
                                "relative to max(|ref|, 1e-3); fp32 summation order is the only difference" % (world, spp)}
        barrier()

    out = None
    if rank == 0:
        # ---- roofline inputs: per-segment work counted by the counter variant of the kernel
        cctx = capi.Context(profile=profile, device=local, seed=1984, flags=D.RT_FLAG_COUNTERS)
        cctx.upload(sc).build_accel(1)
        cctx.render(W, H, max(1, min(4, spp)))
        cs = cctx.stats()
        fp32_peak = cctx.measure_fp32_peak()
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count
        sm_mhz_max = (clocks or {}).get("sm_max_mhz") or 1965.0
        cctx.close()
        n_box = cs["box_tests"] / cs["segments"]
        n_prim = cs["prim_tests"] / cs["segments"]
        h_bar = 1.0 - cs["paths"] / cs["segments"]
        n_prims = {"sphere": len(sc.spheres), "triangle": len(sc.triangles), "quad": len(sc.quads)}
        tot = max(1, sum(n_prims.values()))
        # 16 / 40 / 8 lane-instructions per sphere / triangle / rect test, weighted by the scene's mix
        i_prim = (16.0 * n_prims["sphere"] + 40.0 * n_prims["triangle"] + 8.0 * n_prims["quad"]) / tot
        b_prim = (16.0 * n_prims["sphere"] + 48.0 * n_prims["triangle"] + 32.0 * n_prims["quad"]) / tot
        i_seg = 19.0 * n_box + i_prim * n_prim + 150.0 * h_bar + 40.0
        b_seg = 16.0 * n_box + b_prim * n_prim + 32.0 * h_bar
        seg_per_s = value * 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        k_ms = st["ms_k_render"]  # mean CUDA-event duration of the k_render launches of the TIMED region
        seg_per_launch = segments / K / world
        traffic, traffic_src = ncu_traffic_bytes(cfg["name"])
        smem_peak = 128.0 * sm_count * sm_mhz_max * 1e6 / 1e9
        roofline = {
            "bound": "fp32_issue", "kernel": "k_render<profile %d, smem plan %d>" % (profile, st["smem_plan"]),
            "achieved": seg_per_launch * i_seg * 2.0 / (k_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": seg_per_launch * i_seg * 2.0 / (k_ms * 1e-3) / 1e12 / fp32_peak if fp32_peak > 0 else None,
            "launch_ms": k_ms, "launch_share_of_step": k_ms / ms_per_step,
            "peak_source": "measured in this run by rt_measure_fp32_peak (FMA chain, 2 flops/FMA); "
                           "MEASURED_PEAKS.json has no FP32 entry",
            "model": {"lane_instr_per_segment": i_seg, "box_tests_per_segment": n_box,
                      "prim_tests_per_segment": n_prim, "hits_per_segment": h_bar,
                      "formula": "19*box + (16 sphere | 40 triangle | 8 rect, scene mix)*prim + 150*hit + 40 "
                                 "(SURVEY.md 8d), 2 flop per lane-instruction; segments of one launch / the launch's "
                                 "CUDA-event duration (k_render_ms)"},
            "traffic": traffic, "traffic_source": traffic_src,
            "smem": {"bound": "shared_memory", "unit": "GB/s",
                     "achieved": seg_per_launch * b_seg / (k_ms * 1e-3) / 1e9, "peak": smem_peak,
                     "frac": seg_per_launch * b_seg / (k_ms * 1e-3) / 1e9 / smem_peak,
                     "model": "B_seg = 16 B (quantised node record) x box tests + (16 sphere | 48 triangle | 32 rect) x "
                              "prim tests + 32 B material per hit; peak = nominal 128 B/clk/SM x SMs x max SM clock"},
            "hbm": {"bound": "hbm", "achieved": (W * H * 16.0 * 2) / (ms_per_step * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": (W * H * 16.0 * 2) / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                    "note": "algorithmic framebuffer traffic only (one float4 frame written and read); the scene lives in "
                            "shared memory / L1: not HBM-bound",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                c = cpu_reference_sample(cfg, args.cpu_seconds)
                cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated_s_per_frame")}
                if args.cpu_variants and cfg["scene"] == "weekend":
                    cpu["variants"] = cpu_variants(cfg, args.cpu_seconds)
            except Exception as e:  # pragma: no cover
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
               "ms_per_step": ms_per_step, "s_per_frame": ms_per_step * 1e-3, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": DATA[cfg["scene"]],
               "config": config_dict(cfg),
               "run": {"parallelism": "sample-split x%d + rt_reduce (ncclReduce of R,G,B to rank 0)" % world
                       if world > 1 else "single GPU",
                       "l2_flush_between_steps": True, "inputs": "scene resident in HBM/shared memory"},
               "paths_per_step": paths / K, "bounces_per_step": segments / K,
               "wall_ms_per_step_incl_flush": 1e3 * t_wall / K,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_ms_mean,
                       "host_ms": {"scene_upload": float(e2e_t[1]), "accel_build": float(e2e_t[2]),
                                   "render_reduce_resolve_d2h": float(e2e_t[3])},
                       "what": "host scene -> rt_scene_upload -> rt_accel_build -> render -> rt_reduce -> rt_resolve "
                               "-> 8-bit image in pinned host memory (rt_host_alloc)"},
               "phases": phases, "multi_gpu_check": mg_check,
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
               "kernel": {k: st[k] for k in ("n_nodes", "n_big_prims", "smem_bytes", "smem_plan", "block_threads",
                                             "grid_blocks", "regs_per_thread")}}
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def main():
    args = parse()
    if args.cpu_worker:
        cpu_worker(args)
        return
    # The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version
    # banner on fd 1): point fd 1 at stderr while working and hand the real stdout to print().
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
