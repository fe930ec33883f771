#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: path-bounces/s and s/frame for the
Weekend final scene (~487 spheres), 1200x800, 500 spp, depth 50 (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU)
  python bench.py --impl reference --gpus N ...            the reference's own CPU renderer

A "step" is one complete frame. N > 1 splits the SAMPLES of the frame over the ranks
(strong scaling of one frame, as BASELINE.json's metric asks), each rank accumulating
into its own frame, combined by one NCCL sum-reduce to rank 0.
Prints ONE JSON line on rank 0.
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

WORKLOAD = "weekend_final_scene_487_spheres_1200x800_500spp_depth50"
METRIC = "path_bounces_per_second"
UNIT = "Mpath-bounces/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1200)
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--spp", type=int, default=500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the bounded baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(W, H, target_seconds, threads=None):
    """The reference's own CPU renderer (oracle/_ref/libref_l0.so = the unmodified
    rt_in_one_weekend sources, its worker() + std::thread split, main.cpp:267-334) on a
    bounded sample of the SAME workload: the 1200x800 view, every 16th row, spp chosen so
    the sample takes about target_seconds. Returns dict(value Mbounces/s, ...)."""
    from oracle import pyoracle
    threads = threads or os.cpu_count() or 1
    kind = "reference"
    try:
        l0 = pyoracle.L0()
    except Exception:
        l0 = None
    cam13 = pyoracle.WEEKEND_CAM13(W / H)
    rows = list(range(8, H, 16))
    if l0 is not None:
        # calibrate on two rows at 2 spp
        secs, seg, _ = 0.0, 0, None
        for j in rows[:2]:
            s_, g_, _ = l0.worker_timed(W, H, 2, cam13, j * W, (j + 1) * W, threads, seed=1)
            secs += s_
            seg += g_
        per_row_spp = secs / (2 * 2)
        spp = int(max(1, min(500, round(target_seconds / (per_row_spp * len(rows))))))
        secs, seg = 0.0, 0
        for j in rows:
            s_, g_, _ = l0.worker_timed(W, H, spp, cam13, j * W, (j + 1) * W, threads, seed=1)
            secs += s_
            seg += g_
    else:
        # the compiled reference did not travel: time the plain-C restatement instead
        kind = "port"
        from a_dive_into_ray_tracing_b200 import scenes
        orc = pyoracle.L1(64)
        sc = scenes.weekend(W, H)
        spp = 4
        t0 = time.perf_counter()
        seg = 0
        for j in rows:
            _, _, g_ = orc.render(sc, 0, W, H, spp, seed=1, rows=(j, j + 1), want_sumsq=False)
            seg += g_
        secs = time.perf_counter() - t0
        threads = 1
    paths = len(rows) * W * spp
    return {"value": seg / secs / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "%dx%d view, every 16th row (%d rows), %d spp, depth 50: %d paths, %d bounces in %.2f s"
                      % (W, H, len(rows), spp, paths, seg, secs),
            "seconds": secs, "bounces": seg, "paths": paths,
            "extrapolated_s_per_frame_500spp": secs * (H / len(rows)) * (500.0 / spp)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W, H = args.width, args.height
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_sample(W, H, 1.0)
    vals, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_reference_sample(W, H, args.cpu_seconds)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic (reference random_scene(), glibc default seed)",
           "config": {"workload": WORKLOAD, "width": W, "height": H, "spp": args.spp, "max_depth": 50,
                      "note": "each step = bounded sample of the workload on the host CPU"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
                            "sample": last["sample"]},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "extrapolated_s_per_frame": last["extrapolated_s_per_frame_500spp"]}
    print(json.dumps(out), flush=True)


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


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of k_render per launch, from the committed
    newest `ncu --set full` capture of this kernel (profiles/r1*_k_render_raw.csv); None if absent."""
    import csv
    import glob
    try:
        newest = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1*_k_render_raw.csv")))[-1]
        rows = list(csv.reader(open(newest)))
        d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return sum(float(d[k]) * mult[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        return None


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from a_dive_into_ray_tracing_b200 import capi, ctypes_defs as D, scenes
    from a_dive_into_ray_tracing_b200.dist import render_frame

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    W, H, spp, K, Wm = args.width, args.height, args.spp, args.steps, max(args.warmup, 3)

    sc = scenes.weekend(W, H)
    ctx = capi.Context(profile=0, device=local, seed=1984)
    ctx.upload(sc).build_accel(1)
    accum = torch.zeros(H, W, 4, device=dev, dtype=torch.float32)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)  # > 126 MB L2

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

    # ---- end to end through the public API with HOST buffers: scene in host memory ->
    # upload (H2D) -> BVH build -> render -> reduce -> resolve -> 8-bit image in host memory
    h2d = int(sc.spheres.nbytes + sc.materials.nbytes + sc.triangles.nbytes + sc.quads.nbytes + 96 + 32)
    d2h = W * H * 3
    e2e_ms = []
    for k in range(max(3, min(K, 5)) + 1):
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.upload(sc)
        ctx.build_accel(1)
        frame()
        if rank == 0:
            _, rgb = ctx.resolve_device(W, H, accum.data_ptr(), want_linear=False, want_rgb8=True,
                                        stream_ptr=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        barrier()
        if k > 0:
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
    e2e_t = torch.tensor([float(np.mean(e2e_ms))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms_mean = float(e2e_t[0])
    e2e_value = segments / K / (e2e_ms_mean * 1e-3) / 1e6

    out = None
    if rank == 0:
        # ---- roofline inputs: per-segment work counted by the counter variant of the kernel
        cctx = capi.Context(profile=0, device=local, seed=1984, flags=D.RT_FLAG_COUNTERS)
        cctx.upload(sc).build_accel(1)
        cctx.render(W, H, 4)
        cs = cctx.stats()
        fp32_peak = cctx.measure_fp32_peak()
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count
        sm_mhz_max = (clocks or {}).get("sm_max_mhz") or 1965.0
        cctx.close()
        n_box = cs["box_tests"] / cs["segments"]
        n_prim = cs["prim_tests"] / cs["segments"]
        h_bar = 1.0 - cs["paths"] / cs["segments"]
        i_seg = 19.0 * n_box + 16.0 * n_prim + 150.0 * h_bar + 40.0  # SURVEY.md §8d lane-instructions / segment
        seg_per_s = value * 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        n_chunks_bytes = W * H * 16
        hbm_bytes_per_frame = n_chunks_bytes * (2 * 3 + 2)  # partial write+read (<=3 chunks) + accum rw (approx.)
        roofline = {
            "bound": "fp32_issue", "kernel": "k_render<0,false,2,false>",
            "achieved": seg_per_s * i_seg * 2.0 / 1e12 / world, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": seg_per_s * i_seg * 2.0 / 1e12 / world / fp32_peak if fp32_peak > 0 else None,
            "peak_source": "measured in this run by rt_measure_fp32_peak (FMA chain, 2 flops/FMA); "
                           "MEASURED_PEAKS.json has no FP32 entry",
            "model": {"lane_instr_per_segment": i_seg, "box_tests_per_segment": n_box,
                      "prim_tests_per_segment": n_prim, "hits_per_segment": h_bar,
                      "formula": "19*box + 16*prim + 150*hit + 40 (SURVEY.md 8d), 2 flop per lane-instruction"},
            "traffic": ncu_traffic_bytes(),
            "smem": {"bound": "shared_memory", "unit": "GB/s",
                     "achieved": seg_per_s * (32.0 * n_box + 16.0 * n_prim + 32.0 * h_bar) / 1e9 / world,
                     "peak": 128.0 * sm_count * sm_mhz_max * 1e6 / 1e9,
                     "frac": seg_per_s * (32.0 * n_box + 16.0 * n_prim + 32.0 * h_bar) / 1e9 / world /
                             (128.0 * sm_count * sm_mhz_max * 1e6 / 1e9),
                     "model": "B_seg = 32 B/node x box tests + 16 B x sphere tests + 32 B material per hit "
                              "(SURVEY.md 8d); peak = nominal 128 B/clk/SM x SMs x max SM clock; ncu (profiles/): "
                              "the shared-memory data pipe runs at ~80 % of its wavefront peak"},
            "hbm": {"bound": "hbm", "achieved": hbm_bytes_per_frame / (ms_per_step * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": hbm_bytes_per_frame / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                    "note": "framebuffer traffic only; the scene (~60 KB) lives in shared memory: not HBM-bound",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                c = cpu_reference_sample(W, H, args.cpu_seconds)
                cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
                cpu["extrapolated_s_per_frame"] = c["extrapolated_s_per_frame_500spp"]
            except Exception as e:  # pragma: no cover
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
               "ms_per_step": ms_per_step, "s_per_frame": ms_per_step * 1e-3, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data":
                   "synthetic (the reference's random_scene() under glibc's default seed, float-rounded fixture)",
               "config": {"workload": WORKLOAD, "width": W, "height": H, "spp": spp, "max_depth": 50,
                          "parallelism": "sample-split x%d + NCCL reduce" % world if world > 1 else "single GPU",
                          "l2_flush_between_steps": True, "inputs": "scene resident in HBM/shared memory"},
               "paths_per_step": paths / K, "bounces_per_step": segments / K,
               "wall_ms_per_step_incl_flush": 1e3 * t_wall / K,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_ms_mean,
                       "what": "host scene -> rt_scene_upload -> rt_accel_build -> render -> reduce -> rt_resolve "
                               "-> 8-bit image in host memory"},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
               "kernel": {k: st[k] for k in ("n_nodes", "n_big_prims", "smem_bytes", "block_threads", "grid_blocks",
                                             "regs_per_thread")}}
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def main():
    args = parse()
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
