#!/usr/bin/env python
"""bench.py — Msamples/s (and Gbounces/s) of the per-pixel radiance path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one full render of the workload's frame (all samples of all pixels) through the C ABI of
libipt_b200.so.  At N > 1 the frame's tiles are interleaved statically over the ranks (one process per GPU), every
rank renders its tiles, and the finished tiles are written into rank 0's frame over NVLink peer access (CUDA IPC);
there is no collective on the data path, so torch.distributed is used only for the barrier and the max over ranks.
`value` is device-resident (scene already in HBM, CUDA-event time of the kernels, max over ranks); `e2e` is the same
metric through the host-buffer path: scene upload from pinned host memory + kernels + gather + frame download.

--impl reference times the reference's own CPU routine (oracle/_ref: its unmodified Renderer.cu compiled for the
host) on all host cores, on a bounded sample of the same workload.  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "improved-path-tracer_b200"))

# BASELINE.json configs.  flops = algorithmic fp32 flops per bounce (19 S + 31 R + 60, brute force)
WORKLOADS = {
    "spheres4k": dict(scene="spheres", width=3840, height=2160, depth=32, spp=1024, flops=303,
                      desc="scenes/spheres.json at 3840x2160 -d=32 -s=1024 (BASELINE configs[3], literal reading)"),
    # the evenly loaded variant SURVEY.md §8d asks for beside the literal one: every length x3 (positions, radii, edge
    # vectors, camera position), so the room fills the 3840x2160 frame as it fills 1280x720 upstream
    "spheres4k_x3": dict(scene="spheres_x3", width=3840, height=2160, depth=32, spp=1024, flops=303,
                         desc="spheres.json with all lengths x3 at 3840x2160 -d=32 -s=1024 (configs[3], evenly loaded variant: the room fills the frame)"),
    "spheres": dict(scene="spheres", width=None, height=None, depth=10, spp=40, flops=303, desc="scenes/spheres.json -d=10 -s=40 (configs[0])"),
    "mirrors": dict(scene="mirrors", width=None, height=None, depth=10, spp=40, flops=453, desc="scenes/mirrors.json -d=10 -s=40 (configs[1])"),
    "maze": dict(scene="maze", width=None, height=None, depth=10, spp=40, flops=1786, desc="scenes/maze.json -d=10 -s=40 (configs[2])"),
    # configs[4]: generated on the fly by scripts/make_synthetic_scene.py (seeded); BVH path, flops not modelled
    "synthetic1m": dict(scene="synthetic1m", width=None, height=None, depth=10, spp=256, flops=0,
                        desc="synthetic 1M-primitive scene in the scenes/*.json schema -d=10 -s=256 (configs[4])"),
}
BYTES_PER_BOUNCE = 96          # a wavefront that compacts after every bounce: 48 B ray record read + 48 B written
FP32_LANES_PER_SM = 128


def scene_file(name):
    if name == "synthetic1m":
        import subprocess
        p = "/tmp/ipt_synthetic_1000000.json"
        if not os.path.isfile(p):
            if int(os.environ.get("LOCAL_RANK", "0")) == 0:        # one writer per box, atomic rename; the others wait
                tmp = f"{p}.{os.getpid()}.tmp"
                subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "make_synthetic_scene.py"), tmp, "1000000"], check=True)
                os.replace(tmp, p)
            else:
                t_end = time.time() + 600
                while not os.path.isfile(p):
                    if time.time() > t_end:
                        raise SystemExit(f"{p}: rank 0 did not write the synthetic scene")
                    time.sleep(0.2)
        return p
    if name.endswith("_x3"):
        src = json.load(open(scene_file(name[:-3])))
        for k in ("width", "height"):
            src[k] *= 3
        for o in [src["camera"]] + src["objects"]:
            for k in ("position", "north", "east"):
                if k in o:
                    o[k] = {a: 3.0 * b for a, b in o[k].items()}
            if "radius" in o:
                o["radius"] *= 3.0
        p = f"/tmp/ipt_{name}_{os.getpid()}.json"
        json.dump(src, open(p, "w"))
        return p
    p = os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".json")
    if not os.path.isfile(p):
        raise SystemExit(f"{p} missing: run `python -c 'import __graft_entry__ as g; g.build()'` where the reference is mounted")
    return p


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.power, self.power_limit = [], None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.power_limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1e3
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        pw = sorted(self.power) or [None]
        return {"sm_mhz": s[len(s) // 2], "sm_min_mhz": s[0], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "power_w": pw[len(pw) // 2], "power_w_max": pw[-1], "power_limit_w": self.power_limit}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def spread_cells(n_cells, m):
    """m distinct cells of the reference's 22x22 thread grid spread over the whole frame (multiplicative stride coprime
    with the grid, so no row/column aliasing)."""
    step = 197 if n_cells % 197 else 199
    return sorted({(i * step + 5) % n_cells for i in range(m)})


def run_reference(args, wl, rank):
    """The reference's own per-pixel routine on the host cores: a bounded sample of the workload per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    if not O.ref_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_host.so not built"}))
        return
    cores = args.ref_threads or os.cpu_count() or 1
    W, H = wl["width"] or 1280, wl["height"] or 720
    n_cells = O.ref().ref_num_cells(W, H)
    # bounded sample: every `stride`-th reference cell (a cell = one reference CUDA thread's pixel rectangle,
    # Renderer.cu:33-53), reduced spp; throughput in samples/s does not depend on spp
    stride, spp = args.ref_stride, args.ref_spp
    cells = spread_cells(n_cells, max(1, n_cells // stride))
    px_per_cell = (W * H) / n_cells
    ref_scene = scene_file(wl["scene"]) if wl["scene"].endswith("_x3") else wl["scene"]

    def one_step():
        return O.ref_time_cells(ref_scene, spp, wl["depth"], wl["width"], wl["height"], cells, cores)

    for _ in range(min(args.warmup, 1)):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    samples = len(cells) * px_per_cell * spp
    ms = 1e3 * sum(times) / len(times)
    value = samples / (ms * 1e-3) / 1e6
    sample_desc = f"{len(cells)} of the reference's {n_cells} thread cells, spread over the frame ({int(samples)} samples) at {spp} spp, depth {wl['depth']}"
    line = {"impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": wl["desc"], "sample": sample_desc},
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "reference", "sample": sample_desc},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline(wl, seconds_budget=20.0):
    """Rank 0, N=1: the reference's host-compiled routine (kind "reference") on a bounded sample, all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import oracle as O
        if not O.ref_available():
            raise RuntimeError("oracle/_ref not built")
        cores = os.cpu_count() or 1
        W, H = wl["width"] or 1280, wl["height"] or 720
        n_cells = O.ref().ref_num_cells(W, H)
        cells = spread_cells(n_cells, min(n_cells, 4 * cores))
        spp = 4
        ref_scene = scene_file(wl["scene"]) if wl["scene"].endswith("_x3") else wl["scene"]
        dt = O.ref_time_cells(ref_scene, spp, wl["depth"], wl["width"], wl["height"], cells, cores)
        # scale spp so that the sample takes ~seconds_budget, then time that
        spp = int(max(4, min(256, spp * seconds_budget / max(dt, 1e-3))))
        dt = O.ref_time_cells(ref_scene, spp, wl["depth"], wl["width"], wl["height"], cells, cores)
        samples = len(cells) * (W * H / n_cells) * spp
        return {"value": samples / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
                "sample": f"{len(cells)} of the reference's {n_cells} thread cells (spread over the frame) at {spp} spp, depth {wl['depth']}: {int(samples)} samples in {dt:.1f} s, fp64"}
    except Exception as e:   # the baseline is reported, never required for the GPU number
        return {"value": None, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("IPT_BENCH_WORKLOAD", "spheres4k"), choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (the line then says so)")
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--leaf", type=int, default=4, help="primitives per BVH leaf (A/B knob; BVH scenes only)")
    ap.add_argument("--fp64", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dry-run", action="store_true", help="CPU only (gloo): exercise sharding, handle exchange and reductions without rendering")
    ap.add_argument("--ref-stride", type=int, default=4, help="reference arm: 1/stride of the reference's 484 thread cells are rendered per step")
    ap.add_argument("--ref-spp", type=int, default=4)
    ap.add_argument("--ref-threads", type=int, default=0, help="reference arm: host threads (0 = all cores)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.spp:
        wl["spp"] = args.spp
        wl["desc"] += f" [spp overridden to {args.spp}]"
    if args.depth:
        wl["depth"] = args.depth
        wl["desc"] += f" [depth overridden to {args.depth}]"

    if args.impl == "reference":
        run_reference(args, wl, rank)
        return

    import numpy as np
    import pyipt

    if args.dry_run:
        dry_run(args, wl, rank, world, pyipt)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        token = torch.zeros(1, device="cuda")

    def barrier():
        if dist is not None:
            dist.all_reduce(token)
            torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if pyipt.lib().ipt_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device — there is no CPU path to time (use --impl reference for the CPU baseline)")

    hs = pyipt.HostScene.load(scene_file(wl["scene"]), width=wl["width"], height=wl["height"], leaf_size=args.leaf)
    W, H = hs.width, hs.height
    ctx = pyipt.Context(local)
    ctx.set_scene(hs)
    flags = pyipt.FLAG_FP64 if args.fp64 else 0
    gather = "single GPU"
    if world > 1:
        # rank 0's fp32 frame is the gather target of every rank: its CUDA IPC handle goes round once
        handle = [ctx.export_frame() if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            ctx.set_gather_target_ipc(handle[0])
        gather = "tiles stored into rank 0's frame over NVLink peer access (CUDA IPC), no collective"

    def step():
        return ctx.render(wl["spp"], wl["depth"], seed=123456, flags=flags, rank=rank, world=world, batch=args.batch)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms, bounces, samples, launches, qbytes = 0.0, 0, 0, 0, 0
    for _ in range(args.steps):
        st = step()
        dev_ms += st["render_ms"]; bounces += st["traced_bounces"]; samples += st["samples"]; launches += st["kernel_launches"]
        qbytes += st["queue_bytes"]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.result()
    ms_dev = allmax(dev_ms) / args.steps                 # CUDA events on the rendering stream, max over ranks
    per_rank_ms = [dev_ms / args.steps]
    if dist is not None:
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = dev_ms / args.steps
        dist.all_reduce(t)
        per_rank_ms = [float(x) for x in t.tolist()]
    ms_wall = allmax(t_wall * 1e3) / args.steps
    tot_samples = allsum(samples) / args.steps
    tot_bounces = allsum(bounces) / args.steps
    tot_launches = int(allsum(launches))
    allsum_active = allsum(st["active_pixels"])
    my_bounces_per_step = bounces / args.steps

    # ---- end to end through host buffers: upload from pinned memory + kernels + gather + download, every step
    pinned = pyipt.PinnedArray((H, W, 3), np.float32) if rank == 0 else None   # the caller's frame buffer, page-locked
    frame = pinned.array if pinned else None
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    phase = [0.0, 0.0, 0.0, 0.0]                        # this rank's wall time in set_scene / render / barrier / download
    for _ in range(args.steps):
        ta = time.perf_counter()
        ctx.set_scene(hs)
        tb = time.perf_counter()
        st = step()
        tc = time.perf_counter()
        barrier()
        td = time.perf_counter()
        if rank == 0:
            ctx.download(out=frame)
        te = time.perf_counter()
        for k, v in enumerate((tb - ta, tc - tb, td - tc, te - td)):
            phase[k] += v * 1e3 / args.steps
    barrier()
    e2e_ms = allmax((time.perf_counter() - t0) * 1e3) / args.steps
    h2d = int(ctx_last(pyipt, ctx, "h2d"))
    d2h = int(H * W * 3 * 4)

    if rank == 0:
        sm_count = 148
        clk = (clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fp32_peak = sm_count * FP32_LANES_PER_SM * 2 * (clocks["sm_max_mhz"] or 1965) * 1e6 / 1e12
        # per-GPU figures of rank 0 (the kernels are the same on every rank)
        r0_ms = dev_ms / args.steps
        ach_fp32 = my_bounces_per_step * wl["flops"] / (r0_ms * 1e-3) / 1e12
        # HBM: the ray-queue bytes this design moves (records written + read back, counted on the device); a pass that
        # advances rays k bounces in registers moves 96/k bytes per bounce
        my_qbytes_per_step = qbytes / args.steps
        ach_hbm = my_qbytes_per_step / (r0_ms * 1e-3) / 1e9
        bytes_per_bounce = my_qbytes_per_step / max(1.0, my_bounces_per_step)
        traffic = wi = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
            traffic = tr.get(args.workload, {}).get("dram_bytes_per_queue_byte")
            wi = tr.get(args.workload, {}).get("warp_instructions_per_bounce")
        except Exception:
            pass
        roof_fp32 = {"bound": "fp32", "achieved": ach_fp32, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_fp32 / fp32_peak,
                     "traffic": None, "peak_source": "148 SM x 128 lanes x 2 x max SM clock (no measured fp32 figure in MEASURED_PEAKS.json)",
                     "frac_at_observed_clock": ach_fp32 / (sm_count * FP32_LANES_PER_SM * 2 * clk * 1e6 / 1e12),
                     "algorithmic_flops_per_bounce": wl["flops"]}
        roof_hbm = {"bound": "hbm", "achieved": ach_hbm, "peak": hbm_peak, "unit": "GB/s", "frac": ach_hbm / hbm_peak,
                    "traffic": None if traffic is None else traffic * my_qbytes_per_step / max(1, launches / args.steps),
                    "peak_source": hbm_src, "algorithmic_bytes_per_bounce": bytes_per_bounce,
                    "one_bounce_per_pass_equivalent_gbs": my_bounces_per_step * BYTES_PER_BOUNCE / (r0_ms * 1e-3) / 1e9,
                    "kernel": ("k_extend_bvh + k_bounce<MODE_SHADE>" if wl["scene"] == "synthetic1m" else "k_bounce_fast") +
                              "; achieved = ray-queue bytes written + read (ipt_stats.queue_bytes) / sum of launch durations; "
                              "traffic = measured DRAM bytes (ncu) per queue byte x queue bytes per launch"}
        binding = roof_hbm if roof_hbm["frac"] >= roof_fp32["frac"] else roof_fp32
        # what actually limits the typed-list kernel: warp-instruction issue slots (4 schedulers per SM, one per clock)
        roof_issue = None
        if wi:
            issue_peak = sm_count * 4 * (clocks["sm_max_mhz"] or 1965) * 1e6 / 1e9
            ach_issue = my_bounces_per_step * wi / (r0_ms * 1e-3) / 1e9
            roof_issue = {"bound": "issue", "achieved": ach_issue, "peak": issue_peak, "unit": "G warp-inst/s", "frac": ach_issue / issue_peak,
                          "warp_instructions_per_bounce": wi, "source": "instruction count of the ncu capture in profiles/ (smsp__inst_executed.sum / rays / bounces)"}
        line = {
            "metric": "Msamples/s", "value": tot_samples / (ms_dev * 1e-3) / 1e6, "unit": "Msamples/s",
            "gbounces_per_s": tot_bounces / (ms_dev * 1e-3) / 1e9,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "ms_per_step_wall": ms_wall,
            "ms_per_step_per_rank": per_rank_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.fp64 else "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "frame": [W, H], "spp": wl["spp"], "max_depth": wl["depth"],
                       "samples_per_step": int(tot_samples), "traced_bounces_per_step": int(tot_bounces),
                       "pixels_with_camera_rays": int(allsum_active), "pixels": W * H,
                       "parallelism": f"tiles 64x32 interleaved over {world} GPU(s); {gather}",
                       "l2": "inputs larger than L2: each wavefront batch streams ray queues of up to 2 x 6 GB (64 Mi-sample batches, 48 B per ray), up to 8 bounces per ray between two queue round trips", "rng": "philox4x32-10 keyed by pixel/sample/bounce"},
            "e2e": {"value": tot_samples / (e2e_ms * 1e-3) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms,
                    "rank0_ms": {"set_scene": phase[0], "render_call": phase[1], "wait_for_ranks": phase[2], "download": phase[3]}},
            "gpu_launches": tot_launches,
            "clocks": clocks,
            "roofline": binding, "roofline_fp32": roof_fp32, "roofline_hbm": roof_hbm,
        }
        if roof_issue:
            line["roofline_issue"] = roof_issue
        if not wl["flops"]:
            line["roofline"] = roof_hbm
            line.pop("roofline_fp32")
        if world == 1 and not args.no_cpu_baseline and wl["scene"] != "synthetic1m":
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line))
        frame = None
        pinned.close()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def shard(pyipt, W, H, tile_w, tile_h, rank, world):
    """Tiles (and pixels) of `rank` under the library's static interleaved schedule (ipt_tile_owner)."""
    tiles_x, tiles_y = (W + tile_w - 1) // tile_w, (H + tile_h - 1) // tile_h
    L = pyipt.lib()
    tiles = [(tx, ty) for ty in range(tiles_y) for tx in range(tiles_x) if L.ipt_tile_owner(tx, ty, tiles_x, world) == rank]
    pixels = sum((min(W, (tx + 1) * tile_w) - tx * tile_w) * (min(H, (ty + 1) * tile_h) - ty * tile_h) for tx, ty in tiles)
    return tiles, pixels


def dry_run(args, wl, rank, world, pyipt):
    """The N > 1 host logic on CPU (gloo): every rank computes its shard, the 64-byte frame handle goes round, sums and
    maxima are reduced, rank 0 prints one line.  No rendering, no GPU: used by tests/test_distributed_cpu.py."""
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("gloo")
    W, H = wl["width"] or 1280, wl["height"] or 720
    tiles, pixels = shard(pyipt, W, H, 64, 32, rank, world)
    handle = [bytes(range(64)) if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(handle, src=0)
    assert handle[0] == bytes(range(64))
    tot = torch.tensor([float(pixels) * wl["spp"], float(len(tiles))], dtype=torch.float64)
    mx = torch.tensor([float(rank + 1)], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"dry_run": True, "n_gpus": world, "samples_per_step": int(tot[0].item()), "tiles": int(tot[1].item()),
                          "max_rank_plus_1": int(mx.item()), "my_tiles": len(tiles), "frame": [W, H], "scaling": "strong"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ctx_last(pyipt, ctx, what):
    # bytes uploaded by the last ipt_ctx_set_scene: geometry (fp64+fp32), materials, slot ids, BVH nodes
    v = ctx.scene.view.contents
    n = v.n_objects
    return n * (16 * 8 + 16 * 4 + 8 * 8 + 8 * 4) + ((n + 3) // 4) * 16 + v.n_bvh_nodes * 64


if __name__ == "__main__":
    main()
