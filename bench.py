#!/usr/bin/env python3
"""Benchmark of the path-tracing bounce loop: Mpath-samples/s (+ Grays/s) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA renderer
  python bench.py --impl reference [...]                         the reference's own CPU renderer (oracle/_ref)
  torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU

Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[1] -- Cornell box (scene 5), 1920x1080, 1024 spp, 32 bounces.
`per_config` carries the other four BASELINE configs (C1, C3, C4, C5) on the same GPUs: device-timed Mpaths/s, Grays/s, kernel
time, roofline fraction, live-lane share, end-to-end figure and the steps / sizes actually run (C5 = 4K x 4096 spp is run in full
on 8 GPUs and with a stated, reduced sample count otherwise).

A step = one full render of the workload.  With N GPUs the samples per pixel are split into N slices of disjoint PCG streams
(each rank renders the whole frame for its slice), the float4 accumulators are sum-reduced with one NCCL all-reduce and finalised
(mean over finite samples + luminance clamp) -- total work is fixed, so scaling is "strong".  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene, width, height, spp, depth)
    "C1": (0, 500, 500, 16, 32),
    "C2": (5, 1920, 1080, 1024, 32),
    "C3": (6, 1920, 1080, 1024, 32),
    "C4": (7, 1920, 1080, 4096, 32),
    "C5": (8, 3840, 2160, 4096, 32),
}
WORKLOAD_DESC = {
    "C1": "'In One Weekend' random spheres 500x500, 16 spp, 32 bounces",
    "C2": "Cornell box (rect/box geometry, mixture-pdf light sampling) 1920x1080, 1024 spp, 32 bounces",
    "C3": "Cornell box with smoke 1920x1080, 1024 spp, 32 bounces",
    "C4": "'The Next Week' final scene 1920x1080, 4096 spp, 32 bounces",
    "C5": "triangle meshes (bunny + teapot) 3840x2160, 4096 spp, 32 bounces",
}
FP32_LANES_PER_SM = 128


def alg_flop_per_ray():
    """Algorithmic flop per ray of each config, derived from operation counters by tools/alg_flops.py."""
    d = json.load(open(os.path.join(ROOT, "tools", "alg_flops.json")))
    return {k: float(v["flop_per_ray"]) for k, v in d["configs"].items()}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def config_dict(name, scene, W, H, spp, depth):
    """The SAME dict in both arms (the driver compares them)."""
    return {"workload": name + ": " + WORKLOAD_DESC[name], "scene": scene, "width": W, "height": H, "spp": spp, "max_bounces": depth}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, n_gpus=1):
        super().__init__(daemon=True)
        self.gpu = gpu_index        # the GPU whose median clock is reported as `sm_mhz`
        self.n_gpus = n_gpus        # GPUs 0 .. n_gpus-1 are sampled (one process per GPU: rank 0 watches them all)
        self.samples = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(i) for i in range(max(self.n_gpus, self.gpu + 1))), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons, per_gpu = [], [], set(), {}
        for s in self.samples:
            try:
                per_gpu.setdefault(int(s[0]), []).append(float(s[1]))
                if int(s[0]) == self.gpu:
                    sm.append(float(s[1]))
                mx.append(float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        if len(per_gpu) > 1:   # median SM clock of every GPU of the job
            out["sm_mhz_per_gpu"] = [sorted(v)[len(v) // 2] for _, v in sorted(per_gpu.items())]
        return out


# ----------------------------------------------------------------------------------------------- reference arm
_REF_BIN = None


def reference_binary():
    """The reference built with ITS OWN flags (-O3 -march=native -fno-exceptions -fno-rtti, host libm; oracle/build_ref.sh).
    The -march=native build is this repo's build container's; where the box's CPU cannot run it the x86-64-v3 twin is used."""
    global _REF_BIN
    if _REF_BIN:
        return _REF_BIN
    run_dir = os.path.join(ROOT, "assets", "run")
    os.makedirs(run_dir, exist_ok=True)
    for name in ("mrt_ref_native", "mrt_ref_v3", "mrt_ref"):
        p = os.path.join(ROOT, "oracle", "_ref", name)
        if not os.path.exists(p):
            continue
        try:
            r = subprocess.run([p, "stock", "-scene", "5", "-width", "32", "-height", "18", "-samples", "1", "-depth", "4", "-mode", "0", "-threads", "1"],
                               cwd=run_dir, capture_output=True, text=True, timeout=60)
            if r.returncode == 0:
                _REF_BIN = (p, name)
                return _REF_BIN
        except Exception:
            continue
    raise RuntimeError("oracle/_ref/mrt_ref* missing or not runnable (built by __graft_entry__.build() where /root/reference exists)")


def run_reference_cpu(scene, width, height, sample_spp, depth, threads):
    """The reference's own multithreaded CPU renderer (its main(), `-mode 0`), timed by its own clock
    (main.cpp:375,394-405).  Test/benchmark infrastructure under oracle/."""
    binary, name = reference_binary()
    out = subprocess.run([binary, "stock", "-scene", str(scene), "-width", str(width), "-height", str(height), "-samples", str(sample_spp),
                          "-depth", str(depth), "-mode", "0", "-threads", str(threads)], cwd=os.path.join(ROOT, "assets", "run"), check=True,
                         capture_output=True, text=True)
    r = json.loads(out.stdout.strip().splitlines()[-1])
    r["binary"] = name
    return r


def cpu_sample_spp(name, W, H, spp, want):
    """Samples per pixel of the bounded CPU sample: `want` for the 1080p configs, fewer at 4K, the config's own spp if smaller."""
    s = want if W * H <= 1920 * 1080 else max(4, want // 4)
    return min(s, spp)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1

    def measure(name, steps, warmup):
        scene, W, H, spp, depth = WORKLOADS[name]
        sample_spp = cpu_sample_spp(name, W, H, spp, args.cpu_spp)
        vals, rays_vals, secs, binary = [], [], [], ""
        for i in range(warmup + steps):
            r = run_reference_cpu(scene, W, H, sample_spp, depth, cores)
            binary = r["binary"]
            if i >= warmup:
                vals.append(r["mpaths_per_s"]); rays_vals.append(r["mrays_per_s"]); secs.append(r["trace_seconds"])
        sample = (f"{W}x{H}, {sample_spp} of {spp} spp per step (throughput is spp-independent), reference -mode 0 -threads {cores}, "
                  f"binary {binary} (reference's own flags, host libm)")
        return sum(vals) / len(vals), sum(rays_vals) / len(rays_vals) * 1e-3, 1000.0 * sum(secs) / len(secs), sample

    name = args.workload
    scene, W, H, spp, depth = WORKLOADS[name]
    v, grays, ms, sample = measure(name, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "Mpath-samples/s", "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (reference scene built from the reference's seed)",
        "config": config_dict(name, scene, W, H, spp, depth),
        "grays_per_s": grays,
        "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_per_config:
        per = {}
        for other in WORKLOADS:
            if other == name:
                continue
            ov, og, oms, osample = measure(other, 1, 0)
            per[other] = {"value": ov, "unit": "Mpaths/s", "grays_per_s": og, "ms_per_step": oms, "steps": 1, "sample": osample,
                          "config": config_dict(other, *WORKLOADS[other])}
        line["per_config"] = per
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------ our arm
def scene_bytes(desc):
    d = desc.contents
    return 16 * (3 * d.n_sphere + 2 * d.n_rect + 2 * d.n_list + 2 * d.n_bvh + 4 * d.n_node2 + 6 * d.n_tri + d.n_xlate + 3 * d.n_rot +
                 d.n_vol + d.n_mat + d.n_tex + (256 if d.perlin_vec else 0)) + 4 * (d.n_child + d.n_lights + 2 * d.n_trileaf + (768 if d.perlin_perm else 0)) + \
        int(d.n_image_bytes)


class Workload:
    """One BASELINE config on this rank: resident scene, bound accumulator, device-timed steps and an end-to-end step."""

    def __init__(self, name, rank, world, local_rank, spp_override=0, size_override=""):
        import torch
        from miniraytracer_b200 import api, distributed as mdist
        self.torch, self.api = torch, api
        self.name, self.rank, self.world, self.local_rank = name, rank, world, local_rank
        self.scene, self.W, self.H, self.spp, self.depth = WORKLOADS[name]
        self.reduced = ""
        if spp_override:
            self.spp = spp_override
            self.reduced = f" [REDUCED spp={self.spp}]"
        if size_override:
            self.W, self.H = (int(v) for v in size_override.lower().split("x"))
            self.reduced += f" [REDUCED {self.W}x{self.H}]"
        self.N = api.grid_samples(self.spp)
        self.s_begin, self.s_end = mdist.shard_range(self.N, rank, world)
        self.dev = torch.device("cuda", local_rank)
        self.hs = api.HostScene(self.scene, self.W, self.H)
        self.r = api.Renderer(self.hs, local_rank)
        self.stream = torch.cuda.current_stream()
        self.r.set_stream(self.stream.cuda_stream)
        self.acc = torch.zeros((self.H, self.W, 4), dtype=torch.float32, device=self.dev)
        self.final = torch.empty_like(self.acc)
        self.r.bind_accumulator(self.acc.data_ptr(), self.W, self.H)
        self.host_out = torch.empty((self.H, self.W, 4), dtype=torch.float32, pin_memory=True)
        self.paths_per_step = self.W * self.H * self.N

    def step_device(self, spp_slice=None):
        """Hot path with the scene resident in HBM: render slice -> (all-reduce) -> finalize."""
        import torch.distributed as dist
        b, e = (self.s_begin, self.s_end) if spp_slice is None else spp_slice
        self.r.render_async(self.W, self.H, self.spp, self.depth, sample_begin=b, sample_end=e)
        if self.world > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM)
        self.r.finalize_device(self.acc.data_ptr(), self.final.data_ptr(), self.W, self.H)

    def step_e2e(self):
        """Through the C ABI with HOST buffers: scene tables H2D (mrt_gpu_scene_upload) + render + reduce + finalize + frame D2H.
        Frames are double-buffered: the D2H of frame k runs on a copy stream while the host uploads frame k+1 and the GPU starts
        rendering it; a frame is retired (its copy awaited, its scene handle closed) one step later."""
        import torch.distributed as dist
        torch, api = self.torch, self.api
        if not hasattr(self, "_inflight"):
            self._inflight, self._frame = [], 0
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._final2 = [self.final, torch.empty_like(self.final)]
            self._host2 = [self.host_out, torch.empty((self.H, self.W, 4), dtype=torch.float32, pin_memory=True)]
        k = self._frame & 1
        trace = os.environ.get("MRT_BENCH_TRACE") == "1"   # host-side time of each sub-step (stderr), to see where the host blocks
        t = [time.perf_counter()]
        r2 = api.Renderer(self.hs, self.local_rank)
        t.append(time.perf_counter())
        r2.set_stream(self.stream.cuda_stream)
        r2.bind_accumulator(self.acc.data_ptr(), self.W, self.H)
        r2.render_async(self.W, self.H, self.spp, self.depth, sample_begin=self.s_begin, sample_end=self.s_end)
        t.append(time.perf_counter())
        if self.world > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM)
        t.append(time.perf_counter())
        done = torch.cuda.Event()
        if self.rank == 0:
            r2.finalize_device(self.acc.data_ptr(), self._final2[k].data_ptr(), self.W, self.H)
            ready = torch.cuda.Event()
            ready.record(self.stream)
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ready)
                self._host2[k].copy_(self._final2[k], non_blocking=True)
                done.record(self._copy_stream)
        else:
            done.record(self.stream)
        self._inflight.append((r2, done))
        self._frame += 1
        t.append(time.perf_counter())
        while len(self._inflight) > 1:
            self._retire()
        t.append(time.perf_counter())
        if trace:
            names = ["create+upload", "launch", "all_reduce call", "finalize+copy call", "retire previous"]
            sys.stderr.write(f"[e2e rank {self.rank} frame {self._frame}] " + " ".join(f"{n}={1e3 * (b - a):.2f}ms" for n, a, b in zip(names, t, t[1:])) + "\n")

    def _retire(self):
        r, done = self._inflight.pop(0)
        done.synchronize()
        r.close()

    def drain_e2e(self):
        while getattr(self, "_inflight", []):
            self._retire()
        self.host_out = self._host2[(self._frame - 1) & 1] if getattr(self, "_frame", 0) else self.host_out

    def close(self):
        self.r.close()
        self.hs.close()
        del self.acc, self.final, self.host_out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS), help="headline workload (default: BASELINE configs[1])")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (profiling only; marks the line as reduced)")
    ap.add_argument("--size", default="", help="WxH override (profiling only; marks the line as reduced)")
    ap.add_argument("--cpu-spp", type=int, default=16, help="spp of the bounded CPU-baseline sample (a quarter of it at 4K)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="headline workload only")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the renderer has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == max(1, args.gpus) or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"
    flops = alg_flop_per_ray()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def measure(wl, steps, warmup, e2e_steps, warm_slice=None):
        """Device-timed steps (CUDA events on the launch stream, L2 flushed between steps, max over ranks) + end-to-end steps."""
        for _ in range(warmup):
            wl.step_device(warm_slice)
            flush.zero_()
        barrier()
        stream = wl.stream
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kernel_ms, rays, iters, coop, busy, ssum = [], [], [], [], [], []
        t0 = time.perf_counter()
        for i in range(steps):
            ev[i][0].record(stream)
            wl.step_device()
            ev[i][1].record(stream)
            st = wl.r.stats()                    # waits for the render kernel of this step
            kernel_ms.append(st["kernel_ms"]); rays.append(st["rays"]); iters.append(st["warp_iterations"])
            coop.append((st["coop_trees"], st["coop_node_items"], st["coop_node_steps"], st["coop_leaf_items"], st["coop_leaf_steps"]))
            if st["warps"]:   # mode B: share of the launch the average warp was at work, share of that spent summing staged samples
                busy.append(st["warp_time_sum_ns"] / max(1, st["warps"] * st["warp_span_ns"]))
                ssum.append(st["stage_sum_ns"] / max(1, st["warp_time_sum_ns"]))
            flush.zero_()                        # L2 flush between timed steps (outside the event pairs)
        barrier()
        wall = time.perf_counter() - t0
        total_s = reduce_max(sum(a.elapsed_time(b) for a, b in ev)) / 1000.0
        k_mine = sum(kernel_ms) / len(kernel_ms)
        k_all = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        if world > 1:
            dist.all_gather(k_all, torch.tensor([k_mine], dtype=torch.float64, device=dev))
        rays_total = reduce_sum(float(sum(rays)))
        res = {
            "value": wl.paths_per_step * steps / total_s / 1e6, "ms_per_step": 1000.0 * total_s / steps,
            "grays_per_s": rays_total / total_s / 1e9, "rays_per_path": rays_total / (wl.paths_per_step * steps),
            "kernel_ms": sum(kernel_ms) / len(kernel_ms), "rays_per_launch": sum(rays) / len(rays),
            "live_lane_frac": sum(rays) / max(1.0, 32.0 * sum(iters)), "wall_s": wall, "steps": steps, "warmup": warmup,
        }
        if busy:
            res["warp_busy_frac"], res["stage_sum_frac"] = sum(busy) / len(busy), sum(ssum) / len(ssum)
        if world > 1:
            res["kernel_ms_per_rank"] = [round(float(t.item()), 3) for t in k_all]   # the step waits for the slowest GPU
        if coop[-1][0]:
            c = coop[-1]
            res["coop_trees"] = {"node_step_fill": c[1] / max(1.0, 32.0 * c[2]), "leaf_step_fill": c[3] / max(1.0, 32.0 * c[4])}
        if e2e_steps:
            for _ in range(min(2, e2e_steps)):   # warm: two frames are in flight at a time, so two sets of device buffers / pinned
                wl.step_e2e()                    # staging / streams must exist in the per-process caches before the clock starts
            wl.drain_e2e()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                wl.step_e2e()
            wl.drain_e2e()                       # every frame's D2H has landed in pinned host memory
            torch.cuda.synchronize()
            e2e_s = reduce_max(time.perf_counter() - t0)   # max over ranks (the all-reduce inside every frame keeps them in step;
            barrier()                                      #  a dist.barrier() costs tens of ms by itself and is not part of a frame)
            res["e2e"] = {"value": wl.paths_per_step * e2e_steps / e2e_s / 1e6, "unit": "Mpaths/s", "h2d_bytes_per_step": scene_bytes(wl.hs.desc),
                          "d2h_bytes_per_step": wl.W * wl.H * 16, "steps": e2e_steps,
                          "image_mean": float(wl.host_out[..., :3].double().mean()) if rank == 0 else None}
        return res

    def roofline_of(name, res, sm_count, max_mhz, peaks, clocks=None):
        fp32_peak = sm_count * FP32_LANES_PER_SM * 2 * max_mhz * 1e6 / 1e12
        achieved = res["rays_per_launch"] * flops[name] / (res["kernel_ms"] * 1e-3) / 1e12
        out = {
            "bound": "fp32",   # SM issue / FP32 pipe (north_star): not a dense contraction, scene is cache resident, HBM nearly idle
            "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
            "peak_source": (f"{sm_count} SMs x 128 FP32 lanes x 2 x {max_mhz:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json; no measured FP32 "
                            "figure exists there)") if peaks else "fallback 1965 MHz (MEASURED_PEAKS.json absent)",
            "alg_flop_per_ray": flops[name], "alg_flop_source": "tools/alg_flops.py (operation counters of the reference algorithm x SURVEY 8d weights)",
            "rays_per_launch": res["rays_per_launch"], "kernel_ms": res["kernel_ms"],
        }
        if clocks and clocks.get("sm_mhz"):
            out["frac_at_clock_under_load"] = achieved / (fp32_peak * clocks["sm_mhz"] / max_mhz)
        return out

    # ---------------------------------------------------------------- headline
    head = Workload(args.workload, rank, world, local_rank, args.spp, args.size)
    sampler = ClockSampler(local_rank, world)
    if rank == 0:
        sampler.start()
    hres = measure(head, args.steps, args.warmup, max(1, args.steps))
    clocks = sampler.stop() if rank == 0 else None
    sm_count = head.r.info.sm_count
    peaks = measured_peaks()
    max_mhz = float(peaks.get("sm_max_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)

    line = None
    if rank == 0:
        roof = roofline_of(args.workload, hres, sm_count, max_mhz, peaks, clocks)
        roof["kernel"] = "render_pixel_binned<cornell, 6 blocks/SM>" if args.workload == "C2" else "render_pixel_binned"
        # DRAM traffic of the dominant kernel: from ONE ncu --set full capture of this workload on one GPU (profiles/), not a
        # measurement of this run -- so only reported for the full-size single-GPU line
        traffic_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        traffic = None
        if world == 1 and not head.reduced and os.path.exists(traffic_file):
            traffic = json.load(open(traffic_file)).get(args.workload)
        roof["traffic"] = traffic["dram_bytes_per_launch"] if traffic else None
        roof["traffic_source"] = ("from_ncu_capture: " + traffic["capture"]) if traffic else None
        roof["note"] = ("bound = SM instruction issue on divergent code; 'achieved' counts the reference algorithm's flops per ray "
                        "(SURVEY 8d); the scene is L1/L2 resident, HBM carries the accumulator and the sample staging only")
        line = {
            "metric": "Mpath-samples/s", "value": hres["value"], "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": hres["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (reference scene rebuilt from the reference's seed; no external data)",
            "config": config_dict(args.workload, head.scene, head.W, head.H, head.N, head.depth),
            "reduced": head.reduced or None,
            "parallelism": f"spp-sharded x{world}",
            "l2": "256 MB memset between timed steps; scene tables (<1 MB) are cache resident by design",
            "grays_per_s": hres["grays_per_s"], "rays_per_path": hres["rays_per_path"], "live_lane_frac": hres["live_lane_frac"],
            "wall_s_timed_region": hres["wall_s"],
            "e2e": hres["e2e"],
            "kernel_ms_per_rank": hres.get("kernel_ms_per_rank"),
            # mode B, rank 0: share of the launch the average warp was at work (the rest: warps out of tickets waiting for the last
            # one) and share of that warp time spent adding up the staged samples at chunk ends (mrt_gpu_stats)
            "warp_busy_frac": hres.get("warp_busy_frac"), "stage_sum_frac": hres.get("stage_sum_frac"),
            "gpu_launches": args.steps * 2,   # render kernel + finalize kernel per step (NCCL's kernels not counted)
            "roofline": roof,
            "clocks": clocks,
        }
    head.close()

    # ---------------------------------------------------------------- the other BASELINE configs
    if not args.no_per_config and not args.spp and not args.size:
        per = {}
        for name in WORKLOADS:
            if name == args.workload:
                continue
            spp_override = 0
            if name == "C5" and world < 8:
                spp_override = {1: 256, 2: 576, 4: 1024}.get(world, 256)   # 4K x 4096 spp is ~75 s per frame on one GPU: reduced and SAID so (full on 8 GPUs)
            wl = Workload(name, rank, world, local_rank, spp_override)
            heavy = wl.paths_per_step > 3e9
            # heavy configs: one timed step; the warm-up renders a 64-sample slice (kernel and caches warm, clocks up)
            warm_slice = (wl.s_begin, min(wl.s_end, wl.s_begin + 64)) if heavy else None
            res = measure(wl, 1 if heavy else 3, 1 if heavy else 2, 1 if heavy else 3, warm_slice)
            if rank == 0:
                entry = {"value": res["value"], "unit": "Mpaths/s", "grays_per_s": res["grays_per_s"], "kernel_ms": res["kernel_ms"],
                         "ms_per_step": res["ms_per_step"], "rays_per_path": res["rays_per_path"], "lanes": res["live_lane_frac"],
                         "lanes_note": "share of lanes with a live path per segment step (mrt_gpu_stats.warp_iterations); lanes per instruction are in profiles/",
                         "roofline": {k: v for k, v in roofline_of(name, res, sm_count, max_mhz, peaks).items()
                                      if k in ("achieved", "peak", "frac", "alg_flop_per_ray", "unit")},
                         "e2e": res["e2e"], "steps": res["steps"], "warmup": res["warmup"],
                         "config": config_dict(name, wl.scene, wl.W, wl.H, wl.N, wl.depth), "reduced": wl.reduced or None}
                if "coop_trees" in res:
                    entry["coop_trees"] = res["coop_trees"]
                if "warp_busy_frac" in res:
                    entry["warp_busy_frac"], entry["stage_sum_frac"] = res["warp_busy_frac"], res["stage_sum_frac"]
                per[name] = entry
            wl.close()
        if rank == 0:
            line["per_config"] = per

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                cores = os.cpu_count() or 1
                scene, W, H, spp, depth = WORKLOADS[args.workload]
                cb = run_reference_cpu(scene, head.W, head.H, args.cpu_spp, depth, cores)
                line["cpu_baseline"] = {"value": cb["mpaths_per_s"], "unit": "Mpaths/s", "cores": cores, "kind": "reference",
                                        "grays_per_s": cb["mrays_per_s"] * 1e-3,
                                        "sample": f"{head.W}x{head.H}, {args.cpu_spp} of {head.N} spp, reference -mode 0 -threads {cores}, "
                                                  f"{cb['trace_seconds']:.1f} s, binary {cb['binary']} (reference's own flags, host libm)"}
            except Exception as e:   # the baseline is reported, never substituted
                line["cpu_baseline"] = {"value": None, "unit": "Mpaths/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
