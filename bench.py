#!/usr/bin/env python
"""bench.py -- Mpaths/s of the reinforcement-learned path tracing hot path (BASELINE.json metric).

One *step* = one frame of the reference's method-1 loop (G/main.cu:301-364): trace `spp` samples per pixel with
Expected-SARSA importance sampling and TD accumulation, all-reduce the Q accumulators (N > 1), merge into Q, rebuild the
CDFs. Default workload = BASELINE.json configs[1]: Cornell box 512x512, 12x12 radiance volumes, 32 spp per frame
(32 timed frames = 1024 spp). One *path* = one camera sample of one pixel traced to termination.

  value     whole-job Mpaths/s, device-timed (CUDA events on the library's stream, max over ranks), scene / Q-table /
            path queues already resident in HBM
  e2e       the same frames driven one by one through the C ABI the way the reference's frame loop runs: camera upload,
            render, frame-buffer download to pinned host memory and the statistics read-back, every step (G/main.cu:307-349)
  roofline  dominant kernel = the per-bounce tracing kernel (k_bounce): executed ray-triangle / ray-box tests (counted in
            the kernel) x 72 / 18 flop / its device time, against the FP32 FMA rate measured on this GPU
  cpu_baseline   BASELINE.json configs[0]: the reference's Old_CPU_Rendering_Engine (baseline/_ref/libcpu_engine_b*.so, built by
            oracle/cpu_engine/build.sh), built-in Cornell box, default path tracer, 16 spp, all host cores (and 6 threads as hard-coded)
  cpu_baseline_sarsa / --impl reference   this bench's own workload on the host: the reference's GPU engine sources compiled for
            the host (oracle/_ref/libref_host.so: G/path_tracing/reinforcement_path_tracing.cu etc. behind oracle/host_shim, OpenMP
            over pixels), same scene, method, resolution and 32 spp per frame, all host cores (set explicitly)

N > 1 (torchrun): sample-partitioned, weak scaling -- every rank traces `spp` samples of every pixel per frame
(global frame = N*spp samples, disjoint Philox sample indices), Q accumulators are all-reduced every frame (NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200"))

F_TRI, F_AABB = 72.0, 18.0          # flop per ray-triangle solve (reference form) / per ray-AABB slab test

WORKLOADS = {
    # name: (scene in tests/golden/scenes.npz, method, camera, env_light)
    "cornell_sarsa": ("cornell", 1, (0.0, 0.0, -3.0), 0.0),
    "cornell_default": ("cornell", 0, (0.0, 0.0, -3.0), 0.0),
    # BASELINE.json configs[2]: door_room.obj with its own light quad and colours (tests/golden/make_presets.py)
    "door_room_sarsa": ("door_room_lit", 1, (0.0, 0.5, -0.9), 0.0),
    "archway_sarsa": ("archway", 1, (-1.0, 0.2, -0.99), 0.0),
    # BASELINE.json configs[3]: the large mesh through the BVH. The reference has no preset for it: the
    # importer's commented normalisation (tests/golden/make_presets.py), lit by ENVIRONMENT_LIGHT = 1
    "medieval_default": ("medieval_norm", 0, (0.0, 0.0, -3.0), 1.0),
    "medieval_sarsa": ("medieval_norm", 1, (0.0, 0.0, -3.0), 1.0),
    # the same mesh seen from INSIDE its main room (86 % of the directions from the camera hit a wall; found with the oracle's closest
    # hit): paths bounce ~7 times before they leave through a window, so the BVH is walked by incoherent secondary rays
    "medieval_inside_default": ("medieval_norm", 0, (0.2, 0.3, -0.5), 1.0),
    "medieval_inside_sarsa": ("medieval_norm", 1, (0.2, 0.3, -0.5), 1.0),
    "complex_light_room_default": ("complex_light_room", 0, (0.0, 0.0, -0.9), 0.0),
    # BASELINE.json configs[4]: archway.obj, Neural-Q (method 3): the DQN fc_layer network (K = 918 -> 200 -> 300 -> 200 -> 144) trained online,
    # batch 4096 (G/main.cu:116-118), one pass over the pixels per frame; a step = one training frame. cornell_neuralq: the K = 342 network
    "archway_neuralq": ("archway", 3, (-1.0, 0.2, -0.99), 0.0),
    "cornell_neuralq": ("cornell", 3, (0.0, 0.0, -3.0), 0.0),
}
DQN_FLOP_PER_RAY = {342: 434400.0, 918: 664800.0}        # dense-equivalent forward flop per ray, 2 (K 200 + 200 300 + 300 200 + 200 144)


def load_scene(name):
    for f in ("scenes.npz", "scene_presets.npz"):
        z = np.load(os.path.join(ROOT, "tests", "golden", f))
        s = {k.split("/")[1]: z[k] for k in z.files if k.startswith(name + "/")}
        if s:
            return s
    raise SystemExit("bench.py: no scene %r in tests/golden" % name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        # NVML polled from a thread every ~2 ms (the timed region of the default run is only ~80 ms: nvidia-smi -lms 20 yields five samples);
        # nvidia-smi stays as the fallback when the binding is missing
        self.nvml_rows, self._stop = [], False
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop:
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0; rs = int(reasons_fn(h))
                        self.nvml_rows.append((time.perf_counter(), sm, smax, pw, rs))
                    except Exception:
                        break
                    time.sleep(0.002)
            self.nvml_t = threading.Thread(target=poll, daemon=True)
            self.nvml_t.start()
        except Exception:
            self.nvml_rows = None
        if self.nvml_rows is not None and not os.environ.get("RLPT_BENCH_SMI"):
            return                                             # NVML answers: no nvidia-smi child per rank (eight of them polling beside eight ranks is host noise)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        self._stop = True
        if self.nvml_rows:
            rows = [r for r in self.nvml_rows if t0 <= r[0] <= t1] or self.nvml_rows
            sm = sorted(r[1] for r in rows)
            bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            reasons = [n for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap") if any(r[4] & bits[n] for r in rows)]
            if self.proc:
                self.proc.terminate()
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][2], "power_w_max": max(r[3] for r in rows), "reasons": reasons, "samples": len(rows), "source": "NVML polled every 2 ms"}
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1] or [r for _, r in self.rows]
        if not rows:
            return None
        try:
            sm = sorted(float(r[0]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows), "reasons": reasons, "samples": len(rows)}
        except (ValueError, IndexError):
            return None


class c_stdout_to_stderr:
    """The reference printf()s progress text; keep stdout to the one JSON line by pointing fd 1 at stderr meanwhile."""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def set_host_threads(n):
    """OpenMP team size of the CPU arms, set explicitly: torch.distributed.run exports OMP_NUM_THREADS=1 to every rank."""
    import ctypes
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def cpu_reference_run(frames, warmup, want_seconds=None, spp=32):
    """The reference's SARSA frame loop on the host cores. Returns (Mpaths/s, kind, cores, sample, ms_per_step)."""
    with c_stdout_to_stderr():
        return _cpu_reference_run(frames, warmup, want_seconds, spp)


def _cpu_reference_run(frames, warmup, want_seconds=None, spp=32):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from checkers import Oracle, Reference
    set_host_threads(os.cpu_count() or 1)
    suffix = "" if spp == 32 else "_spp%d" % spp             # oracle/build_ref.sh: libref_host.so is the 512 x 512 x 32 spp build
    if Reference.available("host", suffix):
        R = Reference("host", suffix)
        R.scene_cornell(); R.camera(0.0, 0.0, -3.0)
        R.rmap_build()
        for _ in range(warmup):
            R.render_sarsa(1, 0)
        t0 = time.perf_counter(); done = 0
        for _ in range(frames):
            R.render_sarsa(1, 0); done += 1
            if want_seconds and time.perf_counter() - t0 > want_seconds:
                break
        dt = time.perf_counter() - t0
        paths = done * R.width * R.height * R.spp
        sample = "%d frames x %d spp, Cornell %dx%d, Expected SARSA, after %d warm-up frames (%.2f M paths); unmodified reference sources (GPU_Rendering_Engine) built for the host with g++ -O2 -fopenmp, %d threads" % (
            done, R.spp, R.width, R.height, warmup, paths / 1e6, R.threads())
        return paths / dt / 1e6, "reference", R.threads(), sample, dt / done * 1e3
    orc = Oracle()
    s = load_scene("cornell")
    orc.scene_set(s["sv"], s["srgb"], s["lv"], s["lrgb"]); orc.rmap_build(); orc.rmap_update_distributions(); orc.rmap_merge_frame()
    w = h = 512
    def frame(i):
        orc.render_frame(1, w, h, spp, sample0=i * spp, fma_mode=1, td_mode=1); orc.rmap_merge_frame(); orc.rmap_update_distributions()
    for i in range(warmup):
        frame(i)
    t0 = time.perf_counter(); done = 0
    for i in range(frames):
        frame(warmup + i); done += 1
        if want_seconds and time.perf_counter() - t0 > want_seconds:
            break
    dt = time.perf_counter() - t0
    paths = done * w * h * spp
    return paths / dt / 1e6, "port", orc.threads(), "%d frames x %d spp, Cornell 512x512, Expected SARSA (oracle port, OpenMP)" % (done, spp), dt / done * 1e3


def cpu_engine_run():
    """BASELINE.json configs[0]: the reference's Old_CPU_Rendering_Engine (restored and built by oracle/cpu_engine/build.sh into
    baseline/_ref/), built-in Cornell box, default path tracer; std::chrono around draw_default_path_tracing only."""
    import ctypes
    out = {}
    ncpu = os.cpu_count() or 1
    for key, lib, threads in (("as_committed_all_cores", "libcpu_engine_b2.so", ncpu), ("as_committed_6_threads", "libcpu_engine_b2.so", 6), ("bounces_80_all_cores", "libcpu_engine_b80.so", ncpu)):
        path = os.path.join(ROOT, "baseline", "_ref", lib)
        if not os.path.exists(path):
            return None
        L = ctypes.CDLL(path)
        w, h, spp, b = [ctypes.c_int() for _ in range(4)]
        L.cpu_engine_dims(*[ctypes.byref(x) for x in (w, h, spp, b)])
        sec = np.zeros(1)
        with c_stdout_to_stderr():
            L.cpu_engine_render_default(1, int(threads), sec.ctypes.data_as(ctypes.c_void_p), None)
        paths = w.value * h.value * spp.value
        out[key] = {"value": paths / sec[0] / 1e6, "unit": "Mpaths/s", "cores": int(min(threads, ncpu)), "threads": int(threads), "seconds": float(sec[0]),
                    "sample": "1 frame %dx%d x %d spp = %.2f M paths, MAX_RAY_BOUNCES %d, built-in Cornell box, default path tracer (uniform hemisphere sampling)" % (w.value, h.value, spp.value, paths / 1e6, b.value)}
    return out


def run_reference_arm(args, rank):
    if rank != 0:
        return
    v, kind, cores, sample, ms = cpu_reference_run(args.steps, args.warmup, want_seconds=240.0, spp=args.spp)
    line = {"impl": "reference", "metric": "Mpaths/s", "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "built-in Cornell box scene (no dataset)",
            "config": {"workload": "cornell_sarsa", "scene": "Cornell box (36 surfaces + 2 area lights)", "width": 512, "height": 512, "spp_per_frame": args.spp,
                       "radiance_volumes": 24526, "grid": "12x12", "max_bounces": 80,
                       "note": "CPU arm: the reference's GPU_Rendering_Engine sources compiled for the host (oracle/_ref/libref_host.so), same scene, method, resolution and spp per frame as the GPU arm; "
                               "a step = one full frame; at most 240 s of timed frames"},
            "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def run_neuralq(args, ctx, s, scene_name, cam, rank, world, local, fp32_peak):
    """BASELINE.json configs[4]: online Neural-Q training + rendering (NeuralQPathtracer, G/deep_learning/neural_q_pathtracer.cu:226-600).
    Single GPU here (the gradient all-reduce path is covered by the 2-GPU test)."""
    import torch
    batch = args.batch
    ctx.dqn_init(seed=1984)
    n_par, k_in = ctx.dqn_param_count()
    for _ in range(args.warmup):
        ctx.render_neuralq(1, batch=batch)
    ctx.sync(); torch.cuda.synchronize(); ctx.stats_reset(); ctx.sync()
    clocks = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    loss = ctx.render_neuralq(args.steps, batch=batch)
    ctx.sync(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    st = ctx.stats()
    clk = clocks.stop(t0, t1) if clocks else None
    dev_s = st["device_seconds"]
    value = st["paths"] / dev_s / 1e6
    pinned = torch.empty((args.width * args.height, 3), dtype=torch.float32, pin_memory=True); frame_np = pinned.numpy()
    ctx.camera_set(cam); ctx.render_neuralq(1, batch=batch); ctx.frame_download(frame_np); ctx.stats()
    ctx.sync(); ctx.stats_reset()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.camera_set(cam); ctx.render_neuralq(1, batch=batch); ctx.frame_download(frame_np); est = ctx.stats()
    ctx.sync(); torch.cuda.synchronize()
    e2e_value = est["paths"] / (time.perf_counter() - e0) / 1e6
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    fl = DQN_FLOP_PER_RAY.get(k_in, 2.0 * (k_in * 200 + 200 * 300 + 300 * 200 + 200 * 144))
    fwd_s = max(st["dqn_forward_seconds"], 1e-12)
    achieved = st["dqn_forward_rays"] * fl / fwd_s / 1e12
    steps_n = max(st["train_steps"], 1.0)
    line = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 operands, f32 accumulate (network); f32 (tracing)",
            "data": "bundled .obj geometry (no dataset); network starts from a seeded Glorot draw and is trained online by the frames themselves",
            "config": {"workload": args.workload, "scene": "%s (%d surfaces + %d area lights)" % (scene_name, len(s["sv"]), len(s["lv"])), "method": "Neural-Q (DQN fc_layer), train + render",
                       "width": args.width, "height": args.height, "spp_per_frame": args.spp, "frames": args.steps, "batch": batch, "dqn_inputs": k_in, "dqn_parameters": n_par, "max_bounces": 80,
                       "l2": "the Q matrix of one bounce (144 x W*H floats = %.0f MB) exceeds L2" % (144 * args.width * args.height * 4 / 1e6)},
            "mean_path_length": st["path_length_sum"] / max(st["paths"], 1), "loss_last_frame": loss,
            "optimiser_steps_per_frame": steps_n / args.steps, "us_per_optimiser_step": st["train_seconds"] / steps_n * 1e6,
            "train_share_of_frame": st["train_seconds"] / max(dev_s, 1e-12),
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": 48, "d2h_bytes_per_step": args.width * args.height * 12 + 23 * 8,
                    "timed": "wall clock around K x (camera_set, render 1 training frame, frame download to pinned host memory, stats read-back)"},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": {"kernel": "k_dqn_forward (per-bounce network evaluation of every ray: layer 1 in fp32 as a rank-3 update, layers 2-4 tcgen05.mma bf16 -> fp32 in TMEM)",
                         "bound": "tensor", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s", "frac": achieved / bf16_peak, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if "bf16_tflops_sustained" in peaks else "fallback 1400 TFLOP/s sustained (B200_PROFILING.md)",
                         "flop_per_launch": st["dqn_forward_rays"] * fl / max(st["dqn_forward_launches"], 1), "avg_launch_ms": fwd_s / max(st["dqn_forward_launches"], 1) * 1e3,
                         "launches": st["dqn_forward_launches"], "share_of_step": fwd_s / max(dev_s, 1e-12),
                         "note": "algorithmic flop = dense-equivalent forward flop per ray (SURVEY 8d: %.0f for K = %d) x rays evaluated; duration = CUDA event pairs around every per-bounce launch" % (fl, k_in)},
            "clocks": clk}
    emit(line)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cornell_sarsa", choices=sorted(WORKLOADS))
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--spp", type=int, default=32)
    ap.add_argument("--traversal", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exclusive", action="store_true")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: --spp is the GLOBAL frame's samples per pixel, each rank traces spp / N of them")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    # stdout carries exactly one JSON line: everything libraries print there (NCCL's version banner, the reference's
    # progress text) goes to stderr; the line itself is written to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global emit
    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if args.impl == "reference":
        return run_reference_arm(args, rank)
    if args.warmup < 3:
        args.warmup = 3                                    # timing rule: at least 3 warm-up steps
    import torch
    import torch.distributed as dist
    import rlpt
    from rlpt.dist import torch_allreduce_hook
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene_name, method, cam, env = WORKLOADS[args.workload]
    s = load_scene(scene_name)
    if method == 3 and args.spp == 32:
        args.spp = 1                                           # Neural-Q: one pass over the pixels per training frame
    global_spp = args.spp * world if args.scaling == "weak" else args.spp
    if args.scaling == "strong":
        if args.spp % world:
            raise SystemExit("bench.py --scaling strong: --spp must be a multiple of the GPU count")
        args.spp //= world                                     # this rank's share of the global frame
    ctx = rlpt.Context(local, width=args.width, height=args.height, spp=args.spp, max_bounces=80, env_light=env, traversal=args.traversal,
                       rank=rank, world_size=world)
    ctx.scene_upload(s["sv"], s["srgb"], s["lv"], s["lrgb"])
    ctx.camera_set(cam)
    nv = ctx.radiance_map_build() if method == 1 else 0
    exchange = "none"
    if world > 1:
        ctx.set_allreduce(torch_allreduce_hook(local))
        exchange = "nccl all-reduce of the Q accumulators (torch.distributed), then the merge kernel"
        if method == 1 and os.environ.get("RLPT_EXCHANGE", "p2p") == "p2p":
            from rlpt.dist import p2p_setup
            if p2p_setup(ctx):
                exchange = "fused exchange + merge kernel over peer memory (CUDA IPC, P2P loads/stores over NVLink); no collective call per frame"
    fp32_peak = ctx.measure_fp32_peak()
    render = ctx.render_sarsa if method == 1 else ctx.render_default
    if method == 3:
        return run_neuralq(args, ctx, s, scene_name, cam, rank, world, local, fp32_peak)

    def barrier():
        ctx.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    render(args.warmup)
    # ---- timed region: K frames enqueued back to back, device-timed inside the library (events on its stream)
    barrier(); ctx.stats_reset(); ctx.sync()
    clocks = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    render(args.steps)
    barrier()
    t1 = time.perf_counter()
    st = ctx.stats()
    clk = clocks.stop(t0, t1) if clocks else None
    dev_s = torch.tensor([st["device_seconds"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dev_s, op=dist.ReduceOp.MAX)
    dev_s = float(dev_s.item())
    paths_rank = st["paths"]
    paths_total = paths_rank * world
    value = paths_total / dev_s / 1e6

    # ---- e2e: the reference's per-frame protocol through the C ABI with host buffers
    pinned = torch.empty((args.width * args.height, 3), dtype=torch.float32, pin_memory=True)
    frame_np = pinned.numpy()
    for _ in range(2):                                      # untimed: first use of the download path (staging buffers, page faults of the pinned block)
        ctx.camera_set(cam); render(1); ctx.frame_download(frame_np); ctx.stats()
    barrier(); ctx.stats_reset()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.camera_set(cam)                                 # cudaMemcpy(device_camera, &camera) every frame (G/main.cu:307)
        render(1)
        ctx.frame_download(frame_np)                        # cudaMemcpy(host_buffer, device_buffer) (G/main.cu:349)
        est = ctx.stats()                                   # path lengths / zero-contribution read-back (G/main.cu:322-339)
    barrier()
    e1 = torch.tensor([time.perf_counter() - e0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e1, op=dist.ReduceOp.MAX)
    e2e_value = est["paths"] * world / float(e1.item()) / 1e6
    # ---- exclusive per-kernel times: the two sample lanes (and the next frame's primary scan) overlap on purpose, so the per-launch
    # durations above include sharing the SMs. One short pass with a single lane and no cross-frame overlap gives each kernel's time alone.
    excl = None
    if rank == 0 and world == 1 and not args.no_exclusive:
        os.environ["RLPT_LANES"] = "1"; os.environ["RLPT_PRE"] = "0"
        cx = rlpt.Context(local, width=args.width, height=args.height, spp=args.spp, max_bounces=80, env_light=env, traversal=args.traversal)
        cx.scene_upload(s["sv"], s["srgb"], s["lv"], s["lrgb"]); cx.camera_set(cam)
        if method == 1:
            cx.radiance_map_build()
        rx = cx.render_sarsa if method == 1 else cx.render_default
        rx(max(args.warmup, 4)); cx.sync(); cx.stats_reset(); rx(8); cx.sync()
        sx = cx.stats()
        excl = {"k_isect": sx["isect_seconds"] / 8 * 1e3, "k_shade": sx["shade_seconds"] / 8 * 1e3, "tail": sx["tail_seconds"] / 8 * 1e3,
                "merge": sx["merge_seconds"] / 8 * 1e3, "frame": sx["device_seconds"] / 8 * 1e3,
                "note": "8 frames traced as ONE lane with no cross-frame overlap (RLPT_LANES=1 RLPT_PRE=0): kernels run back to back, so these are exclusive times; the headline runs two overlapping lanes"}
        cx.close()
        del os.environ["RLPT_LANES"]; del os.environ["RLPT_PRE"]
    h2d = 48                                                # FrameDyn: camera position + rotation + sample base, staged by the render call
    d2h = args.width * args.height * 3 * 4 + 8 * 8          # frame buffer + statistics block

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        flops = st["triangle_tests"] * F_TRI + st["box_tests"] * F_AABB
        split = st["isect_launches"] > 0
        # closest-hit kernel: k_isect when the split pipeline runs (its own event pairs), else the fused k_bounce (trace phase)
        # (per-kernel event pairs are taken on every fourth frame: a run of fewer than four steps has none and falls back to the trace phase)
        hit_s = max(st["isect_seconds"] if (split and st["isect_seconds"] > 0) else st["trace_seconds"], 1e-12)
        hit_n = st["isect_launches"] if split else st["kernel_launches"] - (st["frames"] if method == 1 else 0)
        tail_flop_share = 0.0
        if split and st["tail_launches"] > 0:
            # the run-to-completion launches execute triangle tests too; their flops are taken out in proportion to ray casts
            tail_flop_share = min(1.0, st["tail_seconds"] / max(st["tail_seconds"] + st["isect_seconds"] + st["shade_seconds"], 1e-12))
        hit_flops = flops * (1.0 - tail_flop_share)
        achieved = hit_flops / hit_s / 1e12
        kernel_total = max(st["isect_seconds"] + st["shade_seconds"] + st["tail_seconds"], 1e-12)
        # measured DRAM traffic per ray from the committed ncu --set full capture (profiles/r1_traffic.json), scaled to this run's average launch
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        except (OSError, ValueError):
            pass
        rays_per_launch = st["ray_casts"] * (1.0 - tail_flop_share) / max(hit_n, 1)
        def traffic_of(kernel):
            return traffic[kernel]["bytes_per_ray"] * rays_per_launch if (split and kernel in traffic and args.workload == "cornell_sarsa") else None
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "built-in Cornell box scene / bundled .obj geometry (no dataset); radiance volumes start untrained, trained by the %d warm-up frames" % args.warmup,
            "config": {"workload": args.workload, "scene": "%s (%d surfaces + %d area lights)" % (scene_name, len(s["sv"]), len(s["lv"])),
                       "method": "Expected SARSA radiance volumes, train + render" if method == 1 else "default path tracer",
                       "width": args.width, "height": args.height, "spp_per_frame": args.spp, "spp_per_global_frame": global_spp, "frames": args.steps, "spp_total_per_gpu": args.spp * args.steps,
                       "radiance_volumes": nv, "grid": "12x12", "max_bounces": 80, "partition": "samples (rank r traces samples r*spp..(r+1)*spp-1 of each global frame)", "exchange": exchange,
                       "l2": "inputs larger than L2: path queues %.2f GB + Q-table %.0f MB per GPU" % (args.width * args.height * args.spp * 60 * 2 / 1e9, nv * 144 * 20 / 1e6)},
            "radiance_map_build_ms": ({k: v * 1e3 for k, v in ctx.radiance_map_build_seconds().items()} if method == 1 else None),
            "mean_path_length": st["path_length_sum"] / max(st["paths"], 1), "mray_casts_per_s": st["ray_casts"] * world / dev_s / 1e6,
            "zero_contribution_fraction": st["zero_contribution_paths"] / max(st["paths"], 1),
            "kd_search_fallback_fraction": st.get("kd_fallbacks", 0.0) / max(st["ray_casts"], 1),
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "timed": "wall clock around K x (camera_set, render 1 frame, frame download to pinned host memory, stats read-back), synchronised on both sides"},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": {"kernel": "k_isect (closest hit of every live ray, one launch per bounce)" if split else "k_bounce (closest hit + shade + SARSA step + compaction)",
                         "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp32_peak if fp32_peak else None,
                         "traffic": traffic_of("k_isect"),
                         "note": "bound is the FP32 pipe (SURVEY 8d; not hbm/tensor): algorithmic flop = ray-triangle tests x 72 + ray-box tests x 18, counted in the kernel; duration = CUDA event pairs "
                                 "around every launch on its stream inside the timed region (launches of the two sample lanes overlap, so a launch's duration includes sharing the SMs)",
                         "peak_source": "FP32 FMA microbenchmark run by this bench on this GPU (MEASURED_PEAKS.json has no FP32 figure; nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4)",
                         "flop_per_launch": hit_flops / max(hit_n, 1), "avg_launch_ms": hit_s / max(hit_n, 1) * 1e3, "launches": hit_n,
                         "triangle_tests": st["triangle_tests"], "box_tests": st["box_tests"],
                         "share_of_kernel_time": (st["isect_seconds"] / kernel_total) if split else 1.0},
            "clocks": clk,
        }
        if split and st["shade_launches"] > 0:
            casts = st["ray_casts"] * (1.0 - tail_flop_share)
            survive = max(0.0, 1.0 - st["paths"] / max(st["ray_casts"], 1))
            per_ray = 60.0 + 52.0 * survive + (16.0 + 100.0 + 80.0 if method == 1 else 0.0) + 16.0 * (1.0 - survive)
            shade_bytes = casts * per_ray
            line["roofline_shade"] = {"kernel": "k_shade (nearest volume, TD target, direction sampling, compaction)", "bound": "hbm",
                                      "achieved": shade_bytes / max(st["shade_seconds"], 1e-12) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": shade_bytes / max(st["shade_seconds"], 1e-12) / 1e9 / hbm_peak, "traffic": traffic_of("k_shade"), "peak_source": hbm_src,
                                      "bytes_per_ray": per_ray, "avg_launch_ms": st["shade_seconds"] / st["shade_launches"] * 1e3, "launches": st["shade_launches"],
                                      "note": "algorithmic bytes per ray: 60 path state + hit in, 52 x survival out, 16 TD accumulate, 100 CDF (row ends, one row, irradiance), 80 nearest-volume "
                                              "(one table slot + 4 candidates), 16 x termination frame-buffer add; traffic = DRAM bytes per launch from the ncu capture in profiles/ (lower: the tables stay in L2)",
                                      "share_of_kernel_time": st["shade_seconds"] / kernel_total}
            line["tail"] = {"kernel": "k_bounce run-to-completion (paths left after the planned per-bounce launches)", "seconds": st["tail_seconds"], "launches": st["tail_launches"],
                            "share_of_kernel_time": st["tail_seconds"] / kernel_total}
        if method == 1 and st["merge_seconds"] > 0:
            merge_bytes = nv * 144 * 4 * 4.0                 # per frame: read Q + accumulator counts, write Q-derived CDF (+ sums/visits for touched cells): >= 4 arrays
            line["roofline_merge"] = {"kernel": "k_merge_cdf", "bound": "hbm", "achieved": merge_bytes * st["frames"] / st["merge_seconds"] / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": merge_bytes * st["frames"] / st["merge_seconds"] / 1e9 / hbm_peak, "traffic": None, "peak_source": hbm_src,
                                      "note": "duration = the merge phase of the frame on the context stream, during which the next frame's primary k_isect runs on purpose "
                                              "(low-priority stream); k_merge_cdf alone takes 40-48 us per launch = 1.2-1.4 TB/s (profiles/r1_launches_cornell_sarsa.csv)"}
        # image parity and the reference's own kernels on this GPU type: recorded by tests/test_gpu_mape.py on a B200 (bench.py itself
        # runs no checker code outside the cpu_baseline leg)
        for name in ("r2_mape_parity.json", "r1_mape_parity.json"):
            try:
                mp = json.load(open(os.path.join(ROOT, "profiles", name)))
            except (OSError, ValueError):
                continue
            key = "cornell" if "cornell" in mp else None
            m = mp[key] if key else mp
            line["mape"] = {"product_sarsa_8x32spp": m.get("mape_prod_sarsa_8x32spp"), "reference_sarsa_8x32spp": m.get("mape_ref_sarsa_8x32spp"),
                            "product_default_1024spp": m.get("mape_prod_default_1024spp"), "reference_default_1024spp": m.get("mape_refB_1024spp"),
                            "ground_truth": "the reference's default path tracer (its own CUDA kernels on a B200) at 1024 spp; MAPE = Graphing/mape.py:10-21 on 8-bit RGB",
                            "source": "profiles/%s (written by tests/test_gpu_mape.py)" % name}
            line["reference_kernels_on_b200"] = {"sarsa_mpaths_s": m.get("ref_kernels_sarsa_mpaths_s"), "default_mpaths_s": m.get("ref_kernels_default_mpaths_s"),
                                                 "what": "GPU_Rendering_Engine's own kernels recompiled for sm_100a (oracle/_ref/libref_cuda.so), Cornell 512x512x32 spp, CUDA events around its frame loop",
                                                 "source": "profiles/%s" % name}
            break
        if excl:
            line["exclusive_kernel_ms_per_step"] = excl
            if excl["k_isect"] > 0:
                fe = hit_flops / args.steps / (excl["k_isect"] * 1e-3) / 1e12
                line["roofline"]["frac_exclusive"] = fe / fp32_peak if fp32_peak else None
                line["roofline"]["achieved_exclusive"] = fe
            if "roofline_shade" in line and excl["k_shade"] > 0:
                line["roofline_shade"]["frac_exclusive"] = shade_bytes / args.steps / (excl["k_shade"] * 1e-3) / 1e9 / hbm_peak
        if world == 1 and not args.no_cpu_baseline:
            eng = cpu_engine_run()
            v, kind, cores, sample, _ = cpu_reference_run(64, 1, want_seconds=args.cpu_seconds, spp=32)
            sarsa = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": kind, "sample": sample}
            if eng:
                a = eng["as_committed_all_cores"]
                line["cpu_baseline"] = {"value": a["value"], "unit": "Mpaths/s", "cores": a["cores"], "kind": "reference",
                                        "sample": "Old_CPU_Rendering_Engine (BASELINE.json configs[0]; unmodified sources + the two functions it declares but never defines, "
                                                  "oracle/cpu_engine/, g++ -O3 -fopenmp), " + a["sample"] + ", std::chrono around draw_default_path_tracing",
                                        "six_threads_as_hard_coded": eng["as_committed_6_threads"], "bounces_80": eng["bounces_80_all_cores"]}
                line["cpu_baseline_sarsa"] = sarsa                 # this bench's own workload on the host: the reference's GPU engine sources behind a host shim
            else:
                line["cpu_baseline"] = sarsa
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
