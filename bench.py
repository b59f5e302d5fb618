#!/usr/bin/env python
"""bench.py -- DRBA hot path on B200: RIFE-4.26-heavy 1080p 24->60, scale 1.0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one sliding-triplet window of the reference's driver loop (infer.py:112-156):
``RIFE.inference_ts_drba(I0, I1, I2, ts, reuse, linear=True)`` with the 24->60 timestamp schedule
(windows alternate ts = [0.6, 1.0, 1.4] / [0.8, 1.2]: 2.5 output frames per window, 2 of them
interpolated).  The metric is OUTPUT frames per second at net-input size 1088x1920.

* value : whole-job frames/s with the frames already resident in HBM (CUDA events, K steps).
* e2e   : the same through the public API with HOST buffers: every step uploads the window's new
          decoded uint8 frame from pinned host memory (to_inp) and downloads every output frame as
          uint8 (to_out, models/utils/tools.py:59-68), on copy streams next to the compute stream.
* roofline : the dominant kernel family of the step, timed live with CUDA events in an
          instrumented pass of the same steps (drba_b200._lib.LaunchProfiler).
* cpu_baseline : the oracle port of the reference path (oracle/ifnet.py, torch fp32 CPU convs +
          C splat/warp) on the host cores, one window (rank 0, N = 1 only).
* --impl reference : that CPU path as the reference arm (bounded number of windows).

N > 1: frame-window sharding, one replica per GPU, no collective on the data path;
every rank runs K windows of its own shard -> "scaling": "weak".
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
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_SRC, W_SRC = 1080, 1920
METRIC = "output frames/s, RIFE-4.26-heavy 1080p 24->60 (DRBA inference_ts_drba windows)"
UNIT = "frames/s"
TS_PATTERN = [np.array([0.6, 1.0, 1.4]), np.array([0.8, 1.2])]   # calc_t() at 24 -> 60 fps (infer.py:76-91)


def net_size(h, w, scale=1.0, div=64):
    """get_valid_net_inp_size (models/utils/tools.py:41-56)."""
    def up(v):
        return int((v * scale // div + 1) * div / scale) if v * scale % div != 0 else v
    return up(h), up(w)


def synth_clip(n_frames, h, w, seed, device):
    """Seeded low-frequency texture translated/warped smoothly from frame to frame."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    base = torch.rand((1, 3, h // 8 + 16, w // 8 + 16), generator=g)
    base = torch.nn.functional.interpolate(base, scale_factor=8, mode="bicubic", align_corners=False).clamp(0, 1)
    fine = torch.rand((1, 3, h + 128, w + 128), generator=g) * 0.15
    tex = (base[:, :, : h + 128, : w + 128] * 0.85 + fine).to(device)
    frames = []
    for i in range(n_frames):
        dx, dy = 5.5 * i, 2.25 * i
        x0, y0 = int(dx), int(dy)
        ax, ay = dx - x0, dy - y0
        p = tex[:, :, y0:y0 + h + 1, x0:x0 + w + 1]
        f = (p[:, :, :h, :w] * (1 - ax) * (1 - ay) + p[:, :, :h, 1:w + 1] * ax * (1 - ay)
             + p[:, :, 1:h + 1, :w] * (1 - ax) * ay + p[:, :, 1:h + 1, 1:w + 1] * ax * ay)
        frames.append(f.contiguous())
    return frames


def load_state():
    from drba_b200.weights import find_rife_weights, load_ifnet_state, synth_ifnet_state
    wdir = find_rife_weights()
    if wdir is not None:
        return load_ifnet_state(wdir), "reference checkpoint flownet.pkl"
    return synth_ifnet_state(0), "seeded random-init weights (reference checkpoint not on this machine)"


class ClockSampler:
    """SM clocks / throttle reasons while the timed region runs (B200_PROFILING.md).  Sampled in-process through
    NVML (a thread polling every 50 ms); an `nvidia-smi -lms` child process is only the fallback -- its start-up and
    its polling were observed to stall the GPU for ~100 ms once in a while, which doubled a 165 ms timed region."""

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc, self.skip = gpu_index, [], None, 0
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def mark(self):
        """Samples taken before this call (warm-up) are dropped from the report."""
        self.skip = len(self.rows)

    def start(self):
        mode = os.environ.get("DRBA_BENCH_CLOCKS", "smi")
        if mode == "none":
            return
        try:
            if mode != "nvml":
                raise RuntimeError("nvidia-smi sampler selected")
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.idx]) if visible and visible.split(",")[self.idx].isdigit() else self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            R = pynvml
            bits = [(getattr(R, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                    (getattr(R, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                    (getattr(R, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                    (getattr(R, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]

            def poll():
                while not self._stop.is_set():
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        try:
                            rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.rows.append([str(self.idx), str(sm), str(mx)] + ["Active" if rs & b else "Not Active" for b, _ in bits])
                    except Exception:
                        pass
                    self._stop.wait(0.05)

            self._nvml = pynvml
            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self._nvml is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.rows = self.rows[self.skip:] or self.rows
        sm = [int(r[1]) for r in self.rows if len(r) >= 7 and r[1].isdigit()]
        mx = [int(r[2]) for r in self.rows if len(r) >= 7 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return {"hbm": float(d.get("hbm_gbs", 6650.0)), "tensor": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                    "src": "MEASURED_PEAKS.json (sustained bf16; kernel timed inside a long step)"}
        except Exception:
            pass
    return {"hbm": 6650.0, "tensor": 1590.0, "src": "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"}


def softsplat_roofline(dev, peaks):
    """softsplat(C=64, soft) at 1152x1920 through drba_softsplat_f32: the flow is a smooth field with a
    constant offset (what DRBA feeds it: flow x timestep); L2 is flushed between runs."""
    from drba_b200.softsplat import softsplat
    c, h, w = 64, 1152, 1920
    g = torch.Generator(device="cpu").manual_seed(1)
    lo = 2.0 * torch.randn((1, 2, h // 16, w // 16), generator=g)
    flow = (torch.nn.functional.interpolate(lo, size=(h, w), mode="bilinear", align_corners=False) + 6.5).to(dev)
    x = torch.randn((1, c, h, w), device=dev)
    metric = torch.randn((1, 1, h, w), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        softsplat(x, flow, metric, "soft")
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        softsplat(x, flow, metric, "soft")
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = h * w * 4 * ((c + 3) + c)
    ach = nbytes / ms / 1e6
    return {"kernel": "softsplat (count/scan/fill/gather, csrc/splat_gather.cu)", "bound": "hbm", "achieved": round(ach, 1),
            "peak": peaks["hbm"], "unit": "GB/s", "frac": round(ach / peaks["hbm"], 4), "ms": round(ms, 4),
            "workload": "C=64 soft 1152x1920 fp32 NCHW, smooth flow; 6 launches per call", "traffic": None}


# ------------------------------------------------------------------------------------------
def cpu_port_windows(h, w, n_windows, seed):
    """The reference path restated on the CPU (oracle port): returns (seconds, output frames, threads)."""
    from oracle.ifnet import RIFEOracle
    state, _ = load_state()
    torch.set_grad_enabled(False)
    frames = synth_clip(n_windows + 2, h, w, seed, "cpu")
    m = RIFEOracle(state)
    reuse = None
    # warm-up state like the sequential loop (first window computes both pairs); timed as part of the job
    t0 = time.perf_counter()
    nout = 0
    for j in range(n_windows):
        out, reuse = m.inference_ts_drba(frames[j], frames[j + 1], frames[j + 2], TS_PATTERN[j % 2], reuse, True)
        nout += len(out)
    return time.perf_counter() - t0, nout, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    h, w = net_size(H_SRC, W_SRC)
    budget = 150.0
    t1, n1, threads = cpu_port_windows(h, w, 1, 0)      # also the warm-up
    k = max(1, min(args.steps, int(budget // max(t1, 1e-3))))
    secs, nout, threads = cpu_port_windows(h, w, k, 0)
    fps = nout / secs
    line = {"impl": "reference", "metric": METRIC, "value": round(fps, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": k, "warmup": 1, "ms_per_step": round(1e3 * secs / k, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "RIFE-4.26-heavy 1080p 24->60, scale=1.0 (BASELINE.json configs[1]) on host CPU",
                       "net_input": [h, w], "ts_pattern": "[0.6,1.0,1.4]/[0.8,1.2]"},
            "cpu_baseline": {"value": round(fps, 4), "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{k} window(s) ({nout} output frames) of the same 1088x1920 clip; oracle port "
                                       f"(torch fp32 CPU convs + C splat/warp); /root/reference cannot travel to the GPU box"},
            "e2e": {"value": round(fps, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--warmup-seconds", type=float, default=1.0,
                    help="keep running warm-up windows until this much wall time has passed (clock ramp from idle)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--profile-json", default=None, help="write the per-kernel-family breakdown here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from drba_b200 import _lib
    from drba_b200.rife import RIFE
    state, wdesc = load_state()
    model = RIFE(state=state, device=dev, precision=args.precision, graphs=not args.no_graphs)
    h, w = net_size(H_SRC, W_SRC, model.scale, model.pad_size)
    K, Wm = args.steps, args.warmup
    ring = 8
    frames = synth_clip(ring, h, w, 1000 + rank, dev)     # each rank: its own shard of the stream

    def window(j, reuse, src):
        return model.inference_ts_drba(src[j % ring], src[(j + 1) % ring], src[(j + 2) % ring], TS_PATTERN[j % 2], reuse, True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------
    # warm-up: W windows AND at least `--warmup-seconds` of work, so graph capture, allocator growth and
    # the GPU's clock ramp from idle are all outside the timed region (Wm stays even: the ts pattern
    # alternates, so the timed region starts on the same phase for every run)
    # the clock sampler (an nvidia-smi child polling every 100 ms) is started BEFORE the warm-up: its start-up
    # takes a driver lock that stalled the GPU for ~100 ms when it was launched right at the timed region
    clocks = ClockSampler(local_rank)
    clocks.start()
    reuse = None
    t_w = time.perf_counter()
    j = 0
    while j < Wm or (time.perf_counter() - t_w < args.warmup_seconds and j < 2000):
        _, reuse = window(j, reuse, frames)
        j += 1
        if j >= Wm and j % 2 == 0:
            torch.cuda.synchronize()
    if j % 2:
        _, reuse = window(j, reuse, frames)
        j += 1
    # last warm-up phase: K windows enqueued back to back exactly like the timed loop (no intermediate syncs).  The
    # first deep asynchronous burst of a process makes the driver grow its command queues, a one-off 50-150 ms stall
    # that otherwise lands inside the timed region (observed in half of the runs)
    out = None
    for _k in range(K + (K % 2)):
        out, reuse = window(j, reuse, frames)      # same variable as the timed loop: the same two generations of
        j += 1                                     # output tensors stay alive, the allocator sees nothing new
    Wm_done = j
    # the host is only a window or two ahead of the GPU (a graph exec cannot have two launches in flight), so a
    # Python garbage-collection pause inside the timed loop shows up as a GPU stall: collect now, not then
    import gc
    gc.collect()
    gc.disable()
    barrier()
    clocks.mark()
    launches0 = _lib.KERNEL_LAUNCHES
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    for e in ev:
        e.record()            # creates the CUDA events now (torch creates them lazily at the first record)
    barrier()
    nout = 0
    ev[0].record()
    for k, j in enumerate(range(Wm_done, Wm_done + K)):
        out, reuse = window(j, reuse, frames)
        nout += len(out)
        ev[k + 1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[K])
    step_raw = [ev[k].elapsed_time(ev[k + 1]) for k in range(K)]
    step_ms = sorted(step_raw)
    slow_steps = [(k, round(v, 2)) for k, v in enumerate(step_raw) if v > 3.0 * step_ms[K // 2]][:8]
    launches = _lib.KERNEL_LAUNCHES - launches0

    # ---- end to end through the public API with HOST buffers ------------------------------------------
    # host side = what the reference's driver loop holds (infer.py:112-156): decoded uint8 BGR frames
    # [1080,1920,3] in, encoded-ready uint8 frames out.  Every step uploads the window's new frame
    # (to_inp: H2D + /255 + resize to the 1088x1920 net input, one fused kernel) and downloads EVERY output
    # frame (to_out: resize back + *255 + uint8 + D2H); copies ride two copy streams (drba_b200.tools.FrameIO).
    from drba_b200.tools import FrameIO
    io = FrameIO((H_SRC, W_SRC), (h, w), dev)
    host_u8 = []
    for f in frames:
        f8 = torch.nn.functional.interpolate(f, size=(H_SRC, W_SRC), mode="bilinear", align_corners=False)
        host_u8.append((f8[0].permute(1, 2, 0) * 255.0).clamp(0, 255).to(torch.uint8).contiguous().cpu().pin_memory())
    # the new frame of a window is uploaded while the PREVIOUS window computes (one frame ahead, like the reference's
    # reader thread): every step still moves exactly one frame in and all of its output frames out
    win = [io.upload(host_u8[0]), io.upload(host_u8[1]), io.upload(host_u8[2])]
    reuse_e = None

    def e2e_window(j, reuse_e):
        I0, I1, I2 = win[-3], win[-2], win[-1]
        out, reuse_e = model.inference_ts_drba(I0, I1, I2, TS_PATTERN[j % 2], reuse_e, True)
        io.release_inputs((I0, I1, I2))
        for o in out:
            io.download(o)                                  # every output frame goes back to the host
        win.append(io.upload(host_u8[(j + 3) % ring]))     # next window's new frame: H2D + ingest overlap this window
        del win[0]
        return len(out), reuse_e

    for j in range(Wm):
        _, reuse_e = e2e_window(j, reuse_e)
    io.drain()
    barrier()
    io.h2d_bytes = io.d2h_bytes = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nout_e = 0
    e0.record()
    for j in range(Wm, Wm + K):
        n, reuse_e = e2e_window(j, reuse_e)
        nout_e += n
    io.drain()
    e1.record()
    barrier()
    ms_e = e0.elapsed_time(e1)
    h2d, d2h = io.h2d_bytes, io.d2h_bytes
    clk = clocks.stop()
    gc.enable()

    # ---- instrumented pass: per-kernel-family shares and the roofline ------------------------------
    model.graphs = False          # per-kernel events need eager launches
    _, reuse = window(Wm + K - 1, None, frames)
    with _lib.LaunchProfiler() as prof:
        for j in range(Wm + K, Wm + K + min(K, 6)):
            _, reuse = window(j, reuse, frames)
        detail = prof.summary()
    nprof = min(K, 6)
    fam = {}
    for k, v in detail.items():           # aggregate "family/tag" -> family
        d = fam.setdefault(k.split("/")[0], {"calls": 0, "kernels": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for key in d:
            d[key] += v[key]

    # ---- max over ranks -------------------------------------------------------------------------
    t = torch.tensor([ms, ms_e], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(nout), float(nout_e)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_e = float(t[0]), float(t[1])
    nout_all, nout_e_all = float(tot[0]), float(tot[1])

    if rank == 0:
        peaks = measured_peaks()
        total_ms = sum(d["ms"] for d in fam.values()) or 1.0
        top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        name, d = top
        if d["flops"] > 0:
            achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
            roof = {"kernel": name, "bound": "tensor", "achieved": round(achieved, 2), "peak": peaks["tensor"],
                    "unit": "TFLOP/s", "frac": round(achieved / peaks["tensor"], 4), "traffic": None}
        else:
            achieved = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["bytes"] else 0.0
            roof = {"kernel": name, "bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm"],
                    "unit": "GB/s", "frac": round(achieved / peaks["hbm"], 4), "traffic": None}
        try:      # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_conv_tc_traffic.json")))
            if name == "conv_tc_f16":
                roof["traffic"] = tr["dram_bytes_per_launch"]
                roof["traffic_source"] = "profiles/r1_conv_tc_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"
        except Exception:
            pass
        roof.update({"launches_per_step": round(d["kernels"] / nprof, 1), "avg_launch_us": round(1e3 * d["ms"] / max(d["kernels"], 1), 2),
                     "share_of_step": round(d["ms"] / total_ms, 3), "peak_source": peaks["src"]})
        breakdown = {k: {"ms_per_step": round(v["ms"] / nprof, 4), "kernels_per_step": round(v["kernels"] / nprof, 1),
                         "share": round(v["ms"] / total_ms, 3),
                         "TFLOPs": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["flops"] else None,
                         "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] else None}
                     for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        if args.profile_json:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_json)), exist_ok=True)
            per_tag = {k: {"ms_per_step": round(v["ms"] / nprof, 4), "kernels_per_step": round(v["kernels"] / nprof, 1),
                           "avg_us": round(1e3 * v["ms"] / max(v["kernels"], 1), 2),
                           "TFLOPs": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["flops"] else None}
                       for k, v in sorted(detail.items(), key=lambda kv: -kv[1]["ms"])}
            json.dump({"families": breakdown, "detail": per_tag}, open(args.profile_json, "w"), indent=1)
        print("per-kernel-family breakdown (instrumented pass): " + json.dumps(breakdown), file=sys.stderr)

        # second half of BASELINE.json's metric: softsplat achieved HBM GB/s vs peak, timed live through the
        # C ABI on the size SURVEY.md 8d names (C = 64, soft, 1152x1920; algorithmic bytes 4*HW*[(C+3)+C])
        splat = None
        try:
            splat = softsplat_roofline(dev, peaks)
        except Exception as e:       # the headline line must not die on the secondary measurement
            splat = {"error": str(e)[:200]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            secs, n_cpu, threads = cpu_port_windows(h, w, 1, 1000)
            cpu = {"value": round(n_cpu / secs, 4), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"1 window (ts=[0.6,1.0,1.4]: {n_cpu} output frames, cold start) of the same 1088x1920 clip, "
                             f"{secs:.1f} s; oracle port (torch fp32 CPU convs + C splat/warp)"}
        line = {"metric": METRIC, "value": round(nout_all / (ms * 1e-3), 3), "unit": UNIT, "n_gpus": world,
                "steps": K, "warmup": Wm, "ms_per_step": round(ms / K, 4), "higher_is_better": True,
                "ms_per_step_median": round(step_ms[K // 2], 4), "ms_per_step_max": round(step_ms[-1], 4), "slow_steps": slow_steps, "warmup_windows_run": Wm_done,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.precision == "fp16" else "f32",
                "data": "synthetic",
                "config": {"workload": "RIFE-4.26-heavy 1080p 24->60, scale=1.0 (BASELINE.json configs[1])",
                           "net_input": [h, w], "ts_pattern": "[0.6,1.0,1.4]/[0.8,1.2] (2.5 output frames / window)",
                           "weights": wdesc, "precision": args.precision + " convs (tcgen05), fp32 flow/DRM/warp/splat"
                           if args.precision == "fp16" else "fp32",
                           "parallelism": f"frame-window shards x{world}, no collective",
                           "launch": "eager" if args.no_graphs else "one CUDA graph replay per window shape",
                           "l2": "per-step working set (3 fp32 frames 75 MB + 134 MB state + features/activations) exceeds the 126 MB L2; ring of 8 distinct frames"},
                "e2e": {"value": round(nout_e_all / (ms_e * 1e-3), 3), "unit": UNIT,
                        "h2d_bytes_per_step": int(h2d / K), "d2h_bytes_per_step": int(d2h / K)},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof}
        if splat is not None:
            line["softsplat_roofline"] = splat
        if cpu is not None:
            line["cpu_baseline"] = cpu
        _emit(line)
    if dist is not None:
        dist.destroy_process_group()


def _emit(line):
    """The result line goes to the process's ORIGINAL stdout; see _quiet_stdout."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


_REAL_STDOUT = None


def _quiet_stdout():
    """stdout carries exactly one JSON line (the driver parses it).  Libraries write there too (NCCL prints its
    version banner to fd 1 when NCCL_DEBUG is set on the box), so fd 1 is pointed at stderr for the run and the
    result is written to a private duplicate of the original fd."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


if __name__ == "__main__":
    _quiet_stdout()
    main()
