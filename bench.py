#!/usr/bin/env python3
"""bench.py -- the BASELINE.json metric: Mtexels/s of the BCn block-compression hot path.

A "step" is one pass of the hot path over one synthetic surface.  Default workload (N=1) is
BASELINE.json configs[1]: BC7 `GetProfile_slow`, 4096x4096 RGBA8 uniform-random texels.

  value     whole-job Mtexels/s with inputs resident in HBM (device-pointer entry itw_encode_device),
            timed on the device with CUDA events around every step, max over ranks.
  e2e       the same metric through the reference-facing C-ABI call CompressBlocksBC*(host surface,
            host dst): pinned host buffers, H2D + kernel + D2H inside the timed region.
  roofline  algorithmic bytes of the kernel / its CUDA-event duration, against the measured HBM peak
            in MEASURED_PEAKS.json.
  cpu_baseline / --impl reference
            the reference's own encoder (oracle/_ref = kernel.ispc compiled scalar, else the oracle
            port) on the host cores, on a bounded sample of the same workload.

Multi-GPU (torchrun, one rank per GPU): blocks are independent, so each rank encodes its own
surface of the same shape (weak scaling, no data-path collective); NCCL is used only for the
barrier and the max-over-ranks reduction of the step times.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("intel-texture-works-plugin_b200")
binding = importlib.import_module("intel-texture-works-plugin_b200.binding")

READ_BYTES_PER_TEXEL = {"BC1": 4, "BC3": 4, "BC4": 4, "BC5": 4, "BC7": 4, "BC6H": 8}
WRITE_BYTES_PER_TEXEL = {"BC1": 0.5, "BC4": 0.5, "BC3": 1, "BC5": 1, "BC7": 1, "BC6H": 1}
DEFAULT_PROFILE = {"BC7": "slow", "BC6H": "bc6h_slow"}


def make_surface(fmt, size, seed):
    if fmt == "BC6H":
        return pkg.synth.random_rgba16f(size, size, seed=0xB2000003 + seed)
    return pkg.synth.random_rgba8(size, size, seed=0xB2000002 + seed)


def workload_name(fmt, prof, size):
    kind = "RGBA16F random" if fmt == "BC6H" else "RGBA8 random"
    return f"{fmt}{' ' + prof if prof else ''}, {size}x{size} {kind}"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores (test infrastructure under oracle/)
# ------------------------------------------------------------------------------------------------
def load_cpu_reference():
    """oracle/_ref (the reference's kernel.ispc + ispc_texcomp.cpp compiled scalar) if it exists,
    else the handwritten oracle port."""
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libitw_ref.so")
    if os.path.exists(ref_so):
        return binding.EncoderApi(ref_so, ""), "reference"
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return binding.EncoderApi(os.path.join(ROOT, "oracle", "libitw_oracle.so"), "oracle_"), "port"


def cpu_encode_mt(api, fmt, img, settings, threads):
    """Row-band split exactly like CompressImageMT (win32Threads.cpp:217-230): linesPerThread =
    ceil(h/T), bands rounded down to multiples of 4 rows, one band per thread."""
    h, w = img.shape[:2]
    bpb = binding.FORMATS[fmt][1]
    out = np.zeros((h // 4) * (w // 4) * bpb, np.uint8)
    lines = (h + threads - 1) // threads
    jobs = []
    for t in range(threads):
        y0 = (lines * t) // 4 * 4
        y1 = min((lines * (t + 1)) // 4 * 4, h)
        if y1 > y0:
            jobs.append((y0, y1))

    def work(job):
        y0, y1 = job
        api.encode_raw(fmt, img.ctypes.data + y0 * img.strides[0], w, y1 - y0, img.strides[0],
                       out.ctypes.data + (y0 // 4) * (w // 4) * bpb, settings)

    ts = [threading.Thread(target=work, args=(j,)) for j in jobs]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return out


def cpu_sample_rows(api, fmt, img, settings, threads, target_s):
    """Rows of the workload that take about `target_s` seconds on `threads` host threads (calibrated
    on a small band first)."""
    h, w = img.shape[:2]
    probe = min(h, 4 * threads)
    t0 = time.perf_counter()
    cpu_encode_mt(api, fmt, img[:probe], settings, threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    rows = int(probe * target_s / dt)
    rows = max(4 * threads, min(h, rows // (4 * threads) * (4 * threads)))
    return rows


def run_cpu_arm(args, fmt, prof, size, full_json):
    api, kind = load_cpu_reference()
    settings = api.profile(prof) if prof else None
    threads = os.cpu_count() or 1
    img = make_surface(fmt, size, 0)
    per_step_s = 20.0 / max(args.steps + args.warmup, 1) if full_json else 12.0
    rows = cpu_sample_rows(api, fmt, img, settings, threads, max(per_step_s, 0.5))
    band = img[:rows]
    for _ in range(args.warmup if full_json else 0):
        cpu_encode_mt(api, fmt, band, settings, threads)
    steps = args.steps if full_json else 1
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_encode_mt(api, fmt, band, settings, threads)
    dt = time.perf_counter() - t0
    mtexels = rows * size * steps / dt / 1e6
    info = {"value": round(mtexels, 4), "unit": "Mtexels/s", "cores": threads, "kind": kind,
            "sample": f"first {rows} of {size} rows of the workload per step, row-band split over {threads} threads "
                      f"(scalar build of the reference source, not the ISPC SIMD binary)"}
    return info, dt / steps * 1e3


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--format", default="BC7", choices=sorted(binding.FORMATS))
    ap.add_argument("--profile", default=None)
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    fmt = args.format
    prof = args.profile or DEFAULT_PROFILE.get(fmt)
    size = args.size
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": workload_name(fmt, prof, size), "format": fmt, "profile": prof, "width": size, "height": size,
              "sharding": "one surface per GPU, no collective" if world > 1 else "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        info, ms = run_cpu_arm(args, fmt, prof, size, full_json=True)
        line = {"impl": "reference", "metric": "Mtexels/s", "value": info["value"], "unit": "Mtexels/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": info,
                "e2e": {"value": info["value"], "unit": "Mtexels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = pkg.ItwBcn()
    lib.set_device(local_rank)
    settings = lib.profile(prof) if prof else None
    _, bpb, texel_bytes, _ = binding.FORMATS[fmt]
    out_bytes = (size // 4) * (size // 4) * bpb
    in_bytes = size * size * texel_bytes

    # Inputs resident in HBM.  A rotation of distinct surfaces larger than L2 in total, so a step
    # never finds its input in the 126 MB L2 left there by the previous step.
    nrot = max(2, -(-(256 << 20) // in_bytes))
    config["l2"] = f"rotation of {nrot} distinct input surfaces ({nrot * in_bytes >> 20} MiB > 126 MB L2)"
    hosts = [make_surface(fmt, size, 16 * rank + i) for i in range(nrot)]
    d_in = [torch.from_numpy(h.view(np.uint8).reshape(-1)).cuda() for h in hosts]
    d_out = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def step(i):
        lib.encode_device(fmt, d_in[i % nrot].data_ptr(), size, size, size * texel_bytes, d_out.data_ptr(), settings,
                          stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for i in range(args.steps):
        evs[i][0].record(stream)
        step(args.warmup + i)
        evs[i][1].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    launches = lib.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    texels_per_step = size * size * world
    value = texels_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the reference-facing C-ABI with pinned host buffers ----
    h_in = torch.from_numpy(hosts[0].view(np.uint8).reshape(-1)).pin_memory()
    h_out = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
    for _ in range(min(args.warmup, 2)):
        lib.encode_raw(fmt, h_in.data_ptr(), size, size, size * texel_bytes, h_out.data_ptr(), settings)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lib.encode_raw(fmt, h_in.data_ptr(), size, size, size * texel_bytes, h_out.data_ptr(), settings)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = texels_per_step * e2e_steps / float(te.item()) / 1e6
    # parity spot check of what was just produced (first 4096 blocks) is left to tests/; here only
    # make sure the output is not empty
    assert int(h_out[:4096].to(torch.int64).sum()) != 0, "encoder produced an empty output"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    algo_bytes = size * size * (READ_BYTES_PER_TEXEL[fmt] + WRITE_BYTES_PER_TEXEL[fmt])
    achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 3), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 6),
                "traffic": None, "peak_source": peak_src, "kernel": {"BC7": "bc7_kernel", "BC6H": "bc6h_kernel"}.get(fmt, "bc1_bc3_kernel / bc4_bc5_kernel"),
                "algorithmic_bytes_per_launch": int(algo_bytes),
                "note": "read-only variant: %.3f GB/s" % (size * size * READ_BYTES_PER_TEXEL[fmt] / (ms_per_step * 1e-3) / 1e9)}
    traffic_path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_path):
        counters = json.load(open(traffic_path))
        roofline["traffic"] = counters.get(f"{fmt}:{prof}:{size}")
        # The encoders are issue-bound, not HBM-bound: alongside the required HBM roofline,
        # report the warp-instruction issue rate against 4 schedulers x SM count x the SM clock sampled DURING the
        # timed region.  Instructions per launch come from the committed ncu capture of the same workload.
        winst = counters.get("warp_inst", {}).get(f"{fmt}:{prof}:{size}")
        if winst and clocks.get("sm_mhz"):
            sms = torch.cuda.get_device_properties(0).multi_processor_count
            peak_issue = 4.0 * sms * clocks["sm_mhz"] * 1e6
            got = winst / (ms_per_step * 1e-3)
            roofline["issue"] = {"warp_inst_per_launch": int(winst), "achieved_ginst_s": round(got / 1e9, 1),
                                 "peak_ginst_s": round(peak_issue / 1e9, 1), "frac": round(got / peak_issue, 3),
                                 "note": "4 warp schedulers/SM x SMs x median SM clock under load; instruction count from profiles/"}

    cpu_info = None
    if world == 1 and not args.no_cpu:
        cpu_info, _ = run_cpu_arm(args, fmt, prof, size, full_json=False)

    line = {"metric": "Mtexels/s", "value": round(value, 3), "unit": "Mtexels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": round(e2e_value, 3), "unit": "Mtexels/s", "h2d_bytes_per_step": in_bytes * world,
                    "d2h_bytes_per_step": out_bytes * world, "steps": e2e_steps},
            "gpu_launches": int(launches) * world, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_info,
            "wall_ms_per_step": round(wall / args.steps * 1e3, 4)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
