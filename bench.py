#!/usr/bin/env python
"""bench.py - encoder frames/sec decoded by the transducer-search hot path (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic input of cfg2's shape
(zipformer-large-en offline, modified_beam_search beam=4, 256 utterances x 250 frames, vocab 500, E=768):
encoder_proj -> per frame {stateless decoder, fused joiner + log-softmax/top-k, hypothesis merge} -> best
hypothesis per stream. One independent batch per GPU (weak scaling), no data-path collective; NCCL only
gathers the per-stream results for reporting (inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    torchrun ... bench.py --gpus N ...      (one rank per GPU)

`value`  : frames/s with the batch already resident in HBM (device-pointer entry point).
`e2e`    : the same metric through the host-pointer C-ABI call a P/Invoke shim makes, from pinned host
           memory, H2D and D2H copies inside the timed region.
`roofline`: the joiner GEMM launch (dominant kernel): 2*N*J*V flop per launch over its CUDA-event duration,
           against the measured bf16 tensor peak.
`cpu_baseline` / --impl reference: the CPU oracle port (numpy/OpenBLAS, all host threads) on a bounded sample
           of the same workload. The reference C#+ONNX Runtime binary cannot run in this image (no .NET).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from k2transducerasr_b200 import synth  # noqa: E402

CFG4_TRAFFIC = 342.3e6   # dram__bytes_read.sum + dram__bytes_write.sum of one whole-loop launch on cfg4 (337.1 MB + 5.2 MB)
METRIC = "encoder frames/sec decoded (modified_beam_search beam=4, batch 256/GPU)"
UNIT = "frames/s"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, uuid: str | None = None):
        self.index = index
        self.uuid = uuid            # CUDA ordinals and NVML indices differ under CUDA_VISIBLE_DEVICES: prefer the UUID
        self.rows = []
        self._stop = threading.Event()
        self._th = None

    def _run_nvml(self):
        """NVML in-process (no process spawn per sample: ~2 ms per sample instead of ~100 ms). The clock query is the one
        nvidia-smi's clocks.sm reports; the event-reason bits are nvidia-smi's clocks_event_reasons.*"""
        import pynvml as N
        N.nvmlInit()
        hdl = None
        if self.uuid:
            try:
                hdl = N.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
            except Exception:
                hdl = None
        if hdl is None:
            hdl = N.nvmlDeviceGetHandleByIndex(self.index)
        bits = [(getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8), 3), (getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), 4),
                (getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), 5), (getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4), 6)]
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = N.nvmlDeviceGetMaxClockInfo(hdl, N.NVML_CLOCK_SM)
        while not self._stop.is_set():
            row = [str(N.nvmlDeviceGetClockInfo(hdl, N.NVML_CLOCK_SM)), str(mx), str(N.nvmlDeviceGetPowerUsage(hdl) / 1000.0),
                   "", "", "", ""]
            r = get_reasons(hdl)
            for bit, col in bits:
                row[col] = "Active" if (r & bit) else "Not Active"
            self.rows.append(row)
            self._stop.wait(0.004)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample (the ONLY place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------------------
def cpu_oracle_step(model, raw_sample, beam):
    from oracle import k2_oracle as O
    enc = O.encoder_proj(model, raw_sample)
    return O.modified_beam_search(model, enc, beam)


def cpu_sample_setup(cfg, streams):
    from oracle import k2_oracle as O
    w = synth.make_weights(cfg.dims, blank_bias=cfg.blank_bias)
    model = O.Model.from_dict(w)
    raw = synth.make_frames(streams, cfg.frames, cfg.dims.encoder_dim, cfg.seed)
    return model, raw


def run_cpu_baseline(cfg, streams=32, repeats=1):
    model, raw = cpu_sample_setup(cfg, streams)
    cpu_oracle_step(model, raw[:2, :20], cfg.beam)          # warm BLAS
    t0 = time.perf_counter()
    for _ in range(repeats):
        cpu_oracle_step(model, raw, cfg.beam)
    dt = (time.perf_counter() - t0) / repeats
    return {"value": streams * cfg.frames / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"{streams} of {cfg.streams} streams x {cfg.frames} frames, beam {cfg.beam}, oracle/k2_oracle.py "
                      f"(numpy/OpenBLAS fp32, {os.cpu_count()} threads), {dt:.2f} s"}


def run_reference_arm(args, cfg):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the C#+ORT binary is
    not runnable here). Rank 0 only; other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    budget = 150.0
    model, raw = cpu_sample_setup(cfg, 4)
    t0 = time.perf_counter()
    cpu_oracle_step(model, raw, cfg.beam)
    per_stream = (time.perf_counter() - t0) / 4
    streams = int(max(2, min(cfg.streams, budget / max(1, args.steps + args.warmup) / max(per_stream * 0.5, 1e-3))))
    model, raw = cpu_sample_setup(cfg, streams)
    for _ in range(args.warmup):
        cpu_oracle_step(model, raw, cfg.beam)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_step(model, raw, cfg.beam)
    dt = time.perf_counter() - t0
    val = streams * cfg.frames * args.steps / dt
    sample = (f"{streams} of {cfg.streams} streams x {cfg.frames} frames per step, beam {cfg.beam}; CPU oracle port "
              f"(numpy/OpenBLAS fp32) - the C#+ONNX Runtime reference cannot run in this image (no .NET)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg.name, "streams_per_step": streams, "frames": cfg.frames, "beam": cfg.beam,
                       "vocab": cfg.dims.vocab_size},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from k2transducerasr_b200 import _native, build
    from k2transducerasr_b200 import dist as kd

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    build.build()

    d = cfg.dims
    B, T, K, V, J, E = cfg.streams, cfg.frames, cfg.beam, d.vocab_size, d.joiner_dim, d.encoder_dim
    prec = _native.PREC_NAMES[args.precision]
    h = _native.Handle(vocab_size=V, joiner_dim=J, decoder_dim=d.decoder_dim, encoder_dim=E, device=local_rank,
                       max_streams=B, max_frames=T, max_beam=K)
    h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
    if prec != _native.PREC_FP32:
        h.set_precision(prec)
    # a real (non-legacy) stream shared by torch's events and the library's launches: the legacy default stream has
    # handle 0, which k2b_set_stream reads as "use the handle's own stream" and torch events would then see nothing
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    h.set_stream(stream.cuda_stream)

    # two distinct batches per rank, alternated, each 196 MB > the 126 MB L2
    nbuf = 2
    host_in = [torch.from_numpy(synth.make_frames(B, T, E, cfg.seed + 17 * rank + i)).pin_memory() for i in range(nbuf)]
    dev_in = [x.to(dev) for x in host_in]
    cap = T
    d_tok = torch.zeros((B, cap), dtype=torch.int64, device=dev)
    d_ts = torch.zeros((B, cap), dtype=torch.int32, device=dev)
    d_n = torch.zeros((B,), dtype=torch.int32, device=dev)
    d_sc = torch.zeros((B,), dtype=torch.float32, device=dev)
    p_tok = torch.zeros((B, cap), dtype=torch.int64).pin_memory()
    p_ts = torch.zeros((B, cap), dtype=torch.int32).pin_memory()
    p_n = torch.zeros((B,), dtype=torch.int32).pin_memory()
    p_sc = torch.zeros((B,), dtype=torch.float32).pin_memory()

    def step_dev(i):
        h.call("k2b_modified_beam_search_dev", dev_in[i % nbuf], 1, B, T, K, d_tok, d_ts, d_n, d_sc, cap)

    def step_host(i):
        h.call("k2b_modified_beam_search", host_in[i % nbuf], 1, B, T, K, p_tok, p_ts, p_n, p_sc, cap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    try:
        dev_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        dev_uuid = None
    clk = ClockSampler(local_rank, dev_uuid)
    clk.__enter__()                      # sampled from warm-up to the end of the e2e region: all of it is under load
    for i in range(max(args.warmup, 3)):
        step_dev(i)
    torch.cuda.synchronize()

    h.reset_launch_count()
    ms = timed(step_dev, args.steps)
    launches = h.launch_count()
    value = world * B * T * args.steps / (ms * 1e-3)

    # gather of results for reporting (NCCL over NVLink) - once, outside `value`'s kernel loop but shown to work
    gathered = None
    if world > 1:
        n = d_n.cpu().numpy()
        toks = [d_tok[b, :n[b]].cpu().tolist() for b in range(min(B, 8))]
        tss = [d_ts[b, :n[b]].cpu().tolist() for b in range(min(B, 8))]
        allt, _, _ = kd.gather_results(toks, tss, d_sc[:len(toks)].cpu().tolist(), len(toks) * world, cap, device=dev)
        gathered = len(allt)

    # e2e through the host-pointer entry point (pinned host buffers; H2D + D2H inside the timed region)
    for i in range(2):
        step_host(i)
    e2e_steps = max(2, min(args.steps, 10))
    ms_e2e = timed(step_host, e2e_steps)
    e2e_value = world * B * T * e2e_steps / (ms_e2e * 1e-3)
    h2d = B * T * E * 4
    d2h = B * cap * 8 + B * cap * 4 + B * 4 + B * 4
    clk.__exit__(None, None, None)

    # dominant kernel: the joiner GEMM of every frame, bracketed with CUDA events on the launch stream
    h.profile_enable(True)
    prof_steps = max(1, min(args.steps, 3))
    for i in range(prof_steps):
        step_dev(i)
    torch.cuda.synchronize()
    n_l, tot_ms = h.profile_read()
    h.profile_enable(False)
    peaks = load_peaks()
    fused_loop = prec != _native.PREC_FP32          # persistent cluster kernel: one launch runs all T frames
    flops_per_launch = 2.0 * (B * K) * J * V * (T if fused_loop else 1)
    avg_ms = tot_ms / max(n_l, 1)
    achieved_tf = flops_per_launch / (avg_ms * 1e-3) / 1e12 if n_l else 0.0
    # dram__bytes_read.sum + dram__bytes_write.sum of one cluster_beam_kernel launch on this workload, from the
    # ncu --set full capture profiles/r01_cluster_beam_v8_full.ncu-rep (240.5 MB + 6.8 MB); none taken for the fp32 path
    # cfg4 (persistent joiner_topk_kernel, whole time loop): profiles/r01_beam_mega_cfg4_T250_full.ncu-rep
    traffic = {"cfg2": 247.2e6, "cfg4": CFG4_TRAFFIC}.get(args.workload) if fused_loop else None
    kernel_name = ("cluster_beam_kernel: whole time loop (joiner tcgen05 GEMM + log-softmax/top-k + merge)" if args.workload == "cfg2"
                   else "joiner_topk_kernel<MEGA>: whole time loop (persistent tcgen05 joiner with top-k epilogue + merge warps)")
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["tf_sustained"], "traffic": traffic, "kernel": (kernel_name if fused_loop else "joiner GEMM (+log-softmax/top-k epilogue), one launch per frame"),
                "avg_launch_us": avg_ms * 1e3, "launches_timed": n_l, "flop_per_launch": flops_per_launch,
                "peak_source": peaks["src"] + ", sustained bf16",
                "frame_step_roofline_us": 2.0 * (B * K) * J * V / (peaks["tf_sustained"] * 1e12) * 1e6,
                "frame_step_us": ms * 1e3 / args.steps / T,
                "whole_step_frac": (2.0 * (B * K) * J * V * T / (peaks["tf_sustained"] * 1e12)) / (ms * 1e-3 / args.steps)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_baseline(cfg, streams=cfg.streams, repeats=2)     # two full batches: ~10 s of CPU work

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "bf16x3": "bf16x3", "bf16": "bf16"}[args.precision],
                "data": "synthetic",
                "config": {"workload": cfg.name, "streams_per_gpu": B, "frames": T, "beam": K, "vocab": V, "joiner_dim": J,
                           "encoder_dim": E, "precision": args.precision, "parallelism": f"dp{world} (independent batches)",
                           "l2": f"inputs ({B * T * E * 4 / 1e6:.0f} MB/step, 2 alternating batches) larger than the 126 MB L2",
                           "blank_bias": cfg.blank_bias, "regime": args.regime, "weights": "random-init, seed 7"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
                "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if gathered is not None:
            line["gathered_streams"] = gathered
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"])
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--regime", default="speech", choices=["speech", "raw"],
                    help="speech: blank bias calibrated so that 70-80 %% of the frames are blank (default); raw: random-init joiner as is, "
                         "almost every frame emits (worst case for the decoder gather)")
    args = ap.parse_args()
    cfg = synth.CONFIGS[args.workload]
    if args.regime == "raw":
        import dataclasses
        cfg = dataclasses.replace(cfg, blank_bias=0.0)
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
