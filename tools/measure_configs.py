"""Device-timed throughput of every BASELINE.json config (run on the GPU box). One JSON line per config.
cfg2 is bench.py's job; here: cfg1 greedy single, cfg3 online greedy chunks, cfg4 large-vocab beam, cfg5 CTC (HBM roofline)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from k2transducerasr_b200 import _native, build, synth  # noqa: E402

build.build()
PEAK_HBM = 6554.2
try:
    PEAK_HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def handle(cfg, prec="fp32"):
    d = cfg.dims
    h = _native.Handle(vocab_size=d.vocab_size, joiner_dim=d.joiner_dim, decoder_dim=d.decoder_dim, encoder_dim=d.encoder_dim,
                       precision=_native.PREC_NAMES[prec])
    h.load_weights(synth.make_weights(d, blank_bias=cfg.blank_bias))
    h.set_stream(stream.cuda_stream)
    return h


def outbufs(B, cap):
    return (torch.zeros((B, cap), dtype=torch.int64, device=dev), torch.zeros((B, cap), dtype=torch.int32, device=dev),
            torch.zeros((B,), dtype=torch.int32, device=dev), torch.zeros((B,), dtype=torch.float32, device=dev))


def emit(**kw):
    print(json.dumps(kw), flush=True)


which = sys.argv[1:] or ["cfg1", "cfg3", "cfg4", "cfg5"]

if "cfg1" in which:
    cfg = synth.CONFIGS["cfg1"]
    for prec, label in (("fp32", "per-frame fp32 path"), ("bf16x3", "persistent cluster kernel, split-bf16 x3")):
        h = handle(cfg, prec)
        x = torch.from_numpy(synth.make_frames(1, cfg.frames, cfg.dims.encoder_dim, cfg.seed)).to(dev)
        tok, ts, n, sc = outbufs(1, cfg.frames)
        ms = timed(lambda: h.call("k2b_greedy_offline_dev", x, 1, 1, cfg.frames, _native.GREEDY_SINGLE, tok, ts, n, cfg.frames), 5)
        emit(config="cfg1", workload=cfg.name, mode="greedy_search single stream, " + label, frames_per_s=cfg.frames / (ms * 1e-3),
             ms_per_utterance=ms, us_per_frame_step=ms * 1e3 / cfg.frames, emitted=int(n.item()))
        h.close()

for _prec3 in (("fp32", "bf16x3") if "cfg3" in which else ()):
    cfg = synth.CONFIGS["cfg3"]
    h = handle(cfg, _prec3)
    B, Tc, C = cfg.streams, cfg.frames, cfg.chunks
    x = torch.from_numpy(synth.make_frames(B, Tc * C, cfg.dims.encoder_dim, cfg.seed)).to(dev)
    chunks = [x[:, c * Tc:(c + 1) * Tc].contiguous() for c in range(C)]
    hyp = torch.zeros((B, 2), dtype=torch.int64, device=dev)
    tok, ts, n, sc = outbufs(B, Tc)

    def run():
        hyp.zero_()
        for c in range(C):
            h.call("k2b_greedy_online_chunk_dev", chunks[c], 1, B, Tc, hyp, tok, ts, n, Tc)
    ms = timed(run, 2, warm=1)
    emit(config="cfg3", workload=cfg.name, mode="online greedy, 32 chunks x 8 frames, " + ("per-frame fp32 path" if _prec3 == "fp32" else ("16-CTA cluster kernel per chunk" if os.environ.get("K2B_GREEDY_PERSISTENT") == "0" else "persistent beam kernel as beam 1 per chunk") + ", split-bf16 x3"), frames_per_s=B * Tc * C / (ms * 1e-3),
         ms_per_32_chunks=ms, us_per_frame_step=ms * 1e3 / (Tc * C))
    h.close()

for _prec4 in (("fp32", "bf16x3", "bf16") if "cfg4" in which else ()):
    cfg = synth.CONFIGS["cfg4"]
    h = handle(cfg, _prec4)
    B, T = cfg.streams, cfg.frames
    x = torch.from_numpy(synth.make_frames(B, T, cfg.dims.encoder_dim, cfg.seed)).to(dev)
    tok, ts, n, sc = outbufs(B, T)
    ms = timed(lambda: h.call("k2b_modified_beam_search_dev", x, 1, B, T, 4, tok, ts, n, sc, T), 2, warm=1)
    emit(config="cfg4", workload=cfg.name, mode="modified_beam_search V=5537, " + ("per-frame path, fp32 CUDA-core joiner" if _prec4 == "fp32" else "persistent beam kernel (one launch), tcgen05 joiner " + _prec4), frames_per_s=B * T / (ms * 1e-3), ms_per_batch=ms,
         us_per_frame_step=ms * 1e3 / T, roofline_frames_per_s=61.1e6)
    h.close()

if "cfg5" in which:
    cfg = synth.CONFIGS["cfg5"]
    h = _native.Handle(vocab_size=64, joiner_dim=64, decoder_dim=64)
    h.set_stream(stream.cuda_stream)
    for (B, V) in ((128, 2000), (1024, 2000), (128, 5537), (368, 5537)):
        T = cfg.frames
        g = torch.Generator(device=dev); g.manual_seed(cfg.seed)
        logp = torch.randn((B, T, V), device=dev, generator=g) * 3.0
        logp[:, :, 0] += 11.5
        logp = torch.log_softmax(logp, dim=-1).contiguous()
        tok, ts, n, sc = outbufs(B, T)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def run():
            h.call("k2b_ctc_greedy_dev", logp, B, T, V, 0, None, None, tok, ts, n, None, T)
        # inputs >= 256 MB (> 126 MB L2) except the first shape, which gets an explicit L2 flush between iterations
        small = logp.numel() * 4 < 200e6
        if small:
            times = []
            for _ in range(6):
                flush.fill_(1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = float(np.median(times[1:]))
        else:
            ms = timed(run, 10)
        gbs = B * T * V * 4 / (ms * 1e-3) / 1e9
        emit(config="cfg5", workload=cfg.name, B=B, T=T, V=V, mode="ctc_greedy fused kernel", frames_per_s=B * T / (ms * 1e-3), ms=ms,
             roofline={"bound": "hbm", "achieved": gbs, "peak": PEAK_HBM, "unit": "GB/s", "frac": gbs / PEAK_HBM,
                       "algorithmic_bytes": B * T * V * 4, "l2": "explicit flush" if small else "input larger than L2"})
        del logp
    h.close()
