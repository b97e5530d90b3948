"""Seeded synthetic workloads: random-init decoder/joiner weights and encoder frames.

There is no network, no .onnx file and no audio here, so every config of BASELINE.json runs on
synthetic encoder output of the named shape and random-init weights of the named architecture
(SURVEY.md section 8d). Initialisation follows the PyTorch defaults of the upstream modules the
ONNX files are exported from: Embedding ~ N(0,1); Conv1d / Linear weights and biases
~ U(+-1/sqrt(fan_in)).  `blank_bias` is added to the joiner output bias of the blank symbol so
that a chosen fraction of frames decodes to blank ("speech-like" regime); 0 gives the "raw"
regime in which almost every frame emits.

This module is product-side plumbing (bench.py, smoke and the tests all use it); it never imports
the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

WEIGHT_SEED = 7


@dataclass(frozen=True)
class ModelDims:
    """The ONNX custom-metadata contract (ref OfflineModel.cs:31-46) plus the widths the
    reference never sees (decoder_dim, raw encoder_dim)."""

    vocab_size: int = 500
    joiner_dim: int = 512
    decoder_dim: int = 512
    encoder_dim: int = 0
    context_size: int = 2
    blank_id: int = 0
    sos_eos_id: int = 1
    unk_id: int = 2


@dataclass(frozen=True)
class WorkloadConfig:
    """One row of SURVEY.md section 8's config table."""

    name: str
    mode: str  # greedy_single | greedy_batch | greedy_online | mbs | ctc
    dims: ModelDims
    streams: int
    frames: int  # T (offline) or T' per chunk (online)
    chunks: int = 1
    beam: int = 4
    seed: int = 0
    blank_bias: float = 0.0
    extra: dict = field(default_factory=dict)


# cfg1..cfg5 of BASELINE.json; blank_bias values calibrated so that roughly 70-80 % of frames
# are blank under greedy search (see tests/golden/calibrate_blank_bias.py for how they were obtained).
CONFIGS = {
    "cfg1": WorkloadConfig("zipformer-small-en offline greedy B=1", "greedy_single",
                           ModelDims(500, 512, 512, 256), 1, 250, seed=1001, blank_bias=0.99),
    "cfg2": WorkloadConfig("zipformer-large-en offline modified_beam_search K=4 B=256", "mbs",
                           ModelDims(500, 512, 512, 768), 256, 250, beam=4, seed=1002, blank_bias=0.99),
    "cfg3": WorkloadConfig("zipformer multi-zh-hans streaming greedy 512 streams", "greedy_online",
                           ModelDims(2000, 512, 512, 512), 512, 8, chunks=32, seed=1003, blank_bias=1.17),
    "cfg4": WorkloadConfig("zipformer-zh wenetspeech offline modified_beam_search K=4 B=256 V=5537", "mbs",
                           ModelDims(5537, 512, 512, 512), 256, 250, beam=4, seed=1004, blank_bias=1.22),
    "cfg5": WorkloadConfig("zipformer-ctc-large-zh offline CTC greedy B=1024", "ctc",
                           ModelDims(2000, 512, 512, 0), 1024, 250, seed=1005, blank_bias=11.5),
}


def _uniform(rng: np.random.Generator, shape, fan_in: int) -> np.ndarray:
    bound = 1.0 / np.sqrt(float(fan_in))
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def make_weights(dims: ModelDims, seed: int = WEIGHT_SEED, blank_bias: float = 0.0) -> dict:
    """Random-init weights in the layouts include/k2b200.h documents for k2b_load_weights."""
    rng = np.random.default_rng(seed)
    V, J, D, E, ctx = dims.vocab_size, dims.joiner_dim, dims.decoder_dim, dims.encoder_dim, dims.context_size
    w = {
        "emb": rng.standard_normal((V, D), dtype=np.float32),
        "conv_w": _uniform(rng, (D, 4, ctx), 4 * ctx),
        "dec_proj_w": _uniform(rng, (J, D), D),
        "dec_proj_b": _uniform(rng, (J,), D),
        "out_w": _uniform(rng, (V, J), J),
        "out_b": _uniform(rng, (V,), J),
    }
    if E > 0:
        w["enc_proj_w"] = _uniform(rng, (J, E), E)
        w["enc_proj_b"] = _uniform(rng, (J,), E)
    else:
        w["enc_proj_w"] = None
        w["enc_proj_b"] = None
    if blank_bias:
        w["out_b"] = w["out_b"].copy()
        w["out_b"][dims.blank_id] += np.float32(blank_bias)
    return w


def make_frames(streams: int, frames: int, width: int, seed: int) -> np.ndarray:
    """Encoder output [B,T,width] ~ N(0,1) fp32 (raw when width == E, projected when width == J)."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((streams, frames, width), dtype=np.float32)


def make_ctc_logp(streams: int, frames: int, vocab: int, seed: int, blank_id: int = 0,
                  blank_bias: float = 0.0, scale: float = 3.0, repeat_prob: float = 0.3) -> np.ndarray:
    """CTC encoder output [B,T,V]: log_softmax(scale * N(0,1)) with a blank bias, and with each
    frame copying its predecessor's scores with probability `repeat_prob` so that the
    repeat-collapse branch of the search (ref OfflineRecognizer.cs:346) is exercised."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((streams, frames, vocab), dtype=np.float32) * np.float32(scale)
    if blank_bias:
        x[:, :, blank_id] += np.float32(blank_bias)
    if repeat_prob > 0 and frames > 1:
        rep = rng.random((streams, frames)) < repeat_prob
        rep[:, 0] = False
        for t in range(1, frames):
            r = rep[:, t]
            if r.any():
                x[r, t, :] = x[r, t - 1, :]
    m = x.max(axis=-1, keepdims=True)
    z = x - m
    lse = np.log(np.exp(z, dtype=np.float32).sum(axis=-1, keepdims=True, dtype=np.float32))
    return (z - lse).astype(np.float32)
