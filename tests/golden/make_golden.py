"""Generates tests/golden/*.npz from the CPU oracle (oracle/k2_oracle.py).

The reference ships no golden vectors for this path and cannot be executed in this image (no .NET, no
onnxruntime), so these fixtures pin the ORACLE's behaviour on small seeded inputs,
including the crafted quirk cases of SURVEY.md section 3.6. Run from the repo root:
    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from k2transducerasr_b200 import synth          # noqa: E402
from oracle import k2_oracle as O                # noqa: E402

OUT = Path(__file__).resolve().parent
DIMS = synth.ModelDims(vocab_size=53, joiner_dim=32, decoder_dim=32, encoder_dim=48)


def ragged(lists, fill=-7):
    n = max((len(x) for x in lists), default=0)
    a = np.full((len(lists), max(n, 1)), fill, np.int64)
    for i, x in enumerate(lists):
        a[i, :len(x)] = x
    return a, np.array([len(x) for x in lists], np.int64)


def main():
    w = synth.make_weights(DIMS, blank_bias=0.6)
    m = O.Model.from_dict(w)
    raw = synth.make_frames(5, 24, DIMS.encoder_dim, 4242)
    enc = O.encoder_proj(m, raw)
    y = np.array([[-1, 0], [0, 0], [3, 7], [52, 1], [2, 2]], np.int64)
    dec = O.decoder(m, y)
    lg = O.joiner(m, enc[:, 0, :], dec)
    fix = {"raw": raw, "enc": enc, "y": y, "dec": dec, "logits": lg}

    single = O.greedy_search_single(m, enc[0])
    fix["single_tokens"] = np.array(single.tokens, np.int64)
    fix["single_ts"] = np.array(single.timestamps, np.int64)
    for name, res in (("compat", O.greedy_search_batch(m, enc, compat=True)),
                      ("perstream", O.greedy_search_batch(m, enc, compat=False)),
                      ("mbs4", O.modified_beam_search(m, enc, 4)),
                      ("mbs2", O.modified_beam_search(m, enc, 2))):
        fix[f"{name}_tokens"], fix[f"{name}_ntok"] = ragged([r.tokens for r in res])
        fix[f"{name}_ts"], fix[f"{name}_nts"] = ragged([r.timestamps for r in res])
        fix[f"{name}_score"] = np.array([r.score for r in res], np.float32)
        fix[f"{name}_gap"] = np.array([r.min_gap for r in res], np.float64)

    # online: 3 chunks of 8 frames
    hyps = [[0, 0] for _ in range(5)]
    toks = [[0, 0] for _ in range(5)]
    tss = [[] for _ in range(5)]
    for c in range(3):
        res = O.greedy_search_online_chunk(m, enc[:, 8 * c:8 * c + 8], hyps, toks)
        hyps = [r.hyp for r in res]
        toks = [r.tokens for r in res]
        for b, r in enumerate(res):
            tss[b] += r.timestamps
    fix["online_tokens"], fix["online_ntok"] = ragged(toks)
    fix["online_ts"], fix["online_nts"] = ragged(tss)
    fix["online_hyp"] = np.array(hyps, np.int64)

    # CTC incl. crafted ties (Q2), an all-blank stream and a repeat across what will be a chunk edge (Q10)
    lp = synth.make_ctc_logp(4, 20, 37, 99, blank_bias=3.0)
    lp[1, :, :] = -5.0
    lp[1, :, 0] = -0.1                       # all blank
    lp[2, 3, :] = -3.0
    lp[2, 3, 5] = lp[2, 3, 9] = -0.5         # exact tie -> lowest index 5
    lp[3, 9, :] = -4.0; lp[3, 9, 11] = -0.2  # token 11 on both sides of frame 10 (chunk edge for T'=10)
    lp[3, 10, :] = -4.0; lp[3, 10, 11] = -0.2
    res = O.ctc_greedy_search(lp, blank=0, frame_offset=[0, 5, 0, 0], trailing_blank=[0, 2, 0, 0])
    fix["ctc_logp"] = lp
    fix["ctc_tokens"], fix["ctc_ntok"] = ragged([r.appended for r in res])
    fix["ctc_ts"], fix["ctc_nts"] = ragged([r.timestamps for r in res])
    fix["ctc_trailing"] = np.array([r.num_trailing_blank for r in res], np.int64)
    np.savez_compressed(OUT / "small_v53.npz", **fix)
    print("wrote", OUT / "small_v53.npz", {k: v.shape for k, v in fix.items()})


if __name__ == "__main__":
    main()
