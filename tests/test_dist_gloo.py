"""CPU, world_size 2 over gloo: the N>1 path (shard -> independent search -> gather) with the search itself
replaced by the oracle, so that the sharding and the gather are what is under test."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from k2transducerasr_b200 import dist as kd
from k2transducerasr_b200 import synth
from oracle import k2_oracle as O
from tests.helpers import SMALL, model_and_weights


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 256, 1000):
        for w in (1, 2, 3, 8):
            spans = [kd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, _ = model_and_weights(SMALL, blank_bias=0.6)
        enc = O.encoder_proj(m, synth.make_frames(7, 12, SMALL.encoder_dim, 5))
        a, b = kd.shard_range(7, rank, world)
        res = O.modified_beam_search(m, enc[a:b], 2)           # per-stream independent => shardable
        toks, tss, sc = kd.gather_results([r.appended for r in res], [r.timestamps for r in res],
                                          [r.score for r in res], 7, 12)
        if rank == 0:
            q.put((toks, tss, sc))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    toks, tss, sc = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, _ = model_and_weights(SMALL, blank_bias=0.6)
    enc = O.encoder_proj(m, synth.make_frames(7, 12, SMALL.encoder_dim, 5))
    want = O.modified_beam_search(m, enc, 2)
    assert toks == [r.appended for r in want]
    assert tss == [r.timestamps for r in want]
    np.testing.assert_allclose(sc, [r.score for r in want], atol=1e-6)


def test_pack_roundtrip():
    p = kd.pack_results([[1, 2, 3], []], [[0, 4, 9], []], [-1.5, 0.25], 5)
    t, s, sc = kd.unpack_results(p, 5)
    assert t == [[1, 2, 3], []] and s == [[0, 4, 9], []] and sc == [-1.5, 0.25]
