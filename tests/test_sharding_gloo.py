"""Multi-GPU path on CPU: world_size-2 `gloo` processes exercise the utterance sharding and the whole-job
throughput reduction that bench.py uses on the B200 box with NCCL (independent utterances, full
weight replicas, NO data-path collective — the only collective is the final timing reduction)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from supertonic_b200 import scheduler as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = S.shard_for_rank(lens, 5, rank, world)
    # every rank derives the same global plan without communicating
    plan = S.shard_lpt([S.synth_cost(t, 5) for t in lens], world)
    assert plan[rank] == mine
    # "synthesis": audio seconds proportional to tokens, rank 1 is the slow one
    audio = float(sum(lens[i] for i in mine)) * 0.065
    ms = 100.0 + 50.0 * rank
    value, ms_max, audio_all = S.reduce_throughput(audio, ms, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mine, value, ms_max, audio_all, gathered))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    rng = np.random.default_rng(1234)
    lens = [int(x) for x in rng.integers(20, 301, size=64)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, v0, ms0, a0, g0), (r1, s1, v1, ms1, a1, g1) = out
    assert sorted(s0 + s1) == list(range(64)) and not set(s0) & set(s1)          # a partition of the request batch
    assert g0 == g1 == [s0, s1]
    c0 = sum(S.synth_cost(lens[i], 5) for i in s0); c1 = sum(S.synth_cost(lens[i], 5) for i in s1)
    assert abs(c0 - c1) <= max(S.synth_cost(t, 5) for t in lens)                   # LPT balance bound
    assert ms0 == ms1 == 150.0                                                     # max over ranks, not rank 0's clock
    assert abs(a0 - sum(lens) * 0.065) < 1e-9 and v0 == v1 == a0 / 0.150


def test_length_buckets_and_lpt_properties():
    rng = np.random.default_rng(0)
    lens = [int(x) for x in rng.integers(1, 400, size=200)]
    groups = S.length_buckets(lens, max_batch=16, max_pad=1.25)
    assert sorted(i for g in groups for i in g) == list(range(200))
    for g in groups:
        assert len(g) <= 16
        assert max(lens[i] for i in g) <= 1.25 * max(min(lens[i] for i in g), 1)
    for w in (1, 2, 4, 8):
        sh = S.shard_lpt([float(x) for x in lens], w)
        assert sorted(i for s in sh for i in s) == list(range(200))
        loads = [sum(lens[i] for i in s) for s in sh]
        assert max(loads) - min(loads) <= max(lens)
    assert S.shard_lpt([], 4) == [[], [], [], []]


def test_frame_balanced_groups_properties():
    """Launch groups of synthesize_many: every utterance once, at most max_batch per group, as few groups as the utterance and
    predicted-frame budgets allow, predicted frames equal to within the longest utterance, budget respected when it can be."""
    rng = np.random.default_rng(1)
    lens = [int(x) for x in rng.integers(22, 305, size=1024)]
    for fpt, mb in ((0.8, 128), (1.0, 128), (1.0, 64), (2.5, 128), (0.8, 1000)):
        groups = S.frame_balanced_groups(lens, mb, fpt)
        assert sorted(i for g in groups for i in g) == list(range(1024))
        assert all(g == sorted(g) and 0 < len(g) <= mb for g in groups)
        fq = np.ceil(fpt * 16) / 16                      # the ratio only sets the group count, rounded up to 1/16
        assert len(groups) == max(-(-1024 // mb), int(np.ceil(sum(lens) * fq / S.GROUP_ROWS)))
        loads = [sum(lens[i] for i in g) for g in groups]
        assert max(loads) - min(loads) <= max(lens)
        if len(groups) > -(-1024 // mb):          # the frame budget set the group count: it holds up to one utterance
            assert max(loads) * fq <= S.GROUP_ROWS + max(lens) * fq
    # the same request gives the same groups whatever the measured ratio does between calls (CUDA graphs are keyed by group shape)
    assert S.frame_balanced_groups(lens, 128, 0.7891) == S.frame_balanced_groups(lens, 128, 0.7924)
    assert S.frame_balanced_groups([], 8) == []
    assert S.frame_balanced_groups([5, 9, 300], 128, 1.0) == [[0, 1, 2]]                      # a small request is one group
    assert S.frame_balanced_groups([10] * 5, 2, 1.0) == [[0, 3], [1, 4], [2]]                 # the utterance cap alone
    assert S.reduce_throughput(10.0, 500.0, 1) == (20.0, 500.0, 10.0)
