"""Utterance scheduling: launch groups on one GPU (frame_balanced_groups; length_buckets for padded rectangles) and sharding across
the GPUs of a box.

Utterances are independent (reference `_infer` couples them only through padding to the batch maximum,
cpp/helper.cpp:376, 430), so there is no collective on the data path: every GPU holds a full weight replica
and processes its own shard (north_star).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def length_buckets(lengths: Sequence[int], max_batch: int = 64, max_pad: float = 1.35) -> List[List[int]]:
    """Sort by length, then cut greedily: a group closes when it is full or when its longest member would exceed
    `max_pad` x its shortest (bounds the padded-rectangle waste of each sub-batch)."""
    order = np.argsort(np.asarray(lengths), kind="stable")
    groups: List[List[int]] = []
    cur: List[int] = []
    for i in order:
        i = int(i)
        if cur and (len(cur) >= max_batch or lengths[i] > max_pad * max(lengths[cur[0]], 1)):
            groups.append(cur)
            cur = []
        cur.append(i)
    if cur:
        groups.append(cur)
    return groups


# One wave of the fused ConvNeXt MLP at one hidden slice per row tile = one 128-row tile per SM (148 on B200); a launch group a few
# tiles larger takes a second wave (csrc/model.cu mlp_plan), so groups are filled to 140 tiles of PREDICTED latent frames.
GROUP_ROWS = 140 * 128


def frame_balanced_groups(lengths: Sequence[int], max_batch: int = 128, frames_per_token: float = 1.0,
                          max_rows: int = GROUP_ROWS) -> List[List[int]]:
    """Launch groups for the packed path (`synthesize_many`): as few groups as `max_batch` utterances and `max_rows` predicted
    latent frames per group allow, filled to EQUAL predicted frames (longest utterance first into the emptiest group). Packed rows
    carry no padding, so a group need not hold similar lengths; what costs is a group whose row tiles do not fill the SMs (small
    groups run at a fraction of the rate: 32 utterances 40 k audio-s/s, 128 utterances 57 k) or spill into a second wave. Sorting
    by length (length_buckets) made one group of a 1 024-utterance request 29 row tiles and another 223. Indices inside a group
    are ascending; results do not depend on the grouping (noise streams are keyed by the index in the request)."""
    n = len(lengths)
    if n == 0:
        return []
    # The groups themselves are balanced on the integer token counts (predicted frames are proportional to them), so the same
    # request always gives the same groups — and hits the same CUDA graphs — however the measured ratio jitters between calls; the
    # ratio, rounded up to a multiple of 1/16, only decides how many groups there are.
    tok = np.maximum(np.asarray(lengths, dtype=np.int64), 1)
    fpt = np.ceil(max(float(frames_per_token), 1e-3) * 16.0) / 16.0
    n_groups = min(n, max(-(-n // max(1, int(max_batch))), int(np.ceil(float(tok.sum()) * fpt / max(1, int(max_rows))))))
    load = [0] * n_groups
    groups: List[List[int]] = [[] for _ in range(n_groups)]
    for i in sorted(range(n), key=lambda k: (-int(tok[k]), k)):
        g = min((k for k in range(n_groups) if len(groups[k]) < max_batch), key=lambda k: (load[k], k))
        groups[g].append(i)
        load[g] += int(tok[i])
    return [sorted(g) for g in groups if g]


def shard_lpt(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of work items to `world_size` replicas (deterministic)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += costs[i]
    return [sorted(s) for s in shards]


def synth_cost(n_tokens: int, total_step: int) -> float:
    """Relative cost of one utterance: frames ∝ tokens; VE runs `total_step` times at 1x frame rate, the vocoder once
    at 6x with 2x the width (4x the FLOPs per frame)."""
    return float(n_tokens) * (total_step * 1.0 + 6 * 4.0 * 0.35)


def shard_for_rank(n_tokens: Sequence[int], total_step: int, rank: int, world_size: int) -> List[int]:
    """Indices of the utterances rank `rank` synthesises: LPT over `synth_cost`, identical on every rank (pure function of
    the token counts), so no rank needs to talk to another to know its share."""
    return shard_lpt([synth_cost(int(t), total_step) for t in n_tokens], world_size)[rank]


def reduce_throughput(audio_seconds: float, elapsed_ms: float, world_size: int = 1):
    """Whole-job audio-sec/sec from per-rank (audio, device time): sum of the audio over ranks / MAX of the time over ranks.
    Uses torch.distributed when a process group is up (NCCL on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(elapsed_ms), float(audio_seconds)], dtype=torch.float64)
    if world_size > 1 and dist.is_initialized():
        if dist.get_backend() == "nccl":
            t = t.cuda()
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, audio = float(tmax[0]), float(tsum[1])
    else:
        ms, audio = float(t[0]), float(t[1])
    return audio / (ms / 1000.0), ms, audio
