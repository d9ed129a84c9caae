"""Utterance scheduling: length bucketing on one GPU and sharding across the GPUs of a box.

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
