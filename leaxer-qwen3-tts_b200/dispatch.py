"""Request-level data parallelism: utterances are independent, so N GPUs = N engines, each owning
whole utterances; there is NO collective on the data path. This module holds the (pure, CPU-testable) host logic:

  * `assign(costs, world)`  longest-processing-time-first partition of utterances over ranks by expected frames
    (random-init models never emit EOS, so cost = max_new_tokens; src/tts_onnx.cpp:782-849 runs until EOS or the cap);
  * `utterance_key(seed, index)`  the Philox key of an utterance depends on its GLOBAL index only, so results are
    invariant to the GPU count and to which rank ran it;
  * `gather_results`  the only communication: rank 0 collects (index, n_frames, checksum) records after the timed region
    (`torch.distributed` gather_object over gloo or nccl) -- bookkeeping, not data path.
"""
from __future__ import annotations

import heapq
from typing import Iterable, List, Sequence, Tuple


def assign(costs: Sequence[int], world: int) -> List[List[int]]:
    """LPT: sort by cost descending (ties: lower index first), give each utterance to the least-loaded rank
    (ties: lower rank). Deterministic; every rank computes the same table locally (no broadcast needed)."""
    if world <= 0:
        raise ValueError("world must be positive")
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    heap: List[Tuple[int, int]] = [(0, r) for r in range(world)]
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + int(costs[i]), r))
    for lst in out:
        lst.sort()
    return out


def utterance_key(seed: int, index: int) -> Tuple[int, int]:
    """(Philox key word 0, key word 1) = (seed, global utterance index): see lqt_sampling in include/lqt_b200.h"""
    return int(seed) & 0xFFFFFFFF, int(index) & 0xFFFFFFFF


def makespan(costs: Sequence[int], table: Iterable[Iterable[int]]) -> int:
    return max((sum(int(costs[i]) for i in part) for part in table), default=0)


def gather_results(local: list, dist=None, dst: int = 0):
    """local: list of picklable per-utterance records of this rank -> on rank `dst` the concatenation sorted by the
    record's first field (global utterance index); None elsewhere. dist = torch.distributed (initialised) or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return sorted(local, key=lambda r: r[0])
    bucket = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, bucket, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = [rec for part in bucket for rec in part]
    return sorted(merged, key=lambda r: r[0])
