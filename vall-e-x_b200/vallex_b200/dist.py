"""Multi-GPU plumbing for the request-sharded path.

Utterances are independent, so the only collective is the start-up broadcast of the checkpoint from rank 0
(NCCL over NVLink/NVSwitch on the GPU box; gloo in the CPU tests).  After that every rank serves its own
shard of the request list with no steady-state exchange; timing is reduced with a MAX over ranks.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_units(n_units: int, rank: int, world: int) -> List[int]:
    """Round-robin request sharding: GPU g takes utterances g, g+G, g+2G, ...."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_units, world))


def state_dict_manifest(sd: Dict[str, torch.Tensor]) -> List[Tuple[str, Tuple[int, ...]]]:
    return [(k, tuple(v.shape)) for k, v in sd.items()]


def broadcast_state_dict(sd: Optional[Dict[str, torch.Tensor]], src: int = 0, device: str = "cpu",
                         chunk_elems: int = 64 * 1024 * 1024) -> "OrderedDict[str, torch.Tensor]":
    """Rank `src` passes the fp32 state dict, the others pass None; everybody returns the full dict as views of the
    receive buffers ON `device` (cuda: the tensors stay in HBM and `Engine._load` ingests them device-to-device through
    vx_load_weight_mem -- nothing bounces through the host; cpu: the gloo test path).
    The payload travels as a few large flat buffers (bucket size chosen for launch latency, not link count).
    Tied tensors (nar_predict_layers.j <-> nar_audio_embeddings.j+2) are sent once and re-tied."""
    rank = dist.get_rank()
    meta = [None]
    if rank == src:
        assert sd is not None
        seen, manifest = {}, []
        for k, v in sd.items():
            alias = seen.get(v.data_ptr())
            manifest.append((k, tuple(v.shape), alias))
            if alias is None:
                seen[v.data_ptr()] = k
        meta = [manifest]
    dist.broadcast_object_list(meta, src=src)
    manifest = meta[0]
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    pending: List[Tuple[str, Tuple[int, ...]]] = [(k, s) for k, s, a in manifest if a is None]
    i = 0
    while i < len(pending):
        group, n = [], 0
        while i < len(pending) and (not group or n + _numel(pending[i][1]) <= chunk_elems):
            group.append(pending[i])
            n += _numel(pending[i][1])
            i += 1
        flat = torch.empty(n, dtype=torch.float32, device=device)
        if rank == src:
            off = 0
            for k, s in group:
                m = _numel(s)
                flat[off:off + m].copy_(sd[k].reshape(-1))
                off += m
        dist.broadcast(flat, src=src)
        off = 0
        for k, s in group:
            m = _numel(s)
            out[k] = flat[off:off + m].reshape(s)   # a view of the receive buffer, wherever it lives
            off += m
    ordered: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, s, a in manifest:
        ordered[k] = out[k] if a is None else out[a]
    return ordered


def _numel(shape: Sequence[int]) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


def max_over_ranks(value: float, device: str = "cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
