""" One process per GPU; work is sharded only where the maths is independent (cross-validation folds, outputs, Sobol input subsets)
and results are gathered with small collectives (NCCL on GPUs, gloo in the CPU tests).  No collective sits on the data path."""
from __future__ import annotations

import os
from typing import Any, List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def is_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def rank() -> int:
    return dist.get_rank() if is_initialized() else 0


def world_size() -> int:
    return dist.get_world_size() if is_initialized() else 1


def init_from_env(backend: str | None = None) -> bool:
    """ Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / MASTER_*); no-op for a single process."""
    if is_initialized() or int(os.environ.get('WORLD_SIZE', '1')) <= 1:
        return is_initialized()
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group(backend=backend)
    return True


def shard(items: Sequence[Any], r: int | None = None, w: int | None = None) -> List[Any]:
    """ Round-robin shard of ``items`` owned by rank r of w (defaults: this process)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    return [item for i, item in enumerate(items) if i % w == r]


def owner(index: int, w: int | None = None) -> int:
    return index % (world_size() if w is None else w)


def barrier():
    if is_initialized():
        dist.barrier()


def _device_for_collectives() -> torch.device:
    return torch.device('cuda', torch.cuda.current_device()) if is_initialized() and dist.get_backend() == 'nccl' else torch.device('cpu')


def all_gather_rows(local: np.ndarray, total_rows: int) -> np.ndarray:
    """ Each rank holds the rows ``i`` with ``i % world == rank`` of a (total_rows, ...) array, in order; returns the full array
    on every rank.  One all_gather of equally padded blocks."""
    local = np.asarray(local, dtype=np.float64)
    w = world_size()
    if w == 1:
        return local
    per = (total_rows + w - 1) // w
    tail = local.shape[1:]
    block = np.zeros((per,) + tail, dtype=np.float64)
    block[:local.shape[0]] = local
    dev = _device_for_collectives()
    mine = torch.from_numpy(block).to(dev)
    pieces = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(pieces, mine)
    out = np.zeros((total_rows,) + tail, dtype=np.float64)
    for r, piece in enumerate(pieces):
        rows = list(range(r, total_rows, w))
        out[rows] = piece.cpu().numpy()[:len(rows)]
    return out


def shard_blocks(items: Sequence[Any], block: int, r: int | None = None, w: int | None = None) -> List[Any]:
    """ Round-robin shard of consecutive blocks of ``block`` items (block b goes to rank b % w): what keeps the subsets that share their
    high mask bits - and so their partial products in the all-subsets Sobol kernel - on one rank."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    return [item for i, item in enumerate(items) if (i // block) % w == r]


def all_gather_rows_tensor(local: torch.Tensor, total_rows: int, block: int = 1) -> torch.Tensor:
    """ ``all_gather_rows`` for a tensor that already lives where the collective runs (CUDA for NCCL): no host hop, one
    all_gather_into_tensor on the current stream, rows put back in order on the device.  Rank r holds, in order, the rows of the
    blocks b (of ``block`` consecutive rows) with b % world == r (``shard`` for block = 1, ``shard_blocks`` otherwise)."""
    w = world_size()
    if w == 1:
        return local
    nblocks = (total_rows + block - 1) // block
    per = (nblocks + w - 1) // w * block
    mine = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    mine[:local.shape[0]] = local
    gathered = torch.empty((w * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, mine)
    i = torch.arange(total_rows, device=local.device)
    b = i // block
    return gathered[(b % w) * per + (b // w) * block + i % block]


def all_reduce_max(value: float) -> float:
    if not is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_collectives())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_reduce_sum(value: float) -> float:
    if not is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_collectives())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def all_reduce_sum_tensor(t: torch.Tensor) -> torch.Tensor:
    """ In-place sum over the ranks of a device (NCCL) or host (gloo) tensor; identity for a single process."""
    if is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
