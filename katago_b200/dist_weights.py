"""Model-weight broadcast: the only collective on the self-play path.

The reference distributes new nets as files that every process polls for (cpp/command/selfplay.cpp:142-231,336-352);
here rank `src` reads / synthesises the model file once and its bytes travel to the other ranks with one
torch.distributed broadcast (NCCL over NVLink on the GPU box, gloo in the CPU tests).  Games themselves shard across
ranks with no data-path collective."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


def broadcast_model_bytes(data: Optional[bytes], src: int = 0, device: Optional[torch.device] = None) -> bytes:
    """Every rank returns the bytes rank `src` passed in (`data` is ignored on the other ranks)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert data is not None
        return data
    device = device or torch.device("cpu")
    rank = dist.get_rank()
    if rank == src:
        arr = np.frombuffer(data, dtype=np.uint8)
        size = torch.tensor([arr.size], dtype=torch.int64, device=device)
    else:
        size = torch.zeros(1, dtype=torch.int64, device=device)
    dist.broadcast(size, src)
    if rank == src:
        buf = torch.from_numpy(arr.copy()).to(device)
    else:
        buf = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    return buf.cpu().numpy().tobytes()


def shard_games(total_games: int, rank: int, world: int):
    """Game g runs on rank g % world; returns this rank's global game ids."""
    return list(range(rank, total_games, world))
