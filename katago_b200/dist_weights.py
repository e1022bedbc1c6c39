"""Model-weight broadcast: the only collective on the self-play path.

The reference distributes new nets as files that every process polls for (cpp/command/selfplay.cpp:142-231,336-352);
here rank `src` reads and packs the model file once and the PACKED weight arena travels from its device memory into the
other ranks' with one ncclBroadcast issued by the library itself (WeightBroadcaster -> kgb_handle_broadcast_staged_weights,
include/kgb200.h); torch.distributed only carries the 128-byte NCCL id once.  broadcast_model_bytes (file bytes through a
torch.distributed broadcast; gloo in the CPU tests) remains for start-up, when the other ranks have no handle yet.  Games
themselves shard across ranks with no data-path collective."""
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


class WeightBroadcaster:
    """New nets into the live compute handles of all ranks of one node.

    Created once per handle (collective: every rank constructs it): rank `src` makes an NCCL id through the library, the id
    reaches the others by one torch.distributed broadcast, every rank joins the library's own communicator.  update() is
    then, per new net: the source rank packs the model for the kernels and stages it in its shadow arena; one ncclBroadcast
    moves the packed arena (fp16 convolution weights, fp32 scales and head matrices) device to device; every rank commits
    between two waves (ordered on the handle's stream), so games keep running across the swap."""

    def __init__(self, handle, src: int = 0, device: Optional[torch.device] = None, id_source=None):
        self.handle, self.src = handle, src
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        if self.world > 1:
            if id_source is None:
                from .nn_backend import nccl_unique_id as id_source
            device = device or torch.device("cpu")
            raw = bytearray(id_source()) if self.rank == src else bytearray(128)
            t = torch.frombuffer(raw, dtype=torch.uint8).clone().to(device)
            dist.broadcast(t, src)
            handle.comm_init(bytes(t.cpu().numpy().tobytes()), self.rank, self.world)

    def update(self, loaded_model=None, selfplay_loops=()) -> float:
        """Collective.  `loaded_model` (nn_backend.LoadedModel) is read on the source rank only.  Returns the broadcast's
        device time in ms on this rank (0.0 in a single-process run)."""
        if self.rank == self.src:
            if loaded_model is None:
                raise ValueError("WeightBroadcaster.update: the source rank needs the model")
            self.handle.stage_weights(loaded_model)      # host: parse + pack (tens of ms for a b18 net), then one H2D copy
            self.handle.wait_staged()
        if self.world > 1:
            dist.barrier()                               # the ranks enter the collective together: its device time is the transfer
        ms = self.handle.broadcast_staged_weights(self.src) if self.world > 1 else 0.0
        self.handle.commit_weights()
        for sp in selfplay_loops:
            sp.clear_nn_cache()
        return ms


def shard_games(total_games: int, rank: int, world: int):
    """Game g runs on rank g % world; returns this rank's global game ids."""
    return list(range(rank, total_games, world))
