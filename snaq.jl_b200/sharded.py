"""Quartet-sharded objective (case 2): every rank holds a contiguous block of the quartet
list, evaluates the same batch of parameter vectors on its block, and the per-candidate partial sums are
combined with ONE all-reduce of P doubles (NCCL over NVLink on GPUs; gloo in the CPU tests).

Runs / bootstrap replicates need none of this: they are one engine per GPU with no communication.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(nq: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block partition used by ``snaq_set_data`` (csrc/snaq_engine.cu): rank r keeps
    [r*nq/G, (r+1)*nq/G) of the quartet list (order preserved, so descriptor locality is kept)."""
    return nq * rank // world_size, nq * (rank + 1) // world_size


class ShardedObjective:
    """``eval(X)`` returns the full pseudo-deviance of every row of X on every rank.

    local_eval: callable X -> partial sums of this rank's shard.  On a GPU box it is built by
    :meth:`from_engine` (the engine was created with ``rank``/``world_size`` and fills a CUDA tensor that NCCL
    reduces in place); tests inject a CPU evaluator.
    """

    def __init__(self, local_eval: Callable[[np.ndarray], "object"], group=None):
        self.local_eval = local_eval
        self.group = group

    @classmethod
    def from_engine(cls, engine, group=None) -> "ShardedObjective":
        import torch

        dev = torch.device("cuda", engine.device)
        stream = torch.cuda.Stream(dev)

        def local_eval(X):
            dX = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
            out = torch.empty(dX.shape[0], dtype=torch.float64, device=dev)
            torch.cuda.current_stream(dev).synchronize()
            engine.eval_device(dX, out, stream)
            stream.synchronize()
            engine.check_status()
            return out

        return cls(local_eval, group)

    def eval(self, X) -> np.ndarray:
        import torch
        import torch.distributed as dist

        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(1, -1)
        part = self.local_eval(X)
        t = part if isinstance(part, torch.Tensor) else torch.as_tensor(np.asarray(part, dtype=np.float64))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.detach().cpu().numpy()
