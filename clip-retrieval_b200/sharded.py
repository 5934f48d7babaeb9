"""Range-sharded search across the GPUs of one box.

One process per GPU (torchrun); rank g holds rows [g*N/G, (g+1)*N/G) of the index with
`id_base = g*N/G`, every rank sees the same queries, searches its shard, and ONE collective — an
all-gather of the per-shard top-k candidates (fp32 score + int64 id, 12 bytes per candidate) —
precedes a device-side merge.  The reference has no sharded search (its back end serves one CPU
FAISS index per process); the query API stays `search(x, k) -> (D, I)`.
"""
import numpy as np


def shard_range(n_total, world_size, rank):
    """Rows [lo, hi) of rank `rank`: contiguous, sizes differ by at most one (the task split of
    clip_retrieval/clip_inference/slurm_worker.py:16-37 applied to rows)."""
    base, extra = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


class ShardedIndex:
    """Wraps this rank's shard.  `local_index` must already hold rows shard_range(...) and have
    `id_base` set to the shard's first global row.  `merge_fn(Dg, Ig, k)` merges [G, nq, k]
    candidates (default: the CUDA merge kernel)."""

    def __init__(self, local_index, group=None, merge_fn=None, device=None):
        import torch.distributed as dist

        self.local = local_index
        self.group = group
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if merge_fn is None:
            from .index import merge_shard_results

            merge_fn = merge_shard_results
        self.merge_fn = merge_fn
        self.device = device

    @property
    def d(self):
        return self.local.d

    def search_device(self, q, k):
        """q: float32 [nq, d] tensor on this rank's device (identical on every rank).
        Returns (D, I) [nq, k] tensors, identical on every rank."""
        import torch

        D, I = self.local.search_device(q, k)
        if self.world == 1:
            return D, I
        Dg = torch.empty((self.world,) + tuple(D.shape), dtype=D.dtype, device=D.device)
        Ig = torch.empty((self.world,) + tuple(I.shape), dtype=I.dtype, device=I.device)
        # the single exchange step: candidates are packed so that one all-gather moves both arrays
        packed = torch.cat([I.view(torch.int32).reshape(-1), D.view(torch.int32).reshape(-1)])
        flat = torch.empty(self.world * packed.numel(), dtype=torch.int32, device=packed.device)
        self.dist.all_gather_into_tensor(flat, packed, group=self.group)
        gathered = flat.view(self.world, packed.numel())
        nI = 2 * I.numel()
        Ig.copy_(gathered[:, :nI].clone().view(torch.int64).reshape(self.world, *I.shape))
        Dg.copy_(gathered[:, nI:].clone().view(torch.float32).reshape(self.world, *D.shape))
        return self.merge_fn(Dg, Ig, k)

    def search(self, x, k):
        """FAISS-style host call: numpy in, numpy out."""
        import torch

        x = np.ascontiguousarray(x, dtype=np.float32)
        dev = self.device if self.device is not None else "cuda"
        D, I = self.search_device(torch.from_numpy(x).to(dev), k)
        return D.cpu().numpy(), I.cpu().numpy()
