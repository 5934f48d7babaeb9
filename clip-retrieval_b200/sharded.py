"""Range-sharded search across the GPUs of one box.

`ShardedIndex` — one process per GPU (torchrun): rank g holds rows [g*N/G, (g+1)*N/G) of the index with
`id_base = g*N/G` (flat rows, or its own inverted lists over the replicated global centroids — the union
over shards of list l is the global list l, so with equal `nprobe` the sharded result equals the
single-index result); every rank sees the same queries, searches its shard, and ONE collective — an
all-gather of the per-shard top-k candidates (int64 id + fp32 score, 12 bytes per candidate) — precedes a
device-side merge.  The shard's search writes its (I, D) block straight into the collective's send buffer
and the merge kernel reads the receive buffer in place: no pack / unpack kernels around the all-gather.

`B200ShardedIndex` — one process, one host thread, all GPUs (C ABI `b200_sharded_search`): the exchange is
the search epilogue's own peer stores over NVLink into the root's buffer (or an NCCL all-gather when
communicators are passed).

The reference has no sharded search (its back end serves one CPU FAISS index per process,
clip_back.py:781-782); the query API stays `search(x, k) -> (D, I)` (clip_back.py:362)."""
import ctypes as C

import numpy as np


def shard_range(n_total, world_size, rank):
    """Rows [lo, hi) of rank `rank`: contiguous, sizes differ by at most one (the task split of
    clip_retrieval/clip_inference/slurm_worker.py:16-37 applied to rows)."""
    base, extra = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def block_bytes(nq, k):
    """Size of one shard's packed candidate block [I int64 nq*k | D f32 nq*k], padded to 256 bytes."""
    return (nq * k * 12 + 255) // 256 * 256


class ShardedIndex:
    """Wraps this rank's shard.  `local_index` must already hold rows shard_range(...) and have
    `id_base` set to the shard's first global row.  `merge_fn(Dg, Ig, k)` merges [G, nq, k]
    candidates (default: the CUDA merge kernel reading the gathered buffer in place)."""

    def __init__(self, local_index, group=None, merge_fn=None, device=None):
        import torch.distributed as dist

        self.local = local_index
        self.group = group
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.merge_fn = merge_fn
        self.device = device
        self._bufs = {}        # (nq, k) -> (send uint8 [block], recv uint8 [G * block])
        self.last_gathered = None   # (recv buffer, block bytes, nq, k) of the last multi-rank search (parity checks)

    @property
    def d(self):
        return self.local.d

    @property
    def ntotal(self):
        return self.local.ntotal

    @property
    def nprobe(self):
        return self.local.nprobe

    @nprobe.setter
    def nprobe(self, v):
        self.local.nprobe = v

    def _buffers(self, nq, k, device):
        import torch

        key = (nq, k, str(device))
        b = self._bufs.get(key)
        if b is None:
            if len(self._bufs) > 8:
                self._bufs.clear()
            blk = block_bytes(nq, k)
            b = (torch.zeros(blk, dtype=torch.uint8, device=device), torch.empty(self.world * blk, dtype=torch.uint8, device=device))
            self._bufs[key] = b
        return b

    def search_device(self, q, k):
        """q: float32 [nq, d] tensor on this rank's device (identical on every rank).
        Returns (D, I) [nq, k] tensors, identical on every rank."""
        import torch

        if self.world == 1:
            return self.local.search_device(q, k)
        nq = q.shape[0]
        if self.merge_fn is not None or not q.is_cuda:
            # generic path (CPU / gloo tests, custom merges): gather [G, nq, k] arrays, then merge_fn
            D, I = self.local.search_device(q, k)
            packed = torch.cat([I.contiguous().view(torch.int32).reshape(-1), D.contiguous().view(torch.int32).reshape(-1)])
            flat = torch.empty(self.world * packed.numel(), dtype=torch.int32, device=packed.device)
            self.dist.all_gather_into_tensor(flat, packed, group=self.group)
            g = flat.view(self.world, packed.numel())
            nI = 2 * I.numel()
            Ig = g[:, :nI].contiguous().view(torch.int64).reshape(self.world, *I.shape)
            Dg = g[:, nI:].contiguous().view(torch.float32).reshape(self.world, *D.shape)
            if self.merge_fn is not None:
                return self.merge_fn(Dg, Ig, k)
            from .index import merge_shard_results

            return merge_shard_results(Dg, Ig, k)
        from .index import merge_packed_results

        send, recv = self._buffers(nq, k, q.device)
        nI = nq * k * 8
        I = send[:nI].view(torch.int64).view(nq, k)
        D = send[nI:nI + nq * k * 4].view(torch.float32).view(nq, k)
        self.local.search_device(q, k, out=(D, I))                          # the epilogue fills the send buffer
        self.dist.all_gather_into_tensor(recv, send, group=self.group)      # the single exchange step
        self.last_gathered = (recv, send.numel(), nq, k)
        return merge_packed_results(recv, self.world, send.numel(), nq, k)  # reads the receive buffer in place

    def gathered_candidates(self):
        """(Dg, Ig) [G, nq, k] host arrays of the last multi-rank search's all-gather (for parity checks)."""
        recv, blk, nq, k = self.last_gathered
        raw = recv.cpu().numpy().reshape(self.world, blk)
        Ig = np.ascontiguousarray(raw[:, :nq * k * 8]).view(np.int64).reshape(self.world, nq, k)
        Dg = np.ascontiguousarray(raw[:, nq * k * 8:nq * k * 12]).view(np.float32).reshape(self.world, nq, k)
        return Dg, Ig

    def search(self, x, k):
        """FAISS-style host call: numpy in, numpy out."""
        import torch

        x = np.ascontiguousarray(x, dtype=np.float32)
        dev = self.device if self.device is not None else "cuda"
        D, I = self.search_device(torch.from_numpy(x).to(dev), k)
        return D.cpu().numpy(), I.cpu().numpy()


class B200ShardedIndex:
    """All shards of one box behind ONE host call (C ABI `b200_sharded_search`): `shards[g]` is a
    B200FlatIndex / B200IVFFlatIndex on its own device with `id_base` set.  `use_nccl=True` exchanges the
    candidates with one ncclAllGather group (communicators created with ncclCommInitAll); the default is the
    peer-memory exchange (search epilogues store into the root's buffer over NVLink)."""

    def __init__(self, shards, use_nccl=False):
        from ._lib import lib, check

        self._lib, self._check = lib, check
        self.shards = list(shards)
        arr = (C.c_void_p * len(self.shards))(*[s._h for s in self.shards])
        self._h = C.c_void_p()
        check(lib.b200_sharded_create(arr, len(self.shards), C.byref(self._h)), "sharded_create")
        self._comms = None
        if use_nccl:
            devs = (C.c_int * len(self.shards))(*[int(s.device) for s in self.shards])
            self._comms = (C.c_void_p * len(self.shards))()
            check(lib.b200_nccl_comm_init_all(len(self.shards), devs, self._comms), "nccl_comm_init_all")

    def __del__(self):
        try:
            if self._comms is not None:
                for c in self._comms:
                    self._lib.b200_nccl_comm_destroy(c)
                self._comms = None
            if self._h:
                self._lib.b200_sharded_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @property
    def d(self):
        return self.shards[0].d

    @property
    def ntotal(self):
        return sum(s.ntotal for s in self.shards)

    @property
    def peer_mode(self):
        return bool(self._lib.b200_sharded_peer_mode(self._h))

    def search(self, x, k):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2:
            raise ValueError("query must be 2-D [nq, d], got shape %r" % (x.shape,))
        assert x.shape[1] == self.d, "query dimension %d != index dimension %d" % (x.shape[1], self.d)
        nq = x.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        self._check(self._lib.b200_sharded_search(self._h, self._comms, x.ctypes.data, nq, int(k), D.ctypes.data, I.ctypes.data),
                    "sharded_search")
        return D, I
