"""World-size-2 gloo test of the sharded search plumbing: the range split, the one
all-gather of per-shard candidates and its packing, and the merge order — with the per-shard search
and the merge stood in by the CPU oracle (the CUDA kernels are covered by the -m gpu suite)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _OracleShard:
    """Stands in for a B200FlatIndex shard: same search_device contract, CPU oracle inside."""

    def __init__(self, X, id_base):
        self.X, self.id_base, self.d = X, id_base, X.shape[1]

    def search_device(self, q, k):
        from oracle import knn_ref

        D, I = knn_ref.flat_search(self.X, q.numpy(), k, id_base=self.id_base)
        return torch.from_numpy(D), torch.from_numpy(I)


def _merge(Dg, Ig, k):
    from oracle import knn_ref

    D, I = knn_ref.merge_shards(Dg.numpy(), Ig.numpy(), k)
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import clip_retrieval_b200 as m
    from oracle import synth_ref

    n, d, k = 5003, 64, 17
    lo, hi = m.shard_range(n, world, rank)
    X = synth_ref.rows_f16(hi - lo, d, row0=lo, seed=3)
    sh = m.ShardedIndex(_OracleShard(X, lo), merge_fn=_merge, device="cpu")
    Q = synth_ref.rows_f32(5, d, seed=8)
    D, I = sh.search(Q, k)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), D=D, I=I)
    dist.destroy_process_group()


def test_sharded_search_equals_single_index(tmp_path):
    from oracle import knn_ref, synth_ref

    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    X = synth_ref.rows_f16(5003, 64, seed=3)
    Q = synth_ref.rows_f32(5, 64, seed=8)
    Do, Io = knn_ref.flat_search(X, Q, 17)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(got["I"], Io) and np.array_equal(got["D"], Do)  # every rank holds the global result
