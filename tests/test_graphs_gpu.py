"""Serving shapes are replayed from captured CUDA graphs (batch <= 8 forwards, nq <= 4 searches): the replay must
return exactly what the eager launches return, survive index mutations (stale graphs are rebuilt), and work on the
legacy default stream as well as on side streams."""
import os

import numpy as np
import pytest

from oracle import clip_ref, knn_ref, synth_ref

pytestmark = pytest.mark.gpu


def _model(m, eager):
    cfg = clip_ref.CONFIGS["tiny"]
    sd = clip_ref.make_state_dict(cfg, seed=0)
    arch = m.ClipArch(cfg.embed_dim, cfg.image_size, cfg.patch,
                      m.Tower(cfg.vision.width, cfg.vision.layers, cfg.vision.heads, cfg.vision.mlp),
                      m.Tower(cfg.text.width, cfg.text.layers, cfg.text.heads, cfg.text.mlp),
                      cfg.context_length, cfg.vocab_size, cfg.quick_gelu)
    old = os.environ.get("B200_GRAPHS")
    os.environ["B200_GRAPHS"] = "0" if eager else "1"
    try:
        model = m.B200Clip(arch, device=0, max_batch=16).load_state_dict(sd)
    finally:
        if old is None:
            os.environ.pop("B200_GRAPHS", None)
        else:
            os.environ["B200_GRAPHS"] = old
    return model, cfg


@pytest.mark.timeout(300)
def test_graph_replay_of_small_batch_forwards_equals_eager():
    import torch
    import clip_retrieval_b200 as m

    g_model, cfg = _model(m, eager=False)
    e_model, _ = _model(m, eager=True)
    side = torch.cuda.Stream()
    for B in (1, 3, 8, 9):                       # 9 > the graph limit: eager path on both
        for rep in range(4):                     # call 0 eager, call 1 captures, calls 2.. replay
            tk = clip_ref.synth_tokens(B, cfg, seed=10 * B + rep).cuda()
            px = clip_ref.synth_images(B, cfg, seed=10 * B + rep).cuda()
            want_t = e_model.embed_text_device(tk, dtype=torch.float32)
            want_i = e_model.embed_image_device(px)
            got_t = g_model.embed_text_device(tk, dtype=torch.float32)
            got_i = g_model.embed_image_device(px)
            assert torch.equal(got_t, want_t) and torch.equal(got_i, want_i), (B, rep)
            with torch.cuda.stream(side):
                side.wait_stream(torch.cuda.current_stream())
                got_s = g_model.embed_text_device(tk, dtype=torch.float32)
            side.synchronize()
            assert torch.equal(got_s, want_t)
    # the host entry at batch 1 (ClipMapper / compute_query shape) goes through the same graphs
    tk = clip_ref.synth_tokens(1, cfg, seed=99)
    assert np.array_equal(g_model.embed_text(tk), e_model.embed_text(tk))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("kind", ["flat", "ivf"])
def test_graph_replay_of_single_query_search_and_invalidation(kind):
    import torch
    import clip_retrieval_b200 as m

    d, n, k, nlist = 256, 20000, 40, 32
    kw = dict(seed=5, clustered=True, centroid_seed=7, nlist=nlist, cw=3, nw=1)
    X = synth_ref.rows_f16(2 * n, d, **kw)
    C = synth_ref.centroids_f32(nlist, d, 7)

    def make():
        if kind == "flat":
            return m.B200FlatIndex(d)
        idx = m.B200IVFFlatIndex(d, nlist, C)
        idx.nprobe = 4
        return idx

    os.environ["B200_GRAPHS"] = "1"
    try:
        idx = make()
    finally:
        os.environ.pop("B200_GRAPHS", None)
    idx.add(X[:n])
    Q = synth_ref.rows_f32(8, d, seed=77, clustered=True, centroid_seed=7, nlist=nlist)
    qd = torch.from_numpy(Q).cuda()
    first = {}
    for rep in range(4):
        for nq in (1, 2, 4):
            D, I, R = idx.search_device(qd[:nq].contiguous(), k, reconstruct=True)
            if rep == 0:
                first[nq] = (D.clone(), I.clone(), R.clone())            # eager result
            else:
                assert torch.equal(I, first[nq][1]) and torch.equal(D, first[nq][0]) and torch.equal(R, first[nq][2]), (rep, nq)
        # different queries through the same graph
        D1, I1 = idx.search_device(qd[rep + 1:rep + 2].contiguous(), k)
        Dh, Ih = idx.search(Q[rep + 1:rep + 2], k)                        # host entry: eager on the legacy stream
        assert np.array_equal(I1.cpu().numpy(), Ih) and np.array_equal(D1.cpu().numpy(), Dh)
    ms, launches = idx.last_scan_ms()
    assert launches >= 1 and ms > 0                                       # timing events still work in replay
    # mutation: more rows, a new id base, a new nprobe -> stale graphs must not be replayed
    idx.add(X[n:])
    idx.id_base = 5000
    if kind == "ivf":
        idx.nprobe = 7
    for rep in range(3):
        D, I = idx.search_device(qd[:1].contiguous(), k)
        if kind == "flat":
            ok, msg, _ = knn_ref.check_topk(D.cpu().numpy(), I.cpu().numpy(), knn_ref.scores_f64(X, Q[:1]), k, id_base=5000)
            assert ok, msg
        else:
            assign = knn_ref.ivf_assign(X, C.astype(np.float16))
            Do, Io, _ = knn_ref.ivf_search(X, assign, C.astype(np.float16), Q[:1], k, 7, id_base=5000)
            assert np.array_equal(I.cpu().numpy(), Io)
