"""The clip_back query path on the GPU (`B200KnnService`, service.py) against the oracle chain that restates
`KnnService.query` (clip_back.py:419-470): compute_query (:207-255) -> search_and_reconstruct (:362) -> -1
truncation (:370-378) -> normalise (:379) -> post_filter (:326-341) -> id/distance lists (:388-399)."""
import numpy as np
import pytest

from oracle import clip_ref, knn_ref, postfilter_ref, synth_ref

pytestmark = pytest.mark.gpu


def _expected(sd, cfg, X, tokens, k, deduplicate):
    q = clip_ref.query_embedding(sd, cfg, tokens=tokens)
    D, I, R = knn_ref.flat_search_and_reconstruct(X, q, k)
    res = I[0]
    nb = int(np.where(res == -1)[0][0]) if (res == -1).any() else len(res)
    ids, dist, emb = res[:nb], D[0][:nb], R[0][:nb]
    l2 = np.linalg.norm(emb, axis=1)
    l2[l2 == 0] = 1
    emb = emb / l2[:, None]
    to_remove = set(postfilter_ref.get_non_uniques(emb)) if deduplicate else set()
    removed = {ids[i] for i in to_remove}
    out_i, out_d = [], []
    for ind, distance in zip(ids, dist):
        if ind not in removed:
            removed.add(ind)
            out_i.append(int(ind))
            out_d.append(float(distance))
    return out_d, out_i, q


def _setup():
    import clip_retrieval_b200 as m

    cfg = clip_ref.CONFIGS["tiny"]
    sd = clip_ref.make_state_dict(cfg, seed=0)
    arch = m.ClipArch(cfg.embed_dim, cfg.image_size, cfg.patch,
                      m.Tower(cfg.vision.width, cfg.vision.layers, cfg.vision.heads, cfg.vision.mlp),
                      m.Tower(cfg.text.width, cfg.text.layers, cfg.text.heads, cfg.text.mlp),
                      cfg.context_length, cfg.vocab_size, cfg.quick_gelu)
    model = m.B200Clip(arch, device=0, max_batch=64).load_state_dict(sd)
    return m, cfg, sd, model


@pytest.mark.timeout(300)
def test_query_matches_oracle_chain_with_dedup_and_padding():
    m, cfg, sd, model = _setup()
    d, n = cfg.embed_dim, 600
    X = synth_ref.rows_f16(n, d, seed=11)
    toks = clip_ref.synth_tokens(3, cfg, seed=4)
    # plant near-duplicates of the best rows of query 0 so that dedup has work to do
    q0 = clip_ref.query_embedding(sd, cfg, tokens=toks[:1])
    top = np.argsort(-(X.astype(np.float32) @ q0[0]))[:3]
    X[500], X[501], X[502] = X[top[0]], X[top[0]], X[top[1]]
    idx = m.B200FlatIndex(d)
    idx.add(X)
    res = m.ClipResource(model, image_index=idx, text_index=idx)
    svc = m.B200KnnService({"idx": res})
    for qi in range(3):
        for dedup in (False, True):
            for k in (40, 700):                         # 700 > ntotal: the -1 tail is cut
                got = svc.query(text_tokens=toks[qi:qi + 1], modality="image", num_images=k, num_result_ids=k, deduplicate=dedup)
                want_d, want_i, q = _expected(sd, cfg, X, toks[qi:qi + 1], k, dedup)
                # the device query embedding is bf16-tower output: compare through scores recomputed with IT
                qd = svc.compute_query_device(res, text_tokens=toks[qi:qi + 1]).cpu().numpy()
                assert (1 - clip_ref.cosine(qd, q)).max() <= 1e-3
                Dx, Ix, Rx = knn_ref.flat_search_and_reconstruct(X, qd, k)
                nb = int((Ix[0] >= 0).sum())
                emb = Rx[0][:nb] / np.maximum(np.linalg.norm(Rx[0][:nb], axis=1, keepdims=True), 1e-30)
                rm = {Ix[0][i] for i in (postfilter_ref.get_non_uniques(emb) if dedup else [])}
                exp_i = []
                for ind in Ix[0][:nb]:
                    if ind not in rm:
                        rm.add(ind)
                        exp_i.append(int(ind))
                ids = [r["id"] for r in got]
                assert ids == exp_i, (qi, dedup, k)
                assert all(isinstance(r["similarity"], float) for r in got)
                sims = np.array([r["similarity"] for r in got])
                assert np.all(np.diff(sims) <= 1e-7)
                if dedup and qi == 0:
                    # of every planted pack of exact copies only the first hit in result order survives
                    assert len({int(top[0]), 500, 501} & set(ids)) == 1 and len({int(top[1]), 502} & set(ids)) == 1
                    assert len(ids) < min(k, n)
    with pytest.raises(ValueError):
        svc.query()


@pytest.mark.timeout(300)
def test_micro_batcher_returns_per_request_results():
    """Concurrent single queries gathered into batched passes give each caller exactly its own single-query result."""
    import torch

    m, cfg, sd, model = _setup()
    d, n, k = cfg.embed_dim, 5000, 40
    idx = m.B200FlatIndex(d)
    idx.add(synth_ref.rows_f16(n, d, seed=11))
    toks = clip_ref.synth_tokens(50, cfg, seed=9)
    single = []
    for i in range(50):
        q = model.embed_text_device(toks[i:i + 1].cuda(), dtype=torch.float32)
        D, I = idx.search_device(q, k)
        single.append((D[0].cpu().numpy(), I[0].cpu().numpy()))
    mb = m.MicroBatcher(model, idx, max_batch=16, max_wait_ms=2.0, k=k)
    try:
        futs = [mb.submit(toks[i]) for i in range(50)]
        for i, f in enumerate(futs):
            D, I = f.result(timeout=60)
            assert np.array_equal(I, single[i][1])
            np.testing.assert_allclose(D, single[i][0], atol=2e-6)
    finally:
        mb.close()
    assert mb.served == 50 and mb.batches < 50          # requests really were gathered
