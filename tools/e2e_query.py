"""BASELINE.json configs[4] on one GPU: the clip_back query path — text tokens -> encode_text (batch 1) ->
L2-normalise -> kNN over the shard — p50/p99 latency and closed-loop QPS.  `--arch open_clip:ViT-H-14`
uses the H/14 text tower (D=1024, so the index is 1024-d); the default is ViT-L/14 (768-d index)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import clip_retrieval_b200 as m

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="ViT-L/14")
ap.add_argument("--rows", type=int, default=50_000_000)
ap.add_argument("--ivf", type=int, default=0, help="nlist (0 = flat)")
ap.add_argument("--nprobe", type=int, default=16)
ap.add_argument("--queries", type=int, default=200)
args = ap.parse_args()

arch = m.ARCHS[args.arch]
model = m.B200Clip(arch, max_batch=64).load_state_dict(m.synthetic_state_dict(arch, seed=0))
d = arch.embed_dim
if args.ivf:
    from clip_retrieval_b200.index import synth_rows
    cent = synth_rows(args.ivf, d, m.SynthSpec(seed=7), dtype="float32").cpu().numpy()
    idx = m.B200IVFFlatIndex(d, args.ivf, cent)
    idx.add_synthetic(args.rows, m.SynthSpec(seed=5, clustered=True, centroid_seed=7, nlist=args.ivf))
    idx.nprobe = args.nprobe
else:
    idx = m.B200FlatIndex(d)
    idx.reserve(args.rows)
    for r0 in range(0, args.rows, 8_000_000):
        idx.add_synthetic(min(8_000_000, args.rows - r0), m.SynthSpec(seed=1234), row0=r0)
g = torch.Generator().manual_seed(0)
toks = torch.zeros(args.queries, arch.context_length, dtype=torch.int64)
for i in range(args.queries):
    L = int(torch.randint(3, 30, (1,), generator=g))
    toks[i, 0] = arch.vocab_size - 2
    toks[i, 1:1 + L] = torch.randint(1, arch.vocab_size - 2, (L,), generator=g)
    toks[i, 1 + L] = arch.vocab_size - 1
toks = toks.cuda()
lat, lat_embed = [], []
for i in range(args.queries + 10):
    t = toks[i % args.queries: i % args.queries + 1]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    q = model.embed_text_device(t, dtype=torch.float32)       # compute_query: normalised fp32 [1, D]
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    D, I = idx.search_device(q, 40)
    res = I.cpu()                                              # the ids leave the device, as in map_to_metadata
    t2 = time.perf_counter()
    if i >= 10:
        lat.append(t2 - t0)
        lat_embed.append(t1 - t0)
lat = np.array(lat) * 1e3
le = np.array(lat_embed) * 1e3
# closed loop with batching: 64 queries per call
t0 = time.perf_counter()
n = 0
for rep in range(5):
    for s in range(0, args.queries - 63, 64):
        q = model.embed_text_device(toks[s:s + 64], dtype=torch.float32)
        D, I = idx.search_device(q, 40)
        n += 64
torch.cuda.synchronize()
qps = n / (time.perf_counter() - t0)
print(json.dumps({"arch": args.arch, "rows": args.rows, "index": "ivf%d/nprobe%d" % (args.ivf, args.nprobe) if args.ivf else "flat",
                  "p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
                  "embed_p50_ms": float(np.percentile(le, 50)), "batched64_qps": qps}))
