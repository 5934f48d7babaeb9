"""Launch-list target for the batch-1 serving shapes (BASELINE configs[4]): one text query through the ViT-L/14 /
ViT-H/14 text tower and one IVF / flat search at nq=1.  Run under
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/serve_launches.csv python tools/serve_shapes.py
to see which kernels the 0.3 ms IVF query and the 1.5 ms text embed are made of; without ncu it prints wall/event times."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clip_retrieval_b200 as m
from clip_retrieval_b200.index import synth_rows

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="ViT-L/14")
ap.add_argument("--rows", type=int, default=20_000_000)
ap.add_argument("--nlist", type=int, default=16384)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

arch = m.ARCHS[args.arch]
model = m.B200Clip(arch, max_batch=8).load_state_dict(m.synthetic_state_dict(arch, seed=0))
d = arch.embed_dim
cent = synth_rows(args.nlist, d, m.SynthSpec(seed=7), dtype="float32").cpu().numpy()
ivf = m.B200IVFFlatIndex(d, args.nlist, cent)
ivf.add_synthetic(args.rows, m.SynthSpec(seed=5, clustered=True, centroid_seed=7, nlist=args.nlist))
ivf.nprobe = 16
tok = torch.zeros(1, arch.context_length, dtype=torch.int64)
tok[0, 0], tok[0, 1:9], tok[0, 9] = arch.vocab_size - 2, torch.arange(1, 9), arch.vocab_size - 1
tok = tok.cuda()
q = model.embed_text_device(tok, dtype=torch.float32)
ivf.search_device(q, 40)
torch.cuda.synchronize()


def timeit(name, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%s: %.3f ms per call (events), %.3f ms wall" % (name, e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps), flush=True)


torch.cuda.nvtx.range_push("text_embed_b1")
timeit("text embed batch 1 (%s)" % args.arch, lambda: model.embed_text_device(tok, dtype=torch.float32), args.reps)
torch.cuda.nvtx.range_pop()
torch.cuda.nvtx.range_push("ivf_nq1")
timeit("ivf nq=1 nprobe=16 (%d rows, nlist %d)" % (args.rows, args.nlist), lambda: ivf.search_device(q, 40), args.reps)
torch.cuda.nvtx.range_pop()
