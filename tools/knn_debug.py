"""GPU bring-up helper for the flat scan: achieved HBM GB/s of the row-scan kernel per nq."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clip_retrieval_b200 as m
from clip_retrieval_b200.index import synth_rows

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
d = 768
idx = m.B200FlatIndex(d)
idx.reserve(n)
idx.add_synthetic(n, m.SynthSpec(seed=1234))
torch.cuda.synchronize()
print("index ready: ntotal=%d (%.1f GB)" % (idx.ntotal, n * d * 2 / 1e9))
nqs = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (1, 4, 8, 128, 256, 512, 1000)
for nq in nqs:
    q = synth_rows(nq, d, m.SynthSpec(seed=4321), dtype="float32")
    for k in (40,):
        for _ in range(2):
            idx.search_device(q, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D, I = idx.search_device(q, k)
        e1.record()
        torch.cuda.synchronize()
        scan_ms, launches = idx.last_scan_ms()
        tot = e0.elapsed_time(e1)
        print("nq=%d k=%d: total %.3f ms, scan %.3f ms in %d launches -> %.0f GB/s per pass, %.1f QPS" % (
            nq, k, tot, scan_ms, launches, n * d * 2 / (scan_ms / launches) / 1e6, nq / tot * 1e3))
