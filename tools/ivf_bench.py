"""IVF-Flat throughput on one GPU (BASELINE.json configs[3], the per-GPU slice): clustered synthetic
rows bucketed by their generating list, nq queries, top-k; reports QPS and the HBM roofline of the
list-scan kernel (bytes per query = nprobe * N/nlist * d * 2)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import clip_retrieval_b200 as m
from clip_retrieval_b200.index import synth_rows

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000_000)
ap.add_argument("--nlist", type=int, default=65536)
ap.add_argument("--nq", type=int, default=1000)
ap.add_argument("--k", type=int, default=40)
ap.add_argument("--d", type=int, default=768)
args = ap.parse_args()

d, nlist = args.d, args.nlist
spec = m.SynthSpec(seed=5, clustered=True, centroid_seed=7, nlist=nlist, cw=3, nw=1)
cent = synth_rows(nlist, d, m.SynthSpec(seed=7), dtype="float32").cpu().numpy()  # generating centroids
idx = m.B200IVFFlatIndex(d, nlist, cent)
t0 = time.time()
idx.add_synthetic(args.rows, spec)
torch.cuda.synchronize()
print("built %d rows in %.1f s" % (idx.ntotal, time.time() - t0), flush=True)
q = synth_rows(args.nq, d, m.SynthSpec(seed=77, clustered=True, centroid_seed=7, nlist=nlist), dtype="float32")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
out = []
for nprobe in (16, 64):
    idx.nprobe = nprobe
    for nq in (1, args.nq):
        qq = q[:nq].contiguous()
        for _ in range(2):
            idx.search_device(qq, args.k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if nq == 1 else 3
        e0.record()
        for _ in range(reps):
            D, I = idx.search_device(qq, args.k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        scan_ms, launches = idx.last_scan_ms()
        bytes_q = nprobe * (args.rows / nlist) * d * 2
        rec = {"rows": args.rows, "nlist": nlist, "nprobe": nprobe, "nq": nq, "k": args.k, "ms": ms, "qps": nq / ms * 1e3,
               "list_scan_ms": scan_ms, "list_scan_launches": launches,
               "list_scan_GBps": bytes_q * nq / (scan_ms / 1e3) / 1e9, "hbm_frac": bytes_q * nq / (scan_ms / 1e3) / 1e9 / peaks["hbm_gbs"]}
        out.append(rec)
        print(json.dumps(rec), flush=True)
