#!/usr/bin/env python
"""Per-CTA trace of the C3 k-NN kernel: MP2P_KNN_TRACE=1 python scripts/knn_trace.py  (GPU box).
Prints how long the CTAs took, which SM finished last and what it ran."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MP2P_KNN_TRACE", "1")
import bench
import mp2p_icp_b200 as b200

w = bench.make_workload("C3")
ctx = b200.Context(0)
gmap = b200.Map(ctx, *bench.xyz(w["map"]))
cloud = b200.Cloud(ctx, *bench.xyz(w["local"]))
prm = b200.Pt2PlParams(**w["pt2pl"])
for it in range(4):
    gmap.match_pt2pl(cloud, None, None, w["pose"], prm)
    t = ctx.tile_trace().astype(np.int64)
    dur = (t[:, 2] - t[:, 1]) % (1 << 32)
    t0 = t[:, 1].min()
    start, end = (t[:, 1] - t0) % (1 << 32), (t[:, 2] - t0) % (1 << 32)
    print(f"call {it}: CTAs {len(t)}  kernel span {end.max()/1e3:.1f} us  CTA duration us: mean {dur.mean()/1e3:.1f} p50 {np.percentile(dur,50)/1e3:.1f} p90 {np.percentile(dur,90)/1e3:.1f} p99 {np.percentile(dur,99)/1e3:.1f} max {dur.max()/1e3:.1f}")
    order = np.argsort(-end)[:5]
    for k in order:
        print(f"   late CTA blockIdx {k} tile {t[k,3]} SM {t[k,0]} start {start[k]/1e3:.1f} end {end[k]/1e3:.1f} dur {dur[k]/1e3:.1f}")
    sm_end = np.zeros(256)
    for smid in np.unique(t[:, 0]):
        sm_end[smid] = end[t[:, 0] == smid].max()
    se = sm_end[sm_end > 0]
    print(f"   per-SM finish us: min {se.min()/1e3:.1f} mean {se.mean()/1e3:.1f} max {se.max()/1e3:.1f}; CTAs per SM min {np.bincount(t[:,0]).min()} max {np.bincount(t[:,0]).max()}")
    long = np.argsort(-dur)[:5]
    print("   longest CTAs (tile, us, rounds, steps, inserts, probes, levels of warp 0):", [(int(t[k,3]), round(float(dur[k])/1e3,1), int(t[k,4]), int(t[k,5]), int(t[k,6]), int(t[k,7] & 0xffff), int(t[k,7] >> 16)) for k in long])
    print("   all CTAs: rounds mean %.1f max %d  steps mean %.1f max %d  inserts mean %.1f max %d" % (t[:,4].mean(), t[:,4].max(), t[:,5].mean(), t[:,5].max(), t[:,6].mean(), t[:,6].max()))
    typ = np.argsort(dur)[len(dur)//2]
    print("   median CTA:", (int(t[typ,3]), round(float(dur[typ])/1e3,1), int(t[typ,4]), int(t[typ,5]), int(t[typ,6]), int(t[typ,7] & 0xffff), int(t[typ,7] >> 16)))
