"""Wall-clock breakdown of the two-call (plugin) iteration on C2: where do the microseconds go?"""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mp2p_icp_b200 as b200

w = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "C2")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = b200.Context(0, stream=stream.cuda_stream)
gmap = b200.Map(ctx, *bench.xyz(w["map"]))
nq = len(w["local"]); pose = w["pose"]
pt2pt = w["matcher"] == "pt2pt"
rec = 36 if pt2pt else 72
h_l = [torch.from_numpy(a).pin_memory() for a in bench.xyz(w["local"])]
d_l = [t.to(dev) for t in h_l]
h_pairs_t = torch.empty(nq * rec, dtype=torch.uint8).pin_memory()
h_pairs = h_pairs_t.numpy().view(b200.PAIR_PT2PT if pt2pt else b200.PAIR_PT2PL)
d_pairs = torch.empty(nq * rec, dtype=torch.uint8, device=dev)
mprm = b200.Pt2PtParams(**w["pt2pt"]) if pt2pt else b200.Pt2PlParams(**w["pt2pl"])
sprm = b200.HornParams() if pt2pt else b200.GNParams(**w["gn"])
hx, hy, hz = (t.numpy() for t in h_l)
match = gmap.match_pt2pt if pt2pt else gmap.match_pt2pl

def t(fn, n=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

def solve(pairs, **kw):
    if pt2pt: return ctx.solve_horn(pairs, prm=sprm, **kw)
    return ctx.solve_gauss_newton(None, pairs, sprm, pose, **kw)

res = {}
res["match host->host"] = t(lambda: match(hx, hy, hz, pose, mprm, out=h_pairs))
res["match host->device out"] = t(lambda: match(hx, hy, hz, pose, mprm, out=d_pairs.data_ptr(), out_on_device=True, capacity=nq))
res["match device->device out"] = t(lambda: match(d_l[0].data_ptr(), d_l[1].data_ptr(), d_l[2].data_ptr(), pose, mprm, n_local=nq, local_on_device=True, out=d_pairs.data_ptr(), out_on_device=True, capacity=nq))
res["match device->host"] = t(lambda: match(d_l[0].data_ptr(), d_l[1].data_ptr(), d_l[2].data_ptr(), pose, mprm, n_local=nq, local_on_device=True, out=h_pairs))
pairs, _ = match(hx, hy, hz, pose, mprm, out=h_pairs)
n = len(pairs)
res["solve upload"] = t(lambda: solve(pairs))
def two(reuse):
    p, _ = match(hx, hy, hz, pose, mprm, out=h_pairs)
    return solve(p, last_match=reuse)
res["match+solve reuse (spec)"] = t(lambda: two(True))
res["match+solve upload"] = t(lambda: two(False))
step = gmap.make_plugin_step(hx, hy, hz, mprm, sprm, h_pairs, reuse_device_pairs=True)
res["prebound plugin step reuse"] = t(lambda: step(pose))
def copies():
    for a, b in zip(d_l, h_l): a.copy_(b, non_blocking=True)
    torch.cuda.current_stream().synchronize()
res["H2D 3 arrays + sync (torch)"] = t(copies)
def d2h():
    h_pairs_t[: n * rec].copy_(d_pairs[: n * rec], non_blocking=True); torch.cuda.current_stream().synchronize()
res["D2H pairs + sync (torch)"] = t(d2h)
res["empty sync"] = t(lambda: torch.cuda.current_stream().synchronize())
print("pairs", n, "bytes", n * rec)
for k, v in res.items(): print(f"{k:34s} {v:8.1f} us")
