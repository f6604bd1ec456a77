"""One-off probe: where does a fused C2 iteration spend its time? (device kernel vs call vs wall)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import mp2p_icp_b200 as b200
from bench import make_workload, xyz
w = make_workload("C2", shard=0, n_shards=1)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx = b200.Context(0, stream=stream.cuda_stream)
gmap = b200.Map(ctx, *xyz(w["map"]))
d_l = [torch.from_numpy(a).to(dev) for a in xyz(w["local"])]
nq = len(w["local"])
cloud = b200.Cloud(ctx, d_l[0].data_ptr(), d_l[1].data_ptr(), d_l[2].data_ptr(), n=nq, on_device=True)
d_pairs = torch.empty(nq * 36, dtype=torch.uint8, device=dev)
mprm, sprm = b200.Pt2PtParams(**w["pt2pt"]), b200.HornParams()
fused = gmap.make_iterator(cloud, None, None, nq, mprm, sprm, d_pairs.data_ptr(), nq)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pose = w["pose"]
for _ in range(5): fused(pose)
for label, do_flush in (("cold", True), ("warm", False)):
    ev, wall, kern, call = [], [], [], []
    for it in range(20):
        if do_flush: flush.zero_()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); s.record(stream); fused(pose); e.record(stream); torch.cuda.synchronize(); t1 = time.perf_counter()
        ev.append(s.elapsed_time(e)); wall.append((t1 - t0) * 1e3)
    ctx.set_profiling(True, False)
    for it in range(10):
        if do_flush: flush.zero_()
        torch.cuda.synchronize()
        fused(pose); tm = ctx.timings(); kern.append(tm["nn_search"]); call.append(tm["call_total"])
    ctx.set_profiling(False, False)
    print(label, "events %.1f us  wall %.1f us | fused kernel %.1f us  call(device) %.1f us" % (np.median(ev) * 1e3, np.median(wall) * 1e3, np.median(kern) * 1e3, np.median(call) * 1e3))
