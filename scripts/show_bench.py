#!/usr/bin/env python
"""One-screen summary of bench.py JSON lines: show_bench.py file.json [...]"""
import json, sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    c, e, r = d.get("config", {}), d.get("e2e") or {}, d.get("roofline") or {}
    print(f"{f}: {c.get('workload')} N={d.get('n_gpus')} {d.get('scaling')}  value {d.get('value'):.1f} it/s  ms/step {d.get('ms_per_step'):.4f}  launches {d.get('gpu_launches')}")
    if e.get("value"):
        print(f"   e2e {e['value']:.1f} it/s ({e['ms_per_step']:.4f} ms)  assume_unmodified {e.get('assume_unmodified_pairings', {}).get('ms_per_step')}  no_cloud_cache {e.get('local_cloud_uploaded_every_call', {}).get('ms_per_step')}  pageable {e.get('pageable', {}).get('ms_per_step')}  pcie floor {e.get('pcie_floor_ms')}")
    if r:
        print(f"   roofline {r.get('achieved'):.0f} GB/s = {r.get('frac'):.3f} of {r.get('peak')}  kernel_ms {r.get('kernel_ms'):.4f}  parts { {k: round(v, 4) for k, v in r.get('kernel_ms_parts', {}).items() if v} }")
        print(f"   counts {r.get('counts')}  cands/q {r.get('candidates_per_query'):.1f}  probes/q {r.get('probes_per_query'):.1f}  search alone: {r.get('search_kernel_alone')}")
    if d.get("cpu_baseline"):
        cb = d["cpu_baseline"]
        print(f"   cpu {cb['value']:.2f} it/s on {cb['cores']} threads ({cb['kind']}); value/cpu {d['value'] / cb['value']:.1f}x  e2e/cpu {(e.get('value') or 0) / cb['value']:.1f}x")
    for k, a in (d.get("align") or {}).items():
        if isinstance(a, dict) and "plugin_calls_host_buffers" in a:
            p, q = a["plugin_calls_host_buffers"], a["fused_device_resident"]
            print(f"   align {k}: plugin {p['wall_ms']:.2f} ms / {p['iterations']} it ({p['termination']})  fused {q['wall_ms']:.2f} ms  cpu {a.get('cpu_baseline', {}).get('wall_ms')} ms / {a.get('cpu_baseline', {}).get('iterations')} it  pose diff vs cpu {a.get('cpu_baseline', {}).get('pose_diff_vs_gpu')}")
        else:
            print("   align", k, a)
    if d.get("c5"):
        c5 = d["c5"]
        print("   c5:", {k: c5.get(k) for k in ("n_gpus", "ms_per_step", "value", "queries_per_gpu", "gpu_launches_per_step", "kernel_ms_rank0", "parity_vs_n1", "error")})
    print("   clocks", d.get("clocks"))
