"""A few figures of a bench line (tools/gpu_iter.sh)."""
import json
import sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("no bench line:", e)
    sys.exit(0)
r = d["roofline"]
print(f"value {d['value']:.4g}  ms/step {d['ms_per_step']:.4f}  e2e ms {d['e2e']['ms_per_step']:.3f}  "
      f"filter {r['filter_ms']:.3f} compact {r['compact_ms']:.3f} edge {r['kernel_ms']:.3f}  frac {r['frac']:.3f}  "
      f"parity {d.get('parity', {}).get('mismatch')}  cpu {d['cpu_baseline']['value']:.3g}")
if "dbastar" in d:
    print("dbastar:", {k: d["dbastar"].get(k) for k in ("expansions_per_s", "time_search_ms", "time_gpu_ms", "identical_search", "error")})
    for leg in ("quadrotor_window", "roadmap_heuristic"):
        w = d["dbastar"].get(leg) or {}
        print(" ", leg, {k: w.get(k) for k in ("expands", "expansions_per_s", "time_search_ms", "cpu_oracle_time_search_ms", "identical_search", "error")})
