"""Pretty-print a bench.py JSON line (per-stage roofline table)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ["value", "ms_per_step", "n_gpus", "gpu_launches", "clocks", "final_loss"]})
print("e2e", d.get("e2e"))
print("cpu", d.get("cpu_baseline"))
for s in d.get("roofline_stages", []):
    print(f"{s['kernel']:24s} {s['launches']:5d} {s['ms_per_step']:8.2f} ms  share {s['share']:.3f}  "
          f"{s['achieved']:8.1f} {s['unit']:8s} frac {s['frac']:.3f}")
