#!/usr/bin/env python
"""Pretty-print a bench.py JSON line. usage: tools/show_bench.py file.json"""
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def row(name, v):
    if "error" in v: print(f"{name:11s} ERROR {v['error']}"); return
    rf = v.get("roofline", {}); e = v.get("e2e", {}); c = v.get("cpu_baseline", {})
    print(f"{name:11s} value={v['value']:.3e} rows/s  ms/step={v['ms_per_step']:.4f}  roofline={rf.get('achieved',0):.1f} {rf.get('unit','')} "
          f"({100*rf.get('frac',0):.1f}%)  e2e={e.get('value',0):.3e}  cpu={c.get('value',0):.3e} ({c.get('cores','?')} cores)")
import signal; signal.signal(signal.SIGPIPE, signal.SIG_DFL)
row("HEAD", j)
print("  clocks:", j.get("clocks"), " launches:", j.get("gpu_launches"), " timed:", j.get("timed_region"), " numa:", j.get("numa_binding"))
print("  e2e:", {k: (f"{v:.3e}" if isinstance(v, float) else v) for k, v in j.get("e2e", {}).items() if k.endswith("value")})
for k, v in j.get("models", {}).items(): row(k, v)
