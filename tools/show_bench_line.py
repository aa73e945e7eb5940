#!/usr/bin/env python
"""Prints the headline fields of a bench.py JSON line: python tools/show_bench_line.py <file>."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "gpu_launches", "roofline_likelihood", "roofline_likelihood_large", "small_configs", "parity_c4", "strong"):
    if k in d and d[k] is not None:
        v = d[k]
        if k == "parity_c4":
            v = {a: v[a] for a in ("elbo_rel", "grad_rel_max") if a in v}
        print(k, "=", json.dumps(v)[:700])
r = d.get("roofline", {})
print("roofline =", {a: r.get(a) for a in ("bound", "achieved", "peak", "frac", "kernel_ms", "traffic")})
