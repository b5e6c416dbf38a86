"""Reads gpurun_out/sweep_defer_<d>.txt (last line of tools/profile_step.py for each deferral threshold) and prints the environment
assignment with the fastest threshold per traversal kernel, e.g. `VKX_PT_DEFER=8 VKX_PT_DEFER_SHADOW=0`."""
import ast, os, re, sys

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
best = {"trace_primary": (1e9, 0), "trace_shadow": (1e9, 0)}
for d in (0, 1, 8, 12, 16):
    p = os.path.join(out, "sweep_defer_%d.txt" % d)
    if not os.path.exists(p):
        continue
    m = re.findall(r"\{[^{}]*\}", open(p).read())
    if not m:
        continue
    try:
        k = ast.literal_eval(m[-1])
    except Exception:
        continue
    for name in best:
        if name in k and float(k[name]) < best[name][0] * 0.99:  # a later threshold must win by more than 1 %
            best[name] = (float(k[name]), d)
print("VKX_PT_DEFER=%d VKX_PT_DEFER_SHADOW=%d" % (best["trace_primary"][1], best["trace_shadow"][1]))
