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
extra = ""
shade = (1e9, None)  # gpurun_out/sweep_shade_<blocks per SM>.txt: grid size of k_shade_front
import glob
for p in glob.glob(os.path.join(out, "sweep_shade_*.txt")):
    m = re.findall(r"\{[^{}]*\}", open(p).read())
    if not m:
        continue
    try:
        k = ast.literal_eval(m[-1])
    except Exception:
        continue
    b = int(re.search(r"sweep_shade_(\d+)\.txt", p).group(1))
    if "shade" in k and float(k["shade"]) < shade[0]:
        shade = (float(k["shade"]), b)
if shade[1] is not None:
    extra = " VKX_SHADE_BLOCKS_PER_SM=%d" % shade[1]
print("VKX_PT_DEFER=%d VKX_PT_DEFER_SHADOW=%d%s" % (best["trace_primary"][1], best["trace_shadow"][1], extra))
