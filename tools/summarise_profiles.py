"""Turn the ncu reports collect_profiles.sh left in gpurun_out/ into the tracked summaries under profiles/ (no GPU needed):
profiles/<tag>_ncu_full_<kernel>.csv (selected metrics of the `--set full` capture), profiles/<tag>_launches_bench.csv,
profiles/traffic.json (DRAM bytes per launch, read by bench.py) and profiles/<tag>_sass_<kernel>.txt (cuobjdump listings)."""
import csv, glob, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
traffic = {}
issue = {}
_tp = os.path.join(ROOT, "profiles", "traffic.json")
if os.path.exists(_tp):  # kernels captured in earlier sessions keep their entries; this run's captures replace theirs
    _old = json.load(open(_tp))
    traffic = {k: v for k, v in _old.items() if not k.startswith("_")}
    issue = _old.get("_issue", {})
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_%s_k_*.ncu-rep" % tag))):
    kernel = re.search(r"prof_%s_(k_\w+)\.ncu-rep" % tag, rep).group(1)
    traffic_key = kernel[:-len("_alpha")] + "<alpha>" if kernel.endswith("_alpha") else kernel
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("skip", rep); continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = [("metric", "unit", "value"), ("kernel", "", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else kernel)]
    for i, h in enumerate(hdr):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            out.append((h, units[i], vals[i]))
    with open(os.path.join(ROOT, "profiles", "%s_ncu_full_%s.csv" % (tag, kernel)), "w", newline="") as f:
        csv.writer(f).writerows(out)
    def num(name):
        i = hdr.index(name); v = float(vals[i].replace(",", "")); u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    traffic[traffic_key] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    def plain(name):
        return float(vals[hdr.index(name)].replace(",", "")) if name in hdr else None
    issue[traffic_key] = {"issue_slots_busy_pct": plain("smsp__issue_active.avg.pct_of_peak_sustained_active"), "lanes_per_instruction": plain("smsp__thread_inst_executed_per_inst_executed.ratio"),
                          "warp_instructions": plain("smsp__inst_executed.sum"), "duration_us_under_ncu": plain("gpu__time_duration.sum"), "achieved_occupancy_pct": plain("sm__warps_active.avg.pct_of_peak_sustained_active"),
                          "dram_throughput_pct_of_peak": plain("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "capture": tag}
    print(kernel, vals[hdr.index("gpu__time_duration.sum")], units[hdr.index("gpu__time_duration.sum")], "dram bytes", traffic[traffic_key])
if "k_shade_front" in traffic and "k_shade_miss" in traffic:
    traffic["k_shade"] = traffic["k_shade_front"] + traffic["k_shade_miss"]
traffic["_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (and under _issue: issue-slot utilisation, lanes per warp instruction, warp instructions per launch), ncu --set full --clock-control none, tools/collect_profiles.sh "
                      "(DDGI kernels: cfg2, 16384 probes x 256 rays; screen-space kernels: cfg3 at 3840x2160), summaries in profiles/r01*_ncu_full_*.csv (latest capture: %s)" % tag)
traffic["_issue"] = issue
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
src = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
if os.path.exists(src):
    lines = [l for l in open(src) if l.startswith('"') ]
    open(os.path.join(ROOT, "profiles", "%s_launches_bench.csv" % tag), "w").writelines(lines)
    print("launch list:", len(lines) - 1, "rows")
# SASS listings of the hot kernels (built objects, no GPU needed)
objs = {"k_trace_primary": "ddgi.o", "k_trace_shadow": "ddgi.o", "k_blend": "ddgi.o", "k_blend_tc": "blend_tc.o", "k_bin_count": "ddgi.o", "k_bin_scatter": "ddgi.o", "k_shade_front": "ddgi_shade.o", "k_shade_miss": "ddgi_shade.o",
        "k_filter_x": "shadow.o", "k_filter_y": "shadow.o", "k_direct_light": "shadow.o", "k_final_gather": "gather.o", "k_reflect_shade": "reflection.o"}
# template instantiation that ships as the default (name suffix after the kernel name in the mangled symbol)
inst = {"k_trace_primary": "ILi0E", "k_trace_shadow": "ILi12E", "k_shade_front": "ILb0E", "k_direct_light": "ILb0E", "k_reflect_shade": "ILb0E"}
for kernel, obj in objs.items():
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "vulkanexp_b200", "csrc", "build", obj)], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", txt)
    for b in blocks[1:]:
        name = b.split("\n", 1)[0]
        if re.search(r"\d+%s%s" % (kernel, inst.get(kernel, r"(E|I|\d)")), name):
            body = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in b.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
            with open(os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (tag, kernel)), "w") as f:
                f.write("// cuobjdump -sass %s, function %s (%d instructions)\n" % (obj, name, len(body)) + "\n".join(body) + "\n")
            print("sass", kernel, len(body))
            break
