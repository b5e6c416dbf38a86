"""Per CUDA-C source line view of an ncu `--set full --import-source on` report: the lines that execute the most warp instructions and
collect the most stall samples (kernels built with -lineinfo).  usage: ncu_line_summary.py report.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or cur is None: continue
    try:
        if r[hdr["Address"]] != "-": continue  # SASS rows under the line
        out.append((cur.split("/")[-1], int(r[0]), r[1].strip()[:110], int(r[hdr["Instructions Executed"]] or 0), int(r[hdr["# Samples"]] or 0)))
    except (ValueError, IndexError):
        pass
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print("warp instructions %d, samples %d" % (ti, ts))
print("by instructions executed:")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print("%5.1f %% instr %5.1f %% samples  %s:%d  %s" % (o[3] * 100.0 / max(ti, 1), o[4] * 100.0 / max(ts, 1), o[0], o[1], o[2]))
