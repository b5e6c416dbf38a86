"""Per-instruction view of an ncu `--set full --import-source on` report (no GPU needed): groups the kernel's SASS into segments of
equal execution count (loop bodies) with their share of the executed warp instructions, of the stall samples and of the
long-scoreboard (memory wait) samples, and lists the instructions that collect the most samples. This is where the kernel's time
goes, as opposed to where its instructions are.  usage: ncu_source_summary.py report.ncu-rep > profiles/<tag>_src_<kernel>.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1] if len(rows[0]) > 1 else rows[0])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print("warp instructions executed: %d, stall samples: %d, SASS instructions: %d" % (tot_inst, tot_samp, len(data)))
out = [(n, r[ix["Source"]].strip()[:64], int(r[ix["Instructions Executed"]]), float(r[ix["Avg. Threads Executed"]] or 0), int(r[ix["# Samples"]]), int(r[ix["stall_long_sb"]] or 0)) for n, r in enumerate(data)]
seg = []
for o in out:
    if seg and abs(o[2] - seg[-1][-1][2]) <= 0.02 * max(o[2], 1) + 50:
        seg[-1].append(o)
    else:
        seg.append([o])
print("\nsegments of equal execution count (>= 0.7 % of the instructions or samples):")
print(" lines        n   executions  lanes  instr %  samples %  of which memory wait %   first instruction")
for s in seg:
    ins = sum(o[2] for o in s); sm = sum(o[4] for o in s); lsb = sum(o[5] for o in s)
    if ins * 100.0 / tot_inst < 0.7 and sm * 100.0 / tot_samp < 0.7:
        continue
    thr = sum(o[2] * o[3] for o in s) / max(ins, 1)
    print("%4d..%4d  %3d  %11d  %5.1f  %6.1f  %8.1f  %12.1f              %s" % (s[0][0], s[-1][0], len(s), s[0][2], thr, ins * 100.0 / tot_inst, sm * 100.0 / tot_samp, lsb * 100.0 / max(sm, 1), s[0][1]))
print("\ninstructions with the most stall samples:")
for o in sorted(out, key=lambda o: -o[4])[:16]:
    print("line %4d  samples %5.1f %%  memory wait %5.1f %%  executions %10d  lanes %4.1f   %s" % (o[0], o[4] * 100.0 / tot_samp, o[5] * 100.0 / max(o[4], 1), o[2], o[3], o[1]))
