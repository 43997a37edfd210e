"""Per-CUDA-source-line instruction and stall-sample shares of one kernel of an ncu report (needs -lineinfo and
--import-source on).  usage: python tools/ncu_hotspots.py report.ncu-rep kernel_name [top]
Runs `ncu -i report --page source --csv --print-source cuda,sass --kernel-name ...` and aggregates the SASS rows under
each source line; also prints the share of FP64-pipe opcodes (D*) and of the slow-path ones (MUFU, division helpers)."""
import collections, csv, io, os, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
fname, line_no, src = "?", "?", ""
agg = collections.defaultdict(lambda: [0, 0, ""])
ops = collections.Counter()
tot_i = tot_s = 0
hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        func = r[1]; continue
    if r[0] == "Line No":
        hdr = r; ci = r.index("Instructions Executed"); cs = r.index("# Samples"); continue
    if hdr is None:
        continue
    if r[0] != "":
        line_no, src = r[0], r[1].strip(); continue
    if len(r) <= ci or r[2] == "...":
        continue
    try:
        ni, ns = int(r[ci]), int(r[cs])
    except ValueError:
        continue
    a = agg[(fname, line_no)]
    a[0] += ni; a[1] += ns; a[2] = src
    tot_i += ni; tot_s += ns
    ops[r[3].split()[0] if not r[3].strip().startswith("@") else r[3].split()[1]] += ni
print(f"# {func}\n# warp instructions executed {tot_i}, stall samples {tot_s}; top {top} lines by executed instructions")
for (f, ln), (ni, ns, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{f:20s}:{ln:>4s} inst={100 * ni / max(tot_i, 1):5.1f}% samples={100 * ns / max(tot_s, 1):5.1f}%  {s[:120]}")
print("# opcode mix (share of executed warp instructions)")
for o, n in ops.most_common(18):
    print(f"#   {o:12s} {100 * n / max(tot_i, 1):5.1f}%")
