"""Summarises an ncu report (run where ncu is installed, no GPU needed): key raw metrics, stall reasons, dynamic opcode mix
and the hottest SASS regions.  usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
kv = {a: (b, c) for a, b, c in zip(h, u, v)}
print("# ncu summary of", rep)
print("kernel:", kv.get("Kernel Name", ("", ""))[1], " grid:", kv.get("Grid Size", ("", ""))[1], " block:", kv.get("Block Size", ("", ""))[1])
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second"]
for k in keys:
    if k in kv:
        print("%-72s %14s %s" % (k, kv[k][1], kv[k][0]))
print("\n# warp stall reasons (cycles per issued instruction)")
for a in h:
    if a.startswith("smsp__average_warps_issue_stalled") and a.endswith("per_issue_active.ratio"):
        try:
            x = float(kv[a][1])
        except ValueError:
            continue
        if x >= 0.02:
            print("%-40s %.3f" % (a.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), x))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    hdr = rows[hi[0]]
    ia, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot, total, data = collections.Counter(), 0, []
    for r in rows[hi[0] + 1:]:
        try:
            n, s = int(r[ie]), int(r[iss])
        except (ValueError, IndexError):
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ia].strip())
        if not m:
            continue
        tot[m.group(2)] += n
        total += n
        data.append((n, s))
    print("\n# dynamic opcode mix (warp instructions executed: %d)" % total)
    for op, n in tot.most_common(16):
        print("%-8s %5.1f%%" % (op, 100.0 * n / total))
    fp64 = sum(tot[o] for o in ("DFMA", "DMUL", "DADD"))
    print("fp64 arithmetic share of issued instructions: %.1f%%  (DFMA alone %.1f%%)" % (100.0 * fp64 / total, 100.0 * tot["DFMA"] / total))
    print("tensor / TMA opcodes: UTC*MMA %d, HMMA %d, UTMALDG %d, UBLKCP %d" % (sum(n for o, n in tot.items() if o.startswith("UTC")), tot["HMMA"], tot["UTMALDG"], tot["UBLKCP"]))
    segs, cur = [], None
    for i, (n, s) in enumerate(data):
        key = round(n / 1e6)
        if cur is None or abs(key - cur[0]) > max(2, 0.1 * cur[0]):
            cur = [key, i, i, 0, 0]
            segs.append(cur)
        cur[2] = i
        cur[3] += s
        cur[4] += n
    ts = sum(s for _, s in data)
    print("\n# SASS regions by execution count (lines, executions per line in millions, share of stall samples, warp instructions)")
    for s in segs:
        if s[3] > 0.01 * ts:
            print("lines %4d-%4d  exec=%4dM  samples=%5.1f%%  instr=%.2e" % (s[1], s[2], s[0], 100.0 * s[3] / ts, s[4]))
