"""Summarise an ncu report of tri_tile_kernel: per-barrier-segment stall samples and the hottest
SASS instructions.  Usage: python tools/ncu_phases.py report.ncu-rep [launch_skip ...]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
skips = [int(v) for v in sys.argv[2:]] or [0, 1]
for sk in skips:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:tri_tile",
                          "--launch-skip", str(sk), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print("==", rows[0][1])
    hdr = rows[1]
    si, ws, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    data = []
    for n, r in enumerate(rows[2:]):
        try:
            data.append((int(r[ws] or 0), int(r[ie] or 0), n, r[si]))
        except Exception:
            pass
    half = len(data) // 2
    if half and data[0][3] == data[half][3]:
        data = data[:half]
    tot = sum(d[0] for d in data)
    print("total samples", tot, "instructions", len(data))
    seg = [0] + [d[2] for d in data if "BAR.SYNC" in d[3]] + [len(data)]
    for a, b in zip(seg[:-1], seg[1:]):
        ss = sum(d[0] for d in data if a <= d[2] < b)
        ex = sum(d[1] for d in data if a <= d[2] < b)
        print(f"  segment {a:4d}-{b:4d}: samples {ss:6d} ({100 * ss / max(tot, 1):5.1f}%)  warp-instr {ex}")
    for d in sorted(data, reverse=True)[:12]:
        print(f"  {d[0]:7d} {100 * d[0] / max(tot, 1):5.1f}%  exec={d[1]:9d}  #{d[2]:4d}: {d[3][:90]}")
