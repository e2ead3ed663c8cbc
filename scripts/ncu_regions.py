#!/usr/bin/env python
"""Per-region instruction and stall-sample shares of one kernel of an .ncu-rep (source page): consecutive SASS instructions with the
same execution count form a region (loop level). Usage: ncu_regions.py <rep> <launch index> [min share %]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2])
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:150])
hdr = rows[1]
data = [r for r in rows[2:] if r and r[0].startswith("0x")]
data = data[: len(data) // 2] if len(data) > 1 and data[0][0] == data[len(data) // 2][0] else data  # the page lists the kernel twice
ia, isrc, ie, ismp, it = (hdr.index(x) for x in ("Address", "Source", "Instructions Executed", "# Samples", "Avg. Threads Executed"))
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[ismp]) for r in data)
print("instructions executed %.3e, samples %d" % (tot, tots))
base = int(data[0][ia], 16)
groups, cur = [], None
for r in data:
    c, a = int(r[ie]), int(r[ia], 16) - base
    if cur and abs(c - cur["c"]) <= 0.03 * max(c, cur["c"], 1):
        cur["n"] += 1; cur["inst"] += c; cur["smp"] += int(r[ismp]); cur["end"] = a
    else:
        cur = dict(start=a, end=a, c=c, n=1, inst=c, smp=int(r[ismp]), first=r[isrc].strip(), thr=r[it]); groups.append(cur)
for g in groups:
    if g["inst"] > minshare / 100 * tot or g["smp"] > minshare / 100 * tots:
        print("%05x-%05x n=%4d exec=%9.3fM inst=%6.2f%% samples=%6.2f%% thr=%5s  %s" % (g["start"], g["end"], g["n"], g["c"] / 1e6, 100 * g["inst"] / tot,
                                                                                        100 * g["smp"] / tots, g["thr"], g["first"][:48]))
