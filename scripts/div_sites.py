#!/usr/bin/env python
"""Where do the fp64 divisions of a kernel come from? Usage: div_sites.py <inst .cu> <kernel substring> [nvcc -D flags...]"""
import collections, re, subprocess, sys
src, kern, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-ptx", src, "-o", "/tmp/div_sites.ptx"] + flags,
                      stderr=subprocess.DEVNULL)
files, cur, cnt, inkernel = {}, None, collections.Counter(), False
for l in open("/tmp/div_sites.ptx"):
    m = re.match(r'\s*\.file\s+(\d+)\s+"([^"]+)"', l)
    if m: files[m.group(1)] = m.group(2).split("/")[-1]
    if ".entry" in l: inkernel = kern in l
    m = re.match(r"\s*\.loc\s+(\d+)\s+(\d+)", l)
    if m: cur = (m.group(1), int(m.group(2)))
    if inkernel and ("div.rn.f64" in l or "rcp.rn.f64" in l): cnt[cur] += 1
for (f, ln), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:25]:
    print(files.get(f, f), ln, c)
print("total", sum(cnt.values()))
