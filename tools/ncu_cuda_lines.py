"""Per-source-line and per-function totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`
(the capture must have been taken with --import-source on and the library built with -lineinfo).
usage: ncu_cuda_lines.py <csv> [top] [source file as it was when the capture was taken]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file, hdr = None, None
lines = []          # (file, line, text, inst, thread_inst, samples, smem_wavefronts, smem_ideal)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed")
        isamp = hdr.index("# Samples"); iw = hdr.index("L1 Wavefronts Shared"); iwi = hdr.index("L1 Wavefronts Shared Ideal")
        continue
    if hdr and r[0].isdigit():
        f = lambda i: int(r[i]) if r[i].isdigit() else 0
        lines.append((cur_file, int(r[0]), r[1], f(ie), f(it), f(isamp), f(iw), f(iwi)))
tot = sum(l[3] for l in lines)
tots = sum(l[5] for l in lines)
print(f"total warp instructions {tot}, samples {tots}")
# per function of the main file: a function starts at a line matching a definition
main = max(collections.Counter(l[0] for l in lines).items(), key=lambda kv: kv[1])[0]
src = {l[1]: l[2] for l in lines if l[0] == main}
if len(sys.argv) > 3:   # the csv only carries the lines that have instructions; definitions come from the file
    src = {i + 1: t for i, t in enumerate(open(sys.argv[3]).read().split("\n"))}
fn_at, cur = {}, "(file scope)"
for ln in range(1, max(src) + 1):
    t = src.get(ln, "")
    m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__)[^;(]*?(\w+)\(", t)
    if m:
        cur = m.group(1)
    fn_at[ln] = cur
agg = collections.Counter(); aggt = collections.Counter(); aggs = collections.Counter(); aggw = collections.Counter(); aggwi = collections.Counter()
for f, ln, t, e, te, s, w, wi in lines:
    key = fn_at.get(ln, "?") if f == main else f
    agg[key] += e; aggt[key] += te; aggs[key] += s; aggw[key] += w; aggwi[key] += wi
print("\nper function (inlined code is attributed to the function whose source line it carries):")
print(f"{'inst %':>7} {'lanes':>6} {'stall %':>8} {'smem wavefronts (x ideal)':>26}  function")
for k, e in agg.most_common():
    if e == 0:
        continue
    print(f"{100*e/tot:7.2f} {aggt[k]/max(e,1):6.1f} {100*aggs[k]/max(tots,1):8.2f} {aggw[k]:16d} ({aggw[k]/max(aggwi[k],1):4.2f})  {k}")
print(f"\ntop {top} lines:")
for f, ln, t, e, te, s, w, wi in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100*e/tot:5.2f}% inst {te/max(e,1):5.1f} lanes {100*s/max(tots,1):5.2f}% stall  {f}:{ln}  {t.strip()[:90]}")
