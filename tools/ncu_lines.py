"""Per-source-line instruction / stall breakdown: joins an ncu SASS source page with nvdisasm -g line info.
usage: ncu_lines.py <ncu source csv> <nvdisasm -g output> <source file> [top]"""
import csv, re, collections, sys
csv_path, sass_path, src_path = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cur = None; addr2line = {}
for l in open(sass_path):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', l)
    if m and cur: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Address"); ie = hdr.index("Instructions Executed"); ist = hdr.index("Warp Stall Sampling (All Samples)")
conv = lambda x: int(x, 16) if x.startswith('0x') else int(x)
base = conv(data[0][ia])
agg = collections.Counter(); st = collections.Counter(); tot = tots = 0
for r in data:
    a = conv(r[ia]) - base
    e = int(r[ie] or 0); s = int(r[ist] or 0)
    ln = addr2line.get(a, ("?", 0))
    agg[ln] += e; st[ln] += s; tot += e; tots += s
src = open(src_path).read().split('\n')
name = src_path.split('/')[-1]
print("total warp instructions", tot)
for (f, ln), e in agg.most_common(top):
    text = src[ln - 1].strip()[:86] if f == name and ln > 0 else f
    print(f"{100*e/tot:5.1f}% inst {100*st[(f,ln)]/max(tots,1):5.1f}% stall  {f}:{ln}  {text}")
