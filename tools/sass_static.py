"""Static SASS statistics of one kernel, no GPU needed: instructions per source line and per opcode
class, from `cuobjdump -xelf` + `nvdisasm -g` of the built library.  The lz77 kernel is issue bound, so
the static size of its loop bodies is the first thing to compare between two builds.

usage: sass_static.py <kernel substring> [source file name] [first line] [last line]
   e.g. sass_static.py lz77_kernelILi0 zs_lz77.cu 280 380      (the chain walk of the greedy instantiation)
       sass_static.py --totals                                  (every kernel: SASS instructions, registers)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "zlib-streams-ts_b200", "libzsgpu.so")


def totals():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    name = None
    for l in res.split("\n"):
        m = re.match(r"\s*Function (\S+):", l)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)\s+STACK:(\d+)\s+SHARED:(\d+)\s+LOCAL:(\d+)", l)
        if m and name:   # per-thread local memory in use is STACK (the frame: spills and local arrays); LOCAL is static .local data
            regs[name] = (int(m.group(1)), int(m.group(3)), int(m.group(2)) + int(m.group(4)))
    counts = collections.Counter()
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
        for cubin in sorted(os.listdir(tmp)):
            txt = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
            fn = None
            for l in txt.split("\n"):
                m = re.match(r"\s*\.text\.(\S+):", l)
                if m:
                    fn = m.group(1)
                elif fn and re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", l):
                    counts[fn] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.split("\n")
    print(f"{'SASS':>6} {'regs':>5} {'smem':>7} {'stack':>6}  kernel   (stack = STACK + LOCAL bytes per thread)")
    for (fn, c), pretty in zip(counts.items(), demangle):
        r = regs.get(fn, (0, 0, 0))
        print(f"{c:6d} {r[0]:5d} {r[1]:7d} {r[2]:6d}  {pretty.replace('(anonymous namespace)::', '')[:100]}")


def main():
    if sys.argv[1] == "--totals":
        return totals()
    want = sys.argv[1]
    fname = sys.argv[2] if len(sys.argv) > 2 else None
    lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 30
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
        per_line = collections.Counter()
        per_op = collections.Counter()
        total = 0
        for cubin in sorted(os.listdir(tmp)):
            txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
            fn = None
            cur = ("?", 0)
            for l in txt.split("\n"):
                m = re.match(r"\s*\.text\.(\S+):", l)
                if m:
                    fn = m.group(1)
                    continue
                if fn is None or want not in fn:
                    continue
                m = re.search(r'//## File "([^"]+)", line (\d+)', l)
                if m:
                    cur = (m.group(1).split("/")[-1], int(m.group(2)))
                    continue
                m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
                if not m:
                    continue
                total += 1
                if fname and not (cur[0] == fname and lo <= cur[1] <= hi):
                    continue
                per_line[cur] += 1
                per_op[m.group(1).split(".")[0]] += 1
    sel = sum(per_line.values())
    print(f"kernel ~ {want}: {total} SASS instructions, {sel} selected")
    src = {}
    for (f, ln), c in sorted(per_line.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if f not in src:
            p = os.path.join(ROOT, "zlib-streams-ts_b200", "csrc", f)
            src[f] = open(p).read().split("\n") if os.path.exists(p) else None
        text = src[f][ln - 1].strip()[:90] if src[f] and 0 < ln <= len(src[f]) else ""
        print(f"{c:5d}  {f}:{ln}  {text}")
    print("by opcode:", ", ".join(f"{k} {v}" for k, v in per_op.most_common(24)))


if __name__ == "__main__":
    main()
