import sys, os, importlib, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from conftest import make_mixed
    B = importlib.import_module("zlib-streams-ts_b200.batch")
    rng = np.random.default_rng(13)
    lens = rng.integers(0, 300, size=20000)
    lens[rng.integers(0, lens.size, 800)] = 0
    lens[rng.integers(0, lens.size, 50)] = rng.integers(300, 40000, 50)
    off = np.zeros(lens.size + 1, dtype=np.uint64); np.cumsum(lens, out=off[1:])
    data = make_mixed(int(off[-1]), 10)
    np.save("/tmp/off.npy", off)
    B.deflate_batch(data, 0, 1, 0, B.MODE_INDEPENDENT, flags=B.FLAG_PRIME, in_off=off)
    sys.exit(0)
# needs a library built with ZS_NVCC_EXTRA=-DZS_DEBUG_HOOKS
for name, env in (("g1", {"ZS_LZ_GRID": "1"}), ("full", {}), ("full2", {})):
    e = dict(os.environ); e.update(env); e["ZS_DUMP_LZ"] = f"/tmp/lz_{name}.bin"
    subprocess.run([sys.executable, __file__, "child"], env=e, check=True)
off = np.load("/tmp/off.npy").astype(np.int64)
nch = off.size - 1
def load(name):
    a = np.fromfile(f"/tmp/lz_{name}.bin", dtype=np.uint32)
    nblk = a[:nch]; rest = a[nch:]
    n_in = int(off[-1]); desc = rest[: rest.size - n_in].reshape(nch, -1, 4); sym = rest[rest.size - n_in:]
    return nblk, desc, sym
ref = load("g1")
for name in ("full", "full2"):
    cur = load(name)
    print(name, "nblk differ:", int((ref[0] != cur[0]).sum()))
    dd = np.argwhere((ref[1] != cur[1]).any(axis=2))
    dd = [(c, b) for c, b in dd if b < ref[0][c]]
    print("  desc differ (chunk, blk):", len(dd), dd[:6])
    for c, b in dd[:8]:
        print("    chunk", c, "len", int(off[c + 1] - off[c]), "ref desc", [int(np.int32(x)) for x in ref[1][c, b]], "cur desc", [int(np.int32(x)) for x in cur[1][c, b]],
              "prev chunk ref", [int(np.int32(x)) for x in ref[1][c - 1, 0]], "prev cur", [int(np.int32(x)) for x in cur[1][c - 1, 0]])
    # symbols: compare only the slots the reference's blocks use
    bad = []
    for c in range(nch):
        for b in range(int(ref[0][c])):
            s0 = int(off[c]) + int(np.int32(ref[1][c, b, 0])); ns = int(ref[1][c, b, 1])
            if ns and not np.array_equal(ref[2][s0:s0 + ns], cur[2][s0:s0 + ns]):
                w = np.nonzero(ref[2][s0:s0 + ns] != cur[2][s0:s0 + ns])[0]
                bad.append((c, c % 16, ns, int(w[0]), int(w[-1]), len(w)))
    print("  chunks with different symbols (chunk, idx in seg, nsym, first diff, last diff, count):", bad[:12])
    for c, _, ns, w0, w1, _ in bad[:3]:
        s0 = int(off[c]) + int(np.int32(ref[1][c, 0, 0]))
        print("   chunk", c, "ref syms", [hex(int(x)) for x in ref[2][s0 + w0: s0 + w0 + 6]], "cur", [hex(int(x)) for x in cur[2][s0 + w0: s0 + w0 + 6]], "slot", s0 + w0, "chunk byte range", int(off[c]), int(off[c + 1]), "next seg off0", int(off[(c // 16 + 1) * 16]))
