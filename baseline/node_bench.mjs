// baseline/node_bench.mjs -- the reference itself (zlib-streams-ts) under Node, one stream per worker_thread.
//
// bench.py --impl reference runs this when `node` is on PATH (it is not in the build image, see DESIGN.md) and
// a reference checkout exists (--reference DIR, default /root/reference); it uses the reference's own bundle
// DIR/dist/zlib-streams.min.js through its public API, new CompressionStream(format, {level})
// (src/mod/streams.ts:242-251), so no TypeScript loader is needed.
//
//   node baseline/node_bench.mjs --reference /root/reference --workload configs1|configs2 [--mib N]
//        [--steps K] [--warmup W] [--threads T]
// Prints ONE JSON object: {value (input GB/s), ms_per_step, cores, node, sample_bytes, compressed_ratio}.
// Workloads (same shapes as bench.py; the generators are xorshift re-statements, not byte-identical to numpy's):
//   configs1  text-like word salad, deflate-raw level 1, 64 KiB chunks, each chunk its own stream
//   configs2  mixed corpus (text / ramp / random / runs / far repeats), zlib level 6, 256 KiB chunks
// The streams API has no deflateSetDictionary, so chunks are not primed here (the GPU arm and the C port prime
// them with the preceding 32 KiB, which costs the compressor slightly more work per chunk).
import { Worker, isMainThread, parentPort, workerData } from "node:worker_threads";
import os from "node:os";
import path from "node:path";
import { pathToFileURL } from "node:url";

function args() {
  const a = { reference: "/root/reference", workload: "configs1", steps: 3, warmup: 1, threads: os.availableParallelism?.() ?? os.cpus().length, mib: 0 };
  const v = process.argv.slice(2);
  for (let i = 0; i < v.length; i += 2) a[v[i].replace(/^--/, "")] = /^\d+$/.test(v[i + 1]) ? Number(v[i + 1]) : v[i + 1];
  return a;
}

function xorshift(seed) {
  let s = seed >>> 0 || 1;
  return () => { s ^= s << 13; s >>>= 0; s ^= s >>> 17; s ^= s << 5; s >>>= 0; return s; };
}

function textInto(buf, off, n, seed) {
  const rnd = xorshift(seed);
  const vocab = [];
  for (let i = 0; i < 4096; i++) {
    const l = 2 + (rnd() % 8);
    const w = new Uint8Array(l);
    for (let j = 0; j < l; j++) w[j] = 97 + (rnd() % 26);
    vocab.push(w);
  }
  let p = off;
  const end = off + n;
  while (p < end) {
    const w = vocab[rnd() % 4096];
    for (let j = 0; j < w.length && p < end; j++) buf[p++] = w[j];
    if (p < end) buf[p++] = rnd() % 100 < 85 ? 32 : 10;
  }
}

function makeCorpus(workload, n) {
  const buf = new Uint8Array(n);
  if (workload === "configs1") { textInto(buf, 0, n, 0xC0FFEE); return buf; }
  const tile = 4 << 20, rnd = xorshift(0xB200);
  for (let t0 = 0, t = 0; t0 < n; t0 += tile, t++) {
    const size = Math.min(tile, n - t0);
    const a = Math.floor(size * 0.4), b = Math.floor(size * 0.2), c = Math.floor(size * 0.2), d = Math.floor(size * 0.1);
    let p = t0;
    textInto(buf, p, a, 0xB200 + 7 * t + 1); p += a;
    for (let j = 0; j < b; j++) buf[p++] = j % 251;
    for (let j = 0; j < c; j++) buf[p++] = rnd() & 255;
    for (let j = 0; j < d; j++) buf[p++] = j < d / 2 ? 0 : ((j >> 9) * 37) & 255;
    const blk = new Uint8Array(4096);
    for (let j = 0; j < 4096; j++) blk[j] = rnd() & 255;
    for (let j = 0; p < t0 + size; j++) buf[p++] = (rnd() % 2048 === 0) ? rnd() & 255 : blk[j & 4095];
  }
  return buf;
}

async function compressChunk(CompressionStream, format, level, chunk) {
  const cs = new CompressionStream(format, { level });
  const w = cs.writable.getWriter();
  const r = cs.readable.getReader();
  let total = 0;
  const pump = (async () => { for (;;) { const { done, value } = await r.read(); if (done) break; total += value.length; } })();
  await w.write(chunk);
  await w.close();
  await pump;
  return total;
}

if (!isMainThread) {
  const { bundle, shared, n, chunk, format, level, index, stride } = workerData;
  const mod = await import(bundle);
  const data = new Uint8Array(shared, 0, n);
  parentPort.on("message", async () => {
    let total = 0;
    const nChunks = Math.ceil(n / chunk);
    for (let c = index; c < nChunks; c += stride) total += await compressChunk(mod.CompressionStream, format, level, data.subarray(c * chunk, Math.min(n, (c + 1) * chunk)));
    parentPort.postMessage(total);
  });
} else {
  const a = args();
  const conf = a.workload === "configs2" ? { format: "deflate", level: 6, chunk: 262144, mib: a.mib || 256 }
                                          : { format: "deflate-raw", level: 1, chunk: 65536, mib: a.mib || 256 };
  const n = conf.mib << 20;
  const shared = new SharedArrayBuffer(n);
  new Uint8Array(shared).set(makeCorpus(a.workload, n));
  const bundle = pathToFileURL(path.join(a.reference, "dist", "zlib-streams.min.js")).href;
  const T = Math.max(1, a.threads | 0);
  const workers = [];
  for (let i = 0; i < T; i++) workers.push(new Worker(new URL(import.meta.url), { workerData: { bundle, shared, n, chunk: conf.chunk, format: conf.format, level: conf.level, index: i, stride: T } }));
  const step = () => Promise.all(workers.map((w) => new Promise((res) => { w.once("message", res); w.postMessage("go"); })));
  let out = 0;
  for (let i = 0; i < a.warmup; i++) await step();
  const t0 = performance.now();
  for (let i = 0; i < a.steps; i++) out = (await step()).reduce((x, y) => x + y, 0);
  const ms = (performance.now() - t0) / a.steps;
  for (const w of workers) w.terminate();
  console.log(JSON.stringify({ value: n / (ms / 1e3) / 1e9, ms_per_step: ms, cores: T, node: process.version, sample_bytes: n,
                               compressed_ratio: out / n, workload: a.workload, format: conf.format, level: conf.level, chunk: conf.chunk }));
}
