#!/usr/bin/env python
"""bench.py -- headline benchmark of the deflate/inflate hot path (BASELINE.json).

Workload (config.workload): BASELINE.json configs[1] -- a 1 GiB synthetic text-like corpus,
deflate-raw level 1, 64 KiB chunks each primed with the preceding 32 KiB, on one B200 (per rank
under torchrun: weak scaling, every rank holds its own 1 GiB, contiguous chunk ranges, and the
ranks exchange their compressed sizes / checksums with one all_gather -> exclusive scan).

A step is one pass of the hot path (checksum-free raw deflate: LZ77 match finding + parse, Huffman
construction, bit-packing encode with the fused stitch) over the whole corpus.
  value  input GB/s with the corpus resident in HBM (CUDA events on the launching stream)
  e2e    the same metric through the C ABI's host-buffer call zs_deflate_batch: pinned host input,
         H2D, kernels, D2H of the compressed streams + offsets, all inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches as the bench contract asks.
`--impl reference` times the CPU reference arm: the oracle port of the reference's deflate (the
reference itself is TypeScript and there is no JavaScript engine in this image), one independent
stream per host thread, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "deflate_input_GBps"
UNIT = "GB/s"
CHUNK = 65536
LEVEL = 1


def pkg(sub):
    return importlib.import_module("zlib-streams-ts_b200." + sub)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, device_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(device_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].strip().lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def cpu_deflate_baseline(sample, threads=0, repeats=2):
    """Oracle port (oracle/deflate.c), one stream per host thread, same chunk + dictionary plan."""
    from oracle import oracle as O
    import numpy as np
    L = O.lib()
    cores = L.zo_max_threads() if threads <= 0 else threads
    arr = np.ascontiguousarray(sample)
    best = None
    total = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        total = L.zo_deflate_chunks_mt(arr.ctypes.data, arr.size, CHUNK, LEVEL, 0, 1, cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return arr.size / best / 1e9, cores, total, best


def run_reference(args):
    """CPU reference arm: the oracle port on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    corpus = pkg("corpus")
    cores = O.lib().zo_max_threads()
    sample_bytes = min(args.size_mib << 20, max(64 << 20, cores * (8 << 20)))
    sample = corpus.text_numpy(sample_bytes, 0xC0FFEE)
    times = []
    for i in range(args.warmup + args.steps):
        gbs, cores, total, dt = cpu_deflate_baseline(sample, repeats=1)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = sample_bytes / (ms / 1e3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[1]: text-like corpus, deflate-raw level 1, 64 KiB chunks + 32 KiB dictionary priming",
                   "level": LEVEL, "chunk": CHUNK, "sample_bytes": sample_bytes},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample_bytes >> 20} MiB of the corpus per step, one stream per host thread "
                                   "(oracle/deflate.c, byte-exact with C zlib 1.3; the TypeScript reference cannot run: no node)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-mib", type=int, default=1024, help="corpus size per GPU (MiB)")
    ap.add_argument("--no-extra", action="store_true", help="skip the level-6 / inflate side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, capi, corpus = pkg("batch"), pkg("capi"), pkg("corpus")
    ctx = B.default_context(local_rank)

    n = args.size_mib << 20
    data = corpus.text_torch(n, dev, seed=0xC0FFEE + 1000 * rank)
    n_chunks = B.n_chunks_for(n, CHUNK)
    flags = B.FLAG_PRIME
    torch.cuda.synchronize()

    def step(bufs):
        r = B.deflate_batch_dev(data, CHUNK, LEVEL, B.WRAP_RAW, B.MODE_INDEPENDENT, flags, ctx=ctx, reuse=bufs,
                                want_checks=False)
        if world > 1:
            # the path's exchange step (sharded.exchange_meta): per-rank compressed size / length ->
            # all_gather -> exclusive scan of the offsets on every rank
            mine = torch.stack([r.out_off[-1] * 8, torch.zeros((), dtype=torch.int64, device=dev),
                                torch.tensor(n, dtype=torch.int64, device=dev)])
            allv = torch.empty(world * 3, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allv, mine)
            bits = allv.view(world, 3)[:, 0]
            _ = torch.cumsum(bits, 0) - bits
        return r

    bufs = step(None)
    for _ in range(max(args.warmup - 1, 0)):
        step(bufs)
    torch.cuda.synchronize()
    rr = bufs.read_result()
    out_bytes = int(rr.total_out_bytes)

    # ---- timed region: device-resident input ----
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count
    ctx.profile(True)
    ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(bufs)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n / (ms_step / 1e3) / 1e9

    # ---- e2e: host buffers through zs_deflate_batch ----
    lib = capi.load()
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.copy_(data)
    cap = int(lib.zs_deflate_batch_bound(n, n_chunks, CHUNK, B.WRAP_RAW, B.MODE_INDEPENDENT))
    h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    h_off = torch.empty(n_chunks + 1, dtype=torch.int64, pin_memory=True)
    h_bits = torch.empty(n_chunks, dtype=torch.int64, pin_memory=True)
    res = capi.DeflateResult()
    import ctypes as C

    def e2e_step():
        rc = lib.zs_deflate_batch(ctx.handle, C.c_void_p(h_in.data_ptr()), n, None, n_chunks, CHUNK, LEVEL, B.WRAP_RAW,
                                  B.MODE_INDEPENDENT, flags, C.c_void_p(h_out.data_ptr()), cap,
                                  C.c_void_p(h_off.data_ptr()), C.c_void_p(h_bits.data_ptr()), None, C.byref(res))
        ctx.check(rc, "zs_deflate_batch")

    e2e_step()
    e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    wall_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    e2e_ms = max(e2e_ms, wall_ms)  # the call synchronises internally; never report less than wall time
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * n / (e2e_ms / 1e3) / 1e9
    e2e_out = int(res.total_out_bytes)
    # the e2e stream must be the same stream the device path produced
    same = bool(torch.equal(h_out[:e2e_out].to(dev), bufs.out[:out_bytes])) if e2e_out == out_bytes else False

    # ---- roofline of the dominant kernel (lz77_kernel) ----
    peak, peak_src = hbm_peak()
    lz_n, lz_ms = prof.get("lz77_kernel", (0, 0.0))
    lz_avg_ms = lz_ms / max(lz_n, 1)
    algo_bytes = n + out_bytes + 32768 * (n_chunks - 1)   # in + compressed out + 32 KiB dictionary per chunk
    achieved = algo_bytes / (lz_avg_ms / 1e3) / 1e9 if lz_avg_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            # DRAM bytes per input byte from the committed ncu --set full capture, scaled to this launch
            traffic = json.load(open(tpath))["lz77_kernel_dram_bytes_per_input_byte"] * n
        except Exception:
            traffic = None
    # Second yardstick, because the kernel is issue bound and not DRAM bound: warp instructions per second
    # against the SMs' issue rate (4 schedulers per SM, one warp instruction each per clock).  Instructions
    # per input byte come from the committed ncu capture, time and clock are this run's.
    issue = None
    try:
        ipb = json.load(open(tpath))["lz77_kernel_warp_inst_per_input_byte"]
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz")
        if lz_avg_ms > 0 and mhz:
            ach = ipb * n / (lz_avg_ms / 1e3) / 1e9
            pk = sms * 4 * mhz * 1e6 / 1e9
            issue = {"warp_inst_per_input_byte": ipb, "achieved_ginst_per_s": ach, "peak_ginst_per_s": pk,
                     "frac": ach / pk, "sm_mhz": mhz, "source": "profiles/lz77_r1e_summary.md x this run's kernel time"}
    except Exception:
        issue = None
    roofline = {"bound": "hbm", "kernel": "lz77_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms_per_launch": lz_avg_ms, "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_share_of_step": (lz_ms / max(lz_n, 1)) / ms_step if ms_step else None,
                "per_kernel_ms_per_step": {k: v[1] / args.steps for k, v in sorted(prof.items())},
                # what actually limits the kernel (ncu --set full, profiles/lz77_r1e_summary.md): not DRAM
                "limiter": "integer ALU pipe / instruction issue: SM throughput 72 %, IPC 2.86 of 4, "
                           "28 warp-instructions per input byte, DRAM throughput 0.7 %",
                "issue": issue}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "configs[1]: 1 GiB text-like corpus per GPU, deflate-raw level 1, 64 KiB chunks + 32 KiB dictionary priming",
                   "bytes_per_gpu": n, "level": LEVEL, "chunk": CHUNK, "n_chunks": n_chunks, "wrapper": "deflate-raw",
                   "cache": "input (1 GiB) larger than L2 (126 MB)", "parallelism": f"contiguous chunk ranges x{world}"},
        "compressed_ratio": out_bytes / n,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n,
                "d2h_bytes_per_step": e2e_out + 8 * (2 * n_chunks + 1) + 24, "ms_per_step": e2e_ms,
                "identical_to_device_path": same},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "clocks": clocks,
    }

    if rank == 0 and not args.no_cpu:
        from oracle import oracle as O
        O.build()
        cores = O.lib().zo_max_threads()
        sample_bytes = min(n, max(64 << 20, cores * (8 << 20)))
        sample = data[:sample_bytes].cpu().numpy()
        gbs, cores, total, dt = cpu_deflate_baseline(sample)
        line["cpu_baseline"] = {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {sample_bytes >> 20} MiB of the corpus, same chunk+dictionary plan, "
                                          f"one stream per host thread, best of 2 ({dt:.2f} s)",
                                "compressed_ratio": total / sample_bytes}
        # ratio gate on the same sample (GPU bytes for those chunks / reference bytes)
        k = sample_bytes // CHUNK
        gpu_sample = int(bufs.out_off[k].item())
        line["size_vs_reference_level1"] = gpu_sample / total if total > 0 else None

    if rank == 0 and not args.no_extra and world == 1:
        extra = {}
        # level 6 (lazy matching), zlib wrapper, 256 KiB chunks stitched into one stream (configs[2] shape)
        m = min(n, 512 << 20)
        view = data[:m]
        r6 = B.deflate_batch_dev(view, 262144, 6, B.WRAP_ZLIB, B.MODE_STITCHED, 0, ctx=ctx)
        torch.cuda.synchronize()
        e0.record()
        reps = 2
        for _ in range(reps):
            B.deflate_batch_dev(view, 262144, 6, B.WRAP_ZLIB, B.MODE_STITCHED, 0, ctx=ctx, reuse=r6)
        e1.record()
        torch.cuda.synchronize()
        extra["deflate_level6_input_GBps"] = m / (e0.elapsed_time(e1) / reps / 1e3) / 1e9
        rr6 = r6.read_result()
        extra["level6_ratio"] = rr6.total_out_bytes / m
        if not args.no_cpu:
            from oracle import oracle as O
            s = view[: 32 << 20].cpu().numpy().tobytes()
            ref6 = len(O.deflate(s, 6, 1))
            g6 = B.deflate_batch_dev(view[: 32 << 20], 262144, 6, B.WRAP_ZLIB, B.MODE_STITCHED, 0, ctx=ctx).read_result()
            extra["size_vs_reference_level6"] = g6.total_out_bytes / ref6
        # inflate of the level-1 primed chunks (each chunk = independent raw stream + its dictionary)
        off = torch.arange(0, n_chunks + 1, dtype=torch.int64, device=dev) * CHUNK
        off[-1] = n
        starts = off[:-1]
        rng = torch.stack([torch.clamp(starts - 32768, min=0), starts], 1).reshape(-1).contiguous()
        inf = B.inflate_batch_dev(bufs.out, bufs.out_off, off, -15, d_dict=data, dict_rng=rng, out_capacity=n, ctx=ctx)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            B.inflate_batch_dev(bufs.out, bufs.out_off, off, -15, d_dict=data, dict_rng=rng, ctx=ctx, reuse=inf)
        e1.record()
        torch.cuda.synchronize()
        ok = bool((inf.status == 1).all().item()) and bool(torch.equal(inf.out[:n], data))
        extra["inflate_output_GBps"] = n / (e0.elapsed_time(e1) / reps / 1e3) / 1e9
        extra["inflate_roundtrip_bit_exact"] = ok
        # the HBM-bound kernels of the path: per-chunk adler32 / crc32 of the same 1 GiB (64 KiB segments)
        for kind, name in ((0, "adler32"), (1, "crc32")):
            B.checksum_batch_dev(data, off, kind, ctx=ctx)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                B.checksum_batch_dev(data, off, kind, ctx=ctx)
            e1.record()
            torch.cuda.synchronize()
            gbps = n / (e0.elapsed_time(e1) / 3 / 1e3) / 1e9
            extra[f"{name}_GBps"] = gbps
            extra[f"{name}_frac_of_hbm_peak"] = gbps / hbm_peak()[0]
        line["extra"] = extra

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
