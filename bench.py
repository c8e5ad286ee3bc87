#!/usr/bin/env python
"""bench.py -- compress MB/s of input bytes on N B200s, next to the reference's CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4] [--impl reference]

A "step" is one whole pass of the hot path (rzip -> stream blocks -> backend -> .lrz framing) over the
workload's input.  Workloads follow BASELINE.json `configs`; the default at N=1 is the
configuration the metric is quoted on, C2: 1 GiB of enwik-style text, LZMA level 7.  At N>1 every
rank owns one window (chunk) of the per-GPU size -- weak scaling -- and the only collective is the
final gather of the finished chunk blobs to rank 0 (NCCL).

value     = input bytes / time, inputs resident in HBM when the timed region starts
e2e       = the same through the C ABI with HOST buffers (pinned H2D of the input and D2H of the
            archive inside the timed region)
roofline  = the K1 tag-scan kernel timed alone with CUDA events on its launch stream; algorithmic
            bytes = N * (1 + 16 * 2^-initial_freq)  
cpu_baseline / --impl reference = the unmodified reference binary (oracle/_ref/lrzip-next) on the
            host cores, on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

MB = 1e6
# Warm-up steps run the whole path on a prefix of the workload (module load, arena growth, clocks); the
# timed steps always run the full workload.  A full C2 step is minutes of serial commit + LZMA work, so
# full-size warm-ups would not let the default run finish "within minutes".
WARM_BYTES = 8 << 20
WORKLOADS = {
    # name: (generator, per-GPU bytes, backend, level, description)
    "c1": ("rep", 100 << 20, "none", 7, "C1: 100 MiB repeated 1 MiB random block, rzip-only (-n)"),
    "c2": ("text", 1 << 30, "lzma", 7, "C2: 1 GiB enwik-style text, lzma level 7"),
    "c2n": ("text", 1 << 30, "none", 7, "C2 input, rzip-only (-n)"),
    "c3": ("trees", 1000 << 20, "lzma", 7, "C3 shard: 1000 MiB of duplicated source trees, lzma level 7"),
    "c4": ("randzero", 1 << 30, "zstd", 7, "C4 scaled: 512 MiB random + 512 MiB zeros, zstd, lz4 gate on"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def gen_input(kind: str, size: int, seed_shift: int = 0) -> np.ndarray:
    from lrzip_next_b200 import datagen
    if kind == "text":
        return datagen.gen_text_blocks(size, seed=7 + 100 * seed_shift)
    if kind == "rep":
        return datagen.gen_rep(size, seed=1234 + seed_shift)
    if kind == "trees":
        return datagen.gen_trees(size, seed=3 + seed_shift)
    if kind == "randzero":
        return datagen.gen_randzero(size, seed=4 + seed_shift)
    raise ValueError(kind)


def ref_flags(backend: str, level: int, threads: int, ram_units: int, window: int = 0):
    f = ["-Q", "-f", f"-L{level}", f"-p{threads}", f"-m{ram_units}"]
    if backend == "none":
        f.append("-n")
    elif backend == "zstd":
        f.append("-Z")
    if window:
        f.append(f"-w{window}")
    return f


def run_reference(data: np.ndarray, flags, repeats: int = 1):
    """Wall-clock the unmodified reference binary file -> file on tmpfs. Returns (best seconds, out size)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "lrzip-next")
    if not os.path.exists(ref):
        return None, None
    env = dict(os.environ, LRZIP="NOCONFIG")
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    src, dst = os.path.join(d, "in.bin"), os.path.join(d, "out.lrz")
    try:
        data.tofile(src)
        best, size = None, None
        for _ in range(repeats):
            t = time.perf_counter()
            subprocess.run([ref, *flags, "-o", dst, src], check=True, env=env, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
            size = os.path.getsize(dst)
        return best, size
    finally:
        for f in (src, dst):
            if os.path.exists(f):
                os.unlink(f)
        os.rmdir(d)


def cpu_sample_bytes(backend: str, cores: int) -> int:
    # ~10-30 s of reference CPU work: rzip-only runs at 10-150 MB/s on one thread, lzma/zstd level 7 at
    # roughly 0.7 MB/s per core
    if backend == "none":
        return 256 << 20
    # at least one 10 MiB stream block per host core, so that every core has a backend block to work on
    return min(1 << 30, max(64 << 20, (cores * 12) << 20))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lrzgpu", choices=["lrzgpu", "reference"])
    ap.add_argument("--workload", default=os.environ.get("LRZ_BENCH_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--size-mb", type=int, default=0, help="override the per-GPU input size (MiB)")
    ap.add_argument("--threads", type=int, default=0, help="-p given to both arms (default: host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true",
                    help="after timing (N=1): run the unmodified reference on the FULL workload with the same flags and "
                         "compare archive SHA-256 (minutes of CPU time; outside every timed region)")
    a = ap.parse_args()

    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version
    # banner at the first communicator), so everything else is sent to stderr and only emit() uses stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kind, size, backend, level, desc = WORKLOADS[a.workload]
    if a.size_mb:
        size = a.size_mb << 20
    cores = os.cpu_count() or 1
    # -p: stream blocks are limit/threads but never below 10 MiB (src/stream.c:1140-1348), and one block is
    # the unit of backend parallelism on both arms.  -p 160 gives the 10 MiB minimum for 1 GiB at -m 600
    # (105 threads after the reference's own memory-driven reduction); it is passed to both arms.
    threads = a.threads or max(cores, 160)
    ram_units = 600  # -m 600 = 60 GB, pinned for both arms
    window = 0 if world == 1 else size // (100 << 20)
    if world > 1 and size % (100 << 20):
        size = window * (100 << 20)
    config = {"workload": desc, "per_gpu_bytes": size, "backend": backend, "level": level, "threads_p": threads,
              "ram_m": ram_units, "window_w": window, "l2": "inputs larger than L2 (126 MB); no flush needed",
              "warmup_bytes": min(size, WARM_BYTES),
              "sharding": ("one rzip window per GPU, blobs gathered to rank 0; "
                           + ("windows chained through victim_round" if kind == "text" else "victim_round speculated"))
              if world > 1 else "single window"}

    if a.impl == "reference":
        if rank != 0:
            return
        sample = min(size, cpu_sample_bytes(backend, cores))
        data = gen_input(kind, sample)
        flags = ref_flags(backend, level, threads, ram_units)
        times = []
        for i in range(a.warmup + a.steps):
            dt, osz = run_reference(data, flags)
            if dt is None:
                emit({"impl": "reference", "unavailable": "oracle/_ref/lrzip-next was not built"})
                return
            if i >= a.warmup:
                times.append(dt)
        v = sample / (sum(times) / len(times)) / MB
        emit({
            "impl": "reference", "metric": "compress MB/s (input bytes)", "value": v, "unit": "MB/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "MB/s", "cores": min(threads, cores), "kind": "reference",
                             "sample": f"first {sample >> 20} MiB of the workload, lrzip-next {' '.join(flags)}"},
            "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ratio": sample / osz if osz else None})
        return

    import torch
    import torch.distributed as dist
    from lrzip_next_b200 import BACKEND_LZMA, BACKEND_NONE, BACKEND_ZSTD, Context, make_params, sizing
    from lrzip_next_b200 import multigpu

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bk = {"none": BACKEND_NONE, "lzma": BACKEND_LZMA, "zstd": BACKEND_ZSTD}[backend]
    total = size * world
    params = make_params(level=level, backend=bk, threads=threads, window=window, ramsize=ram_units * 100 * 1048576,
                         processors=cores)
    sz = sizing(params, total)
    ctx = Context(local_rank)

    # ---- synthetic input: this rank's window, resident in HBM (16-byte aligned, padded) and pinned on the host
    host = gen_input(kind, size, seed_shift=rank)
    pinned = torch.from_numpy(host).pin_memory()
    dbuf = torch.zeros(size + 8192 + 256, dtype=torch.uint8, device=dev)
    dbuf[256:256 + size].copy_(pinned)
    d_in = dbuf.data_ptr() + 256
    assert d_in % 16 == 0
    whole_md5 = None
    if world > 1:  # rank 0 needs the whole file's MD5: collect it once, outside the timed region
        parts = [torch.empty_like(dbuf[256:256 + size]) for _ in range(world)] if rank == 0 else None
        dist.gather(dbuf[256:256 + size].contiguous(), parts, dst=0)
        if rank == 0:
            h = hashlib.md5()
            for p_ in parts:
                h.update(p_.cpu().numpy().tobytes())
            whole_md5 = h.digest()
            del parts
    plans = multigpu.plan_chunks(total, sz.max_chunk, world) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last = {}

    def step_device():
        if world == 1:
            out, ol, st = ctx.compress_device_raw(d_in, size, params)
            last.update(out_len=ol, stats=st.as_dict())
            ctx.free(out)
        else:  # the shard is re-uploaded from the pinned copy by compress_chunk (the chunk ABI takes host data)
            # text consults the reference's cross-window counter all the time: chain the windows
            # through it (rzip stages in sequence, backends overlapped); other data speculates and shards freely
            fn = multigpu.compress_chained if kind == "text" else multigpu.compress_sharded
            arc, sts = fn(ctx, params, sz, {rank: host}, plans, whole_md5, dev)
            last.update(out_len=len(arc) if arc is not None else 0, stats=sts[0])

    def step_e2e():
        if world == 1:
            out, ol, st = ctx.compress_raw(pinned.data_ptr(), size, params)
            last.update(out_len=ol, stats=st.as_dict())
            if a.verify:
                last["sha256"] = hashlib.sha256(C.string_at(out, ol)).hexdigest()
            ctx.free(out)
        else:
            step_device()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = 0
        e0.record()
        for _ in range(steps):
            fn()
            launches0 += int(last["stats"]["kernel_launches"])
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), launches0

    warm = min(size, WARM_BYTES)
    for _ in range(a.warmup):
        out, ol, st = ctx.compress_device_raw(d_in, warm, params)
        ctx.free(out)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches = timed(step_device, a.steps)
    clocks = sampler.stop()
    value = total * a.steps / (ms / 1e3) / MB
    out_len = last["out_len"]
    stats = last["stats"]

    if world == 1:
        e2e_ms, _ = timed(step_e2e, a.steps)
    else:
        # at N > 1 the timed step above already IS the host-buffer path: every rank's window goes through
        # lrzgpu_compress_chunk from pinned host memory and the blobs come back to the host on rank 0
        e2e_ms = ms
    e2e_value = total * a.steps / (e2e_ms / 1e3) / MB
    e2e_out = last["out_len"]

    # ---- roofline of the dominant HBM kernel of the rzip stage: K1 over the whole window, alone
    roofline = None
    if rank == 0:
        rl = params.rzip_level or params.level
        initial_freq = [4, 4, 4, 4, 4, 4, 2, 1, 1, 1][rl]
        tiles = (size + 511) // 512
        cand = torch.empty(tiles * 512 * 2, dtype=torch.int64, device=dev)
        tcnt = torch.empty(tiles, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        mask = (1 << initial_freq) - 1
        for _ in range(3):
            ctx.k1_launch(d_in, size, mask, cand.data_ptr(), tcnt.data_ptr(), stream)
        torch.cuda.synchronize()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ctx.k1_launch(d_in, size, mask, cand.data_ptr(), tcnt.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        k1_ms = e0.elapsed_time(e1) / reps
        alg = size * (1 + 16 * 2.0 ** -initial_freq)
        peak, how = peaks()
        ach = alg / (k1_ms / 1e3) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture
            with open(tpath) as fh:
                tj = json.load(fh)
            if tj.get("initial_freq") == initial_freq:
                traffic = tj["dram_bytes"] * size / tj["n"]
                traffic_src = (f"{tj['source']}: {tj['dram_bytes']} B for a launch over {tj['n']} input bytes, "
                               f"scaled linearly to this launch")
        roofline = {"kernel": "k1_tagscan_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": how,
                    "ms_per_launch": k1_ms,
                    "algorithmic_bytes_per_launch": alg,
                    "note": "K2 commit / LZMA block encoders are serial, latency-bound stages: no roofline fraction"}
        del cand, tcnt

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sample = min(size, cpu_sample_bytes(backend, cores))
        flags = ref_flags(backend, level, threads, ram_units)
        dt, osz = run_reference(host[:sample], flags)
        if dt is not None:
            cpu_baseline = {"value": sample / dt / MB, "unit": "MB/s", "cores": min(threads, cores), "kind": "reference",
                            "sample": f"first {sample >> 20} MiB of the workload, one run of lrzip-next {' '.join(flags)} on tmpfs",
                            "ratio": sample / osz}
        else:
            cpu_baseline = {"value": None, "unit": "MB/s", "cores": 0, "kind": "reference",
                            "sample": "oracle/_ref/lrzip-next not built"}

    verified = None
    if a.verify and rank == 0 and world == 1 and "sha256" in last:
        ref = os.path.join(ROOT, "oracle", "_ref", "lrzip-next")
        d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        src, dst = os.path.join(d, "in.bin"), os.path.join(d, "out.lrz")
        try:
            host.tofile(src)
            t0 = time.perf_counter()
            subprocess.run([ref, *ref_flags(backend, level, threads, ram_units), "-o", dst, src], check=True,
                           env=dict(os.environ, LRZIP="NOCONFIG"), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            ref_s = time.perf_counter() - t0
            hsh = hashlib.sha256()
            with open(dst, "rb") as fh:
                for blk in iter(lambda: fh.read(1 << 24), b""):
                    hsh.update(blk)
            verified = {"identical_to_reference": hsh.hexdigest() == last["sha256"], "sha256": last["sha256"],
                        "reference_seconds_full_workload": ref_s, "reference_MBps_full_workload": size / ref_s / MB}
        finally:
            for f in (src, dst):
                if os.path.exists(f):
                    os.unlink(f)
            os.rmdir(d)

    if rank == 0:
        emit({
            "verified": verified,
            "metric": "compress MB/s (input bytes)", "value": value, "unit": "MB/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": total, "d2h_bytes_per_step": int(e2e_out),
                    "ms_per_step": e2e_ms / a.steps},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "ratio": total / out_len if out_len else None, "archive_bytes": int(out_len),
            "stage_ms": {k: stats[k] for k in ("ms_h2d", "ms_rzip", "ms_emit", "ms_backend", "ms_d2h", "ms_md5", "ms_total")},
            "rzip": {k: stats[k] for k in ("matches", "match_bytes", "literals", "literal_bytes", "inserts", "lookups", "chain_evictions", "sweeps")},
            "block_size": int(sz.bufsize), "blocks": int(stats["blocks"]), "blocks_stored": int(stats["blocks_stored"]),
        })
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
