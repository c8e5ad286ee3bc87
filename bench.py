#!/usr/bin/env python
"""bench.py -- compress MB/s of input bytes on N B200s, next to the reference's CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--budget-s S] [--workload c1|c2|c3|c4] [--impl reference]

A "step" is one whole pass of the hot path (rzip -> stream blocks -> backend -> .lrz framing) over the
workload's input, called through the reference-facing C ABI with HOST buffers.  Workloads follow
BASELINE.json `configs`; the default is the configuration the metric is quoted on, C2: 1 GB
(1000 MiB) of enwik-style text per GPU, LZMA level 7.  At N>1 every rank owns one rzip window of the same
per-GPU size -- weak scaling -- the file is the concatenation of the windows (reference flag `-w 10`), and the
only collective is the final gather of the finished chunk blobs to rank 0 (NCCL).

ONE timed loop produces both headline numbers (a full C2 step is tens of seconds of serial rzip commit and
LZMA work, so the loop is also bounded by a wall-clock budget, see --budget-s):
  e2e       = input bytes / wall time of the steps; every step does the pinned H2D copy of the input and the
              D2H read of the finished archive inside the timed region
  value     = the same steps with the H2D / D2H time the library measured around its copies taken out
              (= inputs resident in HBM when the step starts)
`steps` is the number of full-size steps actually timed, `steps_requested` what --steps asked for.
roofline    = the K1 tag-scan kernel timed alone with CUDA events on its launch stream; algorithmic
              bytes = N * (1 + 16 * 2^-initial_freq)  
cpu_baseline = the unmodified reference binary (oracle/_ref/lrzip-next) on the host cores, on a bounded
              prefix of the workload (N=1, rank 0)
--impl reference = the unmodified reference binary on the FULL workload (N x per-GPU size, same flags, same
              `-w`), same budget rule; both arms print `archive_sha256`, equal values = bit-identical archives.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

T_START = time.perf_counter()
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before CUDA starts: the backend runs many streams
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

MB = 1e6
MiB = 1 << 20
# Warm-up steps run the whole path (every kernel, arena growth, clocks) on a prefix of the workload; the timed
# steps always run the full workload.  A full C2 step is tens of seconds, so full-size warm-ups would eat the budget.
WARM_BYTES = 1 << 20
REF_WARM_BYTES = 32 << 20
WORKLOADS = {
    # name: (generator, per-GPU bytes, backend, level, description)
    "c1": ("rep", 100 * MiB, "none", 7, "C1: 100 MiB repeated 1 MiB random block, rzip-only (-n)"),
    "c2": ("text", 1000 * MiB, "lzma", 7, "C2: 1 GB (1000 MiB) enwik-style text, lzma level 7"),
    "c2n": ("text", 1000 * MiB, "none", 7, "C2 input, rzip-only (-n)"),
    "c3": ("trees", 1000 * MiB, "lzma", 7, "C3 shard: 1000 MiB of duplicated source trees, lzma level 7"),
    "c4": ("randzero", 1000 * MiB, "zstd", 7, "C4 scaled: 500 MiB random + 500 MiB zeros, zstd, lz4 gate on"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy figure: K1 is timed alone)"
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
                                          "-lms", "500", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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


class RefRunner:
    """The unmodified reference binary, file -> file on tmpfs, wall clock around the whole run
    (test/speedtest.sh:87-99 of the reference times the same way)."""

    def __init__(self, need_bytes: int = 0):
        self.ref = os.path.join(ROOT, "oracle", "_ref", "lrzip-next")
        # tmpfs when it can hold the input, the warm-up file and the archive; else the default temporary directory
        shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
        if shm and need_bytes:
            import shutil
            if shutil.disk_usage(shm).free < int(2.3 * need_bytes) + (256 << 20):
                shm = None
        self.dir = tempfile.mkdtemp(dir=shm)
        self.tmpfs = shm is not None
        self.files = set()

    def available(self) -> bool:
        return os.path.exists(self.ref)

    def put(self, name: str, data: np.ndarray) -> str:
        p = os.path.join(self.dir, name)
        data.tofile(p)
        self.files.add(p)
        return p

    def run(self, src: str, flags) -> tuple[float, str]:
        dst = src + ".lrz"
        self.files.add(dst)
        t = time.perf_counter()
        subprocess.run([self.ref, *flags, "-o", dst, src], check=True, env=dict(os.environ, LRZIP="NOCONFIG"),
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return time.perf_counter() - t, dst

    def close(self):
        for f in self.files:
            if os.path.exists(f):
                os.unlink(f)
        os.rmdir(self.dir)


def sha256_file(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for blk in iter(lambda: fh.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lrzgpu", choices=["lrzgpu", "reference"])
    ap.add_argument("--workload", default=os.environ.get("LRZ_BENCH_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--size-mb", type=int, default=0, help="override the per-GPU input size (MiB, multiple of 100 at N>1)")
    ap.add_argument("--threads", type=int, default=0, help="-p given to both arms (default: max(host cores, 160))")
    ap.add_argument("--budget-s", type=float, default=float(os.environ.get("LRZ_BENCH_BUDGET_S", "540")),
                    help="wall-clock budget of the whole run: full-size steps are timed until --steps are done or "
                         "the next one would overrun it (at least one is always timed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()

    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version
    # banner at the first communicator), so everything else is sent to stderr and only emit() uses stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    def elapsed():
        return time.perf_counter() - T_START

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nwin = world if a.impl == "lrzgpu" else max(1, a.gpus)
    kind, size, backend, level, desc = WORKLOADS[a.workload]
    if a.size_mb:
        size = a.size_mb * MiB
    cores = os.cpu_count() or 1
    # -p: stream blocks are limit/threads but never below 10 MiB (src/stream.c:1140-1348), and one block is
    # the unit of backend parallelism on both arms.  -p 160 gives the 10 MiB minimum for 1 GB at -m 600
    # (the reference itself reduces it to what fits its memory budget); it is passed to both arms.
    # Both scale with the number of windows: the reference sizes blocks as file size / threads while the file is smaller
    # than a sixth of the declared RAM, and caps the threads by that RAM (105 at -m 600), so that N windows at a fixed
    # -p / -m would be cut into N-times larger blocks than one window is.  -p 160 N -m 600 N keeps the per-window
    # block structure (10 MiB blocks, 101 per window) the same at every N; both arms get the same flags.
    threads = a.threads or max(cores, 160) * nwin
    ram_units = 600 * nwin  # -m 600 = 60 GB per window, pinned for both arms
    # the same per-GPU size at every N; N > 1 = N windows of that size (-w in units of 100 MiB)
    window = 0 if nwin == 1 else size // (100 * MiB)
    if nwin > 1 and size % (100 * MiB):
        raise SystemExit("per-GPU size must be a multiple of 100 MiB at N > 1 (-w units)")
    total = size * nwin
    config = {"workload": desc, "per_gpu_bytes": size, "total_bytes": total, "backend": backend, "level": level,
              "threads_p": threads, "ram_m": ram_units, "window_w": window,
              "l2": "inputs larger than L2 (126 MB); no flush needed",
              "budget_s": a.budget_s,
              "sharding": "single window" if nwin == 1 else
              f"{nwin} rzip windows of {size // MiB} MiB, one per GPU, all-values speculation of victim_round; "
              f"blobs gathered to rank 0 (NCCL)"}

    if a.impl == "reference":
        if rank != 0:
            return
        rr = RefRunner(total)
        if not rr.available():
            emit({"impl": "reference", "unavailable": "oracle/_ref/lrzip-next was not built"})
            return
        try:
            data = np.concatenate([gen_input(kind, size, seed_shift=r) for r in range(nwin)]) if nwin > 1 \
                else gen_input(kind, size)
            flags = ref_flags(backend, level, threads, ram_units, window)
            config["warmup_bytes"] = min(total, REF_WARM_BYTES)
            wsrc = rr.put("warm.bin", data[:config["warmup_bytes"]])
            for _ in range(a.warmup):
                rr.run(wsrc, ref_flags(backend, level, threads, ram_units))
            src = rr.put("in.bin", data)
            del data
            times, dst = [], None
            while len(times) < a.steps:
                est = max(times) if times else 0.0
                if times and elapsed() + est > a.budget_s:
                    break
                dt, dst = rr.run(src, flags)
                times.append(dt)
            osz = os.path.getsize(dst)
            sha = sha256_file(dst)
        finally:
            rr.close()
        mean = sum(times) / len(times)
        v = total / mean / MB
        emit({
            "impl": "reference", "metric": "compress MB/s (input bytes)", "value": v, "unit": "MB/s", "n_gpus": a.gpus,
            "steps": len(times), "steps_requested": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * mean,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": v, "unit": "MB/s", "cores": min(threads, cores), "kind": "reference",
                             "sample": f"the full workload ({total // MiB} MiB), lrzip-next {' '.join(flags)}, file -> file on "
                                       f"{'tmpfs' if rr.tmpfs else 'the temporary directory (tmpfs too small)'}"},
            "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ratio": total / osz if osz else None, "archive_bytes": osz, "archive_sha256": sha,
            "step_seconds": times, "wall_s": elapsed()})
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
    params = make_params(level=level, backend=bk, threads=threads, window=window, ramsize=ram_units * 100 * 1048576,
                         processors=cores)
    sz = sizing(params, total)
    ctx = Context(local_rank)

    # ---- synthetic input: this rank's window in pinned host memory (the C ABI takes host buffers)
    host = gen_input(kind, size, seed_shift=rank)
    pinned = torch.from_numpy(host).pin_memory()
    whole = None
    if world > 1:
        # rank 0 plays the process that read the file: it holds all of it on the host, because the trailing
        # MD5 is over the whole file (src/rzip.c:1195-1218) and is hashed INSIDE every timed step, as at N=1
        dbuf = pinned.to(dev)
        parts = [torch.empty_like(dbuf) for _ in range(world)] if rank == 0 else None
        dist.gather(dbuf, parts, dst=0)
        if rank == 0:
            whole = np.concatenate([p_.cpu().numpy() for p_ in parts])
        del parts, dbuf
        torch.cuda.empty_cache()
    plans = multigpu.plan_chunks(total, sz.max_chunk, world) if world > 1 else None
    if world > 1:
        assert len(plans) == world and all(p.size == size for p in plans), "one window per GPU expected"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last = {}

    def step():
        """One pass through the public API with host buffers. Returns (copy_ms, kernel_launches) of this rank."""
        if world == 1:
            out, ol, st = ctx.compress_raw(pinned.data_ptr(), size, params)
            if last.get("out") is not None:
                ctx.free(last["out"])
            last.update(out=out, out_len=ol, stats=st.as_dict())
            return st.ms_h2d + st.ms_d2h, int(st.kernel_launches)
        md5_box = {}
        th = None
        if rank == 0:
            th = threading.Thread(target=lambda: md5_box.update(d=hashlib.md5(whole).digest()))
            th.start()

        def md5():
            th.join()
            return md5_box["d"]
        # text consults the reference's cross-window counter all the time: every window but the
        # first runs its rzip stage for all 16 values of it at once and picks the true one when it arrives;
        # other data passes the counter through unchanged and shards freely (redone only if it did not)
        fn = multigpu.compress_speculated if kind == "text" else multigpu.compress_sharded
        arc, sts = fn(ctx, params, sz, {rank: pinned}, plans, md5 if rank == 0 else None, dev)
        last.update(arc=arc, out_len=len(arc) if arc is not None else 0, stats=sts[0])
        return sum(s["ms_h2d"] + s["ms_d2h"] for s in sts), sum(int(s["kernel_launches"]) for s in sts)

    # warm-up: the whole single-window path on a prefix, on every rank (no collectives)
    config["warmup_bytes"] = min(size, WARM_BYTES)
    warm_params = make_params(level=level, backend=bk, threads=threads, window=0, ramsize=ram_units * 100 * 1048576,
                              processors=cores)
    for _ in range(a.warmup):
        wout, _, _ = ctx.compress_raw(pinned.data_ptr(), config["warmup_bytes"], warm_params)
        ctx.free(wout)

    # ---- roofline of the dominant HBM kernel of the rzip stage: K1 over the whole window, alone (before the
    # long loop, so that the line can be written the moment the loop ends)
    roofline = None
    if rank == 0:
        dbuf = torch.zeros(size + 8192 + 256, dtype=torch.uint8, device=dev)
        dbuf[256:256 + size].copy_(pinned)
        d_in = dbuf.data_ptr() + 256
        assert d_in % 16 == 0
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
                    "ms_per_launch": k1_ms, "algorithmic_bytes_per_launch": alg,
                    "note": "K2 commit / LZMA block encoders are serial, latency-bound stages: no roofline fraction"}
        del cand, tcnt, dbuf
        torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        rr = RefRunner()
        if rr.available():
            try:
                # ~10-30 s of reference CPU work: rzip-only runs at 10-150 MB/s on one thread, lzma level 7 at
                # roughly 1.5 MB/s per core; at least one 10 MiB stream block per core
                sample = min(size, 256 * MiB if backend == "none" else max(64 * MiB, min(cores, 24) * 10 * MiB))
                flags = ref_flags(backend, level, threads, ram_units)
                dt, dst = rr.run(rr.put("sample.bin", host[:sample]), flags)
                cpu_baseline = {"value": sample / dt / MB, "unit": "MB/s", "cores": min(threads, cores), "kind": "reference",
                                "sample": f"first {sample // MiB} MiB of the workload, one run of lrzip-next {' '.join(flags)} "
                                          f"on tmpfs (the --impl reference arm runs the full workload)",
                                "ratio": sample / os.path.getsize(dst)}
            finally:
                rr.close()
        else:
            cpu_baseline = {"value": None, "unit": "MB/s", "cores": 0, "kind": "reference",
                            "sample": "oracle/_ref/lrzip-next not built"}

    # ---- the timed loop: full-size steps until --steps are done or the budget is spent
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms, copy_ms, launches = [], 0.0, 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_loop = time.perf_counter()
    while len(step_ms) < a.steps:
        t0 = time.perf_counter()
        cm, nl = step()
        copy_ms += cm
        launches += nl
        step_ms.append(1e3 * (time.perf_counter() - t0))
        # every rank must take the same decision: the slowest rank's clock decides
        proj = torch.tensor([elapsed() + max(step_ms) / 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(proj, op=dist.ReduceOp.MAX)
        if float(proj.item()) > a.budget_s:
            break
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(e1) - copy_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    clocks = sampler.stop()
    nsteps = len(step_ms)
    e2e_ms, dev_ms = float(ms[0].item()), float(ms[1].item())
    e2e_value = total * nsteps / (e2e_ms / 1e3) / MB
    value = total * nsteps / (dev_ms / 1e3) / MB
    out_len = last["out_len"]
    stats = last["stats"]

    if rank == 0:
        if world == 1:
            import ctypes as C
            sha = hashlib.sha256(C.string_at(last["out"], out_len)).hexdigest()
        else:
            sha = hashlib.sha256(last["arc"]).hexdigest()
        emit({
            "metric": "compress MB/s (input bytes)", "value": value, "unit": "MB/s", "n_gpus": world, "steps": nsteps,
            "steps_requested": a.steps, "warmup": a.warmup, "ms_per_step": e2e_ms / nsteps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": total, "d2h_bytes_per_step": int(out_len),
                    "ms_per_step": e2e_ms / nsteps},
            "value_note": "the e2e loop with the library-measured H2D/D2H copy time removed "
                          f"({copy_ms / nsteps:.0f} ms per step on rank 0)",
            "gpu_launches": int(cnt.item()), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "ratio": total / out_len if out_len else None, "archive_bytes": int(out_len), "archive_sha256": sha,
            "step_ms": step_ms, "wall_s": elapsed(),
            "stage_ms": {k: stats[k] for k in ("ms_h2d", "ms_rzip", "ms_emit", "ms_backend", "ms_d2h", "ms_md5", "ms_total")},
            "rzip": {k: stats[k] for k in ("matches", "match_bytes", "literals", "literal_bytes", "inserts", "lookups",
                                           "chain_evictions", "sweeps")},
            "block_size": int(sz.bufsize), "blocks": int(stats["blocks"]), "blocks_stored": int(stats["blocks_stored"]),
            "zstd_parity": "unpinned" if backend == "zstd" else None,
        })
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
