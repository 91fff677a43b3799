#!/usr/bin/env python3
"""bench.py -- input GB/s scanned by find_overlapping_iter on the UniDic-scale automaton.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (port)

Workload (BASELINE.json configs[2] with the headline iterator): a bytewise automaton of
675 000 UniDic-like patterns; one *step* = one pass of find_overlapping_iter over a batch of
1 Mi haystacks x 4 KiB (4 GiB of synthetic text) per GPU.  Weak scaling: every rank scans its
own seeded batch; for N > 1 the per-rank match buffers are gathered to rank 0 over NCCL inside
the timed step.  `value` is device-resident throughput (inputs already in HBM), `e2e` is the
same metric through the host-buffer C-ABI call (pinned host -> device -> scan -> host).

One JSON line is printed by rank 0.  See DESIGN.md section "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "input GB/s scanned, find_overlapping_iter, 675k-pat bytewise"
UNIT = "GB/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self, skip=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[skip:] if len(self.rows) > skip else self.rows
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[2 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(scale, rank, need_pool_bytes):
    from daachorse_b200 import synth as S

    cfg = S.config("C3", scale)
    ps = S.make_patterns(cfg)
    pool, bounds = S.make_pool(cfg, ps, need_pool_bytes, seed=2 + 1000 * rank)
    starts = S.window_starts(bounds, len(pool), cfg["n_haystacks"], cfg["hay_len"], seed=3 + 1000 * rank)
    return cfg, ps, pool, starts


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (the C port in oracle/,
    the Rust crate cannot be built in this image) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    import oracle_api as O
    from daachorse_b200 import synth as S

    cores = os.cpu_count() or 1
    cfg, ps, pool, starts = build_workload(args.scale, 0, 32 << 20)
    opma = O.OraclePma.build_packed(ps.blob, ps.offs)
    n_sample = max(cores * 64, min(len(starts), args.ref_haystacks))
    text, offs = S.materialise_host(pool, starts[:n_sample], cfg["hay_len"])
    for _ in range(args.warmup):
        opma.scan_batch(O.FIND_OVERLAPPING, text[: offs[cores * 16]], offs[: cores * 16 + 1], nthreads=cores, want_hashes=False)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        r = opma.scan_batch(O.FIND_OVERLAPPING, text, offs, nthreads=cores, want_hashes=False)
        total += r["total"]
    dt = (time.perf_counter() - t0) / args.steps
    val = text.size / dt / 1e9
    sample = "%d haystacks x %d B (%.1f MiB) of the C3 batch per step" % (n_sample, cfg["hay_len"], text.size / 2**20)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C3: bytewise 675k UniDic-like patterns, find_overlapping_iter, 4 KiB haystacks",
                   "n_patterns": len(ps), "hay_len": cfg["hay_len"], "matches_per_byte": total / args.steps / text.size},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of daachorse 4.0.0 CPU path (Rust toolchain unavailable)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the 1 Mi x 4 KiB batch (debug only)")
    ap.add_argument("--pool-mib", type=int, default=128)
    ap.add_argument("--e2e-haystacks", type=int, default=262144, help="haystacks per e2e step (host buffers)")
    ap.add_argument("--ref-haystacks", type=int, default=262144)
    ap.add_argument("--cpu-haystacks", type=int, default=262144)
    ap.add_argument("--option", action="append", default=[], help="kernel option name=value")
    ap.add_argument("--chunks", type=int, default=2, help="N > 1: chunks per step (gather of chunk k overlaps scan of k+1)")
    ap.add_argument("--reserve-sms", type=int, default=8, help="N > 1: SMs left free for the concurrent NCCL gather")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import daachorse_b200 as D
    from daachorse_b200 import synth as S

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: while the job runs, file descriptor 1 points at stderr, so that
    # anything a library writes to stdout (NCCL prints its version banner there under NCCL_DEBUG=VERSION)
    # lands in the log; the descriptor is restored just before rank 0 prints the line.
    saved_stdout = None
    try:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
    except OSError:  # no usable stderr: leave stdout alone
        saved_stdout = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t_setup = time.time()
    cfg, ps, pool, starts = build_workload(args.scale, rank, args.pool_mib << 20)
    hay_len, n = cfg["hay_len"], cfg["n_haystacks"]
    pma = D.DoubleArrayAhoCorasick.new(ps.as_list())
    for kv in args.option:
        k, v = kv.split("=")
        pma.set_option(k, int(v))
    pool_t = torch.from_numpy(pool).to(dev)
    text_t, offs_t = S.materialise_on_device(pool_t, torch.from_numpy(starts).to(dev), hay_len)
    del pool_t
    text_bytes = text_t.numel()
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    # sizing pass (not timed): learn the match counts, allocate the outputs once.  For N > 1 the batch is
    # scanned in chunks so that the NCCL gather of chunk k overlaps the scan of chunk k+1.
    n_chunks = 1 if world == 1 else args.chunks
    bounds = [n * k // n_chunks for k in range(n_chunks + 1)]
    chunk_offs = [offs_t[bounds[k]: bounds[k + 1] + 1] for k in range(n_chunks)]
    outs, out_offs_l, chunk_matches = [], [], []
    for k in range(n_chunks):
        first = pma.scan_batch_device(D.FIND_OVERLAPPING, text_t, chunk_offs[k])
        chunk_matches.append(int(first.matches.shape[0]))
        del first
    total_matches = sum(chunk_matches)
    chunk_cap = list(chunk_matches)
    if world > 1:
        if args.reserve_sms:
            pma.set_option("reserve_sms", args.reserve_sms)
        # every rank gathers equally sized buffers: the largest chunk count over the ranks
        cm = torch.tensor(chunk_matches, dtype=torch.int64, device=dev)
        dist.all_reduce(cm, op=dist.ReduceOp.MAX)
        chunk_cap = [int(x) for x in cm.tolist()]
    gbufs, gobufs = [], []
    for k in range(n_chunks):
        outs.append(torch.zeros((chunk_cap[k] + 1024, 3), dtype=torch.int32, device=dev))
        out_offs_l.append(torch.empty(bounds[k + 1] - bounds[k] + 1, dtype=torch.int64, device=dev))
        if world > 1:
            gbufs.append([torch.empty((chunk_cap[k], 3), dtype=torch.int32, device=dev) for _ in range(world)] if rank == 0 else None)
            gobufs.append([torch.empty_like(out_offs_l[k]) for _ in range(world)] if rank == 0 else None)
    out_offs = out_offs_l[0]

    pending = [None] * n_chunks

    def step():
        r = None
        for k in range(n_chunks):
            if pending[k] is not None:
                # the gather of this chunk issued one step ago must be done before its buffer is rewritten
                for w in pending[k]:
                    w.wait()
                pending[k] = None
            r = pma.scan_batch_device(D.FIND_OVERLAPPING, text_t, chunk_offs[k], out=outs[k], out_offs=out_offs_l[k])
            if world > 1:
                # the one exchange step of the path: this chunk's match buffer and offsets go to rank 0
                # over NVLink straight from the scan's output buffer (all ranks send chunk_cap[k] rows; the
                # rows past a rank's own count are ignored through its offsets).  The gather is asynchronous:
                # it overlaps the scan of the following chunks, also across the step boundary.
                pending[k] = [dist.gather(outs[k][: chunk_cap[k]], gbufs[k], dst=0, async_op=True),
                              dist.gather(out_offs_l[k], gobufs[k], dst=0, async_op=True)]
        return r

    def drain():
        for k in range(n_chunks):
            if pending[k] is not None:
                for w in pending[k]:
                    w.wait()
                pending[k] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = pma.stats()["launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    drain()
    barrier()
    n_before = len(sampler.rows)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = pma.stats()["launches"]
    scan_ms, pipe_ms = [], []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
        st = pma.stats()
        scan_ms.append(st["scan_kernel_ms"])
        pipe_ms.append(st["total_ms"])
    drain()  # every gather issued inside the timed region completes inside it
    ev1.record()
    barrier()
    time.sleep(0.25)  # let nvidia-smi flush its last samples
    clocks = sampler.stop(skip=n_before)
    clocks["note"] = "nvidia-smi -lms 50 started before the warm-up; the %d samples taken before the timed region are dropped" % n_before
    launches2 = pma.stats()["launches"]
    ms = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tb = torch.tensor([text_bytes], dtype=torch.int64, device=dev)
        dist.all_reduce(tb)
        job_bytes = int(tb.item())
    else:
        job_bytes = text_bytes
    value = job_bytes / (ms * 1e-3) / 1e9

    # ---- e2e through the host-buffer C-ABI call (pinned host memory), rank-local ------------
    e2e = None
    if not args.no_e2e:
        ne = min(n, args.e2e_haystacks)
        h_text = torch.empty(ne * hay_len, dtype=torch.uint8).pin_memory()
        h_text.copy_(text_t[: ne * hay_len])
        h_offs = (np.arange(ne + 1, dtype=np.uint64) * np.uint64(hay_len))
        h_text_np = h_text.numpy()
        cap = int(total_matches * (ne / n) * 1.25) + 4096
        # caller-owned result buffers in pinned host memory, reused by every step
        h_out_t = torch.empty(cap * 3, dtype=torch.int32).pin_memory()
        h_out = h_out_t.numpy().view(D.MATCH_DTYPE)
        h_oo_t = torch.empty(ne + 1, dtype=torch.int64).pin_memory()
        h_oo = h_oo_t.numpy().view(np.uint64)
        e2e_ms = []
        for i in range(2 + args.steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = pma.scan_batch_host(D.FIND_OVERLAPPING, h_text_np, h_offs, out=h_out, out_offs=h_oo)
            dt = time.perf_counter() - t0
            if i >= 2:
                e2e_ms.append(dt * 1e3)
        st = pma.stats()
        e2e_val = ne * hay_len / (np.mean(e2e_ms) * 1e-3) / 1e9
        if world > 1:
            t = torch.tensor([e2e_val], dtype=torch.float64, device=dev)
            dist.all_reduce(t)  # sum of per-rank host-buffer throughputs (ranks run concurrently)
            e2e_val = float(t.item())
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(st["h2d_bytes"]),
               "d2h_bytes_per_step": int(st["d2h_bytes"]), "ms_per_step": float(np.mean(e2e_ms)),
               "matches_per_step": int(len(r.matches)),
               "workload": "%d haystacks x %d B per step per GPU through dach_scan_batch_host: pinned host text -> "
                           "device -> scan -> pinned host matches, 64 MiB slices, uploads two slices ahead" % (ne, hay_len)}
        del h_text, h_out_t, h_oo_t

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_scan) ---------------------------------------------
    peak, peak_src = measured_peaks()
    k_ms = float(np.mean(scan_ms))  # the library times its last launch: the last chunk of the step
    launch_bytes = (bounds[-1] - bounds[-2]) * hay_len
    achieved = launch_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "k_scan_machine<StdMachine2<M_OVERLAPPING>, Lane2, 1024, 1> (lane machine)", "kernel_ms": k_ms,
                "pipeline_ms": float(np.mean(pipe_ms)),
                "algorithmic_bytes_per_launch": launch_bytes, "peak_source": peak_src,
                "note": "algorithmic bytes = 1 B read per haystack byte x bytes per launch (DESIGN.md); traffic "
                        "from profiles/ ncu capture when present"}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            roofline["traffic"] = json.load(open(tpath)).get("dram_bytes_per_byte_scanned")
            if roofline["traffic"] is not None:
                roofline["traffic"] = roofline["traffic"] * launch_bytes
        except Exception:
            pass

    # ---- CPU baseline: the oracle port on the host cores, bounded sample ----------------------
    cpu = None
    if not args.no_cpu and world == 1:  # the CPU baseline is reported on rank 0 at N = 1 only
        import oracle_api as O

        cores = os.cpu_count() or 1
        nc = min(n, args.cpu_haystacks)
        ctext, coffs = S.materialise_host(pool, starts[:nc], hay_len)
        opma = O.OraclePma.build_packed(ps.blob, ps.offs)
        opma.scan_batch(O.FIND_OVERLAPPING, ctext[: coffs[min(nc, 256)]], coffs[: min(nc, 256) + 1], nthreads=cores)
        t0 = time.perf_counter()
        reps = 0
        while True:  # repeat the sample until ~5 s of wall time have been spent
            ref = opma.scan_batch(O.FIND_OVERLAPPING, ctext, coffs, nthreads=cores, want_hashes=True)
            reps += 1
            if time.perf_counter() - t0 > 5.0 or reps >= 50:
                break
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": ctext.size / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d haystacks x %d B (%.0f MiB) of rank 0's batch, %d passes, all %d host threads, %.2f s per pass" % (
                   nc, hay_len, ctext.size / 2**20, reps, cores, dt),
               "note": "C restatement of daachorse 4.0.0 CPU path (Rust toolchain unavailable)"}
        # parity spot check in the same run: per-haystack counts of the sample
        oo = out_offs_l[0][: nc + 1].cpu().numpy()
        cpu["parity_counts_equal"] = bool(np.array_equal(np.diff(oo), ref["counts"].astype(np.int64)))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "C3: bytewise 675k UniDic-like patterns (seed 1), find_overlapping_iter, %d haystacks x %d B "
                               "per GPU (seeded windows of a %d MiB text pool, materialised in HBM)" % (n, hay_len, args.pool_mib),
                   "n_patterns": len(ps), "num_states": pma.num_states(), "automaton_heap_mib": pma.heap_bytes() / 2**20,
                   "image_mib": pma.stats()["image_bytes"] / 2**20, "hay_len": hay_len, "haystacks_per_gpu": n,
                   "bytes_per_gpu": text_bytes, "matches_per_step_per_gpu": total_matches,
                   "matches_per_byte": total_matches / text_bytes,
                   "l2": "inputs (%.1f GiB per GPU) are larger than L2; no flush needed" % (text_bytes / 2**30),
                   "parallelism": "haystack shards, one rank per GPU" + (
                       "; NCCL gather of match buffers to rank 0 inside the step, %d chunks pipelined, %d SMs reserved" % (
                           n_chunks, args.reserve_sms) if world > 1 else ""),
                   "setup_s": setup_s},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches2 - launches1), "launches_per_step": (launches2 - launches1) / args.steps,
        "clocks": clocks,
    }
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    if saved_stdout is not None:
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
