#!/usr/bin/env python
"""bench.py -- aligned Gbases/s of the per-window statistics path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] -- `popbam sfs` (Tajima's D, Fay & Wu's H; outgroup-polarised) on a
23 Mb X-like contig, 10 samples + outgroup, 30x per sample, 100 bp reads, 10 kb windows, synthetic data from tools/pbsynth
(SURVEY.md §8(d) data model).  The contig is fed as region shards of --shard-mb (the unit the host feeder hands over in
production: BAM-index chunks of whole windows); ONE STEP = the whole contig = every shard once.

  value      whole-job aligned Gbases/s with every shard's read batch already resident in HBM: the complete device
             pipeline (per-read prep, sample partition, pileup/call/site kernel, window compaction, window statistics,
             result copy) re-run on resident inputs, timed with CUDA events on the library's stream.
  e2e        same job through the public C ABI from PINNED HOST batches: pb_region_begin / pb_push_batch (host->device
             copy inside the timed region) / pb_region_end (device->host result copy inside the timed region).
  roofline   the pileup / call / site stage (k_pile_fast, k_hard_cells, k_fast_sites): algorithmic bytes per region / the
             stage's CUDA-event duration.  The per-base pass that feeds it (k_planes) is timed with the
             preparation, as the base-code pass of the single-kernel formulation was.
  cpu_baseline  the unmodified reference (oracle/_ref/popbam, built from /root/reference by oracle/Makefile) timed on
             one host core on a bounded sample of the same workload; falls back to the oracle port if the binary is absent.

`--impl reference` times the reference CPU implementation with one process per host core over split regions.
Under torchrun every rank runs its own 23 Mb job on its own GPU (weak scaling; no collective on the data path).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402

# DRAM bytes (read + written) of the pileup stage's kernels on a 2.3 Mb shard of this workload, from the committed ncu
# launch list profiles/r1_end_launches.csv (dram__bytes_read.sum + dram__bytes_write.sum per kernel): k_pile_fast, the
# cell-list scan, k_hard_cells, k_fast_sites
TRAFFIC_BYTES_PER_LAUNCH = 488.6e6 + 6.4e6 + 980.9e6 + 30.8e6
WIN = 10000
READ_LEN = 100
DEPTH = 30.0
N_INGROUP = 10


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--contig-mb", type=float, default=23.0)
    ap.add_argument("--shard-mb", type=float, default=2.3)
    ap.add_argument("--threads", type=int, default=0, help="generator threads (0 = auto)")
    ap.add_argument("--inflight", type=int, default=4, help="region shards processed concurrently per GPU (one host thread and one context each)")
    ap.add_argument("--distinct-shards", type=int, default=0,
                    help="generate this many distinct shards and cycle them (0 = auto: all distinct when the host has >= 8 cores per rank)")
    ap.add_argument("--cpu-sample-kb", type=int, default=150)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cli-sample-kb", type=int, default=1000, help="BAM sample for the command-line (from-BAM) tier; 0 = skip")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def shard_plan(args):
    shard_len = int(round(args.shard_mb * 1e6 / WIN)) * WIN
    n_shards = max(1, int(round(args.contig_mb * 1e6 / shard_len)))
    return shard_len, n_shards


def gen_threads(args, world):
    if args.threads > 0:
        return args.threads
    return max(2, min(32, (os.cpu_count() or 8) // max(1, world)))


def algorithmic_bytes(batch, n_cells, n_sites):
    """per read pos 4 + meta 4 + offsets 8 + 4/cigar op + packed seq + qual; per (site,sample) one 8-byte
    consensus word; one reference byte per site."""
    return int(batch.n_reads) * 16 + int(batch.n_cigar) * 4 + int(batch.n_bases) // 2 + int(batch.n_bases) + 8 * n_cells + n_sites


def run_b200(args):
    import torch
    import pbtest
    import popbam_b200
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    shard_len, n_shards = shard_plan(args)
    n = N_INGROUP + 1
    an = pbtest.AN["SFS"]
    shards = []
    t_gen = time.time()
    distinct = args.distinct_shards or (n_shards if (os.cpu_count() or 8) // world >= 8 else min(n_shards, 3))
    try:        # ~1.3 GB pinned + ~1.3 GB transient per distinct 2.3 Mb shard and rank: stay well inside the host's memory
        import psutil
        per_shard = 2.8e9 * (shard_len / 2.3e6)
        fit = int(psutil.virtual_memory().available * 0.5 / max(1, world) / per_shard)
        if not args.distinct_shards:
            distinct = max(1, min(distinct, fit))
    except Exception:
        pass
    for s in range(n_shards):
        if s >= distinct:       # cycle the generated shards (same shape, same work; each still has its own context and copy in HBM)
            shards.append(shards[s % distinct])
            continue
        fx = pbtest.Fixture(contig_len=shard_len + 1, n_ingroup=N_INGROUP, has_outgroup=1, depth=DEPTH, read_len=READ_LEN,
                            snp_density=0.01, seed=1000 * (rank + 1) + s, n_threads=gen_threads(args, world))
        b = fx.batch()
        # pinned host copies of the batch arrays (what the host feeder fills in production)
        pins, pb = {}, pbtest.Batch()
        pb.n_reads, pb.n_cigar, pb.n_bases = b.n_reads, b.n_cigar, b.n_bases
        for name, cnt, dt, ct in (("pos", b.n_reads, torch.int32, C.c_int32), ("meta", b.n_reads, torch.int32, C.c_uint32),
                                  ("cig_off", b.n_reads + 1, torch.int32, C.c_uint32), ("cigar", b.n_cigar, torch.int32, C.c_uint32),
                                  ("base_off", b.n_reads + 1, torch.int32, C.c_uint32), ("seq4", b.n_bases // 2, torch.uint8, C.c_uint8),
                                  ("qual", b.n_bases, torch.uint8, C.c_uint8)):
            t = torch.empty(int(cnt), dtype=dt, pin_memory=True)
            C.memmove(t.data_ptr(), getattr(b, name), t.numel() * t.element_size())
            pins[name] = t
            setattr(pb, name, C.cast(t.data_ptr(), C.POINTER(ct)))
        wb, we = pbtest.window_grid(0, shard_len + 1, WIN)
        p = fx.params(flags=pbtest.FLAG["OUTGROUP"], outidx=n - 1, device=local)
        shards.append(dict(batch=pb, pins=pins, ref=fx.ref(), wb=wb, we=we, params=p, aligned=fx.aligned_bases(),
                           h2d=sum(t.numel() * t.element_size() for t in pins.values())))
        fx.close()
    t_gen = time.time() - t_gen

    # one context per shard: "value" needs every shard's reads resident in HBM at the same time
    ctxs = []
    for sh in shards:
        ctx = popbam_b200.Context(sh["params"])
        ctx.set_contig(0, sh["ref"])
        ctxs.append(ctx)

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=max(1, args.inflight))

    def one_e2e(cs):
        ctx, sh = cs
        ctx.region_begin(an, sh["wb"], sh["we"])
        ctx.push_batch_async(sh["batch"])     # pinned arrays, untouched until region_end returns
        res = ctx.region_end()
        return res.n_windows * (3 * 4 + 8 + 3 * 8 * res.n_pops) + int(res.seg_off[res.n_windows]) * 17

    def e2e_step():
        # shards in flight overlap one shard's host->device copy with another shard's kernels
        return sum(pool.map(one_e2e, zip(ctxs, shards)))

    def one_resident(ctx):
        ctx.relaunch()
        ctx.wait()
        return ctx.stage_times()[1]

    def resident_step():
        # the library allows one context per host thread; with two shards in flight the host round trips inside a
        # region's pipeline (two small synchronisations) are hidden behind the other shard's kernels
        return sum(pool.map(one_resident, ctxs))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- e2e (also leaves every shard resident for the device-only leg); clocks are sampled over both timed legs
    for _ in range(max(1, args.warmup)):
        d2h = e2e_step()
    barrier()
    clk = ClockSampler(local)
    clk.__enter__()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d2h = e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)

    # ---- device-resident leg, CUDA events around the whole step on the library's stream(s)
    total_aligned = sum(int(c.res.aligned_bases) for c in ctxs)
    n_sites_total = sum(int(we[-1] - wb[0]) for wb, we in ((s["wb"], s["we"]) for s in shards))
    for _ in range(args.warmup):
        resident_step()
    launches0 = sum(c.kernel_launches() for c in ctxs)
    pile_ms = 0.0
    barrier()
    if True:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        ev1 = [torch.cuda.Event() for _ in ctxs]
        streams = [torch.cuda.ExternalStream(c.L.pb_stream(c.h)) for c in ctxs]
        cur = torch.cuda.current_stream()
        step_ms = 0.0
        for _ in range(args.steps):
            # device-side bracket: e0 precedes every library stream's work of this step, e1 follows all of it
            e0.record(cur)
            for st in streams:
                st.wait_event(e0)
            resident_step()
            for i, st in enumerate(streams):
                ev1[i].record(st)
                cur.wait_event(ev1[i])
            e1.record(cur)
            torch.cuda.synchronize()
            step_ms += e0.elapsed_time(e1)
        barrier()
        # the hot kernel alone (roofline): one more pass with the shards strictly one after the other, so that no
        # other kernel shares the SMs while k_pileup_call is timed by the library's events
        seq_ms = 0.0
        stage_ms = [0.0, 0.0, 0.0, 0.0]
        for ctx, st in zip(ctxs, streams):
            e0.record(st)
            ctx.relaunch()
            ctx.wait()
            e1.record(st)
            e1.synchronize()
            seq_ms += e0.elapsed_time(e1)
            stt = ctx.stage_times()
            pile_ms += stt[1]
            for i in range(4):
                stage_ms[i] += stt[i]
    clk.__exit__(None, None, None)
    dev_s = max_over_ranks(step_ms / 1e3)
    launches = sum(c.kernel_launches() for c in ctxs) - launches0

    peaks = {}
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = sum(algorithmic_bytes(s["batch"], int(s["we"][-1] - s["wb"][0]) * n, int(s["we"][-1] - s["wb"][0])) for s in shards)
    pile_s = pile_ms / 1e3                        # all shards' hot-kernel launches of one (sequential) step
    achieved = alg_bytes / pile_s / 1e9
    n_windows = sum(len(s["wb"]) for s in shards)

    line = {
        "metric": "aligned Gbases/s", "value": world * total_aligned / (dev_s / args.steps) / 1e9, "unit": "Gbases/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 + f64 (error model)",
        "data": "synthetic (tools/pbsynth, seeded)",
        "config": {"workload": "configs[1]: popbam sfs -p og, %.1f Mb contig as %d region shards of %.2f Mb, 10 samples + outgroup, "
                               "30x, 100 bp reads, 10 kb windows" % (n_shards * shard_len / 1e6, n_shards, shard_len / 1e6),
                   "per_gpu": "each rank runs its own contig", "windows_per_step": n_windows * world,
                   "windows_per_s": world * n_windows / (dev_s / args.steps),
                   "aligned_bases_per_step": world * total_aligned, "l2": "inputs (%.1f GB per step) exceed L2" % (alg_bytes / 1e9),
                   "distinct_shards": distinct, "shards_in_flight": max(1, args.inflight), "generator_s": round(t_gen, 1)},
        "e2e": {"value": world * total_aligned / (e2e_s / args.steps) / 1e9, "unit": "Gbases/s",
                "h2d_bytes_per_step": sum(s["h2d"] for s in shards), "d2h_bytes_per_step": int(d2h),
                "windows_per_s": world * n_windows / (e2e_s / args.steps)},
        "gpu_launches": int(launches),
        "stage_ms_per_shard": {k: v / len(shards) for k, v in zip(["prep_planes_partition", "pileup_call_site", "window_compaction", "window_stats"], stage_ms)},
        "roofline": {"kernel": "pileup/call/site stage: k_pile_fast + k_hard_cells + k_fast_sites", "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": TRAFFIC_BYTES_PER_LAUNCH if abs(shard_len - 2300000) < 1 else None,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the stage's kernels, one ncu capture on a 2.3 Mb shard (profiles/r1_end_launches.csv)",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback (B200_PROFILING.md)",
                     "algorithmic_bytes_per_launch": alg_bytes / len(shards), "launch_ms": pile_s * 1e3 / len(shards),
                     "kernel_share_of_step": pile_s / (seq_ms / 1e3),
                     "timed": "alone (shards one after the other); `value` runs %d shards in flight" % max(1, args.inflight)},
        "clocks": clk.summary(),
    }
    if rank == 0 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(None, args.cpu_sample_kb * 1000)
        if args.cli_sample_kb > 0:
            line["cli_from_bam"] = cli_from_bam(args.cli_sample_kb * 1000)
    if rank == 0:
        print(json.dumps(line))
    for c in ctxs:
        c.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def sample_fixture(length, seed=77, threads=8):
    import pbtest
    return pbtest.Fixture(contig_len=length + 1, n_ingroup=N_INGROUP, has_outgroup=1, depth=DEPTH, read_len=READ_LEN,
                          snp_density=0.01, seed=seed, n_threads=threads)


def cpu_baseline(fx_unused, sample_len):
    """One host core on a bounded sample of the same workload: the unmodified reference binary when it is there
    (kind "reference"), else the oracle port (kind "port")."""
    import pbtest
    fx = sample_fixture(sample_len)
    aligned = fx.aligned_bases()
    nwin = len(pbtest.window_grid(0, fx.contig_len, WIN)[0])
    if pbtest.have_ref():
        with tempfile.TemporaryDirectory() as td:
            bam, fa = fx.write_files(Path(td) / "s")
            t0 = time.perf_counter()
            pbtest.run_ref(["sfs", "-w", "10", "-p", "og", "-f", fa, bam, "chr1"])
            dt = time.perf_counter() - t0
        kind = "reference"
    else:
        p = fx.params(flags=pbtest.FLAG["OUTGROUP"], outidx=N_INGROUP)
        wb, we = pbtest.window_grid(0, fx.contig_len, WIN)
        t0 = time.perf_counter()
        pbtest.OracleRun(p, fx.batch(), fx.ref(), pbtest.AN["SFS"], wb, we).close()
        dt = time.perf_counter() - t0
        kind = "port"
    fx.close()
    return {"value": aligned / dt / 1e9, "unit": "Gbases/s", "cores": 1, "kind": kind, "windows_per_s": nwin / dt,
            "sample": "%d kb of the same workload (%d windows, %.3f Gbases), one process, %.1f s" % (sample_len // 1000, nwin, aligned / 1e9, dt)}


def cli_from_bam(sample_len):
    """Third timing tier ("E"): the `popbam` command line on BAM/BAI/FASTA files -- BAI slicing, BGZF
    inflate and record decode on host threads, then the GPU path -- wall time of the whole process (CUDA start-up and
    table construction included), on a bounded sample of the same workload."""
    import pbtest
    import popbam_b200
    exe = popbam_b200.capi.PKG / "_build" / "popbam"
    if not exe.exists():
        return {"unavailable": "popbam executable not built"}
    fx = sample_fixture(sample_len, seed=78, threads=min(16, os.cpu_count() or 4))
    aligned = fx.aligned_bases()
    nwin = len(pbtest.window_grid(0, fx.contig_len, WIN)[0])
    with tempfile.TemporaryDirectory() as td:
        bam, fa = fx.write_files(Path(td) / "c")
        fx.close()
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            r = subprocess.run([str(exe), "sfs", "-w", "10", "-p", "og", "--shard-mb", "0.05", "-f", fa, bam, "chr1"],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                return {"unavailable": "popbam failed: " + r.stderr.decode()[-200:]}
            best = dt if best is None else min(best, dt)
        rows = r.stdout.count(b"\n")
    return {"value": aligned / best / 1e9, "unit": "Gbases/s", "windows_per_s": nwin / best, "wall_s": best, "rows": rows,
            "host_threads": min(16, os.cpu_count() or 4),
            "sample": "%d kb BAM of the same workload (%d windows, %.3f Gbases), whole process incl. CUDA start-up" % (
                sample_len // 1000, nwin, aligned / 1e9)}


def run_reference(args):
    """Reference arm: the reference's CPU implementation with every host core, one `popbam sfs` process per core over
    split regions chr1:a+1-a+mW+1 (BASELINE.md CPU-baseline plan); a step = all regions once."""
    import pbtest
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    win_per_proc = 4
    length = procs * win_per_proc * WIN
    fx = sample_fixture(length, threads=min(32, cores))
    aligned = fx.aligned_bases()
    use_ref = pbtest.have_ref()
    with tempfile.TemporaryDirectory() as td:
        times = []
        if use_ref:
            bam, fa = fx.write_files(Path(td) / "r")
            regions = ["chr1:%d-%d" % (i * win_per_proc * WIN + 1, (i + 1) * win_per_proc * WIN + 1) for i in range(procs)]

            def step():
                ps = [subprocess.Popen([str(pbtest.REF_BIN), "sfs", "-w", "10", "-p", "og", "-f", fa, bam, r], stdout=subprocess.DEVNULL,
                                       stderr=subprocess.DEVNULL) for r in regions]
                for q in ps:
                    if q.wait() != 0:
                        raise RuntimeError("reference popbam failed")
            kind = "reference"
        else:
            p = fx.params(flags=pbtest.FLAG["OUTGROUP"], outidx=N_INGROUP)
            b, ref = fx.batch(), fx.ref()
            grids = [pbtest.window_grid(i * win_per_proc * WIN, (i + 1) * win_per_proc * WIN + 1, WIN) for i in range(procs)]

            def one(i):
                pbtest.OracleRun(p, b, ref, pbtest.AN["SFS"], grids[i][0], grids[i][1]).close()

            def step():
                th = [threading.Thread(target=one, args=(i,)) for i in range(procs)]     # ctypes releases the GIL
                [t.start() for t in th]
                [t.join() for t in th]
            kind = "port"
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
    value = aligned / dt / 1e9
    nwin = procs * win_per_proc
    sample = "%d regions of %d kb (%d windows, %.3f Gbases) of the same workload per step, one process per core" % (
        procs, win_per_proc * WIN // 1000, nwin, aligned / 1e9)
    print(json.dumps({
        "impl": "reference", "metric": "aligned Gbases/s", "value": value, "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/u64 + f64 (error model)", "data": "synthetic (tools/pbsynth, seeded)",
        "config": {"workload": "configs[1]: popbam sfs -p og, 10 samples + outgroup, 30x, 100 bp reads, 10 kb windows; bounded sample: " + sample,
                   "windows_per_s": nwin / dt},
        "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
