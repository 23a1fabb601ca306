#!/usr/bin/env python
"""Benchmark of the 2D-SFS + T2D/T1D window scan (BASELINE.json metric: SNPs/s, % of HBM roofline).

  python bench.py [--gpus N --steps K --warmup W]                 our CUDA path (libtdsfs.so through the C ABI)
  python bench.py --impl reference [--gpus N --steps K --warmup W] the reference's CPU algorithm (oracle port) on host cores

One step = one pass of the hot path over the whole workload: genotype matrix + positions -> per-SNP counts/keys and
genome-wide background spectra -> (all-reduce of the background across GPUs) -> window boundaries -> per-window
T2D / T1D(pop1) / T1D(pop2).  `value` times it with the inputs resident in HBM; `e2e` times the same pass through the
C-ABI call with pinned HOST buffers (host->device copy of the genotype matrix + positions and device->host read of
the window results inside the timed region).  Workload: BASELINE.json configs[4] (50 M SNPs x 2 pops x 500 diploids,
20 kb windows) -- the configuration the metric's target is quoted on; it fits one B200 (12.8 GB), and at N > 1 it is
sharded by contiguous chromosome ranges (total work fixed: "scaling": "strong").
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "2dsfs-scan_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

WORKLOADS = {
    # BASELINE.json configs[4] / SURVEY.md 8(d) row 5
    "config5": dict(S=50_000_000, n1=500, n2=500, C=32, W=20000, seed=20241005, mean_gap=50,
                    name="synthetic 50M SNPs x 2 pops x 500 diploids, 20 kb windows, genome-wide background"),
    # BASELINE.json configs[3] / SURVEY.md 8(d) row 4
    "config4": dict(S=10_000_000, n1=200, n2=200, C=32, W=20000, seed=20241004, mean_gap=50,
                    name="synthetic 10M SNPs x 2 pops x 200 diploids, 20 kb windows, genome-wide background"),
    "tiny": dict(S=400_000, n1=200, n2=200, C=8, W=20000, seed=7, mean_gap=50, name="tiny smoke workload"),
}


def words_for(n):
    """uint32 words per SNP of one population: a (lo plane, hi plane) pair per 32 samples (tdsfs_pack.words_for)"""
    return 2 * max(1, (n + 31) // 32)


def chrom_sizes(S, C):
    base = S // C
    return [base + (1 if i < S - base * C else 0) for i in range(C)]


def positions_for(cfg, chroms):
    """cumsum of Geometric(1/mean_gap) gaps from 1, seeded per chromosome (identical for any sharding)."""
    sizes = chrom_sizes(cfg["S"], cfg["C"])
    out = []
    for c in chroms:
        rng = np.random.default_rng(cfg["seed"] * 1000 + c)
        out.append(np.cumsum(rng.geometric(1.0 / cfg["mean_gap"], size=sizes[c]), dtype=np.int64).astype(np.int32))
    return out


def shard_chroms(C, world, rank):
    """contiguous chromosome ranges, balanced by SNP count (equal-size chromosomes here)"""
    lo, hi = C * rank // world, C * (rank + 1) // world
    return list(range(lo, hi))


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self._stop, self.th = index, [], threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:  # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                pw = nv.nvmlDeviceGetPowerUsage(self.dev) / 1000.0
                self.samples.append((time.time(), sm, rs, pw))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv:
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()

    def stop(self):
        self._stop.set()
        if self.th:
            self.th.join(timeout=1)

    def summary(self, windows):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        nv = self.nv
        sel = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)] or self.samples
        sm = sorted(s[1] for s in sel)
        bits = 0
        for s in sel:
            bits |= s[2]
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2, "display_clock_setting": 0x100}
        reasons = [k for k, v in names.items() if bits & v]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.dev, nv.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            mx = None
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sel),
                "power_w_max": round(max(s[3] for s in sel), 1)}


# ------------------------------------------------------------------------------------------------ CPU arms (oracle)
def numpy_panel(cfg, rows, seed):
    """Host-side synthetic panel with the distribution family of the device generator (log-uniform ancestral frequency,
    per-population drift with F = 0.05, Binomial(2, p) calls, 2 % missing), generated in chunks.  Used by the reference arm,
    which must not touch libtdsfs at all."""
    from tdsfs_pack import _pack_block, to_b32
    rng = np.random.default_rng(seed)
    n1, n2 = cfg["n1"], cfg["n2"]
    lo, hi = 1.0 / (4.0 * (n1 + n2)), 1.0 - 1.0 / (4.0 * (n1 + n2))
    out = []
    for r0 in range(0, rows, 32768):
        m = min(32768, rows - r0)
        pa = lo * np.exp(rng.random(m) * np.log(hi / lo))
        blocks = []
        for ns in (n1, n2):
            p = np.clip(pa + rng.normal(0, 1, m) * np.sqrt(0.05 * pa * (1 - pa)), 0, 1).astype(np.float32)[:, None]
            a = (rng.random((m, ns), dtype=np.float32) < p).astype(np.uint8) + (rng.random((m, ns), dtype=np.float32) < p).astype(np.uint8)
            c = np.where(a == 2, 3, a).astype(np.uint8)
            c[rng.random((m, ns), dtype=np.float32) < 0.02] = 2
            blocks.append(_pack_block(c))
        out.append(np.concatenate(blocks, axis=1))
    return to_b32(np.concatenate(out, axis=0))


def sample_panel(cfg, rows, allow_gpu=True):
    """First `rows` SNP rows of chromosome 0 of the workload: device generator when allowed and a GPU is present (identical
    bytes to the GPU arm's input), numpy otherwise.  Input generation only -- never part of a timed CPU region."""
    w1, w2 = words_for(cfg["n1"]), words_for(cfg["n2"])
    if allow_gpu:
        try:
            import torch
            if torch.cuda.is_available():
                import tdsfs_capi as T
                h = T.Handle(0)
                g = torch.empty(((rows + 31) // 32 * (w1 + w2) * 32,), dtype=torch.int32, device="cuda:0")
                h.synth_genotypes(g.data_ptr(), rows, 0, w1, w2, cfg["n1"], cfg["n2"], cfg["seed"])
                G = g.cpu().numpy().view(np.uint32)
                h.close()
                return G, "device generator"
        except Exception:  # noqa: BLE001
            pass
    return numpy_panel(cfg, rows, cfg["seed"]), "numpy generator"


def oracle_modules():
    """The only place bench.py touches oracle/: the CPU baseline / reference arm."""
    import subprocess
    opath = os.path.join(ROOT, "oracle")
    if not os.path.exists(os.path.join(opath, "libsfs_oracle.so")):
        subprocess.check_call(["make", "-s", "-C", opath])
    if opath not in sys.path:
        sys.path.insert(0, opath)
    import sfs_oracle_c
    return sfs_oracle_c


def cpu_pass(OC, G, pos, cfg, nthreads):
    """One pass of the reference algorithm on a sample: decode calls -> counts, background over the sample, windows, T."""
    w1, w2 = words_for(cfg["n1"]), words_for(cfg["n2"])
    t = time.perf_counter()
    cnt = OC.decode(G, len(pos), w1, w2, cfg["n1"], cfg["n2"], nthreads=nthreads)
    r = OC.scan(cnt, pos, [0, len(pos)], cfg["n1"], cfg["n2"], W=cfg["W"], bg="genome", nthreads=nthreads)
    return time.perf_counter() - t, len(r["start"])


def cpu_sample_rate(cfg, budget_s, nthreads, steps=1, warmup=0, allow_gpu=True):
    """Time the oracle port on a bounded sample sized (by a pilot) to ~budget_s per step.  Returns dict."""
    OC = oracle_modules()
    nthreads = nthreads or OC.max_threads()
    pos_all = positions_for(cfg, [0])[0]
    per_win = cfg["W"] // cfg["mean_gap"]
    pilot_rows = min(len(pos_all), per_win * max(2 * nthreads, 8))
    G, gen = sample_panel(cfg, pilot_rows, allow_gpu)
    t_pilot, nwin = cpu_pass(OC, G, pos_all[:pilot_rows], cfg, nthreads)
    rate = pilot_rows / t_pilot
    rows = int(min(len(pos_all), max(pilot_rows, rate * budget_s)))
    if rows != pilot_rows:
        G, gen = sample_panel(cfg, rows, allow_gpu)
    pos = pos_all[:rows]
    times = []
    for i in range(warmup + steps):
        t, nwin = cpu_pass(OC, G, pos, cfg, nthreads)
        if i >= warmup:
            times.append(t)
    t_step = float(np.mean(times))
    return dict(value=rows / t_step, unit="SNPs/s", cores=nthreads, kind="port", seconds_per_step=t_step,
                sample=f"first {rows} SNPs ({nwin} windows of {cfg['W']} bp) of chromosome 0 of the workload, background over the "
                       f"sample, all {cfg['n1']}+{cfg['n2']} diploids; C restatement of the reference algorithm "
                       f"(oracle/sfs_oracle.c, dense per-window spectra), {nthreads} threads; input from the {gen}"), rows, times


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total_budget = 150.0
    per_step = max(2.0, total_budget / (args.steps + args.warmup + 1))
    cb, rows, times = cpu_sample_rate(cfg, per_step, 0, steps=args.steps, warmup=args.warmup, allow_gpu=False)
    line = {"impl": "reference", "metric": "SNPs/sec, 2D-SFS + T2D/T1D 20 kb window scan", "value": cb["value"], "unit": "SNPs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "S": cfg["S"], "n1": cfg["n1"], "n2": cfg["n2"], "window_bp": cfg["W"],
                       "note": "reference is single-process Python; this arm runs its algorithm restated in C on all host cores, "
                               "on a bounded sample (SNPs/s is per-SNP throughput of the same pass)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "SNPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
class _DevBuf:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def run_b200(args, cfg):
    import torch
    import torch.distributed as dist
    import tdsfs_capi as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")  # keeps NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    n1, n2, W = cfg["n1"], cfg["n2"], cfg["W"]
    w1, w2 = words_for(n1), words_for(n2)
    RW = w1 + w2

    # ---- this rank's shard: contiguous chromosomes (strong) or a full copy of the workload per rank (weak)
    if args.scaling == "strong":
        chroms = shard_chroms(cfg["C"], world, rank)
    else:
        chroms = list(range(cfg["C"]))
    sizes_all = chrom_sizes(cfg["S"], cfg["C"])
    starts_all = np.concatenate([[0], np.cumsum(sizes_all)])
    pos_list = positions_for(cfg, chroms)
    pos_host = np.concatenate(pos_list) if pos_list else np.zeros(0, np.int32)
    S_local = int(len(pos_host))
    off = np.concatenate([[0], np.cumsum([len(p) for p in pos_list])]).astype(np.int64)
    snp0 = int(starts_all[chroms[0]]) if chroms else 0
    if args.scaling == "weak":
        snp0 += rank * cfg["S"]  # different SNPs on every rank
    S_total = cfg["S"] if args.scaling == "strong" else cfg["S"] * world

    h = T.Handle(local_rank)
    # a real (non-NULL) stream shared by torch (events, NCCL ordering) and libtdsfs
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    h.set_panel(n1, n2, True)
    g_words = max((S_local + 31) // 32, 1) * RW * 32  # B32 layout: blocks of 32 SNPs, zero padded
    G = torch.empty((g_words,), dtype=torch.int32, device=dev)
    h.synth_genotypes(G.data_ptr(), S_local, snp0, w1, w2, n1, n2, cfg["seed"])
    pos_dev = torch.from_numpy(pos_host).to(dev)
    torch.cuda.synchronize()

    hist_cache = {}

    def hist_tensor():
        ptr, n, _ = h.background_device()
        if ptr not in hist_cache:
            hist_cache[ptr] = torch.as_tensor(_DevBuf(ptr, n), device=dev)
        return hist_cache[ptr]

    peer = {"on": False}

    def exchange():
        """Sum the background histogram over ranks in place: the library's peer-memory kernel, else NCCL."""
        if peer["on"]:
            h.peer_allreduce_background()
        else:
            dist.all_reduce(hist_tensor())

    def device_step():
        """K1 (+ all-reduce of the background) + finalize + K2 + K3/K4, all enqueued on one stream, no host sync."""
        h.plan(W)                  # K2 on a side stream: overlaps K1 and the all-reduce
        h.background(T.BG_GENOME)
        if world > 1:
            exchange()
        h.finalize_background()
        h.scan(W, fetch=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    windows = []

    # ---- (A) device-resident throughput
    h.load_genotypes(G, S_local, w1, w2, n1, n2, pos_dev, off)
    if world > 1 and args.exchange == "peer":
        from tdsfs_dist import peer_setup
        h.background(T.BG_GENOME)  # allocates the histogram that the peers map
        peer["on"] = peer_setup(h)
    h.set_sync(False)
    launches0 = h.launch_count()
    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    launches_per_step = (h.launch_count() - launches0) // max(args.warmup, 3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        device_step()
    ev1.record(stream)
    barrier()
    t_a1 = time.time()
    windows.append((t_a0, t_a1))
    h._check(h._L.tdsfs_check(h._h))
    ms_total = ev0.elapsed_time(ev1)
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    value = S_total / (ms_step * 1e-3)

    # ---- (B) kernel-level times (CUDA events inside the library), K synchronous steps
    h.set_sync(True)
    kt = []
    t_b0 = time.time()
    for _ in range(args.steps):
        h.plan(W)
        h.background(T.BG_GENOME)
        if world > 1:
            exchange()
            torch.cuda.synchronize()
        h.finalize_background()
        h.scan(W, fetch=False)
        kt.append(h.timings())
    windows.append((t_b0, time.time()))
    k1_ms = float(np.mean([k["k1_count"] for k in kt]))
    kernel_ms = {k: float(np.mean([x[k] for x in kt])) for k in kt[0]}
    n_windows = h.candidates(W)
    nw_t = torch.tensor([n_windows], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nw_t)
    n_windows_total = int(nw_t.item())

    # ---- (C) end to end through the C ABI with pinned HOST buffers
    e2e = None
    if not args.no_e2e:
        G_host = torch.empty((g_words,), dtype=torch.int32, pin_memory=True)
        G_host.copy_(G)
        pos_pin = torch.from_numpy(pos_host).pin_memory()
        torch.cuda.synchronize()
        cap = n_windows

        def e2e_step():
            h.load_genotypes(G_host, S_local, w1, w2, n1, n2, pos_pin, off)
            if world > 1:
                h.plan(W)
                h.background(T.BG_GENOME)
                exchange()
                torch.cuda.synchronize()
                h.finalize_background()
                return h.scan(W, fetch=True)
            return h.run_bp(T.BG_GENOME, W, fetch=True)

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for _ in range(1):
            res = e2e_step()
        barrier()
        t_c0 = time.time()
        ev0.record(stream)
        for _ in range(e2e_steps):
            res = e2e_step()
        ev1.record(stream)
        barrier()
        t_c1 = time.time()
        windows.append((t_c0, t_c1))
        wall = (t_c1 - t_c0) / e2e_steps  # host clock around synchronous C-ABI calls (copies run on the library's copy stream)
        tw = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        d2h = sum(v.nbytes for v in res.values())
        e2e = {"value": S_total / float(tw.item()), "unit": "SNPs/s", "h2d_bytes_per_step": int(g_words * 4 + S_local * 4),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(tw.item()) * 1e3, "steps": e2e_steps,
               "api": "tdsfs_load_genotypes(host) + tdsfs_run_bp(host results) via ctypes",
               "note": "per-rank bytes; PCIe host->device copy of the 2-bit matrix dominates"}
    sampler.stop()
    if peer["on"]:
        from tdsfs_dist import peer_teardown
        peer_teardown(h)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    geno_bytes = (n1 + n2) / 4.0
    achieved = S_local * geno_bytes / (k1_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json"))).get(args.workload)
    except Exception:  # noqa: BLE001
        pass
    step_bytes = S_local * (geno_bytes + 4.0)
    roof = {"bound": "hbm", "kernel": "k1_genotypes (count kernel)", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": S_local * geno_bytes, "kernel_ms": k1_ms,
            "whole_step": {"algorithmic_bytes": step_bytes, "ms": ms_step if world == 1 else kernel_ms["pass_total"],
                           "achieved": step_bytes / ((ms_step if world == 1 else kernel_ms["pass_total"]) * 1e-3) / 1e9},
            "kernel_ms_all": kernel_ms,
            "how": "CUDA events recorded by libtdsfs on the launching stream around each kernel, mean of K synchronous steps "
                   "run right after the timed region; K1 bytes = S*(n1+n2)/4 (2-bit calls), whole step adds 4 B/SNP positions"}
    roof["whole_step"]["frac"] = roof["whole_step"]["achieved"] / peak

    cb = None
    if world == 1 and not args.no_cpu:
        cb, _, _ = cpu_sample_rate(cfg, args.cpu_seconds, 0)
        cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": "SNPs/sec, 2D-SFS + T2D/T1D 20 kb window scan", "value": value, "unit": "SNPs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u32 popcount/histograms + f64 likelihoods", "data": "synthetic",
            "config": {"workload": cfg["name"], "S": S_total, "n1": n1, "n2": n2, "window_bp": W, "windows": n_windows_total,
                       "chromosomes": cfg["C"], "sharding": f"contiguous chromosome ranges over {world} rank(s)", "background": "genome-wide"
                       + ((", all-reduced in place by the library's peer-memory kernel (CUDA IPC over NVLink, uint32 sum)" if peer["on"]
                          else ", all-reduced (NCCL, uint32 sum)") if world > 1 else ""), "row_bytes": RW * 4,
                       "l2": "inputs (per-rank genotype matrix %.1f GB) larger than L2" % (S_local * RW * 4 / 1e9)},
            "roofline": roof, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step), "clocks": sampler.summary(windows)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config5", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: background all-reduce through the library's peer-memory kernel (default) or NCCL")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--snps", type=int, default=0, help="override the SNP count (profiling runs only; not a bench value)")
    args = ap.parse_args()
    cfg = dict(WORKLOADS[args.workload])
    if args.snps:
        cfg["S"] = args.snps
        cfg["name"] += f" [REDUCED to {args.snps} SNPs: profiling run, not a bench value]"
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
