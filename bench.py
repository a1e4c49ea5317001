#!/usr/bin/env python3
"""bench.py -- the hot path's headline metric on B200: variant pairs/sec (r2 + D', 5008 haplotypes).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ld_triangle|ld_area]

Workload at N=1 is BASELINE.json configs[1]: ld_triangle, all-pairs r2/D' for 2,000 variants x
5008 haplotypes (1,999,000 pairs).  One *step* = one all-pairs pass over one 2,000-variant set.
With N GPUs every rank owns its own 2,000-variant set (the reference parallelises over source
files, ld_triangle.py:406-408) -- weak scaling, no data-path collective; value = pairs of all
ranks / max-over-ranks device time.

Printed JSON (one line, rank 0):
  value        pairs/s with the bit planes already resident in HBM and results left in HBM
  e2e          pairs/s through the public API with HOST buffers: planes H2D (pinned) + mask/count
               kernel + all-pairs kernel + packed results D2H, every step
  roofline     dominant kernel (the all-pairs kernel) against the pipe that bounds it; its duration is
               measured live with CUDA events recorded around that kernel's launches on the stream it
               runs on (ldx_kernel_timing), over a repeat of the timed steps (the events would serialise the
               programmatic dependent launches of the timed region itself)
  steady_state the same kernel on a 32,768-variant set (221 tiles per SM instead of one): what the engine
               sustains once tile quantisation and launch latency stop dominating (BASELINE configs[3] regime)
  cpu_baseline the reference algorithm (pure-Python port, oracle/calc_ld_port.py) on the host cores
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_VARIANTS, N_HAP = 2000, 5008
WORKLOAD = "ld_triangle: all-pairs r2/D' matrix, 2000 variants x 5008 haplotypes (BASELINE configs[1])"
OPS_PER_PAIR_I8 = 2 * N_HAP               # int8 Gram-matrix formulation: one MAC per haplotype
POPC_PER_PAIR = (N_HAP + 31) // 32        # AND+POPC formulation: 32-bit popcounts per pair


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "sm_max_mhz": float(p.get("sm_max_mhz", 1965.0)), "source": "measured"}
    except Exception:
        # B200_PROFILING.md fallback figures
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0, "source": "fallback"}


def ncu_traffic(profile_name):
    """DRAM bytes (read + write) of one launch of the dominant kernel, from the committed summary of an
    `ncu --set full` capture of this same command (profiles/, written by tools/ncu_summarize.py); None if absent."""
    units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        total = 0.0
        with open(os.path.join(ROOT, "profiles", profile_name)) as fh:
            for line in fh:
                f = line.split()
                if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(f[1]) * units[f[2]]
        return total or None
    except Exception:
        return None


# --------------------------------------------------------------------------- CPU baseline
_W = {}


def _cpu_init(seed_base):
    """Per-process state: genotype lists of 64 synthetic variants, extracted once."""
    import multiprocessing as mp
    from ld_tools_b200.synth import synth_haplotypes
    ident = mp.current_process()._identity
    seed = seed_base + (ident[0] if ident else 0)
    h = synth_haplotypes(64, N_HAP, seed=seed)
    _W["lists"] = [list(map(int, row)) for row in h]
    _W["rng"] = np.random.default_rng(seed)


def _cpu_step(n_pairs):
    from oracle.calc_ld_port import calc_ld
    lists, rng = _W["lists"], _W["rng"]
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, len(lists), size=(n_pairs, 2))]
    t0 = time.perf_counter()
    for a, b in pairs:
        calc_ld(lists[a], lists[b])
    return n_pairs, time.perf_counter() - t0


class CpuArm:
    """The reference algorithm (pure-Python port of backend/calc_ld.py, oracle/calc_ld_port.py) on
    all host cores, genotype lists pre-extracted -- generous to the reference, whose drivers also
    pay two tabix fetches and 2 x 2504 pysam lookups per pair (ld_triangle.py:158-186)."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_init, initargs=(1000,))
        self.pool.map(_cpu_step, [2] * self.cores)           # force start-up + initialisers

    def step(self, pairs_per_core):
        res = self.pool.map(_cpu_step, [pairs_per_core] * self.cores, chunksize=1)
        n = sum(r[0] for r in res)
        return n, max(r[1] for r in res)                     # concurrent workers: time = slowest

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(pairs_per_core, cores=None):
    arm = CpuArm(cores)
    n, busy = arm.step(pairs_per_core)
    arm.close()
    return {"value": n / busy, "unit": "pairs/s", "cores": arm.cores, "kind": "port",
            "sample": f"{n} calc_ld calls (pure-Python port of backend/calc_ld.py) on random pairs of "
                      f"synthetic 5008-haplotype variants, {pairs_per_core} per process x {arm.cores} "
                      f"processes, genotype lists pre-extracted",
            "per_core": n / busy / arm.cores}


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm for the same metric on the host cores."""
    if rank != 0:
        return
    # each step is a bounded sample sized so that the whole run stays within ~2 minutes
    budget_s = 120.0 / max(args.steps + args.warmup, 1)
    per_core = int(min(2000, max(20, budget_s / 1.0e-3)))
    arm = CpuArm()
    for _ in range(args.warmup):
        arm.step(per_core)
    n_tot, t_tot = 0, 0.0
    for _ in range(args.steps):
        n, t = arm.step(per_core)
        n_tot += n
        t_tot += t
    arm.close()
    v = n_tot / t_tot
    n_pairs_step = per_core * arm.cores
    last = {"unit": "pairs/s", "cores": arm.cores, "kind": "port", "per_core": v / arm.cores,
            "sample": f"{n_pairs_step} calc_ld calls per step (pure-Python port of backend/calc_ld.py) on random "
                      f"pairs of synthetic 5008-haplotype variants, {per_core} per process x {arm.cores} "
                      f"processes, genotype lists pre-extracted"}
    line = {"impl": "reference", "metric": "variant pairs/sec (r2+D', 5008 haplotypes)", "value": v,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * n_pairs_step / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step": f"bounded sample: {n_pairs_step} pairs per step"},
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index, period_s=0.005):
        super().__init__(daemon=True)
        self.period_s = period_s
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = get(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period_s)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ld_tools_b200 import Context, Store
    from ld_tools_b200.engine import ENGINE_AUTO, ENGINE_MMA, ENGINE_POPC
    from ld_tools_b200.synth import pack_bits, synth_haplotypes

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    engine = {"auto": ENGINE_AUTO, "popc": ENGINE_POPC, "mma": ENGINE_MMA}[args.engine]

    # synthetic 1000G-shaped input: this rank's own 2,000-variant set
    h = synth_haplotypes(N_VARIANTS, N_HAP, seed=20130502 + rank)
    planes_np = pack_bits(h)
    n_pairs = N_VARIANTS * (N_VARIANTS - 1) // 2
    rows = np.arange(N_VARIANTS, dtype=np.int64)
    mask_np = np.zeros(planes_np.shape[1], dtype="<u8")
    mask_np[:N_HAP // 64] = ~np.uint64(0)
    mask_np[N_HAP // 64] = np.uint64((1 << (N_HAP % 64)) - 1)

    ctx = Context(local_rank)
    stream = torch.cuda.Stream(device=dev)       # kernels, copies and timing events all on this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if args.tile_n:
        from ld_tools_b200._lib import TUNE_MMA_TILE_N
        ctx.set_tuning(TUNE_MMA_TILE_N, args.tile_n)
    store = Store.from_planes(ctx, planes_np, N_HAP)
    store.set_mask(mask_np)
    depth = max(1, args.pipeline)
    d_out = [torch.empty(n_pairs, dtype=torch.int32, device=dev) for _ in range(depth)]   # one result buffer per call in flight
    d_packed = d_out[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: device-resident ("value").  A step = one all-pairs pass over the variant set: bit gather +
    #      all-pairs kernel + deferred-pairs kernel, results left in HBM.  Calls are enqueued asynchronously, up
    #      to `depth` in flight (each with its own result buffer); ldx_resolve() then settles the near-tie pairs
    #      of all of them (the one host round trip of the path) -- inside the timed region.
    host = {"enqueue_s": 0.0, "resolve_s": 0.0}           # host-side cost of the loop (diagnostic: is the GPU ever starved?)

    def run_steps(n, timed):
        marks = []
        t_h = time.perf_counter()
        for k in range(n):
            f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
            s1 = torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            flush.zero_()                       # L2 flush between iterations; its duration is subtracted below
            f1.record(stream)
            store.triangle_dev(rows, d_out[k % depth].data_ptr(), engine=engine)
            s1.record(stream)                   # f1..s1 = the step's kernels
            marks.append((f0, f1, s1))
            if (k + 1) % depth == 0 or k == n - 1:
                t_r = time.perf_counter()
                ctx.resolve()
                if timed:
                    host["resolve_s"] += time.perf_counter() - t_r
        if timed:
            host["enqueue_s"] += time.perf_counter() - t_h - host["resolve_s"]
        return marks

    run_steps(args.warmup, False)
    barrier()
    sampler = ClockSampler(local_rank, args.sampler_ms * 1e-3)
    sampler.start()
    launches0 = ctx.launch_count
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record(stream)
    marks = run_steps(args.steps, True)
    t_end.record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    # The dominant kernel alone: the same steps once more, now with CUDA events recorded around every all-pairs
    # kernel launch on the stream it runs on (ldx_kernel_timing).  Those events sit between the kernels of a call
    # and serialise their programmatic dependent launches (+12 us per step), which is why the timed region above
    # does not carry them.
    ctx.kernel_timing(True)
    run_steps(min(args.steps, 100), False)
    barrier()
    dom_ms, dom_launches = ctx.kernel_timing(False)
    flush_ms = float(sum(f0.elapsed_time(f1) for f0, f1, _ in marks))
    step_ms = float(t_begin.elapsed_time(t_end)) - flush_ms     # the whole timed region minus the L2 flushes
    kern_ms = float(sum(f1.elapsed_time(s1) for _, f1, s1 in marks))   # gather + all-pairs + deferred-pairs kernels

    # ---- leg 2: end to end through the host API: pinned planes H2D + mask/count kernel + all-pairs kernel + packed
    #      results D2H, every step, through the blocking calls a driver makes (Store.upload / set_mask / triangle).
    #      Like the reference, which works on several source files at once (multiprocessing.Pool over files,
    #      ld_triangle.py:406-408), the job runs `--e2e-contexts` independent contexts -- each with its own stream,
    #      store and pinned result buffer, driven by its own host thread (the library releases the GIL) -- so that one
    #      variant set's 8 MB result copy overlaps the next set's upload and kernels.  Steps alternate between them.
    planes_pin = torch.from_numpy(planes_np.view(np.int64)).pin_memory()
    planes_host = planes_pin.numpy().view("<u8")
    n_ctx = max(1, args.e2e_contexts)
    lanes = []
    for c in range(n_ctx):
        if c == 0:
            l_ctx, l_stream, l_store = ctx, stream, store
        else:
            l_ctx = Context(local_rank)
            l_stream = torch.cuda.Stream(device=dev)
            l_ctx.set_stream(l_stream.cuda_stream)
            l_store = Store(l_ctx, N_VARIANTS, N_HAP)
        out_pin = torch.empty(n_pairs, dtype=torch.int32).pin_memory()
        lanes.append({"ctx": l_ctx, "stream": l_stream, "store": l_store, "out_pin": out_pin, "out": out_pin.numpy().view(np.uint32)})
    out_host = lanes[0]["out"]

    def lane_steps(lane, n, end_event=None):
        torch.cuda.set_device(local_rank)
        for _ in range(n):
            lane["store"].upload(0, planes_host)
            lane["store"].set_mask(mask_np)
            lane["store"].triangle(rows, engine=engine, out=lane["out"])
        if end_event is not None:
            end_event.record(lane["stream"])

    def run_lanes(n_each, events=None):
        ths = [threading.Thread(target=lane_steps, args=(lane, n_each, events[k] if events else None)) for k, lane in enumerate(lanes[1:], 1)]
        for t in ths:
            t.start()
        lane_steps(lanes[0], n_each, events[0] if events else None)
        for t in ths:
            t.join()

    run_lanes(max(args.warmup // 2, 3))
    barrier()
    e_each = max(args.steps // (4 * n_ctx), 3)
    e_steps = e_each * n_ctx
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in lanes]
    e0.record(stream)
    t0 = time.perf_counter()
    run_lanes(e_each, ends)
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = max(float(e0.elapsed_time(e)) for e in ends)      # device clock, start of the first step to the end of the last copy
    clocks = sampler.stop()

    # ---- max over ranks
    t = torch.tensor([step_ms, kern_ms, e2e_ms, dom_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms, e2e_ms, dom_ms = t.tolist()
    total_pairs = n_pairs * world
    value = total_pairs * args.steps / (step_ms * 1e-3)
    e2e_value = total_pairs * e_steps / (e2e_ms * 1e-3)
    kern_s = dom_ms * 1e-3 / max(dom_launches, 1)        # average duration of one all-pairs kernel launch

    # parity spot-check of what was just timed (device-resident result vs host-API result)
    same = all(bool((d.cpu().numpy().view(np.uint32) == lane["out"]).all()) for d in d_out[:4] for lane in lanes)

    used_mma = engine in (ENGINE_MMA, ENGINE_AUTO)      # AUTO picks tcgen05 from 256 variants up
    if used_mma:
        peak = 2.0 * peaks["bf16_tflops"]      # dense int8 = 2 x dense bf16 on sm_100a
        roof = {"bound": "tensor", "achieved": n_pairs * OPS_PER_PAIR_I8 / kern_s / 1e12, "peak": peak,
                "unit": "TFLOP/s", "peak_source": f"2 x {peaks['source']} cuBLAS bf16 ({peaks['bf16_tflops']} TFLOP/s)"}
    else:
        # AND+POPC engine: bounded by the integer POPC pipe (16 lanes/clk/SM), not by HBM or tensor
        peak = 148 * 16 * peaks["sm_max_mhz"] * 1e6 / 1e12
        roof = {"bound": "popc", "achieved": n_pairs * POPC_PER_PAIR / kern_s / 1e12, "peak": peak,
                "unit": "TPOPC32/s", "peak_source": "148 SM x 16 POPC/clk x clocks.max.sm"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = ncu_traffic("r01_ncu_full_mma_v10_v2000.txt") if used_mma else None     # bytes per launch (L2-resident operands)
    roof["kernel"] = "triangle_mma_kernel" if used_mma else "triangle_popc_kernel"
    roof["kernel_ms"] = kern_s * 1e3
    roof["kernel_launches_timed"] = int(dom_launches)
    roof["all_kernels_ms_per_step"] = kern_ms / args.steps
    roof["algorithmic_per_launch"] = (f"{n_pairs} pairs x {OPS_PER_PAIR_I8} int8 ops" if used_mma
                                      else f"{n_pairs} pairs x {POPC_PER_PAIR} POPC32")

    # ---- steady state: one large variant set, kernel-only (rank 0's GPU; every rank runs it so clocks stay loaded)
    steady = None
    if used_mma and not args.no_steady:
        from ld_tools_b200.synth import random_planes
        v_big = 32768
        big = Store.from_planes(ctx, random_planes(v_big, N_HAP, seed=11 + rank), N_HAP)
        big.set_mask(mask_np)
        rows_big = np.arange(v_big, dtype=np.int64)
        pairs_big = v_big * (v_big - 1) // 2
        out_big = torch.empty(pairs_big, dtype=torch.int32, device=dev)
        for _ in range(2):
            big.triangle_dev(rows_big, out_big.data_ptr(), engine=engine)
            ctx.resolve()
        ctx.kernel_timing(True)
        for _ in range(3):
            big.triangle_dev(rows_big, out_big.data_ptr(), engine=engine)
            ctx.resolve()
        big_ms, big_n = ctx.kernel_timing(False)
        pps = pairs_big / (big_ms * 1e-3 / big_n)
        steady = {"workload": f"ld_triangle, {v_big} variants x {N_HAP} haplotypes ({pairs_big} pairs), kernel only",
                  "value": pps, "unit": "pairs/s", "kernel_ms": big_ms / big_n,
                  "roofline_frac": pps * OPS_PER_PAIR_I8 / 1e12 / (2.0 * peaks["bf16_tflops"])}
        del out_big
        big.close()

    line = {"metric": "variant pairs/sec (r2+D', 5008 haplotypes)", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8 tensor (exact) + f32 screen / f64 settle" if used_mma else "u64 popcount + f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": n_pairs, "engine": args.engine,
                       "l2": "flushed (256 MiB write) between timed iterations; flush time excluded",
                       "calls_in_flight": depth, "sharding": "one variant set per GPU"},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(planes_np.nbytes + mask_np.nbytes + rows.nbytes),
                    "d2h_bytes_per_step": int(out_host.nbytes), "steps": e_steps, "wall_s": e2e_wall, "contexts": n_ctx,
                    "api": "Store.upload + Store.set_mask + Store.triangle (blocking host-buffer calls), one host thread per context"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "parity_selfcheck": same,
            "host": {"enqueue_us_per_step": 1e6 * host["enqueue_s"] / args.steps, "resolve_wait_us_per_step": 1e6 * host["resolve_s"] / args.steps,
                     "flush_us_per_step": 1e3 * flush_ms / args.steps}}
    if steady:
        line["steady_state"] = steady
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_pairs_per_core)
        print(json.dumps(line), flush=True)
    for lane in lanes[1:]:
        lane["store"].close()
        lane["ctx"].close()
    store.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- ld_area workload (BASELINE configs[2])
AREA_VARIANTS, AREA_QUERIES, AREA_FLANK, AREA_EUR_SAMPLES = 1_100_000, 1000, 500_000, 503
AREA_WORKLOAD = ("ld_area: 1,000 query variants, +/-500 kb flanks, r2 >= 0.8, EUR subset (503 samples = 1006 of 5008 "
                 "haplotypes, by mask), synthetic 1.1M-variant chr22 (BASELINE configs[2])")


def run_area(args, rank, world, local_rank):
    """The window scan (K4, ld_area.py:215-249): value = candidate pairs scanned per second with the store
    resident in HBM; e2e = through the host API (query arrays H2D, kept hits D2H and sorted).  With N GPUs
    every rank owns its own chromosome-sized store (region sharding: ld_tools_b200/shard.py)."""
    import torch
    import torch.distributed as dist
    from ld_tools_b200 import Context, Store, shard
    from ld_tools_b200.engine import threshold_e4
    from ld_tools_b200.synth import fill_store_grouped
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    ctx = Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(77 + rank)
    nv = AREA_VARIANTS
    st = Store(ctx, nv, N_HAP)
    fill_store_grouped(st, dev, 22 + rank, 0, nv)                     # generated on the GPU: groups of 8 neighbours in LD
    pos0 = np.sort(rng.integers(16_050_000, 51_200_000, size=nv)).astype(np.int32)
    end0 = pos0 + 1
    st.set_annotations(pos0, end0, np.arange(nv, dtype=np.int64), np.ones(nv, np.uint8))
    hap = np.sort(rng.choice(N_HAP // 2, AREA_EUR_SAMPLES, replace=False))
    st.select_haplotypes(np.concatenate([2 * hap, 2 * hap + 1]))      # get_sample_names -> mask plane
    q_row = np.sort(rng.choice(nv, AREA_QUERIES, replace=False)).astype(np.int64)
    lo, hi, ws, we = shard.window_bounds(pos0, 1, pos0[q_row].astype(np.int64) + 1, AREA_FLANK)
    n_cand = int((hi - lo).sum())
    thres = threshold_e4(0.8)
    cap = 1 << 22
    d_hits = torch.empty(cap * 4, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        st.window_dev(q_row, lo, hi, ws, we, "r_square", thres, d_hits.data_ptr(), cap, d_cnt.data_ptr())
        ctx.resolve()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    ctx.kernel_timing(True)
    ev = []
    for _ in range(args.steps):
        flush.zero_()                       # the 704 MB store is larger than L2 anyway; flushed for good measure
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        ev.append((e0, e1))
    barrier()
    launches = ctx.launch_count - launches0
    dom_ms, dom_n = ctx.kernel_timing(False)
    step_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    scanned = int(d_cnt.cpu()[1])
    n_hits = int(d_cnt.cpu()[0])
    # e2e: host API
    for _ in range(2):
        st.window(q_row, lo, hi, ws, we, "r_square", thres, cap=cap)
    barrier()
    e_steps = max(args.steps // 4, 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(e_steps):
        hits, sc = st.window(q_row, lo, hi, ws, we, "r_square", thres, cap=cap)
    e1.record(stream)
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = float(e0.elapsed_time(e1))
    clocks = sampler.stop()
    t = torch.tensor([step_ms, e2e_ms, dom_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms, dom_ms = t.tolist()
    row_bytes = st.stride_words * 8
    kern_s = dom_ms * 1e-3 / max(dom_n, 1)
    roof = {"bound": "hbm", "achieved": scanned * row_bytes / kern_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "peak_source": f"{peaks['source']} copy bandwidth", "traffic": ncu_traffic("r01_ncu_full_window_mq_configs2.txt"), "kernel": "window_mq_kernel",
            "kernel_ms": kern_s * 1e3,
            "kernel_launches_timed": int(dom_n),
            "algorithmic_per_launch": f"{scanned} pairs x {row_bytes} B (one candidate row per pair; the multi-query kernel loads a row once per "
                                      f"four queries and re-reads it from L1 for the other groups, hence frac > 1)"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    line = {"metric": "variant pairs/sec (r2+D', 5008 haplotypes)", "value": scanned * world * args.steps / (step_ms * 1e-3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 popcount + f64", "data": "synthetic",
            "config": {"workload": AREA_WORKLOAD, "pairs_per_step_per_gpu": scanned, "candidate_rows_per_step": n_cand,
                       "hits_per_step": n_hits, "l2": "store (704 MB) larger than L2; 256 MiB flush between timed iterations",
                       "sharding": "one chromosome store per GPU"},
            "e2e": {"value": scanned * world * e_steps / (e2e_ms * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": int(q_row.nbytes + lo.nbytes + hi.nbytes + ws.nbytes + we.nbytes + 8 * (len(q_row) + 1)),
                    "d2h_bytes_per_step": int(16 * len(hits) + 16), "steps": e_steps, "wall_s": e2e_wall},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
    if rank == 0:
        print(json.dumps(line), flush=True)
    st.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["ld_triangle", "ld_area"], default="ld_triangle",
                    help="ld_triangle = BASELINE configs[1] (the headline); ld_area = configs[2], the HBM-bound window scan")
    ap.add_argument("--engine", choices=["auto", "popc", "mma"], default="auto")
    ap.add_argument("--tile-n", type=int, default=0, help="tcgen05 tile width override (0 = heuristic)")
    ap.add_argument("--cpu-pairs-per-core", type=int, default=2000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline", type=int, default=32, help="device-resident calls in flight per ldx_resolve()")
    ap.add_argument("--e2e-contexts", type=int, default=3,
                    help="independent contexts (stream + store + host thread) the end-to-end leg pipelines its steps over")
    ap.add_argument("--sampler-ms", type=float, default=5.0, help="period of the NVML clock sampler thread")
    ap.add_argument("--no-steady", action="store_true", help="skip the 32,768-variant steady-state leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    elif args.workload == "ld_area":
        run_area(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
