#!/usr/bin/env python3
"""bench.py -- the hot path's headline metric on B200: variant pairs/sec (r2 + D', 5008 haplotypes).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ld_triangle|ld_area]

Workload at N=1 is BASELINE.json configs[1]: ld_triangle, all-pairs r2/D' for 2,000 variants x
5008 haplotypes (1,999,000 pairs).  One *step* = one all-pairs pass over one 2,000-variant set.
With N GPUs `value` is N replicas of that step (one variant set per rank: the reference parallelises over source
files, ld_triangle.py:406-408; weak scaling, no data-path collective), so that N=1 is the same number in every
record; the north_star's own partitioning is measured next to it, at every N, in `sharded`.

Printed JSON (one line, rank 0):
  value         pairs/s with the bit planes already resident in HBM and results left in HBM, L2 flushed between steps
  back_to_back  the same steps without the flush (and without its 74 us of cover for the host's enqueue loop)
  e2e           pairs/s through the public API with HOST buffers: planes H2D (pinned) + mask/count kernel + all-pairs
                kernel + packed results D2H, every step; `e2e.d2h_probe` = what plain pinned D2H copies reach in the same run
  roofline      dominant kernel (the all-pairs kernel) against the pipe that bounds it; its duration is
                measured live with CUDA events recorded around that kernel's launches on the stream it
                runs on (ldx_kernel_timing), over a repeat of the timed steps
  batched       >= 16 variant sets of the same shape in ONE launch (ldx_triangle_batch_dev): what a job of many
                (source file, chromosome) matrices sustains, against the same roofline
  steady_state  the same kernel on a 32,768-variant set (221 tiles per SM instead of one)
  ld_area       BASELINE configs[2] through the window kernel (1,000 queries, +/-500 kb, r2 >= 0.8, EUR subset store)
  sharded       north_star (3) at this N: configs[3] (100,000-variant triangle, row-range sharded, strong scaling) and
                configs[4] (genome-scale ld_area: 50,000 queries over 80 M variants) region-sharded with halos, the kept pairs
                gathered over NCCL and sorted on the device inside the timed region
  cpu_baseline  the reference algorithm (pure-Python port, oracle/calc_ld_port.py) on the host cores
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
METRIC = "variant pairs/sec (r2+D', 5008 haplotypes)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "sm_max_mhz": float(p.get("sm_max_mhz", 1965.0)), "source": "measured"}
    except Exception:
        # B200_PROFILING.md fallback figures
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0, "source": "fallback"}


def ncu_traffic(profile_names):
    """DRAM bytes (read + write) of one launch of the dominant kernel, from the committed summary of an
    `ncu --set full` capture of this same command (profiles/, written by tools/ncu_summarize.py); None if absent."""
    units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for name in profile_names:
        try:
            total = 0.0
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                for line in fh:
                    f = line.split()
                    if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        total += float(f[1]) * units[f[2]]
            if total:
                return total
        except Exception:
            pass
    return None


# --------------------------------------------------------------------------- CPU baseline
_W = {}


def _reference_calc_ld():
    """The reference's own backend/calc_ld.py when its tree is reachable (this container; LD_TOOLS_REFERENCE elsewhere),
    else the pinned pure-Python port.  -> (function, kind)."""
    ref = os.environ.get("LD_TOOLS_REFERENCE", "/root/reference")
    path = os.path.join(ref, "backend", "calc_ld.py")
    if os.path.exists(path):
        import importlib.util
        sys.dont_write_bytecode = True
        spec = importlib.util.spec_from_file_location("ld_tools_reference_calc_ld", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.calc_ld, "reference"
    from oracle.calc_ld_port import calc_ld
    return calc_ld, "port"


def _cpu_init(seed_base):
    """Per-process state: genotype lists of 64 synthetic variants, extracted once."""
    import multiprocessing as mp
    from ld_tools_b200.synth import synth_haplotypes
    ident = mp.current_process()._identity
    seed = seed_base + (ident[0] if ident else 0)
    h = synth_haplotypes(64, N_HAP, seed=seed)
    _W["lists"] = [list(map(int, row)) for row in h]
    _W["rng"] = np.random.default_rng(seed)
    _W["calc_ld"], _W["kind"] = _reference_calc_ld()


def _cpu_step(n_pairs):
    calc_ld = _W["calc_ld"]
    lists, rng = _W["lists"], _W["rng"]
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, len(lists), size=(n_pairs, 2))]
    t0 = time.perf_counter()
    for a, b in pairs:
        calc_ld(lists[a], lists[b])
    return n_pairs, time.perf_counter() - t0, _W["kind"]


class CpuArm:
    """The reference algorithm (backend/calc_ld.py itself where the reference tree is reachable, else its pinned pure-Python
    port oracle/calc_ld_port.py) on all host cores, genotype lists pre-extracted -- generous to the reference, whose drivers
    also pay two tabix fetches and 2 x 2504 pysam lookups per pair (ld_triangle.py:158-186)."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_init, initargs=(1000,))
        self.kind = self.pool.map(_cpu_step, [2] * self.cores)[0][2]           # force start-up + initialisers

    def step(self, pairs_per_core):
        res = self.pool.map(_cpu_step, [pairs_per_core] * self.cores, chunksize=1)
        n = sum(r[0] for r in res)
        return n, max(r[1] for r in res)                     # concurrent workers: time = slowest

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self):
        return ("backend/calc_ld.py of the reference tree" if self.kind == "reference"
                else "pure-Python port of backend/calc_ld.py (oracle/calc_ld_port.py)")


def cpu_baseline(pairs_per_core, cores=None):
    arm = CpuArm(cores)
    n, busy = arm.step(pairs_per_core)
    arm.close()
    return {"value": n / busy, "unit": "pairs/s", "cores": arm.cores, "kind": arm.kind,
            "sample": f"{n} calc_ld calls ({arm.describe()}) on random pairs of "
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
    last = {"unit": "pairs/s", "cores": arm.cores, "kind": arm.kind, "per_core": v / arm.cores,
            "sample": f"{n_pairs_step} calc_ld calls per step ({arm.describe()}) on random "
                      f"pairs of synthetic 5008-haplotype variants, {per_core} per process x {arm.cores} "
                      f"processes, genotype lists pre-extracted"}
    line = {"impl": "reference", "metric": METRIC, "value": v,
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


class Env:
    """What every leg needs: device, stream, context, the process group."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        from ld_tools_b200 import Context
        self.torch, self.dist = torch, dist
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = load_peaks()
        self.ctx = Context(local_rank)
        self.stream = torch.cuda.Stream(device=self.dev)       # kernels, copies and timing events all on this stream
        torch.cuda.set_stream(self.stream)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.peak_i8 = 2.0 * self.peaks["bf16_tflops"]       # dense int8 = 2 x dense bf16 on sm_100a (TOP/s)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def min_flag(self, ok):
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def close(self):
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def full_mask(stride_words):
    mask = np.zeros(stride_words, dtype="<u8")
    mask[:N_HAP // 64] = ~np.uint64(0)
    mask[N_HAP // 64] = np.uint64((1 << (N_HAP % 64)) - 1)
    return mask


# --------------------------------------------------------------------------- configs[1]: value, back_to_back, roofline, e2e
def leg_triangle(env, engine):
    torch, args, ctx, stream, dev = env.torch, env.args, env.ctx, env.stream, env.dev
    from ld_tools_b200 import Context, Store
    from ld_tools_b200._lib import PAIR_HIT_DTYPE, R2_MASK
    from ld_tools_b200.engine import ENGINE_AUTO, ENGINE_MMA, threshold_e4
    from ld_tools_b200.synth import pack_bits, synth_haplotypes

    # synthetic 1000G-shaped input: this rank's own 2,000-variant set
    h = synth_haplotypes(N_VARIANTS, N_HAP, seed=20130502 + env.rank)
    planes_np = pack_bits(h)
    n_pairs = N_VARIANTS * (N_VARIANTS - 1) // 2
    rows = np.arange(N_VARIANTS, dtype=np.int64)
    mask_np = full_mask(planes_np.shape[1])
    if args.tile_n:
        from ld_tools_b200._lib import TUNE_MMA_TILE_N
        ctx.set_tuning(TUNE_MMA_TILE_N, args.tile_n)
    store = Store.from_planes(ctx, planes_np, N_HAP)
    store.set_mask(mask_np)
    depth = max(1, args.pipeline)
    d_out = [torch.empty(n_pairs, dtype=torch.int32, device=dev) for _ in range(depth)]   # one result buffer per call in flight
    flush = env.flush

    # ---- device-resident ("value").  A step = one all-pairs pass over the variant set, results left in HBM.  Calls are
    #      enqueued asynchronously, up to `depth` in flight (each with its own result buffer); ldx_resolve() then settles
    #      the near-tie pairs of all of them (the one host round trip of the path) -- inside the timed region.
    host = {"enqueue_s": 0.0, "resolve_s": 0.0}           # host-side cost of the loop (diagnostic: is the GPU ever starved?)

    def run_steps(n, timed):
        marks = []
        t_h = time.perf_counter()
        for k in range(n):
            f0, f1, s1 = env.event(), env.event(), env.event()
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
    env.barrier()
    sampler = ClockSampler(env.local_rank, args.sampler_ms * 1e-3)
    sampler.start()
    launches0 = ctx.launch_count
    t_begin, t_end = env.event(), env.event()
    t_begin.record(stream)
    marks = run_steps(args.steps, True)
    t_end.record(stream)
    env.barrier()
    launches = ctx.launch_count - launches0
    flush_ms = float(sum(f0.elapsed_time(f1) for f0, f1, _ in marks))
    step_ms = float(t_begin.elapsed_time(t_end)) - flush_ms     # the whole timed region minus the L2 flushes
    kern_ms = float(sum(f1.elapsed_time(s1) for _, f1, s1 in marks))   # every kernel of the steps

    # ---- back to back: the same steps with nothing between them (operands stay in L2; the host loop has no cover)
    def run_b2b(n):
        t_h = time.perf_counter()
        for k in range(n):
            store.triangle_dev(rows, d_out[k % depth].data_ptr(), engine=engine)
            if (k + 1) % depth == 0 or k == n - 1:
                ctx.resolve()
        return time.perf_counter() - t_h
    run_b2b(max(args.warmup, 3))
    env.barrier()
    b0, b1 = env.event(), env.event()
    b0.record(stream)
    b2b_host_s = run_b2b(args.steps)
    b1.record(stream)
    env.barrier()
    b2b_ms = float(b0.elapsed_time(b1))

    # ---- the dominant kernel alone: the same steps once more, now with CUDA events recorded around every all-pairs
    #      kernel launch on the stream it runs on (ldx_kernel_timing).
    ctx.kernel_timing(True)
    run_steps(min(args.steps, 100), False)
    env.barrier()
    dom_ms, dom_launches = ctx.kernel_timing(False)

    # ---- end to end through the host API: pinned planes H2D + mask/count kernel + all-pairs kernel + packed
    #      results D2H, every step, through the blocking calls a driver makes (Store.upload / set_mask / triangle).
    #      Like the reference, which works on several source files at once (multiprocessing.Pool over files,
    #      ld_triangle.py:406-408), the job runs `--e2e-contexts` independent contexts -- each with its own stream,
    #      store and pinned result buffer, driven by its own host thread (the library releases the GIL) -- so that one
    #      variant set's 8 MB result copy overlaps the next set's upload and kernels.  The threads exist and wait at a
    #      barrier before the timed region starts.
    planes_pin = torch.from_numpy(planes_np.view(np.int64)).pin_memory()
    planes_host = planes_pin.numpy().view("<u8")
    n_ctx = max(1, args.e2e_contexts)
    lanes = []
    for c in range(n_ctx):
        if c == 0:
            l_ctx, l_stream, l_store = ctx, stream, store
        else:
            l_ctx = Context(env.local_rank)
            l_stream = torch.cuda.Stream(device=dev)
            l_ctx.set_stream(l_stream.cuda_stream)
            l_store = Store(l_ctx, N_VARIANTS, N_HAP)
        out_pin = torch.empty(n_pairs, dtype=torch.int32).pin_memory()
        lanes.append({"ctx": l_ctx, "stream": l_stream, "store": l_store, "out_pin": out_pin, "out": out_pin.numpy().view(np.uint32),
                      "out16": out_pin.numpy().view(np.uint16)[:n_pairs], "hits": np.zeros(1 << 16, dtype=PAIR_HIT_DTYPE), "n_hits": 0})
    e_each = max(args.e2e_steps // n_ctx, 3)
    e_steps = e_each * n_ctx
    # the result forms a caller can ask for: 2 bytes per pair of the one measure the drivers print (the headline: ld_triangle.py:230
    # only ever prints one), the 4-byte words with both measures, and -- with -z 0.8 -- only the pairs that pass
    thres_z = threshold_e4(0.8)
    forms = [("values16", e_each), ("packed32", max(e_each // 2, 3)), ("hits_z", max(e_each // 2, 3))]
    gate = threading.Barrier(n_ctx + 1)
    ends = {f: [env.event() for _ in lanes] for f, _ in forms}

    def lane_main(k):
        lane = lanes[k]
        st = lane["store"]
        torch.cuda.set_device(env.local_rank)
        for form, n_steps in [(f, max(args.warmup // 2, 3)) for f, _ in forms] + forms:      # a warm-up pass of every form, then the timed passes
            gate.wait()
            for _ in range(n_steps):
                st.upload(0, planes_host, wait=False)          # pinned source, enqueued; the step's one wait is in the result call
                st.set_mask(mask_np)
                if form == "values16":
                    st.triangle_values(rows, "r_square", engine=engine, out=lane["out16"])
                elif form == "packed32":
                    st.triangle(rows, engine=engine, out=lane["out"])
                else:
                    lane["n_hits"] = len(st.triangle_hits(rows, "r_square", thres_z, engine=engine, out=lane["hits"]))
            ends[form][k].record(lane["stream"])
            gate.wait()

    threads = [threading.Thread(target=lane_main, args=(k,)) for k in range(n_ctx)]
    for t in threads:
        t.start()
    for _ in forms:
        gate.wait(); gate.wait()                 # warm-up passes
    form_ms, form_wall = {}, {}
    for form, _ in forms:
        env.barrier()
        e0 = env.event()
        e0.record(stream)
        for s_ in (lane["stream"] for lane in lanes[1:]):
            s_.wait_event(e0)
        t0 = time.perf_counter()
        gate.wait(); gate.wait()                 # timed pass: every lane runs its steps
        form_wall[form] = time.perf_counter() - t0
        env.barrier()
        form_ms[form] = max(float(e0.elapsed_time(e)) for e in ends[form])      # device clock, start of the first step to the end of the last copy
    for t in threads:
        t.join()
    e2e_ms, e2e_wall = form_ms["values16"], form_wall["values16"]
    # the 2-byte values are the 4-byte words narrowed (checked on what the timed passes left in the lanes' buffers)
    w32 = lanes[0]["out"].copy()
    lanes[0]["store"].triangle_values(rows, "r_square", engine=engine, out=lanes[0]["out16"])
    narrow_ok = bool((lanes[0]["out16"] == ((w32 & 0xBFFF) | ((w32 >> 16) & 0x4000)).astype(np.uint16)).all())
    hits_ok = lanes[0]["n_hits"] == int(((w32 & R2_MASK) >= thres_z).sum())
    lanes[0]["store"].triangle(rows, engine=engine, out=lanes[0]["out"])

    # ---- what plain pinned D2H copies reach here, same run: every rank at once, three streams, 4 MB pieces (the size of a step's result)
    probe_streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
    probe_src = [torch.empty(n_pairs // 2, dtype=torch.int32, device=dev) for _ in range(3)]
    probe_dst = [torch.empty(n_pairs // 2, dtype=torch.int32).pin_memory() for _ in range(3)]
    probe_reps = 20
    for rep in range(2):                         # the first pass is the buffers' and streams' first use: untimed
        torch.cuda.synchronize()
        env.barrier()
        p0 = [env.event() for _ in probe_streams]
        p1 = [env.event() for _ in probe_streams]
        for s, a in zip(probe_streams, p0):
            a.record(s)
        for _ in range(probe_reps):
            for s, a, b in zip(probe_streams, probe_src, probe_dst):
                with torch.cuda.stream(s):
                    b.copy_(a, non_blocking=True)
        for s, a in zip(probe_streams, p1):
            a.record(s)
        torch.cuda.synchronize()
        probe_ms = max(float(a.elapsed_time(b)) for a, b in zip(p0, p1))
    env.barrier()
    clocks = sampler.stop()

    # ---- max over ranks
    step_ms, kern_ms, e2e_ms, dom_ms, b2b_ms, probe_ms, p32_ms, hz_ms = env.max_over_ranks(
        [step_ms, kern_ms, e2e_ms, dom_ms, b2b_ms, probe_ms, form_ms["packed32"], form_ms["hits_z"]])
    world = env.world
    total_pairs = n_pairs * world
    value = total_pairs * args.steps / (step_ms * 1e-3)
    e2e_value = total_pairs * e_steps / (e2e_ms * 1e-3)
    kern_s = dom_ms * 1e-3 / max(dom_launches, 1)        # average duration of one all-pairs kernel launch

    # parity spot-check of what was just timed (device-resident result vs host-API result)
    same = all(bool((d.cpu().numpy().view(np.uint32) == lane["out"]).all()) for d in d_out[:4] for lane in lanes)

    used_mma = engine in (ENGINE_MMA, ENGINE_AUTO)      # AUTO picks tcgen05 from 256 variants up
    if used_mma:
        roof = {"bound": "tensor", "achieved": n_pairs * OPS_PER_PAIR_I8 / kern_s / 1e12, "peak": env.peak_i8,
                "unit": "TFLOP/s", "peak_source": f"2 x {env.peaks['source']} cuBLAS bf16 ({env.peaks['bf16_tflops']} TFLOP/s)"}
    else:
        # AND+POPC engine: bounded by the integer POPC pipe (16 lanes/clk/SM), not by HBM or tensor
        peak = 148 * 16 * env.peaks["sm_max_mhz"] * 1e6 / 1e12
        roof = {"bound": "popc", "achieved": n_pairs * POPC_PER_PAIR / kern_s / 1e12, "peak": peak,
                "unit": "TPOPC32/s", "peak_source": "148 SM x 16 POPC/clk x clocks.max.sm"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = ncu_traffic(["r02_ncu_full_mma_direct_v2000.txt", "r01_ncu_full_mma_v10_v2000.txt"]) if used_mma else None
    roof["kernel"] = "triangle_mma_kernel" if used_mma else "triangle_popc_kernel"
    roof["kernel_ms"] = kern_s * 1e3
    roof["kernel_launches_timed"] = int(dom_launches)
    roof["all_kernels_ms_per_step"] = kern_ms / args.steps
    roof["frac_of_whole_step"] = n_pairs * OPS_PER_PAIR_I8 / (step_ms * 1e-3 / args.steps) / 1e12 / roof["peak"] if used_mma else None
    roof["algorithmic_per_launch"] = (f"{n_pairs} pairs x {OPS_PER_PAIR_I8} int8 ops" if used_mma
                                      else f"{n_pairs} pairs x {POPC_PER_PAIR} POPC32")
    h2d = int(planes_np.nbytes + mask_np.nbytes + rows.nbytes)
    d2h = int(lanes[0]["out16"].nbytes)
    other = {f: n * n_ctx for f, n in forms}
    out = {
        "value": value, "ms_per_step": step_ms / args.steps, "launches": int(launches), "clocks": clocks, "roofline": roof,
        "parity_selfcheck": same, "used_mma": used_mma, "n_pairs": n_pairs, "depth": depth,
        "back_to_back": {"value": total_pairs * args.steps / (b2b_ms * 1e-3), "unit": "pairs/s", "ms_per_step": b2b_ms / args.steps,
                         "host_us_per_step": 1e6 * b2b_host_s / args.steps,
                         "note": "no L2 flush between steps (the 1.3 MB operand set stays in L2), nothing hides the host's enqueue loop"},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps,
                "wall_s": e2e_wall, "contexts": n_ctx, "ms_per_step": e2e_ms / e_steps,
                "d2h_gbs_per_gpu": d2h * e_each * n_ctx / (e2e_ms * 1e-3) / 1e9,
                "d2h_probe": {"gbs_per_gpu": 3 * probe_reps * probe_dst[0].numel() * 4 / (probe_ms * 1e-3) / 1e9, "ranks_at_once": world,
                              "what": "plain cudaMemcpyAsync D2H of 4 MB pinned buffers on three streams, every rank at the same time"},
                "api": "Store.upload(wait=False) + Store.set_mask + Store.triangle_values (ldx_triangle_values: 2 bytes per pair of the measure asked for; "
                       "blocking host-buffer calls), one host thread per context, threads started before the timed region",
                "narrow_equals_words": narrow_ok,
                "packed32": {"value": total_pairs * other["packed32"] / (p32_ms * 1e-3), "unit": "pairs/s", "d2h_bytes_per_step": int(lanes[0]["out"].nbytes),
                             "steps": other["packed32"], "api": "Store.triangle: 4-byte words with both measures (the round-1 e2e)"},
                "hits_z": {"value": total_pairs * other["hits_z"] / (hz_ms * 1e-3), "unit": "pairs/s", "steps": other["hits_z"],
                           "d2h_bytes_per_step": int(lanes[0]["n_hits"]) * 12 + 8, "hits_per_step": int(lanes[0]["n_hits"]), "hits_equal_words": hits_ok,
                           "api": "Store.triangle_hits: -z 0.8, only the pairs that pass cross PCIe"}},
        "host": {"enqueue_us_per_step": 1e6 * host["enqueue_s"] / args.steps, "resolve_wait_us_per_step": 1e6 * host["resolve_s"] / args.steps,
                 "flush_us_per_step": 1e3 * flush_ms / args.steps},
    }
    for lane in lanes[1:]:
        lane["store"].close()
        lane["ctx"].close()
    store.close()
    del d_out
    return out


# --------------------------------------------------------------------------- batched: many sets, one launch
def leg_batched(env, engine, n_sets):
    torch, ctx, stream, dev = env.torch, env.ctx, env.stream, env.dev
    from ld_tools_b200 import Store
    from ld_tools_b200.synth import random_planes
    n_pairs = N_VARIANTS * (N_VARIANTS - 1) // 2
    mask_np = None
    stores = []
    for k in range(4):                                    # four distinct variant sets, cycled through the batch
        s = Store.from_planes(ctx, random_planes(N_VARIANTS, N_HAP, seed=700 + 10 * env.rank + k), N_HAP)
        mask_np = full_mask(s.stride_words)
        s.set_mask(mask_np)
        stores.append(s)
    rows = np.arange(N_VARIANTS, dtype=np.int64)
    outs = [torch.empty(n_pairs, dtype=torch.int32, device=dev) for _ in range(n_sets)]
    sets = [(stores[k % 4], rows, o.data_ptr()) for k, o in enumerate(outs)]
    for _ in range(3):
        ctx.triangle_batch_dev(sets, engine=engine)
        ctx.resolve()
    env.barrier()
    reps = 10
    ev = []
    for _ in range(reps):
        env.flush.zero_()
        e0, e1 = env.event(), env.event()
        e0.record(stream)
        ctx.triangle_batch_dev(sets, engine=engine)
        ctx.resolve()
        e1.record(stream)
        ev.append((e0, e1))
    env.barrier()
    call_ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    ctx.kernel_timing(True)
    for _ in range(reps):
        env.flush.zero_()
        ctx.triangle_batch_dev(sets, engine=engine)
        ctx.resolve()
    env.barrier()
    k_ms, k_n = ctx.kernel_timing(False)
    # parity: every set of the batch equals a single-set call on the same store
    single = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    ok = True
    for k in (0, 1, n_sets - 1):
        stores[k % 4].triangle_dev(rows, single.data_ptr(), engine=engine)
        ctx.resolve()
        ok &= bool((single == outs[k]).all())
    call_ms, kern_ms = env.max_over_ranks([call_ms, k_ms / max(k_n, 1)])
    pairs = n_sets * n_pairs
    out = {"workload": f"{n_sets} variant sets of {N_VARIANTS} variants x {N_HAP} haplotypes in one launch (ldx_triangle_batch_dev)",
           "value": pairs * env.world / (call_ms * 1e-3), "unit": "pairs/s", "call_ms": call_ms, "kernel_ms": kern_ms,
           "roofline_frac": pairs * OPS_PER_PAIR_I8 / (kern_ms * 1e-3) / 1e12 / env.peak_i8,
           "roofline_frac_whole_call": pairs * OPS_PER_PAIR_I8 / (call_ms * 1e-3) / 1e12 / env.peak_i8,
           "equals_single_set_calls": ok, "l2": "flushed between calls"}
    for s in stores:
        s.close()
    return out


# --------------------------------------------------------------------------- steady state: one large set
def leg_steady(env, engine, v_big=32768):
    torch, ctx, dev = env.torch, env.ctx, env.dev
    from ld_tools_b200 import Store
    from ld_tools_b200.synth import random_planes
    big = Store.from_planes(ctx, random_planes(v_big, N_HAP, seed=11 + env.rank), N_HAP)
    big.set_mask(full_mask(big.stride_words))
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
    del out_big
    big.close()
    return {"workload": f"ld_triangle, {v_big} variants x {N_HAP} haplotypes ({pairs_big} pairs), kernel only",
            "value": pps, "unit": "pairs/s", "kernel_ms": big_ms / big_n,
            "roofline_frac": pps * OPS_PER_PAIR_I8 / 1e12 / env.peak_i8}


# --------------------------------------------------------------------------- sharded: configs[3], row-range sharded triangle
def leg_sharded_triangle(env, v=100_000, reps=2, check=200_000):
    """BASELINE configs[3]: one 100,000-variant triangle, row ranges balanced by tile count (shard.triangle_row_ranges), one
    ldx_triangle_rows_dev call per rank, results left in HBM (20 GB / N).  No data-path collective.  Strong scaling."""
    torch, ctx, stream, dev = env.torch, env.ctx, env.stream, env.dev
    from ld_tools_b200 import Store, shard
    from ld_tools_b200.engine import ENGINE_AUTO
    from ld_tools_b200.synth import random_planes
    planes = random_planes(v, N_HAP, seed=4)              # the same variant set on every rank
    st = Store.from_planes(ctx, planes, N_HAP)
    st.set_mask(full_mask(st.stride_words))
    rows = np.arange(v, dtype=np.int64)
    begin, end = shard.triangle_row_ranges(v, env.world)[env.rank]
    n_mine = shard.tri(end) - shard.tri(begin)
    out = torch.empty(max(n_mine, 1), dtype=torch.int32, device=dev)

    def step():
        st.triangle_rows_dev(rows, begin, end, out.data_ptr(), engine=ENGINE_AUTO)
        ctx.resolve()
    step()
    times, mine = [], []
    for _ in range(reps):
        env.barrier()
        e0, e1 = env.event(), env.event()
        e0.record(stream)
        step()
        e1.record(stream)
        env.barrier()
        mine.append(float(e0.elapsed_time(e1)))
        times.append(env.max_over_ranks([mine[-1]])[0])
    # parity on a seeded sample of this rank's pairs: numpy popcounts + the oracle's finalisation
    from oracle import ld_oracle
    rng = np.random.default_rng(100 + env.rank)
    n_chk = min(check, n_mine)
    ok = True
    if n_chk:
        r = rng.integers(max(begin, 1), end, size=n_chk)
        c = (rng.random(n_chk) * r).astype(np.int64)
        idx = r * (r - 1) // 2 + c - shard.tri(begin)
        got = out[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint32)
        words = planes[:, : (N_HAP + 63) // 64]
        n1 = np.bitwise_count(words).sum(axis=1).astype(np.int32)
        n11 = np.bitwise_count(words[r] & words[c]).sum(axis=1).astype(np.int32)
        ok = bool((got == ld_oracle.packed_words(N_HAP, n11, n1[r], n1[c])).all())
    ok = env.min_flag(ok)
    per_rank = [0.0] * env.world
    per_rank[env.rank] = min(mine)
    t = torch.tensor(per_rank, dtype=torch.float64, device=dev)
    if env.world > 1:
        env.dist.all_reduce(t)
    best = min(times)
    total = v * (v - 1) // 2
    del out
    st.close()
    return {"workload": f"BASELINE configs[3]: ld_triangle large, {v} variants x {N_HAP} haplotypes, {total} pairs, row-range sharded",
            "scaling": "strong", "ms": best, "value": total / (best * 1e-3), "unit": "pairs/s", "ms_per_rank": [round(x, 3) for x in t.tolist()],
            "rows_per_rank": [list(x) for x in shard.triangle_row_ranges(v, env.world)],
            "roofline_frac_per_gpu": total * OPS_PER_PAIR_I8 / (best * 1e-3) / 1e12 / env.peak_i8 / env.world,
            "collective": "none on the data path (max-over-ranks of the device time only)",
            "sample_checked_per_rank": int(n_chk), "parity_sample_ok": ok}


# --------------------------------------------------------------------------- configs[2]: ld_area, one GPU and region-sharded
AREA_VARIANTS, AREA_QUERIES, AREA_FLANK, AREA_EUR_SAMPLES = 1_100_000, 1000, 500_000, 503
AREA_WORKLOAD = ("ld_area: 1,000 query variants, +/-500 kb flanks, r2 >= 0.8, EUR subset (503 samples = 1006 of 5008 "
                 "haplotypes), synthetic 1.1M-variant chr22 (BASELINE configs[2])")


def area_job(seed=77):
    """The configs[2] job, identical on every rank: positions, query rows, candidate ranges, the EUR haplotype columns."""
    from ld_tools_b200 import shard
    rng = np.random.default_rng(seed)
    nv = AREA_VARIANTS
    pos0 = np.sort(rng.integers(16_050_000, 51_200_000, size=nv)).astype(np.int32)
    hap = np.sort(rng.choice(N_HAP // 2, AREA_EUR_SAMPLES, replace=False))
    sel = np.sort(np.concatenate([2 * hap, 2 * hap + 1]))
    q_row = np.sort(rng.choice(nv, AREA_QUERIES, replace=False)).astype(np.int64)
    lo, hi, ws, we = shard.window_bounds(pos0, 1, pos0[q_row].astype(np.int64) + 1, AREA_FLANK)
    return {"nv": nv, "pos0": pos0, "sel": sel, "q_row": q_row, "lo": lo, "hi": hi, "ws": ws, "we": we}


def build_area_store(env, job, row_begin, row_end, subset=True):
    """Rows row_begin..row_end-1 of the synthetic chromosome (generated on the GPU, the same genome on every rank), annotated;
    the EUR columns gathered into a 128-byte-row subset store (ldx_store_subset) unless subset=False (mask on 640-byte rows)."""
    from ld_tools_b200 import Store
    from ld_tools_b200.synth import fill_store_grouped
    n = row_end - row_begin
    full = Store(env.ctx, n, N_HAP)
    fill_store_grouped(full, env.dev, 22, row_begin, row_end)
    pos0 = job["pos0"][row_begin:row_end]
    full.set_annotations(pos0, pos0 + 1, np.arange(row_begin, row_end, dtype=np.int64), np.ones(n, np.uint8))
    full.select_haplotypes(job["sel"])
    if not subset:
        return full
    sub = full.subset(job["sel"])
    full.close()
    return sub


def check_area_queries(st, hits, q_row, lo, hi, thres, n_check, rng, n_hap_sel):
    """Whole windows of a few queries recomputed on the host (numpy popcounts + the oracle's finalisation)."""
    from oracle import ld_oracle
    from ld_tools_b200._lib import BELOW_THRES, R2_MASK
    words = (n_hap_sel + 63) // 64
    ok = True
    for _ in range(n_check):
        k = int(rng.integers(len(q_row)))
        a, b, q = int(lo[k]), int(hi[k]), int(q_row[k])
        win = st.download(a, b - a)[:, :words]
        qr = st.download(q, 1)[0, :words]
        n1 = np.bitwise_count(win).sum(axis=1).astype(np.int32)
        n11 = np.bitwise_count(win & qr[None, :]).sum(axis=1).astype(np.int32)
        n1q = np.full(b - a, int(np.bitwise_count(qr).sum()), dtype=np.int32)
        want_w = ld_oracle.packed_words(n_hap_sel, n11, n1q, n1)             # var_1 = query, var_2 = window row (ld_area.py:242)
        keep = ((want_w & R2_MASK) >= thres) & (np.arange(a, b) != q)          # rounded measure >= thres (:248), not the query (:222)
        got = hits[hits["query"] == k]
        got = got[np.argsort(got["row"])]
        ok &= bool(got["row"].tolist() == (np.flatnonzero(keep) + a).tolist() and (got["n11"] == n11[keep]).all()
                   and ((got["packed"] & ~np.uint32(BELOW_THRES)) == want_w[keep]).all())
    return ok


def leg_area(env, subset=True, steps=5, warmup=3):
    """configs[2] on one GPU per rank: the window scan kernel-level (value, roofline) and through the host API (e2e)."""
    torch, ctx, stream, dev = env.torch, env.ctx, env.stream, env.dev
    from ld_tools_b200.engine import threshold_e4
    job = area_job()
    st = build_area_store(env, job, 0, job["nv"], subset=subset)
    q_row, lo, hi, ws, we = job["q_row"], job["lo"], job["hi"], job["ws"], job["we"]
    thres = threshold_e4(0.8)
    cap = 1 << 22
    d_hits = torch.empty(cap * 4, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(2, dtype=torch.int64, device=dev)

    def step():
        st.window_dev(q_row, lo, hi, ws, we, "r_square", thres, d_hits.data_ptr(), cap, d_cnt.data_ptr())
        ctx.resolve()
    for _ in range(warmup):
        step()
    env.barrier()
    launches0 = ctx.launch_count
    ctx.kernel_timing(True)
    ev = []
    for _ in range(steps):
        env.flush.zero_()
        e0, e1 = env.event(), env.event()
        e0.record(stream)
        step()
        e1.record(stream)
        ev.append((e0, e1))
    env.barrier()
    launches = ctx.launch_count - launches0
    dom_ms, dom_n = ctx.kernel_timing(False)
    step_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    scanned, n_hits = int(d_cnt.cpu()[1]), int(d_cnt.cpu()[0])
    for _ in range(2):
        st.window(q_row, lo, hi, ws, we, "r_square", thres, cap=cap)
    env.barrier()
    e_steps = max(steps, 3)
    e0, e1 = env.event(), env.event()
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(e_steps):
        hits, _ = st.window(q_row, lo, hi, ws, we, "r_square", thres, cap=cap)
    e1.record(stream)
    env.barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = float(e0.elapsed_time(e1))
    ok = check_area_queries(st, hits, q_row, lo, hi, thres, 3, np.random.default_rng(5 + env.rank), st.n_hap if subset else N_HAP) if subset else None
    step_ms, e2e_ms, dom_ms = env.max_over_ranks([step_ms, e2e_ms, dom_ms])
    kern_s = dom_ms * 1e-3 / max(dom_n, 1)
    row_bytes = st.stride_words * 8
    # the kernel counts every (candidate row, query) pair once; windows overlap ~28-fold, rows are re-used from L1 / registers:
    # the bound is the integer pipes (AND + carry-save popcount), not HBM.  SURVEY 8d's POPC roofline: 16 POPC32 / clk / SM.
    popc32_per_pair = (st.n_hap + 31) // 32
    popc_peak_pairs = 148 * 16 * env.peaks["sm_max_mhz"] * 1e6 / popc32_per_pair
    pps = scanned / kern_s
    roof = {"bound": "int-alu/popc", "achieved": pps / 1e9, "peak": popc_peak_pairs / 1e9, "unit": "Gpairs/s",
            "frac": pps / popc_peak_pairs, "peak_source": f"148 SM x 16 POPC32/clk x {env.peaks['sm_max_mhz']:.0f} MHz / {popc32_per_pair} POPC32 per pair "
            "(SURVEY 8d; the carry-save adders move half (128-byte rows) to two thirds (640-byte rows) of the counting to the 64-lane ALU pipe, so frac may exceed 1)",
            "kernel": "window_rows1_kernel (+ mq_extend_kernel, inside the timed pair)" if row_bytes in (128, 256) else "window_mq_kernel (+ mq_extend_kernel)", "kernel_ms": kern_s * 1e3, "kernel_launches_timed": int(dom_n),
            "hbm_frac_if_every_pair_read_its_row": scanned * row_bytes / kern_s / 1e9 / env.peaks["hbm_gbs"],
            "hbm_frac_one_pass_over_the_store": st.n_variants * row_bytes / kern_s / 1e9 / env.peaks["hbm_gbs"],
            "traffic": ncu_traffic(["r02_ncu_full_window_rows1_configs2.txt"] if subset else ["r01_ncu_full_window_mq_configs2.txt"]),
            "algorithmic_per_launch": f"{scanned} pairs x {popc32_per_pair} POPC32 ({row_bytes}-byte rows)"}
    out = {"workload": AREA_WORKLOAD, "store": f"{'subset store' if subset else 'mask on the full store'}: {st.n_hap} haplotype columns, {row_bytes}-byte rows, "
           f"{st.n_variants * row_bytes / 1e6:.0f} MB", "value": scanned * env.world * steps / (step_ms * 1e-3), "unit": "pairs/s",
           "ms_per_step": step_ms / steps, "pairs_per_step_per_gpu": scanned, "hits_per_step": n_hits, "gpu_launches": int(launches),
           "e2e": {"value": scanned * env.world * e_steps / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms / e_steps, "wall_s": e2e_wall,
                   "h2d_bytes_per_step": int(q_row.nbytes + lo.nbytes + hi.nbytes + ws.nbytes + we.nbytes + 8 * (len(q_row) + 1)),
                   "d2h_bytes_per_step": int(16 * len(hits) + 16)},
           "roofline": roof, "parity_windows_ok": ok}
    del d_hits
    st.close()
    return out


# GRCh38 autosome lengths (Mb), chr1..chr22
CHR_MB = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09,
          133.28, 114.36, 107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82]


def leg_sharded_genome(env, n_variants=80_000_000, n_queries=50_000, flank=1_000_000, reps=3, check=3):
    """BASELINE configs[4] (north_star (3)): genome-scale ld_area, region-sharded.  The genome-wide query list, ordered by
    (chromosome, position), is cut into `world` pieces with equal candidate-pair counts (shard.genome_pieces); a rank holds,
    per chromosome it touches, only the rows its queries' windows reach (slab + halo) and makes one ldx_window_dev call per
    piece.  The kept pairs are re-numbered job-wide ON THE DEVICE, gathered with NCCL (shard.gather_hits_tensor: counts,
    padded records) and sorted by (query, row) on the device -- all inside the timed region.  Strong scaling."""
    torch, ctx, stream, dev = env.torch, env.ctx, env.stream, env.dev
    from ld_tools_b200 import Store, shard
    from ld_tools_b200._lib import BELOW_THRES, HIT_DTYPE
    from ld_tools_b200.engine import threshold_e4
    from ld_tools_b200.synth import fill_store_grouped
    thres = threshold_e4(0.8)
    tot_mb = sum(CHR_MB)
    chroms = []
    for c, mb in enumerate(CHR_MB):              # the job, identically on every rank
        nv = int(round(n_variants * mb / tot_mb))
        nq = max(1, int(round(n_queries * mb / tot_mb)))
        rng = np.random.default_rng(9000 + c)
        pos0 = np.sort(rng.integers(10_000, int(mb * 1e6), size=nv, dtype=np.int64)).astype(np.int32)
        q_row = np.sort(rng.choice(nv, nq, replace=False)).astype(np.int64)
        lo, hi, ws, we = shard.window_bounds(pos0, 1, pos0[q_row].astype(np.int64) + 1, flank)
        chroms.append({"nv": nv, "pos0": pos0, "q_row": q_row, "lo": lo, "hi": hi, "ws": ws, "we": we})
    q_first = np.concatenate([[0], np.cumsum([len(ch["q_row"]) for ch in chroms])])
    t_build = time.perf_counter()
    pieces, store_bytes, cap_total = [], 0, 0
    for pc in shard.genome_pieces(chroms, env.world)[env.rank]:
        c, qa, qb, rb, re = pc["chrom"], pc["qa"], pc["qb"], pc["row_begin"], pc["row_end"]
        ch = chroms[c]
        st = Store(ctx, re - rb, N_HAP)
        fill_store_grouped(st, dev, c, rb, re)
        pos0 = ch["pos0"][rb:re]
        st.set_annotations(pos0, pos0 + 1, (np.int64(c) << 32) + np.arange(rb, re, dtype=np.int64), np.ones(re - rb, np.uint8))
        st.set_mask(full_mask(st.stride_words))
        cap = 64 * (qb - qa) + 65536
        pieces.append({"chrom": c, "st": st, "rb": rb, "qa": qa, "qb": qb, "cap": cap, "off": cap_total,
                       "q": ch["q_row"][qa:qb] - rb, "lo": ch["lo"][qa:qb] - rb, "hi": ch["hi"][qa:qb] - rb, "ws": ch["ws"][qa:qb], "we": ch["we"][qa:qb]})
        cap_total += cap
        store_bytes += (re - rb) * st.stride_words * 8
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build
    d_hits = torch.empty((max(cap_total, 1), 4), dtype=torch.int32, device=dev)
    d_cnt = torch.zeros((max(len(pieces), 1), 2), dtype=torch.int64, device=dev)

    def job():
        for k, p in enumerate(pieces):
            p["st"].window_dev(p["q"], p["lo"], p["hi"], p["ws"], p["we"], "r_square", thres, d_hits[p["off"]:].data_ptr(), p["cap"], d_cnt[k].data_ptr())
        ctx.resolve()                                            # near-tie pairs settled (hits the exact rounding rejects get BELOW_THRES)
        g0 = env.event()
        g0.record(stream)
        counts = d_cnt[:, 0].tolist() if pieces else []
        parts = []
        for k, p in enumerate(pieces):                           # job-wide numbering, on the device
            h = d_hits[p["off"]:p["off"] + min(int(counts[k]), p["cap"])]
            h = h[(h[:, 3] & BELOW_THRES) == 0]
            h = h + torch.tensor([p["qa"] + int(q_first[p["chrom"]]), p["rb"], 0, 0], dtype=torch.int32, device=dev)
            parts.append(h)
        mine = torch.cat(parts) if parts else torch.zeros((0, 4), dtype=torch.int32, device=dev)
        allh = shard.gather_hits_tensor(mine, ranks_own_ordered_query_ranges=True)     # genome_pieces: contiguous query ranges in rank order
        return mine, allh, g0
    job()
    times, gathers = [], []
    for _ in range(reps):
        env.barrier()
        e0, e1 = env.event(), env.event()
        e0.record(stream)
        mine, allh, g0 = job()
        e1.record(stream)
        env.barrier()
        m = env.max_over_ranks([float(e0.elapsed_time(e1)), float(g0.elapsed_time(e1))])
        times.append(m[0]); gathers.append(m[1])
    scanned = int(d_cnt[:, 1].sum().item()) if pieces else 0
    overflow = sum(max(0, int(n) - p["cap"]) for n, p in zip(d_cnt[:, 0].tolist(), pieces)) if pieces else 0
    # parity: whole windows of a few of this rank's queries against numpy popcounts + the oracle's finalisation
    ok = overflow == 0
    mine_np = mine.cpu().numpy().reshape(-1).view(HIT_DTYPE)
    rng = np.random.default_rng(500 + env.rank)
    for _ in range(check if pieces else 0):
        p = pieces[int(rng.integers(len(pieces)))]
        local = mine_np[(mine_np["query"] >= p["qa"] + q_first[p["chrom"]]) & (mine_np["query"] < p["qb"] + q_first[p["chrom"]])].copy()
        local["query"] -= p["qa"] + q_first[p["chrom"]]
        local["row"] -= p["rb"]
        ok &= check_area_queries(p["st"], local, p["q"], p["lo"], p["hi"], thres, 1, rng, N_HAP)
    ok = env.min_flag(ok)
    tot = torch.tensor([scanned, store_bytes], dtype=torch.int64, device=dev)
    per_rank = torch.zeros(env.world, dtype=torch.int64, device=dev)
    per_rank[env.rank] = scanned
    if env.world > 1:
        env.dist.all_reduce(tot)
        env.dist.all_reduce(per_rank)
    best = min(times)
    out = {"workload": f"BASELINE configs[4]: genome-scale ld_area, {sum(len(ch['q_row']) for ch in chroms)} queries over synthetic chr1-22 "
                       f"({sum(ch['nv'] for ch in chroms)} variants x {N_HAP} haplotypes), +/-{flank} bp, r2 >= 0.8, region-sharded with halos",
           "scaling": "strong", "ms": best, "ms_all": [round(x, 3) for x in times], "value": int(tot[0].item()) / (best * 1e-3), "unit": "pairs/s",
           "pairs_scanned": int(tot[0].item()), "pairs_per_rank": per_rank.tolist(), "kept_pairs": int(allh.shape[0]),
           "gather_ms": min(gathers), "store_GB_total": round(int(tot[1].item()) / 1e9, 2), "build_s": build_s,
           "collective": "device sort of the rank's own records by (query, row), then NCCL all_gather of the kept-pair counts and of the padded 16-byte records "
                         "(shard.gather_hits_tensor; ranks own ordered query ranges, so the concatenation is the sorted list); "
                         "inside the timed region, reported separately as gather_ms (renumbering + collective + sort)",
           "timed": "CUDA events on the launching stream around window scans + ldx_resolve + renumbering + gather + sort, max over ranks",
           "parity_windows_ok": ok}
    for p in pieces:
        p["st"].close()
    del d_hits
    return out


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    from ld_tools_b200.engine import ENGINE_AUTO, ENGINE_MMA, ENGINE_POPC
    env = Env(args, rank, world, local_rank)
    engine = {"auto": ENGINE_AUTO, "popc": ENGINE_POPC, "mma": ENGINE_MMA}[args.engine]
    tri = leg_triangle(env, engine)
    used_mma = tri["used_mma"]
    legs = {}
    if used_mma and not args.no_batched:
        legs["batched"] = leg_batched(env, engine, args.batch_sets)
    if used_mma and not args.no_steady:
        legs["steady_state"] = leg_steady(env, engine)
    if not args.no_area:
        legs["ld_area"] = leg_area(env)
    if not args.no_sharded:
        legs["sharded"] = {"n_gpus": world, "triangle_configs3": leg_sharded_triangle(env), "ld_area_genome_configs4": leg_sharded_genome(env)}
    line = {"metric": METRIC, "value": tri["value"], "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tri["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 tensor (exact) + f32 screen / f64 settle" if used_mma else "u64 popcount + f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": tri["n_pairs"], "engine": args.engine,
                       "l2": "flushed (256 MiB write) between timed iterations; flush time excluded",
                       "calls_in_flight": tri["depth"], "sharding": "value: one variant set per GPU (replicas); the north_star's partitioning: see `sharded`"},
            "e2e": tri["e2e"], "gpu_launches": tri["launches"], "clocks": tri["clocks"], "roofline": tri["roofline"],
            "parity_selfcheck": tri["parity_selfcheck"], "back_to_back": tri["back_to_back"], "host": tri["host"]}
    line.update(legs)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_pairs_per_core)
        print(json.dumps(line), flush=True)
    env.close()


def run_area(args, rank, world, local_rank):
    """--workload ld_area: configs[2] as the headline line (one chromosome per GPU: weak scaling)."""
    global AREA_EUR_SAMPLES
    if args.area_samples:
        AREA_EUR_SAMPLES = args.area_samples
    env = Env(args, rank, world, local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()
    a = leg_area(env, subset=not args.area_full_store, steps=args.steps, warmup=args.warmup)
    clocks = sampler.stop()
    line = {"metric": METRIC, "value": a["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": a["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 popcount + f64",
            "data": "synthetic",
            "config": {"workload": a["workload"], "store": a["store"], "pairs_per_step_per_gpu": a["pairs_per_step_per_gpu"], "hits_per_step": a["hits_per_step"],
                       "l2": "store larger than L2 (full store) or flushed by a 256 MiB write between timed iterations", "sharding": "one chromosome store per GPU"},
            "e2e": a["e2e"], "gpu_launches": a["gpu_launches"], "clocks": clocks, "roofline": a["roofline"], "parity_windows_ok": a["parity_windows_ok"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["ld_triangle", "ld_area"], default="ld_triangle",
                    help="ld_triangle = BASELINE configs[1] (the headline); ld_area = configs[2], the window scan")
    ap.add_argument("--engine", choices=["auto", "popc", "mma"], default="auto")
    ap.add_argument("--tile-n", type=int, default=0, help="tcgen05 tile width override (0 = heuristic)")
    ap.add_argument("--cpu-pairs-per-core", type=int, default=2000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline", type=int, default=32, help="device-resident calls in flight per ldx_resolve()")
    ap.add_argument("--e2e-contexts", type=int, default=4,
                    help="independent contexts (stream + store + host thread) the end-to-end leg pipelines its steps over")
    ap.add_argument("--e2e-steps", type=int, default=240, help="steps of the end-to-end leg (all contexts together)")
    ap.add_argument("--batch-sets", type=int, default=16, help="variant sets per launch of the batched leg")
    ap.add_argument("--sampler-ms", type=float, default=5.0, help="period of the NVML clock sampler thread")
    ap.add_argument("--no-steady", action="store_true", help="skip the 32,768-variant steady-state leg")
    ap.add_argument("--no-batched", action="store_true", help="skip the batched leg")
    ap.add_argument("--no-area", action="store_true", help="skip the ld_area (configs[2]) leg")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded legs (configs[3] row ranges, configs[4] regions)")
    ap.add_argument("--area-full-store", action="store_true", help="--workload ld_area: scan the 640-byte rows under a mask instead of the subset store")
    ap.add_argument("--area-samples", type=int, default=0, help="--workload ld_area: samples in the selection instead of the 503 of EUR (e.g. 661 = AFR: 256-byte rows)")
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
