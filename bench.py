#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric : 2D SWE cell-updates/s  (interior cells x full RK2 steps / second; one step = perform_step!:
         2 stages incl. friction, dry clean-up, ghost-cell BC and the CFL reduction)
config : configs[4] -- 16384 x 16384 Float32 synthetic bathymetry, WallBC, minmod(1.2), bottom source,
         bh2021 friction, RK2, cfl 0.25; at N > 1 the SAME grid is cut into y-strips (strong scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one rank per GPU)
  python bench.py --impl reference [...]                          the reference algorithm on the host CPU
                                                                  (oracle restatement, all host threads)
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def package():
    """the product package (its directory name contains a dot: imported through importlib, see __graft_entry__)"""
    from __graft_entry__ import _load_package
    return _load_package()

METRIC = "2D SWE cell-updates/s"
UNIT = "cell-updates/s"
WORDS_STAGE1, WORDS_STAGE2 = 7, 10        # algorithmic words per cell: K1 reads U(3)+B(1) writes dU(3);
                                          # K2 reads U(3)+dU(3)+B(1) writes U'(3)   


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons / power DURING the timed region (B200_PROFILING.md recipe).  NVML is polled
    from a thread every ~2 ms (the timed region at N = 8 is only ~15 ms, shorter than one nvidia-smi -lms period);
    nvidia-smi is the fall-back when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.nvml, self.samples = index, None, [], None, []
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)),
                                     n.nvmlDeviceGetPowerUsage(self.h) / 1000.0, int(reasons_fn(self.h))))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.th.join(timeout=1.0)
            sm = [s[0] for s in self.samples]
            reasons = set()
            for s in self.samples:
                for bit, nm in self.BITS.items():
                    if s[2] & bit:
                        reasons.add(nm)
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.mx,
                    "sm_min_mhz": min(sm) if sm else None,
                    "power_w_max": max((s[1] for s in self.samples), default=None), "samples": len(sm),
                    "reasons": sorted(reasons), "source": "nvml, 2 ms polling during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi -lms 100"}


# BASELINE.json configs[1..4] (SURVEY 8d "Config 2..5"): (builder, default n, dtype, algorithmic words per cell of
# K1 / K2 in the state's element size, description).  Config 5 is the bench line; 2-4 are side measurements.
CONFIGS = {
    2: ("swashes", 1024, "float64", 6, 9,
        "configs[1]: 2D SWASHES 4.1.2 dam break on a dry bed, WallBC, flat bottom (no bathymetry words), "
        "LinearReconstruction(1.2), RungeKutta2, cfl 0.1, depth_cutoff = kappa = 1e-7"),
    3: ("terrain", 4096, "float32", 9, 12,
        "configs[2]: terrain.jl-shaped synthetic terrain, lake at level 1, bottom source, ConstantRain, Horton infiltration "
        "(Float64 factor plane = 2 words), ImplicitFriction(bsa2012), LinearReconstruction(1.2), RungeKutta2, cfl 0.25"),
    4: ("urban", 8192, "float32", 7, 10,
        "configs[3]: urban.jl-style synthetic DEM with buildings, lake at level 1, NoReconstruction, rain + constant "
        "infiltration, bottom source, ImplicitFriction(bsa2012), RungeKutta2, cfl 0.1, depth_cutoff = kappa = 8e-2"),
    5: ("lake", 16384, "float32", 7, 10, None),
}


def build_case(n: int, dtype: str, ny=None, config: int = 5):
    W = package().workloads
    kind = CONFIGS[config][0]
    if kind == "swashes":
        return W.swashes_dambreak_2d(n=n, dtype=dtype)
    if kind == "terrain":
        return W.terrain_case(n=n, dtype=dtype, level=1.0, bc=BC)
    if kind == "urban":
        return W.urban_case(n=n, dtype=dtype, level=1.0, bc=BC)
    case, state = W.lake_case(n=n, ny=ny, dtype=dtype, level=LAKE_LEVEL)
    case["bc"] = BC
    return case, state


BC = "wall"          # --bc: "neumann" = the BC of the reference's own terrain / urban examples (examples/terrain.jl:85)
LAKE_LEVEL = 0.0     # --level: still-water level of the config-5 lake (0 = the bench line, 54 % wet; 60 = every cell wet)


def wet_fraction(case, state):
    if not isinstance(case.get("B"), np.ndarray):
        return float(np.mean(state[0].astype(np.float64) >= case["depth_cutoff"]))
    Bc = package().synthetic.cell_average(np.asarray(case["B"]))[2:-2, 2:-2]
    return float(np.mean(state[0].astype(np.float64) - Bc >= case["depth_cutoff"]))


def wet_fraction_rows(case, state):
    """wet fraction of every interior row of a strip (input of the strip balancing cost model)"""
    Bc = package().synthetic.cell_average(np.asarray(case["B"]))[2:-2, 2:-2]
    return np.mean(state[0].astype(np.float64) - Bc >= case["depth_cutoff"], axis=1)


def run_reference(args, rank):
    """The reference's CPU algorithm (oracle restatement: the Julia package cannot run here) on the
    host cores.  Every step is a bounded sample of the workload: config 5 -> a FULL-WIDTH band of `--cpu-rows` rows
    from the middle of the same 16384-wide domain (same row length, same generator, streaming working set far beyond
    the CPU caches: 16 words per cell), other configs -> an `--cpu-n`^2 instance of the same generator."""
    if rank != 0:
        return
    from oracle import oracle
    threads = oracle.set_threads(os.cpu_count() or 1)
    if args.config == 5:
        n, rows = args.n, min(args.cpu_rows, args.n)
        case, state, y0 = package().workloads.lake_strip_case(n, 0, 1, dtype=args.dtype, level=LAKE_LEVEL,
                                                              rows=((args.n - rows) // 2, rows))
        case["extent"] = [0.0, float(n), 0.0, float(rows)]          # a stand-alone domain with the workload's dx = dy = 1
        shape = [n, rows]
        what = f"full-width band: rows {y0}..{y0 + rows} of the {n}x{n} domain"
    else:
        n = args.cpu_n
        case, state = build_case(n, args.dtype, config=args.config)
        shape = [n, n]
        what = f"{n}x{n} instance of the same generator"
    o = oracle.OracleSim(case)
    o.set_state_interior(*state)
    mts = case.get("maximum_timestep") or float("inf")
    for _ in range(args.warmup):
        o.step(mts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.step(mts)
    dt = time.perf_counter() - t0
    cells = shape[0] * shape[1]
    value = cells * args.steps / dt
    sample = (f"{what} ({shape[0]}x{shape[1]} cells per step), {args.steps} RK2 steps, {threads} OpenMP threads; "
              "C++ restatement of the reference algorithm (oracle/), not the Julia package")
    cfg = workload_config(args, args.n, wet_fraction(case, state))
    cfg["reference_sample_grid"] = shape      # the grid this arm actually steps; `grid` is the workload it samples
    cfg["parallelism"] = f"{threads} host threads (OpenMP)"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def workload_config(args, n, wet):
    if args.config != 5:
        return {"workload": f"{CONFIGS[args.config][5].replace('WallBC', BC + ' BC') if BC != 'wall' else CONFIGS[args.config][5]}"
                            f"{'' if BC == 'wall' or 'BC' in CONFIGS[args.config][5] else ', ' + BC + ' BC'}; {n}x{n} {args.dtype}, {wet:.0%} of cells wet",
                "grid": [n, n], "arithmetic": "fast" if not args.strict else "strict", "parallelism": "single GPU",
                "l2": "3 volumes of %.0f MB" % (3 * 3 * (n + 4) ** 2 * np.dtype(args.dtype).itemsize / 1e6)
                      + (" fit the 126 MB L2: an L2 flush (256 MB memset) runs between timed steps" if args.flush_l2 else "")}
    return {"workload": f"configs[4]: {n}x{n} {args.dtype} synthetic bathymetry (separable 5-octave hills, seed 1234), "
                        f"lake at level {LAKE_LEVEL:g} + Gaussian hump ({wet:.0%} of cells wet), WallBC, LinearReconstruction(1.2), "
                        "CentralUpwind, SourceTermBottom, ImplicitFriction(bh2021), RungeKutta2, cfl 0.25, "
                        "depth_cutoff = kappa = 1e-3 (SURVEY 8d sketched 1e-5 and lake level 5 on a 40 m terrain; this "
                        "generator's hills are +-8 m, and 1e-5 is 10 Float32 ulps of w: workloads.py)",
            "grid": [n, n], "arithmetic": "fast" if not args.strict else "strict",
            "parallelism": (f"y-strips x{args.gpus}; halo rows + CFL maxima by peer stores over NVLink from the library's "
                            "own kernels (torch.distributed NCCL group only for set-up and barriers)") if args.gpus > 1 else "single GPU",
            "l2": "working set (state 3 x 3 planes + B, > 10 GB at 16384^2) far exceeds the 126 MB L2; no flush needed"}


def run_ours(args, rank, world):
    import torch
    P = package()
    W = P.workloads
    n = args.n
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t_setup = time.perf_counter()
    balance = None
    strip = None

    def gather_rows(vec, sizes):
        """all-gather of per-row vectors of unequal length -> one global per-row vector"""
        buf = torch.zeros(max(sizes), dtype=torch.float64, device="cuda")
        buf[:len(vec)] = torch.from_numpy(np.asarray(vec, dtype=np.float64)).cuda()
        allv = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(allv, buf)
        return np.concatenate([allv[r][:sizes[r]].cpu().numpy() for r in range(world)])

    def make_strip(y0, rows):
        case, state, y0 = W.lake_strip_case(n, rank, world, dtype=args.dtype, level=LAKE_LEVEL, rows=(y0, rows))
        pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in state]
        s = P.multigpu.StripSimulator(case, [p.numpy() for p in pinned], rank, world, n, y0, P.multigpu.torch_all_gather,
                                      device=local, strict=args.strict)
        return case, state, pinned, s

    if world > 1:
        offs, sizes = P.multigpu.partition_rows(n, world)
        case, state, y0 = W.lake_strip_case(n, rank, world, dtype=args.dtype, level=LAKE_LEVEL)
        if not args.equal_strips:
            # Cost-balanced strips: every rank waits for the slowest one at the CFL all-gather, so the strips are cut
            # by work, not by rows.  Pass 1: a model (a dry cell costs `dry_cost` of a wet one) on the per-row wet
            # fractions.  Pass 2: the model is rescaled per rank by the MEASURED stage-kernel time of a short run on the
            # pass-1 strips (clock / power differences between the GPUs and everything the model misses), and the rows
            # are cut again.  All ranks derive the same cuts from the same all-gathered vectors.
            model = P.multigpu.row_cost_from_wet_fraction(gather_rows(wet_fraction_rows(case, state), sizes), args.dry_cost)
            offs, sizes = P.multigpu.partition_rows(n, world, model)
            balance = {"dry_cost": args.dry_cost, "rows_model": list(sizes)}
            if args.rebalance:
                case, state, pinned, strip = make_strip(offs[rank], sizes[rank])
                strip.sim.run_steps(3)
                tm0 = strip.sim.run_steps_timed(args.rebalance)
                mine = torch.tensor([tm0["stage_kernel_ms"]], dtype=torch.float64, device="cuda")
                allt = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(allt, mine)
                measured = np.array([float(x.item()) for x in allt])
                new_offs, new_sizes = P.multigpu.partition_rows(
                    n, world, P.multigpu.rescale_cost_by_measured(model, offs, sizes, measured))
                balance.update(rows_measured=list(new_sizes),
                               stage_ms_before=[round(float(m) / args.rebalance, 4) for m in measured])
                if list(new_sizes) != list(sizes):
                    strip.close()               # collective: unmap the peers, barrier, destroy
                    del strip, pinned
                    strip = None
                    offs, sizes = new_offs, new_sizes
            balance["rows_per_rank"] = list(sizes)
        if strip is None:
            case, state, pinned, strip = make_strip(offs[rank], sizes[rank])
        y0, rows = offs[rank], sizes[rank]
        sim = strip.sim
    else:
        case, state = build_case(n, args.dtype, ny=args.ny, config=args.config)
        y0, rows = 0, (args.ny or n)
        pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in state]
    wet = wet_fraction(case, state)
    # pinned host buffers for the end-to-end leg
    host_in = [p.numpy() for p in pinned]
    out_pinned = [torch.empty_like(p).pin_memory() for p in pinned]
    if world > 1:
        w = torch.tensor([wet * rows], dtype=torch.float64, device="cuda")
        dist.all_reduce(w)
        wet = float(w.item()) / n
    else:
        sim = P.simulator_from_case(case, strict=args.strict, fused=not args.unfused, device=local)
        sim.set_state_interior(*host_in)
    family = sim.kernel_family()
    t_setup = time.perf_counter() - t_setup

    def barrier():
        sim.sync()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    # ---- device-resident leg: inputs already in HBM when the timed region starts -------------------
    sim.run_steps(args.warmup, case.get("maximum_timestep"))
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    mts = case.get("maximum_timestep")
    if args.flush_l2 and world == 1:
        # the working set fits the 126 MB L2: write a 256 MB buffer between timed steps, time every step on its own
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        tm = None
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            one = sim.run_steps_timed(1, mts)
            tm = one if tm is None else {k: tm[k] + one[k] for k in tm}
    else:
        tm = sim.run_steps_timed(args.steps, mts)       # CUDA events on the library's stream around the K steps
    barrier()
    clocks = sampler.stop()
    cells = n * (args.ny or n) if world == 1 else n * n
    sec = max_over_ranks(tm["total_ms"]) * 1e-3
    value = cells * args.steps / sec

    # ---- end-to-end leg through the public API with HOST buffers -----------------------------------
    # set_current_state!(host state) -> K x perform_step! -> collect(current_interior_state):
    # the H2D upload of the inputs and the D2H read of the result are inside the timed region.
    L = P.api._lib.lib()
    import ctypes as C
    ptr = lambda t: C.c_void_p(t.data_ptr())
    barrier()
    t0 = time.perf_counter()
    sim._check(L.swe_b200_set_state_interior(sim._h, *[ptr(p) for p in pinned]))
    sim.run_steps(args.steps, case.get("maximum_timestep"))
    sim._check(L.swe_b200_get_state_interior(sim._h, *[ptr(p) for p in out_pinned]))
    e2e_sec = max_over_ranks(time.perf_counter() - t0)
    itemsize = np.dtype(args.dtype).itemsize
    io_bytes = 3 * cells * itemsize
    assert np.isfinite(out_pinned[0].numpy()).all()
    launches = sum_over_ranks(float(tm["launches"]))
    cells_local = n * rows
    per_rank = None
    if dist is not None:      # per-strip kernel times: shows the load imbalance behind the strong-scaling number
        ns = max(tm["stage_launches"] // 2, 1)
        mine = torch.tensor([tm["stage1_ms"] / ns, tm["stage2_ms"] / ns, tm["total_ms"] / args.steps, float(rows)],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"k1_ms": round(float(t[0]), 4), "k2_ms": round(float(t[1]), 4), "step_ms": round(float(t[2]), 4),
                     "rows": int(t[3])} for t in allr]
    if rank != 0:
        barrier()
        if dist is not None:
            strip.close()
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    W1, W2 = CONFIGS[args.config][3], CONFIGS[args.config][4]
    ns2 = max(tm["stage_launches"] // 2, 1)
    k2_ms = tm["stage2_ms"] / ns2 or float("nan")      # (per-operator path: no fused stage kernels to time)
    k1_ms = tm["stage1_ms"] / ns2 or float("nan")
    achieved = W2 * itemsize * cells_local / (k2_ms * 1e-3) / 1e9
    # dram__bytes_read + write of ONE K2 launch from an `ncu --set full` capture of exactly this launch shape
    # (profiles/traffic.json, keyed by dtype, grid and config); null when no capture exists for it (N > 1 strips, --ny,
    # --level, other sizes): a number measured on another launch would not describe this one
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and world == 1 and not args.ny and LAKE_LEVEL == 0.0 and BC == "wall" and not args.strict and not args.unfused:
        try:
            traffic = json.loads(tf.read_text()).get(f"k_stage2_{args.dtype}_{n}" + ("" if args.config == 5 else f"_config{args.config}"))
        except Exception:
            traffic = None
    kname = {"fused-x2": "x2::k_stage_x2<" + ("RECON=1, HASB=1, STAGE=2, FRIC=1" if args.config == 5 else "..., STAGE=2") + "> (K2: RK2 stage 2, fused, packed Float32)",
             "fused-d1": "d1::k_stage_d1<..., STAGE=2> (K2: RK2 stage 2, fused, Float64 rings)",
             "fused-generic": "k_stage<..., STAGE=2> (K2, generic fused kernel)"}.get(family, family)
    roofline = {"bound": "hbm", "kernel": kname, "kernel_family": family, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": W2 * itemsize * cells_local,
                "avg_launch_ms": k2_ms, "per": "GPU (rank 0's strip)" if world > 1 else "GPU",
                "k1": {"achieved": W1 * itemsize * cells_local / (k1_ms * 1e-3) / 1e9, "avg_launch_ms": k1_ms},
                "step": {"achieved": (W1 + W2) * itemsize * cells_local * args.steps / sec / 1e9,
                         "frac": (W1 + W2) * itemsize * cells_local * args.steps / sec / 1e9 / peak}}

    # ---- CPU baseline on this box's host cores: bounded sample of the same workload ----------------
    cpu = None
    if not args.no_cpu and world == 1:
        from oracle import oracle
        oracle.set_threads(1)
        cn = args.cpu_n
        ccase, cstate = build_case(cn, args.dtype, config=args.config)     # (config 5 at --level, like the GPU arm)
        o = oracle.OracleSim(ccase)
        o.set_state_interior(*cstate)
        o.step()
        csteps = args.cpu_steps      # ~10 s of scalar CPU work at 1024^2 (about 3 M cell-updates/s per core)
        t0 = time.perf_counter()
        for _ in range(csteps):
            o.step()
        cdt = time.perf_counter() - t0
        cpu = {"value": cn * cn * csteps / cdt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{cn}x{cn} instance of the same generator, {csteps} RK2 steps, scalar oracle restatement "
                         "(the Julia reference cannot run here: no julia binary)"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tm["total_ms"] / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
        "config": workload_config(args, n, wet),
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "e2e": {"value": cells * args.steps / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": io_bytes / args.steps,
                "d2h_bytes_per_step": io_bytes / args.steps,
                "pattern": f"set_current_state!(pinned host) + {args.steps} x perform_step! + collect(state) to pinned host",
                "seconds": e2e_sec},
        "gpu_launches": int(launches), "setup_seconds": t_setup,
    }
    if per_rank is not None:
        out["per_rank"] = per_rank
    if balance is not None:
        out["config"]["strip_balance"] = balance
    print(json.dumps(out), flush=True)
    barrier()
    if dist is not None:
        strip.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[2, 3, 4, 5],
                    help="BASELINE config (SURVEY 8d numbering): 5 = the bench line (16384^2 Float32); 2-4 = side measurements, N = 1")
    ap.add_argument("--n", "--size", dest="n", type=int, default=None,
                    help="grid size (use --size under torchrun, whose own parser finds --n ambiguous)")
    ap.add_argument("--dtype", default=None, choices=["float32", "float64"])
    ap.add_argument("--cpu-n", type=int, default=1024, dest="cpu_n")
    ap.add_argument("--cpu-rows", type=int, default=1024, dest="cpu_rows",
                    help="--impl reference, config 5: rows of the full-width band each CPU step processes")
    ap.add_argument("--rebalance", type=int, default=4,
                    help="N > 1: timed steps of the short run whose measured per-rank stage-kernel times re-cut the strips (0: model only)")
    ap.add_argument("--cpu-steps", type=int, default=30, dest="cpu_steps", help="RK2 steps of the cpu_baseline sample")
    ap.add_argument("--strict", action="store_true")
    ap.add_argument("--unfused", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--ny", type=int, default=None, help="N = 1 only: rows of a non-square grid (strip-shaped tuning runs)")
    ap.add_argument("--equal-strips", action="store_true", dest="equal_strips",
                    help="N > 1: equal-height y-strips instead of the cost-balanced partition")
    ap.add_argument("--level", type=float, default=0.0,
                    help="config 5: still-water level of the lake (0 = the bench line, ~54 %% of the cells wet; 60 floods every cell)")
    ap.add_argument("--bc", default="wall", choices=["wall", "periodic", "neumann"],
                    help="boundary condition of the workload (N = 1; the bench line uses WallBC)")
    ap.add_argument("--dry-cost", type=float, default=0.6, dest="dry_cost",
                    help="cost of a dry cell relative to a wet one in the strip balancing model")
    args = ap.parse_args()
    args.n = args.n or CONFIGS[args.config][1]
    args.dtype = args.dtype or CONFIGS[args.config][2]
    args.flush_l2 = 3 * 3 * (args.n + 4) * ((args.ny or args.n) + 4) * np.dtype(args.dtype).itemsize < 4 * 126e6
    global LAKE_LEVEL, BC
    LAKE_LEVEL = args.level
    BC = args.bc
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        if args.cpu_n == 1024:
            args.cpu_n = 2048
        run_reference(args, rank)
        return
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
