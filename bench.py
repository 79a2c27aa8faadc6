#!/usr/bin/env python
"""bench.py -- inflow cell-updates/s per filter() step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

One "step" = one DIGITAL_FILTER::filter(dt) (df.cpp:449-468) over one synthetic plane: noise generation,
y-sweep, z-sweep, temporal blend, RST scaling, SRA -- all five outputs of every cell.
  value    whole-job cell-updates/s, state resident in HBM, CUDA events on the library's stream
  e2e      the same through the reference-facing call (dfb_filter_to_host: filter(dt) + the five
           fields copied into pinned HOST arrays, as the C++/Fortran facades do every step)
  roofline the dominant kernel against the fp64-FMA roof measured live (DFMA microbenchmark in the
           library; MEASURED_PEAKS.json has no fp64 entry) and the step against the HBM roof
  cpu_baseline  the reference's own df.cpp (oracle/_ref) on the box's host, bounded sample
N > 1 (torchrun): every rank filters its own independent plane (distinct RNG stream group = rank):
weak scaling, no data-path collective; time = max over ranks.
--impl reference: the reference's CPU filter() on all host cores (independent processes; the
reference is single-threaded with a process-wide RNG), same metric/config, bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_WORKLOAD = "1024x2048_profile_N128"
METRIC = "inflow cell-updates/sec per filter() step"
DT = 1e-7


def ncu_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full summary
    (profiles/ncu_full_r01c_summary.csv, captured on the 1024x2048 profile workload); None for other workloads."""
    if workload != DEFAULT_WORKLOAD:
        return None
    try:
        import csv
        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "ncu_full_r01c_summary.csv"))))
        col = [i for i, h in enumerate(rows[0]) if kernel.split("<")[0] in h][0]
        tot = 0.0
        for r in rows[1:]:
            if r[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[1]]
                tot += float(r[col]) * scale
        return tot or None
    except Exception:
        return None


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML every
    2 ms from a thread (pynvml ships in the image); nvidia-smi -lms as the fall-back."""
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("hw_power_brake_slowdown", 0x80), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.idx, self.samples, self.stop_flag, self.t, self.h, self.nv = gpu_index, [], False, None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)),
                                     float(nv.nvmlDeviceGetPowerUsage(h)) / 1e3))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if not self.nv or not self.t:
            return None
        self.stop_flag = True
        self.t.join(timeout=1)
        if not self.samples:
            return None
        sm = [s[0] for s in self.samples]
        mask = 0
        for s in self.samples:
            mask |= s[1]
        reasons = [n for n, bit in self.REASONS if mask & bit]
        return dict(sm_mhz=float(np.median(sm)), sm_min_mhz=float(min(sm)), sm_max_mhz=self.max_mhz, reasons=reasons,
                    power_w_max=max(s[2] for s in self.samples), samples=len(sm))


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own df.cpp on a bounded sample of the workload
# --------------------------------------------------------------------------------------------------
def sample_plane(plane, target_taps=4.0e8):
    """A spanwise sub-slab of the same plane (same rows, same half-width profile, fewer columns)
    sized for ~1 s of single-core CPU work per step."""
    from oracle import oracle as O
    p = dict(plane)
    O.half_widths(p)
    taps_per_col = float(sum((2 * p["N_y"][f][:, 0].astype(np.int64) + 1).sum() + (2 * p["N_z"][f][:, 0].astype(np.int64) + 1).sum() for f in range(3)))
    ncol = int(max(8, min(plane["Nz"], target_taps // taps_per_col)))
    s = dict(plane)
    s["Nz"] = ncol
    s["name"] = plane["name"] + f"[sample:{plane['Ny']}x{ncol}]"
    return s, taps_per_col * ncol


def _ref_worker(args):
    plane, steps, warmup = args
    from oracle import oracle as O
    R = O.RefFilter()
    R.reshape(plane)
    R.time_steps(DT, warmup)
    tot, st = R.time_steps(DT, steps)
    R.close()
    return tot


def cpu_reference_run(plane, steps, warmup, nproc):
    """nproc independent reference objects, one process each; returns (seconds for `steps` steps, kind)."""
    from oracle import oracle as O
    if O.have_ref():
        import multiprocessing as mp
        if nproc == 1:
            return _ref_worker((plane, steps, warmup)), "reference"
        with mp.get_context("fork").Pool(nproc) as pool:
            ts = pool.map(_ref_worker, [(plane, steps, warmup)] * nproc)
        return max(ts), "reference"
    # fall-back checker: the plain-C restatement (single instance, OpenMP over rows)
    p = dict(plane)
    O.half_widths(p)
    Ny, Nz = p["Ny"], p["Nz"]
    rys = [O.noise_rys(1, 0, f, 0, Ny, p["Ny_max"][f], Nz) for f in range(3)]
    hal = [O.noise_halo(1, 0, f, 0, Ny, p["Nz_max"][f]) for f in range(3)]
    fo = np.zeros((3, Ny, Nz))
    for _ in range(warmup):
        O.step(p, rys, hal, fo, DT)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.step(p, rys, hal, fo, DT)
    return time.perf_counter() - t0, "port"


def run_reference_arm(args, plane):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    nproc = max(1, min(ncores, 32))
    # bounded sample: ~100 s of wall clock for the whole --steps/--warmup run at ~0.4 G tap/s per core
    wu = min(args.warmup, 2)
    target = min(2.0e8, max(1.0e7, 0.4e9 * 100.0 / (args.steps + wu)))
    s, taps = sample_plane(plane, target_taps=target)
    secs, kind = cpu_reference_run(s, args.steps, wu, nproc)
    cells = s["Ny"] * s["Nz"] * nproc * args.steps
    value = cells / secs
    line = dict(impl="reference", metric=METRIC, value=value, unit="cell-updates/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * secs / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=plane["name"], plane=[plane["Ny"], plane["Nz"]], dt=DT),
                cpu_baseline=dict(value=value, unit="cell-updates/s", cores=nproc if kind == "reference" else ncores, kind=kind,
                                  sample=f"{nproc} independent instances x ({s['Ny']}x{s['Nz']} sub-slab of the {plane['Ny']}x{plane['Nz']} plane, same rows/half-widths)"
                                         f" per step; the reference's five stage calls (df.cpp:453-461), no print/CSV"),
                e2e=dict(value=value, unit="cell-updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args, plane):
    import torch
    import _dfb_import  # noqa: F401
    import digital_filtering_b200 as dfb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner / warnings: not on stdout, where ONE JSON line goes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    K, Wm = args.steps, max(args.warmup, 3)
    df = dfb.DIGITAL_FILTER(dfb.DFConfig.from_plane(plane, seed=20261018, plane_id=rank, device=local), fetch=False)
    assert df.tuned, "tuned sm_100a kernels must be the measured path"
    stream = torch.cuda.ExternalStream(df.stream(), device=local)
    cells = df.n_cells
    N_y = [df.half_widths(f, 0) for f in range(3)]
    N_z = [df.half_widths(f, 1) for f in range(3)]
    taps_y = int(sum((2 * a.astype(np.int64) + 1).sum() for a in N_y))
    taps_z = int(sum((2 * a.astype(np.int64) + 1).sum() for a in N_z))
    ws_bytes = 8 * (sum((plane["Ny"] + 2 * int(a.max())) * plane["Nz"] for a in N_y) + 3 * cells + 11 * cells)
    l2_bytes = 126 * 2 ** 20
    flush = ws_bytes < 2 * l2_bytes
    fbuf = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda") if flush else None

    # ---- value: device-resident steps, CUDA events on the library's stream ----
    for _ in range(Wm):
        df.filter(DT)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(K):
        if flush:
            with torch.cuda.stream(stream):
                fbuf.zero_()
        ev[i][0].record(stream)
        df.filter(DT)
        ev[i][1].record(stream)
    df.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * cells * K / (ms_total * 1e-3)

    # ---- e2e: filter(dt) + five fields into pinned host arrays, every step ----
    host = [torch.empty(cells, dtype=torch.float64).pin_memory() for _ in range(5)]
    hp = [h.data_ptr() for h in host]
    L = dfb.lib()
    for _ in range(2):
        dfb._check(L.dfb_filter_to_host(df._h, DT, *hp))
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        dfb._check(L.dfb_filter_to_host(df._h, DT, *hp))
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * cells * K / float(te.item())

    # ---- e2e, pipelined variant of the same call (dfb_filter_to_host_begin / _end, two sets of pinned arrays): the copy of step t
    # runs under the compute of step t+1; every step's five fields still arrive in host memory.  Reported beside, not instead of, e2e.
    host2 = [torch.empty(cells, dtype=torch.float64).pin_memory() for _ in range(5)]
    sets = [hp, [h.data_ptr() for h in host2]]
    dfb._check(L.dfb_filter_to_host_begin(df._h, DT, *sets[0]))
    dfb._check(L.dfb_filter_to_host_begin(df._h, DT, *sets[1]))
    dfb._check(L.dfb_filter_to_host_end(df._h)); dfb._check(L.dfb_filter_to_host_end(df._h))
    barrier()
    t0 = time.perf_counter()
    dfb._check(L.dfb_filter_to_host_begin(df._h, DT, *sets[0]))
    for i in range(1, K):
        dfb._check(L.dfb_filter_to_host_begin(df._h, DT, *sets[i & 1]))
        dfb._check(L.dfb_filter_to_host_end(df._h))            # step i-1 is in host memory
    dfb._check(L.dfb_filter_to_host_end(df._h))
    barrier()
    tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    e2e_pipelined = world * cells * K / float(tp.item())

    if rank != 0:
        df.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- per-kernel times (CUDA events inside the library, one sync per step: explains, not the headline) ----
    df.set_timing(True)
    stage = []
    for _ in range(max(5, min(K, 20))):
        if flush:
            with torch.cuda.stream(stream):
                fbuf.zero_()
        df.filter(DT)
        stage.append(df.last_ms())
    df.set_timing(False)
    med = {k: float(np.median([s[k] for s in stage])) for k in stage[0]}
    fp64_peak, mhz = dfb.measure_fp64_peak(local)
    peaks = load_peaks()
    hbm_peak = peaks["hbm_gbs"] if peaks and "hbm_gbs" in peaks else 6650.0
    hbm_src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    alg_bytes = 88 * cells                                    # 5 outputs + filt_old r/w x3
    zmode = "recursive" if df.info(7) == 1 else "direct"
    y_rec = df.info(8) > 0
    y_bytes = 3 * 16 * cells                                  # y-sweep: r_ys read once, r_zs interior written once, per field
    kern = {
        # y-sweep.  Direct form (dense band matrices): executes exactly the reference's 2*(2N_y+1) flops per cell -> fp64 roof.
        # Recursive form (uniform planes): ~6x fewer flops -> its roof is the memory system; equivalent_tflops = reference-
        # formulation flops / time, for comparison only.
        "ysweep_tma_kernel": dict(ms=med["ysweep"], form=("recursive (%d tiles) + direct (%d tiles)" % (df.info(8), df.info(9))) if y_rec else "direct"),
        # z-sweep + epilogue: in recursive mode it no longer executes the reference's 2*(2N_z+1) flops per cell (about 13x fewer at
        # N = 128): its roof is the memory system
        "zsweep_epilogue_kernel": dict(ms=med["zsweep_epilogue"], form=zmode, equivalent_tflops=2 * taps_z / (med["zsweep_epilogue"] * 1e-3) / 1e12,
                                       hbm_gbs=alg_bytes / (med["zsweep_epilogue"] * 1e-3) / 1e9),
        "noise_kernel": dict(ms=med["noise"]),
    }
    ky, kz = kern["ysweep_tma_kernel"], kern["zsweep_epilogue_kernel"]
    if y_rec:
        ky.update(equivalent_tflops=2 * taps_y / (med["ysweep"] * 1e-3) / 1e12, hbm_gbs=y_bytes / (med["ysweep"] * 1e-3) / 1e9)
        ky["frac_hbm"] = ky["hbm_gbs"] / hbm_peak
    else:
        ky.update(tflops=2 * taps_y / (med["ysweep"] * 1e-3) / 1e12)
        ky["frac_fp64"] = ky["tflops"] / fp64_peak
    kz["frac_hbm"] = kz["hbm_gbs"] / hbm_peak
    if zmode == "direct":
        kz["frac_fp64"] = kz["equivalent_tflops"] / fp64_peak
    step_ms = ms_total / K
    eq_tf = 2 * (taps_y + taps_z) / (step_ms * 1e-3) / 1e12
    step = dict(equivalent_tflops=eq_tf, equivalent_frac_fp64=eq_tf / fp64_peak,
                note="reference-formulation flops / step time; the recursive sweeps execute far fewer, so this is a speed-up figure, not a utilisation",
                hbm_gbs=alg_bytes / (step_ms * 1e-3) / 1e9, frac_hbm=alg_bytes / (step_ms * 1e-3) / 1e9 / hbm_peak, hbm_peak=hbm_peak, hbm_peak_source=hbm_src)
    peak_src = "DFMA microbenchmark measured live in this run (dfb_measure_fp64_peak); MEASURED_PEAKS.json has no fp64 entry"
    tsrc = "profiles/ncu_full_r01c_summary.csv (ncu --set full, one launch)"
    dom = "ysweep_tma_kernel" if med["ysweep"] >= med["zsweep_epilogue"] else "zsweep_epilogue_kernel"
    kd = kern[dom]
    if "frac_fp64" in kd and kd["form"] == "direct":
        roofline = dict(bound="fp64", kernel=dom, achieved=kd.get("tflops", kd.get("equivalent_tflops")), peak=fp64_peak, unit="TFLOP/s",
                        frac=kd["frac_fp64"], traffic=ncu_traffic(dom, plane["name"]), traffic_source=tsrc, peak_source=peak_src, step=step, kernels=kern)
    else:
        roofline = dict(bound="hbm", kernel=dom, achieved=kd["hbm_gbs"], peak=hbm_peak, unit="GB/s", frac=kd["frac_hbm"],
                        traffic=ncu_traffic(dom, plane["name"]), traffic_source=tsrc, peak_source=hbm_src, step=step, kernels=kern)

    # ---- cpu baseline: the reference's own df.cpp, 1 core, bounded sample ----
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            s, _ = sample_plane(plane)
            nst = 25                  # ~10 s of one host core
            secs, kind = cpu_reference_run(s, nst, 1, 1)
            cpu = dict(value=s["Ny"] * s["Nz"] * nst / secs, unit="cell-updates/s", cores=1 if kind == "reference" else (os.cpu_count() or 1), kind=kind,
                       sample=f"{nst} steps of a {s['Ny']}x{s['Nz']} spanwise sub-slab of the {plane['Ny']}x{plane['Nz']} plane (same rows, same half-widths); "
                              f"the reference's five stage calls (df.cpp:453-461), 1 of {os.cpu_count()} host cores", ms_per_step_sample=1e3 * secs / nst)
        except Exception as e:   # the checker is optional for the bench; say so rather than die
            cpu = dict(value=None, unit="cell-updates/s", cores=0, kind="unavailable", sample=str(e)[:200])

    # ---- the other single-GPU configurations of BASELINE.json, briefly (parity-test cases; reported for context) ----
    sweep = None
    if world == 1 and not args.no_sweep:
        from digital_filtering_b200 import workloads as WL
        sweep = {}
        cases = [("1024x2048_saturated_N128", lambda: dfb.DFConfig.from_plane(WL.NAMED["1024x2048_saturated_N128"](), seed=1, device=local)),
                 ("512x512_N32", lambda: dfb.DFConfig.from_plane(WL.NAMED["512x512_N32"](), seed=1, device=local))]
        rst, ln = os.path.join(ROOT, "oracle", "_ref", "files", "RST.dat"), os.path.join(ROOT, "oracle", "_ref", "line.dat")
        if os.path.exists(rst) and os.path.exists(ln):
            cases.append(("reference_default_plane_510x400", lambda: dfb.DFConfig(vel_fluc_file=rst, line_file=ln, seed=1, device=local)))
        for name, mk in cases:
            try:
                d2 = dfb.DIGITAL_FILTER(mk(), fetch=False)
                st2 = torch.cuda.ExternalStream(d2.stream(), device=local)
                for _ in range(5):
                    d2.filter(DT)
                d2.sync()
                n2 = 200
                a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a2.record(st2)
                for _ in range(n2):
                    d2.filter(DT)
                b2.record(st2)
                d2.sync()
                ms2 = a2.elapsed_time(b2) / n2
                taps2 = d2.taps_per_step
                sweep[name] = dict(cells=d2.n_cells, ms_per_step=ms2, cell_updates_per_s=d2.n_cells / (ms2 * 1e-3),
                                   equivalent_tflops=2 * taps2 / (ms2 * 1e-3) / 1e12,
                                   l2_note="back-to-back steps, no flush (context only)")
                d2.close()
            except Exception as e:
                sweep[name] = dict(error=str(e)[:160])

    line = dict(metric=METRIC, value=value, unit="cell-updates/s", n_gpus=world, steps=K, warmup=Wm, ms_per_step=step_ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=plane["name"], plane=[plane["Ny"], plane["Nz"]], max_half_width=[int(max(a.max() for a in N_y)), int(max(a.max() for a in N_z))],
                            taps_per_cell=(taps_y + taps_z) / cells, dt=DT, noise="generate (counter-based pcg32, spec v1)",
                            l2="flushed between steps (256 MiB write)" if flush else f"per-step working set {ws_bytes / 2**20:.0f} MiB > 126 MiB L2, no flush",
                            parallelism=f"{world} independent plane(s), one per GPU"),
                clocks=clocks, gpu_launches=3 * K, wall_ms_per_step=1e3 * t_wall / K,
                e2e=dict(value=e2e_value, unit="cell-updates/s", h2d_bytes_per_step=8, d2h_bytes_per_step=40 * cells,
                         call="dfb_filter_to_host (filter(dt) + u',v',w',T',rho' into pinned host arrays); input is the scalar dt",
                         pipelined=dict(value=e2e_pipelined, unit="cell-updates/s",
                                        call="dfb_filter_to_host_begin/_end with two sets of pinned arrays: the copy of step t under the compute of step t+1")),
                roofline=roofline, cpu_baseline=cpu, other_configs=sweep)
    print(json.dumps(line), flush=True)
    df.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the brief runs of the other named configurations")
    args = ap.parse_args()
    import _dfb_import  # noqa: F401
    from digital_filtering_b200 import workloads as W
    if args.workload not in W.NAMED:
        raise SystemExit(f"unknown workload {args.workload}; choose from {sorted(W.NAMED)}")
    plane = W.NAMED[args.workload]()
    if args.impl == "reference":
        run_reference_arm(args, plane)
    else:
        run_b200(args, plane)


if __name__ == "__main__":
    main()
