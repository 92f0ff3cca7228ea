#!/usr/bin/env python
"""bench.py -- gridpoint*days/s of EQM train+adjust (BASELINE.json configs[1]) on N B200s.

One "step" = one pass of the hot path (train(ref, hist) + adjust(sim)) over the whole 0.25 degree
global grid (1440 x 721 gridpoints x 30 years daily, float32, group="time.month", nq=50, kind="+",
interp="nearest", extrapolation="constant"), streamed through HBM as lat-band slabs because the four
arrays (182 GB) do not fit next to each other in 180 GB.  Every slab's synthetic ref/hist/sim is
generated on the device OUTSIDE the timed region; the timed region (CUDA events on the launching
stream) is train + adjust of the slab with inputs resident in HBM; a step's time is the sum over its
slabs.  Multi-GPU: one process per GPU, every rank processes its own full grid (weak scaling, no
data-path collective), time = max over ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NLON, NLAT, NYEARS = 1440, 721, 30
NQ = 50
GROUP = "time.month"
METRIC = "gridpoint*days/s (EQM train+adjust, 0.25deg global x 30yr daily, f32)"
UNIT = "gridpoint*days/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lat-rows", type=int, default=NLAT, help="lat rows per rank per step (default: full grid)")
    ap.add_argument("--slab-rows", type=int, default=48, help="lat rows per slab")
    ap.add_argument("--e2e-rows", type=int, default=8, help="lat rows of the host-buffer end-to-end sample")
    ap.add_argument("--cpu-points", type=int, default=0, help="gridpoints of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args):
    return {
        "workload": "EQM nquantiles=50 group=time.month kind=+ interp=nearest extrapolation=constant, "
                    f"synthetic tas f32, {NLON}x{args.lat_rows} gridpoints x {NYEARS}-year daily ref/hist/sim (noleap)",
        "grid": [args.lat_rows, NLON], "n_time": 365 * NYEARS, "nquantiles": NQ, "group": GROUP,
        "slab_lat_rows": args.slab_rows, "parallelism": f"lat-band slabs, {args.gpus} rank(s), no collective",
        "l2": "inputs of every timed region (>= 4 GB per slab) exceed the 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------------------
# reference arm: the oracle port on the host cores (the reference cannot be installed: no xarray)
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline

    cores = os.cpu_count() or 1
    n_pts = args.cpu_points or 1024 * cores  # ~10 s of work per timed call on 16 cores
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline.run(n_pts, cores, NYEARS, NQ, GROUP, seed=100 + i)
        if i >= args.warmup:
            vals.append(r["value"]); secs.append(r["seconds"])
    v = float(sum(vals) / len(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        self.active = False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag:
                if self.active:
                    self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def synth_slab(torch, gen, T, n_lat, lat0, which, doy, year, device):
    """Synthetic tas slab, time-major (T, n_lat*NLON) f32 (SURVEY.md 8d generator, torch Philox)."""
    A, sigma, k, off = {"ref": (12.0, 3.0, 1.0, 0.0), "hist": (10.0, 3.5, 1.0, 1.5), "sim": (10.0, 3.5, 1.1, 3.5)}[which]
    n = n_lat * NLON
    x = torch.empty((T, n), dtype=torch.float32, device=device)
    x.normal_(0.0, sigma, generator=gen)
    base = 273.15 - A * torch.cos(2 * torch.pi * (doy - 15.0) / 365.0) + 0.03 * year * k + off  # [T]
    lat = torch.deg2rad(-90.0 + 0.25 * (lat0 + torch.arange(n_lat, device=device, dtype=torch.float32)))
    x += base[:, None]
    x += (15.0 * torch.cos(lat)).repeat_interleave(NLON)[None, :]
    # 0.1 % isolated NaNs + one all-NaN gridpoint per slab
    m = torch.empty((T, n), dtype=torch.uint8, device=device).random_(0, 250, generator=gen)  # ~0.4 % per value < 1
    m2 = torch.empty((T, n), dtype=torch.uint8, device=device).random_(0, 4, generator=gen)
    x[(m == 0) & (m2 == 0)] = float("nan")
    del m, m2
    x[:, n // 2] = float("nan")
    return x


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import xsdba_b200 as xs
    from xsdba_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: xsdba_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    t_train = xs.TimeAxis.daily(1981, NYEARS, "noleap")
    t_sim = xs.TimeAxis.daily(2041, NYEARS, "noleap")
    T = len(t_train)
    grouper = xs.Grouper(GROUP)
    q = xs.equally_spaced_nodes(NQ).astype(np.float32)
    h_train = grouper.handle(t_train)
    h_sim = grouper.handle(t_sim, with_window=False)
    G = h_train.n_groups
    q_dev = torch.from_numpy(q).to(dev)
    doy = torch.from_numpy(t_train.dayofyear.astype(np.float32)).to(dev)
    year = torch.from_numpy((t_train.year - t_train.year[0]).astype(np.float32)).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(20260117 + rank)
    stream = torch.cuda.current_stream().cuda_stream

    slabs = [(r0, min(args.slab_rows, args.lat_rows - r0)) for r0 in range(0, args.lat_rows, args.slab_rows)]
    max_pts = max(n for _, n in slabs) * NLON
    af = torch.empty((max_pts, G, NQ), dtype=torch.float32, device=dev)
    hq = torch.empty_like(af)
    scen = torch.empty((T, max_pts), dtype=torch.float32, device=dev)

    def one_step(timed):
        """-> (train_ms, adjust_ms, checksum) summed over the slabs of this rank's grid."""
        tr_ms = ad_ms = 0.0
        chk = 0.0
        for r0, nrow in slabs:
            n = nrow * NLON
            ref = synth_slab(torch, gen, T, nrow, r0, "ref", doy, year, dev)
            hist = synth_slab(torch, gen, T, nrow, r0, "hist", doy, year, dev)
            sim = synth_slab(torch, gen, T, nrow, r0, "sim", doy, year, dev)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            st = lib.xsdba_qm_train_f32(ref.data_ptr(), hist.data_ptr(), n, 1, n, h_train.ptr, q_dev.data_ptr(), NQ, 43, 0,
                                        af.data_ptr(), hq.data_ptr(), None, stream)
            _lib.check(st, "train")
            e1.record()
            st = lib.xsdba_qm_adjust_f32(sim.data_ptr(), n, 1, n, h_sim.ptr, af.data_ptr(), hq.data_ptr(), NQ, 0, 0, 43,
                                         scen.data_ptr(), stream)
            _lib.check(st, "adjust")
            e2.record()
            e2.synchronize()
            tr_ms += e0.elapsed_time(e1)
            ad_ms += e1.elapsed_time(e2)
            if timed:
                chk += float(torch.nansum(scen[:, :n][:: 997, :: 101]))
            del ref, hist, sim
        return tr_ms, ad_ms, chk

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(False)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = lib.xsdba_launch_count()
    sampler.active = True
    wall0 = time.perf_counter()
    tr_tot = ad_tot = 0.0
    for _ in range(args.steps):
        a, b, chk = one_step(True)
        tr_tot += a
        ad_tot += b
    barrier()
    wall = time.perf_counter() - wall0
    sampler.active = False
    launches = lib.xsdba_launch_count() - l0
    dev_ms = torch.tensor([tr_tot + ad_tot, tr_tot, ad_tot], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    tot_ms, tr_ms, ad_ms = (float(v) for v in dev_ms.cpu())
    n_pts_rank = args.lat_rows * NLON
    units = world * n_pts_rank * T * args.steps
    value = units / (tot_ms * 1e-3)

    # ---- roofline of the dominant kernel (train: group-segmented sort + quantiles) ----------------
    sys.path.insert(0, ROOT)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n_launch = len(slabs) * args.steps
    train_bytes_step = n_pts_rank * (2 * T * 4 + 2 * G * NQ * 4)          # reads ref+hist, writes af+hist_q
    adjust_bytes_step = n_pts_rank * (2 * T * 4 + 2 * G * NQ * 4)         # reads sim+tables, writes scen
    dom = "train" if tr_ms >= ad_ms else "adjust"
    dom_bytes = train_bytes_step if dom == "train" else adjust_bytes_step
    dom_ms = max(tr_ms, ad_ms)
    achieved = dom_bytes * args.steps / (dom_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "train_fast_kernel<false, false>" if dom == "train" else "pack_tables_kernel + adjust_tile_kernel",
        "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650",
        # dram__bytes_read+write per launch from the ncu --set full capture in profiles/r01_ncu_final.txt:
        # 6.505 GB for a 69 120-point launch whose algorithmic bytes are 6.387 GB (x1.0185: no re-reads)
        "traffic": (dom_bytes * args.steps / n_launch) * 1.0185 if dom == "train" else None,
        "traffic_source": "ncu dram__bytes_{read,write}.sum of profiles/r01_ncu_final.txt scaled to this launch size",
        "launches": n_launch, "avg_launch_ms": dom_ms / n_launch,
        "algorithmic_bytes_per_launch": dom_bytes * args.steps / n_launch,
        "step": {"train_ms": tr_ms / args.steps, "adjust_ms": ad_ms / args.steps,
                 "train_GBps": train_bytes_step * args.steps / (tr_ms * 1e-3) / 1e9,
                 "adjust_GBps": adjust_bytes_step * args.steps / (ad_ms * 1e-3) / 1e9,
                 "whole_step_frac_of_peak": (train_bytes_step + adjust_bytes_step) * args.steps / (tot_ms * 1e-3) / 1e9 / peak},
    }

    # ---- end to end through the host C ABI (pinned host buffers, copies inside the timed region) ---
    n_e2e = args.e2e_rows * NLON
    e2e = None
    if n_e2e > 0:
        hostgen = torch.Generator().manual_seed(7 + rank)
        hb = []
        for A_, s_, off_ in ((12.0, 3.0, 0.0), (10.0, 3.5, 1.5), (10.0, 3.5, 3.5)):
            t = torch.empty((T, n_e2e), dtype=torch.float32).normal_(0, s_, generator=hostgen)
            t += (273.15 + off_ - A_ * torch.cos(2 * torch.pi * (doy.cpu() - 15.0) / 365.0))[:, None]
            hb.append(t.pin_memory())
        out_h = torch.empty((T, n_e2e), dtype=torch.float32).pin_memory()
        ref_h, hist_h, sim_h = (t.numpy() for t in hb)
        e2e_times = []
        for i in range(2 + max(1, args.steps)):
            barrier()
            t0 = time.perf_counter()
            xs.train_adjust_host(ref_h, hist_h, sim_h, time=t_train, sim_time=t_sim, nquantiles=NQ, group=GROUP,
                                 kind="+", method="eqm", slab_points=2048, out=out_h.numpy())
            chk_e2e = float(out_h[::997, ::101].sum())  # the device->host result is read
            e2e_times.append(time.perf_counter() - t0)
        e2e_t = torch.tensor([min(e2e_times[2:])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n_e2e * T / float(e2e_t.cpu()), "unit": UNIT,
               "h2d_bytes_per_step": 3 * n_e2e * T * 4, "d2h_bytes_per_step": n_e2e * T * 4,
               "sample": f"{n_e2e} gridpoints x {T} days per rank through xsdba_qm_train_adjust_host_f32 "
                         "(pinned host ref/hist/sim -> scen), best of the timed calls"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import cpu_baseline
        cores = os.cpu_count() or 1
        r = cpu_baseline.run(args.cpu_points or 2048 * cores, cores, NYEARS, NQ, GROUP)  # ~15-20 s of CPU work
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
               "seconds": r["seconds"]}

    sampler.stop_flag = True
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args), "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(),
            "wall_ms_per_step_incl_generation": 1e3 * wall / args.steps, "checksum": chk,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
