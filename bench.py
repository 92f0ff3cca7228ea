#!/usr/bin/env python
"""bench.py -- gridpoint*days/s of EQM train+adjust (BASELINE.json configs[1]) on N B200s.

One "step" = one pass of the hot path (train(ref, hist) + adjust(sim)) over the whole 0.25 degree
global grid (1440 x 721 gridpoints x 30 years daily, float32, group="time.month", nq=50, kind="+",
interp="nearest", extrapolation="constant").  The four arrays (182 GB) do not fit next to each other
in 180 GB, so the grid streams through HBM as lat-band slabs; every slab's synthetic ref/hist/sim is
generated on the device OUTSIDE the timed region; the timed region (CUDA events on the launching
stream) is train + adjust of the slab with inputs resident in HBM; a step's time is the sum over its
slabs.

Multi-GPU (--gpus N under torchrun): STRONG scaling -- the ONE 721-row grid is cut into N lat bands
(`xsdba_b200.sharding.lat_band`), one process per GPU, no data-path collective; `value` = gridpoint*days of
the whole grid / max over ranks of the device time.  `weak` (every rank a full grid) is an extra key.

Besides the headline the JSON line carries
  roofline      dominant kernel: algorithmic bytes / its CUDA-event time vs the measured HBM peak
  cpu_baseline  the oracle port on the host cores (rank 0, N = 1)
  e2e           the same metric through the host-buffer C ABI entry (pinned host ref/hist/sim -> scen,
                copies inside the timed region), MEAN over the timed calls, one full slab per rank
  pcie          concurrent pinned H2D bandwidth per rank: the ceiling e2e scales against
  configs       BASELINE.json configs[2..4] (QDM doy x 31, DQM + LOESS, MBCn): device-resident value,
                roofline fraction on SURVEY 8d's algorithmic bytes, e2e / cpu_baseline for cfg3

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--configs cfg3,cfg4,cfg5|none]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import warnings

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
    ap.add_argument("--lat-rows", type=int, default=NLAT, help="lat rows of the whole grid (default: 721)")
    ap.add_argument("--slab-rows", type=int, default=48, help="lat rows per slab")
    ap.add_argument("--e2e-rows", type=int, default=48, help="lat rows of the host-buffer end-to-end sample per rank")
    ap.add_argument("--cpu-points", type=int, default=0, help="gridpoints of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="cfg3,cfg4,cfg5", help="other BASELINE configs to measure (or 'none')")
    ap.add_argument("--cfg-rows", type=int, default=8, help="lat rows per rank of the cfg3 / cfg4 samples")
    ap.add_argument("--no-weak", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "EQM nquantiles=50 group=time.month kind=+ interp=nearest extrapolation=constant, "
                    f"synthetic tas f32, {NLON}x{args.lat_rows} gridpoints x {NYEARS}-year daily ref/hist/sim (noleap)",
        "grid": [args.lat_rows, NLON], "n_time": 365 * NYEARS, "nquantiles": NQ, "group": GROUP,
        "slab_lat_rows": args.slab_rows,
        "parallelism": f"one grid cut into {world} lat band(s), one rank per GPU, no collective",
        "l2": "inputs of every timed region (>= 2.5 GB per slab) exceed the 126 MB L2; no flush needed",
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
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": cores, "gpu_launches": 0,
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
# synthetic inputs (SURVEY.md 8d generators, torch Philox on the device)
# ------------------------------------------------------------------------------------------------
def synth_slab(torch, gen, T, n_lat, lat0, which, doy, year, device):
    """Synthetic tas slab, time-major (T, n_lat*NLON) f32, NaN-seeded."""
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


def synth_pr(torch, gen, T, n, which, device):
    """Synthetic pr [mm/d]: Bernoulli(p_wet) * Gamma(shape, scale), dry days exactly 0, 0.1 % NaNs + an all-NaN
    block of gridpoints (land / sea mask).  The Gamma variates come from torch's global CUDA generator."""
    p_wet, shape, scale = {"ref": (0.45, 0.8, 7.5), "hist": (0.60, 0.9, 5.0), "sim": (0.60, 0.9, 5.5)}[which]
    conc = torch.full((T, n), shape, dtype=torch.float32, device=device)
    x = torch._standard_gamma(conc) * scale
    del conc
    wet = torch.empty((T, n), dtype=torch.float32, device=device).uniform_(0, 1, generator=gen) < p_wet
    x *= wet
    m = torch.empty((T, n), dtype=torch.int16, device=device).random_(0, 1000, generator=gen)
    x[m == 0] = float("nan")
    del m, wet
    x[:, n // 2: n // 2 + 16] = float("nan")
    return x


def ncu_traffic_ratio(path, kernel):
    """dram bytes (read + write) of the captured launch of `kernel` in a profiles/ncu_summary.py text file, divided by
    the algorithmic bytes of that launch (grid.x tiles of 32 points x grid.y month groups); None if not found."""
    try:
        txt = open(path).read()
    except OSError:
        return None
    import re
    m = re.search(r"Kernel Name\s+[^\n]*" + kernel + r"[^\n]*\nGrid Size\s+\((\d+), (\d+), 1\)[^\n]*\n(?:[^\n]*\n){1,3}?"
                  r"dram__bytes_read\.sum\s+([\d.]+) (\w+)\ndram__bytes_write\.sum\s+([\d.]+) (\w+)", txt)
    if not m:
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = float(m.group(3)) * unit[m.group(4)] + float(m.group(5)) * unit[m.group(6)]
    pts, groups = int(m.group(1)) * 32, int(m.group(2))
    return total / (pts * (2 * NYEARS * 365 * 4 + 2 * groups * NQ * 4))


def time_device(torch, fn, steps, warmup=2):  # (two: the second call still grows torch's caching allocator)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[2..4]: rank-local samples, device-resident (weak over ranks)
# ------------------------------------------------------------------------------------------------
def run_other_configs(args, torch, dist, xs, lib, dev, rank, world, peak, want):
    import numpy as np

    out = {}
    tt = xs.TimeAxis.daily(1981, NYEARS, "noleap")
    ts = xs.TimeAxis.daily(2041, NYEARS, "noleap")
    T = len(tt)
    n = args.cfg_rows * NLON
    gen = torch.Generator(device=dev)
    gen.manual_seed(777 + rank)
    torch.manual_seed(4242 + rank)
    steps = max(1, min(args.steps, 2))

    def reduce_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.cpu())

    warnings.simplefilter("ignore")
    if "cfg3" in want:
        # QDM kind='*' pr, Grouper('time.dayofyear', 31), nq=100, nearest / constant; pr jittered once before timing
        g = xs.Grouper("time.dayofyear", 31)
        ref, hist, sim = (xs.jitter_under_thresh(synth_pr(torch, gen, T, n, w, dev), "0.01 mm/d", seed=11 + i)
                          for i, w in enumerate(("ref", "hist", "sim")))
        res = {"workload": f"QDM kind=* pr f32, Grouper(time.dayofyear, window=31), nq=100, nearest/constant, "
                           f"{n} gridpoints x {T} days per rank (weak over ranks)"}
        for rw in (False, True):
            state = {}

            def train():
                state["obj"] = xs.QuantileDeltaMapping.train(ref, hist, time=tt, nquantiles=100, group=g, kind="*")

            def adjust():
                state["scen"] = state["obj"].adjust(sim, time=ts, interp="nearest", extrapolation="constant", rank_window=rw)
            tr = reduce_max(time_device(torch, train, steps))
            ad = reduce_max(time_device(torch, adjust, steps))
            key = "rank_window_true" if rw else "rank_window_false"
            # 56.0 B per gp*day for cfg3 with the factors materialised (af + hist_q written, af read)
            bytes_step = n * (3 * T * 4 + T * 4 + 3 * 365 * 100 * 4)
            res[key] = {"train_ms": tr, "adjust_ms": ad, "value": world * n * T / ((tr + ad) * 1e-3), "unit": UNIT,
                        "roofline": {"bound": "hbm", "achieved": bytes_step / ((tr + ad) * 1e-3) / 1e9, "peak": peak,
                                     "unit": "GB/s", "frac": bytes_step / ((tr + ad) * 1e-3) / 1e9 / peak,
                                     "algorithmic_bytes_per_gp_day": bytes_step / (n * T)}}
            del state
        # end to end through the host entry (QDM mode), 2 lat rows per rank
        n_e = 2 * NLON
        hb = [t[:, :n_e].contiguous().cpu().pin_memory() for t in (ref, hist, sim)]
        out_h = torch.empty((T, n_e), dtype=torch.float32).pin_memory()
        ts_e = []
        for i in range(1 + steps):
            t0 = time.perf_counter()
            xs.train_adjust_host(hb[0].numpy(), hb[1].numpy(), hb[2].numpy(), time=tt, sim_time=ts, nquantiles=100,
                                 group="time.dayofyear", window=31, kind="*", method="qdm", slab_points=1024,
                                 out=out_h.numpy())
            _ = float(out_h[::997, ::101].nansum())
            if i:
                ts_e.append(time.perf_counter() - t0)
        res["e2e"] = {"value": world * n_e * T / (reduce_max(1e3 * sum(ts_e) / len(ts_e)) * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": 3 * n_e * T * 4, "d2h_bytes_per_step": n_e * T * 4,
                      "sample": f"{n_e} gridpoints per rank, mean of {len(ts_e)} calls"}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import cpu_baseline
            cores = os.cpu_count() or 1
            r = cpu_baseline.run_cfg3(4 * cores, cores)
            res["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "seconds")}
        out["cfg3"] = res
        del ref, hist, sim, hb, out_h
        torch.cuda.empty_cache()

    if "cfg4" in want:
        g = xs.Grouper("time.dayofyear", 31)
        doy = torch.from_numpy(tt.dayofyear.astype(np.float32)).to(dev)
        year = torch.from_numpy((tt.year - tt.year[0]).astype(np.float32)).to(dev)
        ref, hist, sim = (synth_slab(torch, gen, T, args.cfg_rows, 300, w, doy, year, dev) for w in ("ref", "hist", "sim"))
        state = {}
        loess = xs.LoessDetrend(group="time", f=0.2, niter=1, d=0)

        def train():
            state["obj"] = xs.DetrendedQuantileMapping.train(ref, hist, time=tt, nquantiles=50, group=g, kind="+")

        def adjust():
            state["scen"] = state["obj"].adjust(sim, time=ts, detrend=loess)
        tr = reduce_max(time_device(torch, train, steps))
        ad = reduce_max(time_device(torch, adjust, steps))
        res = {"workload": f"DQM kind=+ tas f32, Grouper(time.dayofyear, 31), nq=50, LoessDetrend(group=time, f=0.2, "
                           f"niter=1, d=0), {n} gridpoints x {T} days per rank (weak over ranks)",
               "tas": {"train_ms": tr, "adjust_ms": ad, "value": world * n * T / ((tr + ad) * 1e-3), "unit": UNIT}}
        del ref, hist, sim, state
        torch.cuda.empty_cache()
        ref, hist, sim = (synth_pr(torch, gen, T, n, w, dev) for w in ("ref", "hist", "sim"))
        state = {}

        def train_pr():
            state["obj"] = xs.DetrendedQuantileMapping.train(ref, hist, time=tt, nquantiles=50, group=g, kind="*",
                                                             jitter_under_thresh_value="0.01 mm/d")

        def adjust_pr():
            state["scen"] = state["obj"].adjust(sim, time=ts, detrend=loess)
        tr = reduce_max(time_device(torch, train_pr, steps))
        ad = reduce_max(time_device(torch, adjust_pr, steps))
        res["pr_jitter_under_thresh"] = {"train_ms": tr, "adjust_ms": ad, "value": world * n * T / ((tr + ad) * 1e-3),
                                         "unit": UNIT}
        out["cfg4"] = res
        del ref, hist, sim, state
        torch.cuda.empty_cache()

    if "cfg5" in want:
        # MBCn, 5 variables, n_iter=20; the reference refuses group='time.month' for MBCn (adjustment.py:1851-1852), so
        # the block structure is group='time' (SURVEY section 8 note)
        nm = NLON
        doy = torch.from_numpy(tt.dayofyear.astype(np.float32)).to(dev)
        year = torch.from_numpy((tt.year - tt.year[0]).astype(np.float32)).to(dev)

        def mk(which):
            return torch.stack([synth_slab(torch, gen, T, 1, 300 + v, which, doy, year, dev).nan_to_num_(280.0)
                                for v in range(5)])
        ref5, hist5, sim5 = mk("ref"), mk("hist"), mk("sim")
        state = {}

        def train():
            state["obj"] = xs.MBCn.train(ref5, hist5, time=tt, base_kws={"nquantiles": 20, "group": "time"}, n_iter=20,
                                         seed=1)

        def adjust():
            state["scen"] = state["obj"].adjust(sim5, ref5, hist5, time=tt)
        tr = reduce_max(time_device(torch, train, 1))
        ad = reduce_max(time_device(torch, adjust, 1))
        out["cfg5"] = {"workload": f"MBCn 5 variables f32, n_iter=20, nq=20, group=time, {nm} gridpoints x {T} days per "
                                   "rank (weak over ranks)",
                       "train_ms": tr, "adjust_ms": ad, "value": world * nm * T / ((tr + ad) * 1e-3), "unit": UNIT}
        del ref5, hist5, sim5, state
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import xsdba_b200 as xs
    from xsdba_b200 import _lib
    from xsdba_b200.sharding import lat_band, slabs as slab_list

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: xsdba_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to the C stdout when the first communicator is created: keep stdout for the
        # ONE JSON line and send whatever the libraries say during initialisation to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
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

    row_lo, row_hi = lat_band(args.lat_rows, rank, world)              # this rank's band of the ONE grid
    band = [(row_lo + r0, nr) for r0, nr in slab_list(row_hi - row_lo, args.slab_rows)]
    full = slab_list(args.lat_rows, args.slab_rows)                     # (weak figure: every rank the full grid)
    max_pts = args.slab_rows * NLON
    af = torch.empty((max_pts, G, NQ), dtype=torch.float32, device=dev)
    hq = torch.empty_like(af)
    scen = torch.empty((T, max_pts), dtype=torch.float32, device=dev)

    def one_step(slabs, timed):
        """-> (train_ms, adjust_ms, checksum) summed over the given slabs."""
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
                chk += float(torch.nansum(scen.view(-1)[: T * n].view(T, n)[:: 997, :: 101]))
            del ref, hist, sim
        return tr_ms, ad_ms, chk

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    for _ in range(args.warmup):
        one_step(band, False)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = lib.xsdba_launch_count()
    sampler.active = True
    wall0 = time.perf_counter()
    tr_tot = ad_tot = 0.0
    chk = 0.0
    for _ in range(args.steps):
        a, b, chk = one_step(band, True)
        tr_tot += a
        ad_tot += b
    barrier()
    wall = time.perf_counter() - wall0
    sampler.active = False
    launches = lib.xsdba_launch_count() - l0
    tot_ms, tr_ms, ad_ms = reduce_max([tr_tot + ad_tot, tr_tot, ad_tot])
    n_pts_grid = args.lat_rows * NLON
    n_pts_rank = (row_hi - row_lo) * NLON
    value = n_pts_grid * T * args.steps / (tot_ms * 1e-3)

    weak = None
    if world > 1 and not args.no_weak:
        barrier()
        a, b, _ = one_step(full, False)
        (w_ms,) = reduce_max([a + b])
        weak = {"value": world * n_pts_grid * T / (w_ms * 1e-3), "unit": UNIT, "ms_per_step": w_ms,
                "what": "every rank processes its own full 721-row grid (one step)"}

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    ncu_ratio = ncu_traffic_ratio(os.path.join(ROOT, "profiles", "r02_ncu_train_bucket.txt"), "train_bucket_kernel")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n_launch = len(band) * args.steps
    train_bytes_step = n_pts_rank * (2 * T * 4 + 2 * G * NQ * 4)          # reads ref+hist, writes af+hist_q
    adjust_bytes_step = n_pts_rank * (2 * T * 4 + 2 * G * NQ * 4)         # reads sim+tables, writes scen
    dom = "train" if tr_ms >= ad_ms else "adjust"
    dom_bytes = train_bytes_step if dom == "train" else adjust_bytes_step
    dom_ms = max(tr_ms, ad_ms)
    achieved = dom_bytes * args.steps / (dom_ms * 1e-3) / 1e9
    train_kernel = "train_fast_kernel<false, false>" if os.environ.get("XSDBA_B200_TRAIN_ALGO") == "sort" \
        else "train_bucket_kernel<false, false>"
    roofline = {
        "bound": "hbm", "kernel": train_kernel if dom == "train" else "pack_tables_kernel + adjust_tile_kernel",
        "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650",
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE `ncu --set full` capture of this kernel (the committed
        # summary is parsed at run time, so the figure follows the capture and is null without it), scaled from the
        # captured launch to the mean launch of this run; bench.py itself never runs under a profiler
        "traffic": (ncu_ratio * dom_bytes * args.steps / n_launch) if ncu_ratio and dom == "train" and "bucket" in train_kernel else None,
        "traffic_source": "profiles/r02_ncu_train_bucket.txt, parsed at run time: dram bytes per launch = "
                          + (f"{ncu_ratio:.3f}" if ncu_ratio else "n/a") + " x algorithmic on the captured 69 120-point launch",
        "launches": n_launch, "avg_launch_ms": dom_ms / n_launch,
        "algorithmic_bytes_per_launch": dom_bytes * args.steps / n_launch,
        "step": {"train_ms": tr_ms / args.steps, "adjust_ms": ad_ms / args.steps,
                 "train_GBps": train_bytes_step * args.steps / (tr_ms * 1e-3) / 1e9,
                 "adjust_GBps": adjust_bytes_step * args.steps / (ad_ms * 1e-3) / 1e9,
                 "whole_step_frac_of_peak": (train_bytes_step + adjust_bytes_step) * args.steps / (tot_ms * 1e-3) / 1e9 / peak,
                 "fused_floor_frac_of_peak": n_pts_rank * 16 * T * args.steps / (tot_ms * 1e-3) / 1e9 / peak},
    }

    # ---- PCIe ceiling: concurrent pinned host -> device copies on all ranks -----------------------
    pin = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dst.copy_(pin, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        dst.copy_(pin, non_blocking=True)
    torch.cuda.synchronize()
    (h2d_s,) = reduce_max([time.perf_counter() - t0])
    pcie = {"h2d_GBps_per_rank_concurrent": 4 * (256 << 20) / h2d_s / 1e9, "ranks": world,
            "what": "4 x 256 MiB pinned host -> device copies issued by all ranks at once (slowest rank)"}
    del pin, dst

    # ---- end to end through the host C ABI (pinned host buffers, copies inside the timed region) ---
    mem_gb = 0.0
    try:
        mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_AVPHYS_PAGES") / 2**30
    except (ValueError, OSError):
        pass
    e2e_rows = args.e2e_rows
    while e2e_rows > 1 and 4 * e2e_rows * NLON * T * 4 * world / 2**30 > 0.5 * mem_gb:
        e2e_rows //= 2   # pinned ref/hist/sim/scen of all ranks must fit comfortably in host memory
    n_e2e = e2e_rows * NLON
    e2e = None
    if n_e2e > 0:
        hb = []
        for w in ("ref", "hist", "sim"):      # the NaN-seeded generator of the device path, copied to pinned memory
            t = torch.empty((T, n_e2e), dtype=torch.float32).pin_memory()
            for r0 in range(0, e2e_rows, 8):
                nr = min(8, e2e_rows - r0)
                t[:, r0 * NLON:(r0 + nr) * NLON].copy_(synth_slab(torch, gen, T, nr, 300 + r0, w, doy, year, dev))
            hb.append(t)
        out_h = torch.empty((T, n_e2e), dtype=torch.float32).pin_memory()
        ref_h, hist_h, sim_h = (t.numpy() for t in hb)
        e2e_times = []
        for i in range(1 + max(1, args.steps)):
            barrier()
            t0 = time.perf_counter()
            xs.train_adjust_host(ref_h, hist_h, sim_h, time=t_train, sim_time=t_sim, nquantiles=NQ, group=GROUP,
                                 kind="+", method="eqm", slab_points=2048, out=out_h.numpy())
            chk_e2e = float(out_h[::997, ::101].nansum())  # the device->host result is read
            if i:
                e2e_times.append(time.perf_counter() - t0)
        (e2e_s,) = reduce_max([sum(e2e_times) / len(e2e_times)])
        e2e = {"value": world * n_e2e * T / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": 3 * n_e2e * T * 4, "d2h_bytes_per_step": n_e2e * T * 4,
               "sample": f"{n_e2e} gridpoints x {T} days per rank (NaN-seeded) through xsdba_qm_train_adjust_host_f32 "
                         f"(pinned host ref/hist/sim -> scen), mean of {len(e2e_times)} timed calls after 1 warm-up",
               "pcie_floor_value": world * n_e2e * T / (3 * n_e2e * T * 4 / (pcie["h2d_GBps_per_rank_concurrent"] * 1e9)),
               "checksum": chk_e2e}
        del hb, out_h

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import cpu_baseline
        cores = os.cpu_count() or 1
        r = cpu_baseline.run(args.cpu_points or 2048 * cores, cores, NYEARS, NQ, GROUP)  # ~15-20 s of CPU work
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
               "seconds": r["seconds"]}

    del af, hq, scen
    torch.cuda.empty_cache()
    want = [] if args.configs in ("", "none") else args.configs.split(",")
    configs = run_other_configs(args, torch, dist, xs, lib, dev, rank, world, peak, want) if want else None

    sampler.stop_flag = True
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world), "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "pcie": pcie, "weak": weak, "configs": configs,
            "gpu_launches": int(launches), "clocks": sampler.summary(), "host_cores": os.cpu_count(),
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
