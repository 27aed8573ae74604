#!/usr/bin/env python
"""bench.py -- simulated events/s of the NEST detector-response chain on B200.

Workload (BASELINE.json configs[1]): NR, flat 0-100 keV, LUX_Run03 at 180 V/cm, random vertex with
the drift-time fiducial retry, Full S1 (per-hit loop, spike counting, position-dependent LCE),
Full S2, energy reconstruction and the log10(S2/S1)-vs-S1 band -- 1e8 events per step per GPU
(weak scaling: rank r simulates event ids [r*n, (r+1)*n), band accumulators summed by a NCCL
all-reduce every step).

One JSON line on stdout (rank 0).  `value` = events/s with everything resident on the device
(band-only output); `e2e` = the same chain through the C-ABI call nb200_run with HOST buffers:
the per-event record the reference CLI prints (execNEST.cpp:1386-1405) copied back to pinned
host memory plus the band.  `--impl reference` times the unmodified reference binary
(oracle/_ref/execNEST_ref_lux) on the host cores, one process per core.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "simulated events/sec (yields+S1+S2)"
UNIT = "events/s"
WORKLOAD = "NR flat 0-100 keV, LUX_Run03 180 V/cm, Full S1 (spike counting, position-dependent LCE) + Full S2 + recon + band"
# nominal DFMA-equivalent work of the reference algorithm per event
F_FIX_NR = 3600.0
F_HIT = 192.0
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "execNEST_ref_lux")




def bind_to_gpu_cpus(local):
    """Multi-rank runs: keep this rank (and the pinned buffers it first-touches) on the CPUs NVML reports as
    closest to its GPU -- what `numactl --cpunodebind` would do for a user.  Best effort; returns the set."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        mine = near & os.sched_getaffinity(0)
        if mine:
            os.sched_setaffinity(0, mine)
            return sorted(mine)
    except Exception:
        pass
    return None

def ncu_traffic():
    """DRAM bytes of one k_hits launch from the committed `ncu --set full` capture (profiles/): the kernel
    has no HBM stream of its own -- it reads hit counts and writes three words per event."""
    import glob
    import re
    best = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_hits_ncu_full.txt")))
    if not best:
        return {"traffic": None}
    tot = 0.0
    for ln in open(best[-1]):
        m = re.match(r"dram__bytes_(read|write)\.sum\s+(\S+)\s+([0-9.]+)", ln)
        if m:
            tot += float(m.group(3)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
    return {"traffic": tot, "traffic_unit": "bytes per launch (dram read + write)",
            "traffic_source": os.path.relpath(best[-1], ROOT)}

def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nest_b200", choices=["nest_b200", "reference"])
    ap.add_argument("--events", type=float, default=1e8, help="events per step per GPU")
    ap.add_argument("--e2e-events", type=float, default=2e7, help="events per end-to-end step per GPU")
    ap.add_argument("--ref-events", type=float, default=2e5, help="events per reference process per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---- reference arm ---------------------------------------------------------------------------

def run_reference_processes(nproc, n_events, seed0):
    """one unmodified execNEST process per core, distinct seeds (the RandomGen singleton forbids
    threads); returns wall seconds"""
    t0 = time.perf_counter()
    procs = []
    devnull = open(os.devnull, "w")
    for p in range(nproc):
        procs.append(subprocess.Popen([REF_BIN, str(int(n_events)), "NR", "0", "100", "180", "-1", str(seed0 + p)],
                                      stdout=devnull, stderr=devnull))
    for p in procs:
        p.wait()
    dt = time.perf_counter() - t0
    if any(p.returncode != 0 for p in procs):
        raise RuntimeError("reference process failed")
    return dt


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/execNEST_ref_lux not built"}))
        return 0
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = int(a.ref_events)
    for w in range(a.warmup):
        run_reference_processes(cores, max(n // 10, 1000), 1000 + w * cores)
    t = 0.0
    for s in range(a.steps):
        t += run_reference_processes(cores, n, 1 + s * cores)
    total = float(cores) * n * a.steps
    v = total / t
    sample = "%d processes x %d events x %d steps of `execNEST_ref_lux N NR 0 100 180 -1 seed` (stdout to /dev/null)" % (
        cores, n, a.steps)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "events_per_step": cores * n, "detector": "LUX_Run03"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))
    return 0


# ---- clocks ------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, text=True,
                                     timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---- our arm -------------------------------------------------------------------------------------

def main():
    a = parse()
    if a.impl == "reference":
        return reference_arm(a)

    import numpy as np
    import torch
    import torch.distributed as dist
    from nest_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the chain has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_cpus(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    capi.lib()
    ctx = capi.Context(local)
    # one non-default stream carries the chain kernel, torch's fills, the NCCL all-reduce and
    # the timing events (a null handle would select the library's own stream instead)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    mo = capi.model(1)
    ana = capi.analysis(0)
    src = capi.source(capi.SRC_FLAT, capi.NR, 0.0, 100.0, 180.0)
    n = int(a.events)
    seed = 1

    nwords = capi.lib().nb200_band_hist_bytes(C.byref(ana)) // 8
    band_t = torch.zeros(nwords, dtype=torch.int64, device="cuda")
    ctx.band_attach(ana, band_t.data_ptr())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step(i):
        band_t.zero_()
        first = (i * world + rank) * n
        ctx.run_async(det, mo, ana, src, rc, seed, first, n)
        if world > 1:
            dist.all_reduce(band_t)  # int64 sum == the u64 words' sum
        flush.fill_(i & 0xff)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(a.warmup):
        step(i)
    barrier()
    ctx.timing_enable(True)
    ctx.timing_read()
    sampler = ClockSampler(local)
    sampler.start()
    launches1 = ctx.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(a.warmup + i)
    e1.record()
    barrier()
    sampler.stop_flag = True
    stage_ms, stage_launches = ctx.timing_read()
    ctx.timing_enable(False)
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    gpu_launches = ctx.launch_count() - launches1
    value = float(n) * world * a.steps / (ms_total * 1e-3)

    # band sanity of the last step (all ranks' events of that step)
    band = capi.Band(ana)
    ctx.band_fetch(band)
    accepted = int(band.accepted.sum())

    # ---- end to end through nb200_run with pinned host buffers -------------------------------
    ne = int(a.e2e_events)
    cli_cols = ["signalE", "field", "driftTime", "smear_x", "smear_y", "z_mm", "photons", "electrons", "s1_2",
                "s1_5", "s1_7", "s2_0", "s2_4", "s2_7"]
    soa = capi.SoaOut()
    soa.mem_kind = capi.MEM_HOST
    pinned = {}
    d2h = 0
    for nme in cli_cols:
        is_int = nme in capi.SOA_I32
        t = torch.empty(ne, dtype=torch.int32 if is_int else torch.float64, pin_memory=True)
        pinned[nme] = t
        d2h += t.numel() * t.element_size()
        if nme.startswith("s1_") or nme.startswith("s2_"):
            getattr(soa, nme[:2])[int(nme[3:])] = t.data_ptr()
        else:
            setattr(soa, nme, t.data_ptr())
    hband = capi.Band(ana)
    d2h += nwords * 8
    h2d = C.sizeof(capi.Detector) + C.sizeof(capi.Model) + C.sizeof(capi.Analysis) + C.sizeof(capi.Source) + \
        C.sizeof(capi.RunConsts)

    def e2e_step(i):
        first = (10_000 + i * world + rank) * ne
        ctx._check(capi.lib().nb200_run(ctx.h, C.byref(det), C.byref(mo), C.byref(ana), C.byref(src), C.byref(rc),
                                        seed, first, ne, C.byref(soa), C.byref(hband.c)))

    e2e_steps = max(2, min(a.steps, 3))
    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    g0 = torch.cuda.Event(enable_timing=True)
    g1 = torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(e2e_steps):
        e2e_step(1 + i)
    g1.record()
    barrier()
    wall = time.perf_counter() - t0
    tt = torch.tensor([max(g0.elapsed_time(g1) * 1e-3, wall)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = float(ne) * world * e2e_steps / float(tt.item())

    # mean S1 hits per event of this workload (for the algorithmic flop count)
    smp = ctx.run(det, mo, ana, src, rc, seed, 200000, first_event=77_000_000_000, columns=["s1_0"])
    mean_hits = float(np.abs(smp["s1_0"]).mean())
    flops_per_event = F_FIX_NR + F_HIT * mean_hits

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "events_per_step_per_gpu": n, "detector": "LUX_Run03", "seed": seed,
                   "l2": "no HBM-resident input (events come from the counter-based RNG); a 256 MiB buffer is "
                         "rewritten between steps to flush L2 anyway",
                   "band_accepted_last_step": accepted},
        "gpu_launches": int(gpu_launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "events_per_step_per_gpu": ne, "steps": e2e_steps,
                "api": "nb200_run (C-ABI), pinned host SoA of the 14 CLI columns + band",
                "cpu_affinity": numa},
        "clocks": sampler.summary(),
    }

    if rank == 0 and world == 1:
        # the other BASELINE configurations on the same device (band-only, device resident), for information:
        # config 1 (ER flat 0.1-20 keV, the reference's own CPU-runnable case) and config 3 (CH3T, DD)
        others = {}
        for name, s_, er_model in [
                ("config1_ER_flat_0.1-20keV", capi.source(capi.SRC_FLAT, capi.beta, 0.1, 20.0, 180.0), 0),
                ("config3_CH3T", capi.source(capi.SRC_CH3T, capi.CH3T, 0.0, 18.5898, 180.0), -1),
                ("config3_DD", capi.source(capi.SRC_DD, capi.DD, 0.0, 80.0, 180.0, par=(13., 0.12, 71.2, 20., -20.5)), -1)]:
            m_ = capi.model(1)
            m_.er_model = er_model
            if s_.species >= 6:
                m_.massNum = det.molarMass
            n_ = 20_000_000
            for rep in range(2):  # first pass warms up
                band_t.zero_()
                t0_, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0_.record()
                ctx.run_async(det, m_, ana, s_, rc, seed, 0, n_)
                t1_.record()
                torch.cuda.synchronize()
            others[name] = n_ / (t0_.elapsed_time(t1_) * 1e-3)
        # config 5: LArNEST NR mean yields + default fluctuations, device-resident in and out (120 B/event:
        # E and field in, 9 yield columns and 4 fluctuation columns out) -- the HBM-bound member of the family
        n_ = 20_000_000
        E_ = torch.rand(n_, dtype=torch.float64, device="cuda") * 999.9 + 0.1
        F_ = torch.rand(n_, dtype=torch.float64, device="cuda") * 999.0 + 1.0
        Y_ = torch.empty((9, n_), dtype=torch.float64, device="cuda")
        Fl_ = torch.empty((4, n_), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        for rep in range(2):
            ctx.synchronize()
            t0w = time.perf_counter()
            ctx.lar_device(0, n_, E_.data_ptr(), F_.data_ptr(), 1.393, Y_.data_ptr(), Fl_.data_ptr(), seed=seed)
            ctx.synchronize()
            dtw = time.perf_counter() - t0w
        others["config5_LAr_NR_yields+fluctuations"] = n_ / dtw
        out["other_workloads_events_per_s"] = others
        out["lar_hbm_gbs"] = 120.0 * n_ / dtw / 1e9
        del E_, F_, Y_, Fl_
    if rank == 0:
        peak = ctx.fp64_peak()
        # dominant kernel: k_hits, the per-hit S1 loop.  Algorithmic work per launch = SURVEY 8(d)'s
        # F_hit = 192 DFMA-equivalent flop per PMT hit x the hits one launch processes.
        n_hit_launches = max(stage_launches[1], 1)
        hits_ms = stage_ms["k_hits"] / n_hit_launches
        events_per_launch = float(n) * a.steps / n_hit_launches
        achieved = events_per_launch * mean_hits * F_HIT / (hits_ms * 1e-3) / 1e12
        chain_ms = sum(stage_ms.values()) / a.steps
        out["roofline"] = {
            "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            **ncu_traffic(), "kernel": "k_hits", "kernel_ms_per_launch": hits_ms,
            "kernel_launches": int(n_hit_launches), "hits_per_launch": events_per_launch * mean_hits,
            "flops_per_hit": F_HIT, "mean_s1_hits": mean_hits,
            "stage_ms_per_step": {k: v / a.steps for k, v in stage_ms.items()},
            "kernel_share_of_step": stage_ms["k_hits"] / max(sum(stage_ms.values()), 1e-9),
            "chain_flops_per_event": flops_per_event,
            "chain_achieved_tflops": float(n) * flops_per_event / (chain_ms * 1e-3) / 1e12,
            "peak_source": "measured on this device by nb200_fp64_peak (register-resident DFMA chain, CUDA "
                           "events); MEASURED_PEAKS.json carries HBM and bf16 figures only (nominal FP64 37.2)",
            "note": "FP64 CUDA-core roofline, not HBM or tensor: no contraction, and in band-only mode the only HBM "
                    "stream is the 140 B/event stage workspace (see hbm_workspace_frac).  `achieved` counts the "
                    "REFERENCE algorithm's per-hit work (SURVEY 8d: 192 DFMA-equivalent flop per PMT hit: two "
                    "Box-Muller normals, a Bernoulli, ...) delivered per second; the kernel itself executes far "
                    "fewer FP64 instructions per hit (ziggurat normal, one normal per photo-electron, per-event "
                    "double-phe count): ncu shows its FP64 pipe ~11 % busy and 65 % of issue slots used "
                    "(profiles/r01_k_hits_ncu_full.txt, DESIGN.md section 5)",
            "hbm_workspace_frac": None}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                mp = json.load(f)
            out["roofline"]["hbm_workspace_frac"] = (value / world) * 140.0 * 2 / 1e9 / mp["hbm_gbs"]
            out["roofline"]["hbm_output_stream_frac_e2e"] = (e2e_value / world) * (d2h / ne) / 1e9 / mp["hbm_gbs"]
        except Exception:
            pass
        if world == 1 and not a.no_cpu_baseline and os.path.exists(REF_BIN):
            nref = 500000
            t0 = time.perf_counter()
            run_reference_processes(1, nref, 1)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": nref / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                                   "sample": "1 process, %d events of `execNEST_ref_lux N NR 0 100 180 -1 1` "
                                             "(unmodified reference, stdout to /dev/null)" % nref}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
