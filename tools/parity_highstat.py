#!/usr/bin/env python
"""Statistical parity at the north star's size: N GPU events against N reference events, with a reference-vs-reference
null control of the same size and the same estimators.

For BASELINE config 1 (ER flat 0.1-20 keV) and config 2 (NR flat 0-100 keV), LUX_Run03 at 180 V/cm:
  * band mean and width per S1 bin: chi^2 with p-values.  The error of a width is taken from the bin's own fourth
    moment, se(s) = s sqrt((m4/s^4 - 1) / (4 N)) -- NOT the Gaussian s/sqrt(2N): GetBand pushes y = 0 for every
    event whose log10(S2/S1) falls outside (logMin, logMax) (execNEST.cpp:1551-1555), which fattens the tails of
    some bins, and with the Gaussian error even reference-vs-reference gives chi^2/dof > 1 there.
  * KS distance of Nph, Ne, S1c (spike count, corrected) and S2c from fine histograms, p from the Kolmogorov law
  * ER leakage below the NR band's centroid per S1 bin (examples/rootNEST.cpp:880-915), GPU vs reference

The reference side = the unmodified reference (oracle/_ref/libnestprobe.so: the loop body of execNEST() over the
reference's own objects), one worker process per host core with distinct seeds (its RandomGen is a process-wide
singleton).  TEST INFRASTRUCTURE: nothing here is on the product path.

usage: parity_highstat.py [--events 1e8] [--out profiles/r02_parity_1e8.txt] [--procs P]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # name: (species, source kind, eMin, eMax, er_model)
    "config1_ER_0.1-20keV": ("beta", "FLAT", 0.1, 20.0, 0),
    "config2_NR_0-100keV": ("NR", "FLAT", 0.0, 100.0, -1),
}
YB = 600          # log10(S2/S1) histogram bins over (logMin, logMax): leakage counting
H1 = {            # 1-D histograms for the KS distances: (lo, hi, bins)
    "Nph": (0.0, 8000.0, 8000), "Ne": (0.0, 4000.0, 4000), "S1c": (-5.0, 245.0, 5000), "S2c": (-200.0, 49800.0, 5000)}


class Acc:
    """what one sample accumulates: per-S1-bin moment sums of y, a (bin, y) histogram, 1-D histograms"""

    def __init__(self, nb):
        self.n = 0
        self.m = np.zeros((nb, 6))            # N, sum y, y^2, y^3, y^4, rejected
        self.h2 = np.zeros((nb, YB), np.int64)
        self.sx = np.zeros(nb)
        self.h1 = {k: np.zeros(v[2] + 2, np.int64) for k, v in H1.items()}

    def add(self, o):
        self.n += o.n
        self.m += o.m
        self.h2 += o.h2
        self.sx += o.sx
        for k in self.h1:
            self.h1[k] += o.h1[k]
        return self


def accumulate(ana, coin_level, nph, ne, s1c, s2c, sig1, sig2):
    """GetBand's binning and accept rules (execNEST.cpp:1508-1580) on per-event signals -> Acc"""
    nb = ana.numBins
    a = Acc(nb)
    a.n = len(sig1)
    width, border = (ana.maxS1 - ana.minS1) / nb, ana.minS1
    aX = np.abs(sig1)
    j = np.ceil((aX - border) / width).astype(np.int64) - 1       # aX in (lo, lo + width]
    inb = (j >= 0) & (j < nb) & (aX > border)
    good = inb & (sig1 >= 0.0) & (sig2 >= 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.log10(sig2[good] / sig1[good])
    ok = (sig1[good] != 0) & (sig2[good] != 0) & (v > ana.logMin) & (v < ana.logMax)
    y = np.where(ok, v, 0.0)
    jg = j[good]
    for p in range(5):
        a.m[:, p] = np.bincount(jg, weights=y ** p, minlength=nb)[:nb]
    a.m[:, 5] = np.bincount(j[inb & ~good], minlength=nb)[:nb]
    a.sx = np.bincount(jg, weights=sig1[good], minlength=nb)[:nb]
    yy, jj = y[ok], jg[ok]
    kb = np.clip(((yy - ana.logMin) / (ana.logMax - ana.logMin) * YB).astype(np.int64), 0, YB - 1)
    a.h2 = np.bincount(jj * YB + kb, minlength=nb * YB).reshape(nb, YB)
    for k, x in (("Nph", nph), ("Ne", ne), ("S1c", s1c), ("S2c", s2c)):
        lo, hi, bins = H1[k]
        idx = np.clip(np.floor((np.asarray(x, np.float64) - lo) / (hi - lo) * bins).astype(np.int64) + 1, 0, bins + 1)
        a.h1[k] = np.bincount(idx, minlength=bins + 2)
    return a


def ref_worker(task):
    """one chunk of the reference: (config name, seed, n) -> Acc"""
    import refprobe
    from helpers import ref_chain
    from nest_b200 import capi
    name, seed, n = task
    sp, kind, e0, e1, erm = CONFIGS[name]
    ref = refprobe.reference()
    r = ref_chain(ref, seed, n, getattr(capi, sp), getattr(capi, "SRC_" + kind), e0, e1, er_model=erm)
    C = refprobe.COL
    ana = capi.analysis(0)
    return name, accumulate(ana, 2, r[:, C["Nph"]], r[:, C["Ne"]], r[:, refprobe.S1_0 + 7], r[:, refprobe.S2_0 + 7],
                            r[:, refprobe.SIG1], r[:, refprobe.SIG2])


def gpu_sample(ctx, name, n_total, seed, chunk=4_000_000):
    from nest_b200 import capi
    sp, kind, e0, e1, erm = CONFIGS[name]
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    mo = capi.model(1)
    mo.er_model = erm
    if getattr(capi, sp) >= 6:
        mo.massNum = det.molarMass
    ana = capi.analysis(0)
    src = capi.source(getattr(capi, "SRC_" + kind), getattr(capi, sp), e0, e1, 180.0)
    tot = Acc(ana.numBins)
    done = 0
    while done < n_total:
        m = min(chunk, n_total - done)
        c = ctx.run(det, mo, ana, src, rc, seed, m, first_event=done,
                    columns=["photons", "electrons", "s1_7", "s2_7", "signal1", "signal2"])
        tot.add(accumulate(ana, det.coinLevel, c["photons"], c["electrons"], c["s1_7"], c["s2_7"], c["signal1"], c["signal2"]))
        done += m
    return tot


def band_stats(a):
    """per bin: N, mean, sd, se(mean), se(sd) from the fourth moment"""
    N = a.m[:, 0]
    with np.errstate(divide="ignore", invalid="ignore"):
        mu = a.m[:, 1] / N
        m2 = a.m[:, 2] / N - mu ** 2
        m4 = (a.m[:, 4] - 4 * mu * a.m[:, 3] + 6 * mu ** 2 * a.m[:, 2] - 4 * mu ** 3 * a.m[:, 1]) / N + mu ** 4
        sd = np.sqrt(m2 * N / (N - 1))
        se_mu = sd / np.sqrt(N)
        se_sd = sd * np.sqrt(np.maximum(m4 / m2 ** 2 - 1.0, 0.0) / (4.0 * N))
        se_sd_gauss = sd / np.sqrt(2.0 * N)
    return N, mu, sd, se_mu, se_sd, se_sd_gauss


def chi2_pair(a, b, min_n=2000):
    from scipy import stats
    Na, ma, sa, ema, esa, ega = band_stats(a)
    Nb, mb, sb, emb, esb, egb = band_stats(b)
    ok = (Na >= min_n) & (Nb >= min_n)
    dof = int(ok.sum())
    cm = float((((ma - mb) ** 2) / (ema ** 2 + emb ** 2))[ok].sum())
    cw = float((((sa - sb) ** 2) / (esa ** 2 + esb ** 2))[ok].sum())
    cwg = float((((sa - sb) ** 2) / (ega ** 2 + egb ** 2))[ok].sum())
    acc = [float(Na.sum() / a.n), float(Nb.sum() / b.n)]
    return {"dof": dof, "chi2_mean": cm, "p_mean": float(stats.chi2.sf(cm, dof)), "chi2_width": cw,
            "p_width": float(stats.chi2.sf(cw, dof)), "chi2_width_gaussian_error": cwg,
            "p_width_gaussian_error": float(stats.chi2.sf(cwg, dof)),
            "max_abs_mean_diff": float(np.abs(ma - mb)[ok].max()), "max_rel_width_diff": float(np.abs(sa / sb - 1)[ok].max()),
            "accepted_fraction": acc,
            "accepted_z": float((acc[0] - acc[1]) / np.sqrt(acc[0] * (1 - acc[0]) / a.n + acc[1] * (1 - acc[1]) / b.n))}


def ks_pair(a, b):
    from scipy import stats
    out = {}
    for k in H1:
        ca, cb = np.cumsum(a.h1[k]) / a.h1[k].sum(), np.cumsum(b.h1[k]) / b.h1[k].sum()
        D = float(np.abs(ca - cb).max())
        ne = a.h1[k].sum() * b.h1[k].sum() / (a.h1[k].sum() + b.h1[k].sum())
        out[k] = {"D": D, "p": float(stats.kstwobign.sf(D * np.sqrt(ne)))}
    return out


def leakage(er, nr, ana):
    """fraction of ER events below the NR band mean per S1 bin, counted on the (bin, y) histogram with the bin holding
    the centroid split linearly; events pushed as y = 0 count as below (rootNEST.cpp:880-915, NRbandCenter = 1)"""
    _, mu_nr, *_ = band_stats(nr)
    nb = ana.numBins
    w = (ana.logMax - ana.logMin) / YB
    below = np.zeros(nb)
    tot = er.m[:, 0].copy()
    for j in range(nb):
        if not np.isfinite(mu_nr[j]):
            continue
        t = (mu_nr[j] - ana.logMin) / w
        k = int(np.clip(np.floor(t), 0, YB - 1))
        inhist = er.h2[j].sum()
        below[j] = er.h2[j, :k].sum() + er.h2[j, k] * (t - k) + (tot[j] - inhist)   # y = 0 pushes: below
    return below, tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", type=float, default=1e8)
    ap.add_argument("--chunk", type=float, default=250000)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_parity_1e8.txt"))
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    from nest_b200 import capi
    n, chunk = int(a.events), int(a.chunk)
    procs = a.procs or len(os.sched_getaffinity(0))
    ana = capi.analysis(0)
    t0 = time.time()
    # reference: two independent samples (A, B) of n events per configuration, in chunks over the worker pool
    tasks, nchunks = [], (n + chunk - 1) // chunk
    for name in CONFIGS:
        for s, base in (("A", 1_000_000), ("B", 2_000_000)):
            for i in range(nchunks):
                tasks.append(((name, s), (name, base + i, min(chunk, n - i * chunk))))
    ref = {(name, s): Acc(ana.numBins) for name in CONFIGS for s in "AB"}
    with mp.get_context("spawn").Pool(procs) as pool:
        key_of = [k for k, _ in tasks]
        for i, (name, acc) in enumerate(pool.imap(ref_worker, [t for _, t in tasks], chunksize=1)):
            ref[key_of[i]].add(acc)
    t_ref = time.time() - t0
    t1 = time.time()
    ctx = capi.Context(0)
    gpu = {name: gpu_sample(ctx, name, n, seed=20261020) for name in CONFIGS}
    ctx.close()
    t_gpu = time.time() - t1

    res = {"events_per_sample": n, "host_cores": procs, "reference_seconds": t_ref, "gpu_seconds_incl_host_binning": t_gpu}
    for name in CONFIGS:
        res[name] = {"gpu_vs_ref": {"band": chi2_pair(gpu[name], ref[(name, "A")]), "ks": ks_pair(gpu[name], ref[(name, "A")])},
                     "gpu_vs_refB": {"band": chi2_pair(gpu[name], ref[(name, "B")])},
                     "ref_vs_ref": {"band": chi2_pair(ref[(name, "A")], ref[(name, "B")]),
                                    "ks": ks_pair(ref[(name, "A")], ref[(name, "B")])}}
    er, nr = "config1_ER_0.1-20keV", "config2_NR_0-100keV"
    lk = {}
    for tag, e_, n_ in (("gpu", gpu[er], gpu[nr]), ("refA", ref[(er, "A")], ref[(nr, "A")]), ("refB", ref[(er, "B")], ref[(nr, "B")])):
        lk[tag] = leakage(e_, n_, ana)
    from scipy import stats
    def leak_cmp(x, y):
        (bx, tx), (by, ty) = lk[x], lk[y]
        ok = (tx >= 2000) & (ty >= 2000)
        px, py = bx[ok] / tx[ok], by[ok] / ty[ok]
        var = px * (1 - px) / tx[ok] + py * (1 - py) / ty[ok]
        z2 = float(((px - py) ** 2 / np.maximum(var, 1e-300)).sum())
        return {"bins": int(ok.sum()), "chi2": z2, "p": float(stats.chi2.sf(z2, int(ok.sum()))),
                "total": [float(bx[ok].sum() / tx[ok].sum()), float(by[ok].sum() / ty[ok].sum())]}
    res["ER_leakage_below_NR_mean"] = {"gpu_vs_refA": leak_cmp("gpu", "refA"), "refA_vs_refB": leak_cmp("refA", "refB")}

    lines = ["python tools/parity_highstat.py --events %g   (B200 box, %d host cores; reference %.0f s, GPU + host binning %.0f s)" %
             (n, procs, t_ref, t_gpu),
             "%d GPU events vs %d reference events per configuration, and reference (seeds A) vs reference (seeds B) as the null" % (n, n),
             "band: bins with >= 2000 events on both sides; width error from each bin's fourth moment (see the script's header);",
             "      the Gaussian-error chi^2 (sigma/sqrt(2N), what round 1 used) is listed to show why it read ~1.25 per dof", ""]
    for name in CONFIGS:
        lines.append("== %s" % name)
        for pair in ("gpu_vs_ref", "gpu_vs_refB", "ref_vs_ref"):
            b = res[name][pair]["band"]
            lines.append("  %-11s dof %3d | mean chi2 %7.1f p=%.3f | width chi2 %7.1f p=%.3f | width chi2 (Gaussian error) %7.1f p=%.2g | "
                         "max |d mean| %.5f dex, max |d width| %.3f %% | in-window fraction %.5f / %.5f (z = %+.2f)" %
                         (pair, b["dof"], b["chi2_mean"], b["p_mean"], b["chi2_width"], b["p_width"],
                          b["chi2_width_gaussian_error"], b["p_width_gaussian_error"], b["max_abs_mean_diff"],
                          100 * b["max_rel_width_diff"], b["accepted_fraction"][0], b["accepted_fraction"][1], b["accepted_z"]))
        for pair in ("gpu_vs_ref", "ref_vs_ref"):
            lines.append("  %-11s KS  " % pair + "  ".join("%s D=%.2e p=%.3f" % (k, v["D"], v["p"]) for k, v in res[name][pair]["ks"].items()))
        lines.append("")
    lines.append("== ER leakage below the NR band mean, per S1 bin (rootNEST.cpp:880-915)")
    for k, v in res["ER_leakage_below_NR_mean"].items():
        lines.append("  %-13s bins %3d chi2 %7.1f p=%.3f | total leakage %.6f / %.6f" % (k, v["bins"], v["chi2"], v["p"], v["total"][0], v["total"][1]))
    txt = "\n".join(lines) + "\n"
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write(txt)
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)
    print(txt)


if __name__ == "__main__":
    main()
