#!/usr/bin/env python
"""Follow-up to profiles/r02_parity_1e8.txt: the in-window fraction of config 1 (ER 0.1-20 keV) read 0.83960 on the GPU
against 0.83947 on both reference samples (z = +2.4 each).  Here the per-S1-bin accepted and rejected counts of N
reference events (fresh seeds) are compared with those of 2e9 GPU events, bin by bin, to tell a fluctuation of the
reference samples from a bias and, if it is one, where it sits (S1 threshold, coincidence, upper S1 cut).
TEST INFRASTRUCTURE.  usage: acceptance_study.py [--events 4e8] [--config config1_ER_0.1-20keV]"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import parity_highstat as ph
    from nest_b200 import capi
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", type=float, default=4e8)
    ap.add_argument("--gpu-events", type=float, default=2e9)
    ap.add_argument("--config", default="config1_ER_0.1-20keV")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_acceptance_study.txt"))
    a = ap.parse_args()
    name, n, chunk = a.config, int(a.events), 250000
    ana = capi.analysis(0)
    nb = ana.numBins
    ref = ph.Acc(nb)
    tasks = [(name, 7_000_000 + i, chunk) for i in range(n // chunk)]
    with mp.get_context("spawn").Pool(len(os.sched_getaffinity(0))) as pool:
        for _, acc in pool.imap_unordered(ph.ref_worker, tasks, chunksize=1):
            ref.add(acc)
    # GPU: band accumulators only (device resident), 4 seeds
    sp, kind, e0, e1, erm = ph.CONFIGS[name]
    ctx = capi.Context(0)
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    mo = capi.model(1)
    mo.er_model = erm
    if getattr(capi, sp) >= 6:
        mo.massNum = det.molarMass
    src = capi.source(getattr(capi, "SRC_" + kind), getattr(capi, sp), e0, e1, 180.0)
    ng = int(a.gpu_events)
    acc_g, rej_g = np.zeros(nb), np.zeros(nb)
    for seed in (21, 22, 23, 24):
        ctx.band_reset(ana)
        done = 0
        while done < ng // 4:
            m = min(100_000_000, ng // 4 - done)
            ctx.run_async(det, mo, ana, src, rc, seed, done, m)
            done += m
        b = capi.Band(ana)
        ctx.band_fetch(b)
        acc_g += b.accepted
        rej_g += b.rejected
    ctx.close()
    acc_r, rej_r = ref.m[:, 0], ref.m[:, 5]
    lines = ["python tools/acceptance_study.py --events %g --config %s   (B200 box, %d host cores)" % (n, name, len(os.sched_getaffinity(0))),
             "%d reference events (fresh seeds) vs %d GPU events: per-S1-bin accepted fractions of all simulated events" % (ref.n, ng), ""]
    fr, fg = acc_r.sum() / ref.n, acc_g.sum() / ng
    se = np.sqrt(fr * (1 - fr) / ref.n + fg * (1 - fg) / ng)
    lines.append("in-window fraction: GPU %.6f  reference %.6f  difference %+.2e +- %.1e (z = %+.2f)" % (fg, fr, fg - fr, se, (fg - fr) / se))
    rr, rg = rej_r.sum() / ref.n, rej_g.sum() / ng
    se = np.sqrt(rr / ref.n + rg / ng)
    lines.append("binned-but-rejected fraction (negative S1 or S2 flag): GPU %.6f  reference %.6f  (z = %+.2f)" % (rg, rr, (rg - rr) / se))
    pr, pg = acc_r / ref.n, acc_g / ng
    z = (pg - pr) / np.sqrt(pr / ref.n + pg / ng)
    lines.append("per-bin z of the accepted fraction (98 S1 bins, 1.5..99.5 spikes): chi2 = %.1f" % float((z ** 2).sum()))
    for lo in range(0, nb, 14):
        lines.append("  bins %2d-%2d: " % (lo, min(lo + 13, nb - 1)) + " ".join("%+5.2f" % v for v in z[lo:lo + 14]))
    cum = np.cumsum(pg - pr)
    lines.append("cumulative difference of the accepted fraction after bins 7, 21, 49, 97: " + ", ".join("%+.2e" % cum[k] for k in (7, 21, 49, 97)))
    txt = "\n".join(lines) + "\n"
    open(a.out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
