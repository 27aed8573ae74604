"""High-statistics parity with a null control: 4e7 GPU events against 1.6e7 events of the unmodified reference (two
independent samples A, B of 8e6, one worker per core), config 1 (ER 0.1-20 keV) and config 2 (NR 0-100 keV).

Same estimators as tools/parity_highstat.py (whose 1e8-vs-1e8 run is profiles/r02_parity_1e8.txt): band mean and width
chi^2 with the width error from each bin's fourth moment (GetBand's pushed zeros fatten the tails: with the Gaussian
sigma/sqrt(2N) even reference-vs-reference reads >1 per dof), KS distances of Nph / Ne / S1c / S2c, the in-window
fraction -- gated on p-values, and the reference-vs-reference comparison run beside it as the null."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
P_GATE = 1e-3


@pytest.fixture(scope="module")
def samples(ctx, ref):
    import parity_highstat as ph
    from nest_b200 import capi
    ana = capi.analysis(0)
    cores = min(len(os.sched_getaffinity(0)), 32)
    chunk, n_ref = 250_000, 8_000_000
    tasks = []
    for name in ph.CONFIGS:
        for s, base in (("A", 3_000_000), ("B", 4_000_000)):
            for i in range(n_ref // chunk):
                tasks.append(((name, s), (name, base + i, chunk)))
    acc = {(name, s): ph.Acc(ana.numBins) for name in ph.CONFIGS for s in "AB"}
    with mp.get_context("spawn").Pool(cores) as pool:
        for (key, _), (_, a) in zip(tasks, pool.imap(ph.ref_worker, [t for _, t in tasks], chunksize=1)):
            acc[key].add(a)
    gpu = {name: ph.gpu_sample(ctx, name, 40_000_000, seed=77) for name in ph.CONFIGS}
    return ph, ana, acc, gpu


@pytest.mark.parametrize("name", ["config1_ER_0.1-20keV", "config2_NR_0-100keV"])
def test_band_and_spectra_at_high_statistics(samples, name):
    ph, ana, acc, gpu = samples
    both = ph.Acc(ana.numBins).add(acc[(name, "A")]).add(acc[(name, "B")])      # 1.6e7 reference events
    b = ph.chi2_pair(gpu[name], both)
    null = ph.chi2_pair(acc[(name, "A")], acc[(name, "B")])
    msg = "%s gpu-vs-ref %r | null (ref-vs-ref) %r" % (name, b, null)
    assert b["dof"] >= 90, msg
    assert b["p_mean"] > P_GATE and b["p_width"] > P_GATE, msg
    assert null["p_mean"] > 1e-4 and null["p_width"] > 1e-4, msg          # the estimator itself is sound
    assert abs(b["accepted_z"]) < 4.0, msg
    assert b["max_abs_mean_diff"] < 2e-3 and b["max_rel_width_diff"] < 0.02, msg
    ks = ph.ks_pair(gpu[name], both)
    for k, v in ks.items():
        assert v["p"] > P_GATE, (name, k, v)


def test_er_leakage_below_the_nr_band_mean(samples):
    """examples/rootNEST.cpp:880-915: the fraction of ER events below the NR centroid per S1 bin, GPU vs reference"""
    from scipy import stats
    ph, ana, acc, gpu = samples
    er, nr = "config1_ER_0.1-20keV", "config2_NR_0-100keV"
    ref_er = ph.Acc(ana.numBins).add(acc[(er, "A")]).add(acc[(er, "B")])
    ref_nr = ph.Acc(ana.numBins).add(acc[(nr, "A")]).add(acc[(nr, "B")])
    bg, tg = ph.leakage(gpu[er], gpu[nr], ana)
    br, tr = ph.leakage(ref_er, ref_nr, ana)
    ok = (tg >= 2000) & (tr >= 2000)
    pg, pr = bg[ok] / tg[ok], br[ok] / tr[ok]
    var = pg * (1 - pg) / tg[ok] + pr * (1 - pr) / tr[ok]
    chi2 = float(((pg - pr) ** 2 / np.maximum(var, 1e-300)).sum())
    assert stats.chi2.sf(chi2, int(ok.sum())) > P_GATE, (chi2, int(ok.sum()))
    tot_g, tot_r = bg[ok].sum() / tg[ok].sum(), br[ok].sum() / tr[ok].sum()
    assert 1e-3 < tot_r < 2e-2 and abs(tot_g - tot_r) < 5 * np.sqrt(tot_r / tr[ok].sum() + tot_g / tg[ok].sum())
