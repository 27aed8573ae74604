"""WIMP source on the device (BASELINE config 4) against the compiled reference: TestSpectra::WIMP_spectrum
(src/TestSpectra.cpp:456-499) energies, the whole chain's 36 columns, and the command line with its
Poisson-fluctuated event count (src/execNEST.cpp:474-488).

The reference's spectrum table is itself random (one Xe isotope per WIMP_dRate call, :278): both sides use the SAME
table -- the reference's, rebuilt by nb200_wimp_prep_iso from the reference's own isotope draws (bit-identical:
tests/test_wimp_cpu.py)."""
import os

import numpy as np
import pytest
from scipy import stats

import refprobe
from helpers import gpu_chain, ks_report, ref_chain
from nest_b200 import capi
from test_gpu_cli import CLI, rows_of, run

pytestmark = pytest.mark.gpu
N_ISO = 1_000_000 + 1100


def table_of_reference_seed(ref, seed, mass, eStep=5.0):
    iso = ref.sampler(6, seed, N_ISO).astype(np.int32)
    return capi.WimpTable(mass, eStep, 0.0, isotopes=iso)


@pytest.mark.parametrize("mass,seed", [(50.0, 5), (8.0, 3), (200.0, 2)])
def test_wimp_energies_match_reference(ctx, ref, mass, seed):
    n = 200_000
    w = table_of_reference_seed(ref, seed, mass)
    E_ref = ref.wimp_spectrum(seed, 1000 + seed, n, mass)
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    src = w.attach(capi.source(capi.SRC_WIMP, capi.WIMP, mass, 1e-36, 180.0))
    E = ctx.run(det, capi.model(1), capi.analysis(0), src, rc, 17, n, columns=["energy_kev"])["energy_kev"]
    assert E.min() >= 0.0 and E.max() <= w.xMax and E_ref.max() <= w.xMax
    assert stats.ks_2samp(E, E_ref).pvalue > 1e-3
    # the 100-try bail-out returns exactly 0 keV (:491-495): same rate on both sides
    z, zr = int((E == 0).sum()), int((E_ref == 0).sum())
    assert abs(z - zr) < 5 * np.sqrt(max(z + zr, 25)), (z, zr)
    # quantiles of the steeply falling spectrum, to the sampling error of both
    for q in (0.1, 0.5, 0.9, 0.99):
        a, b = np.quantile(E, q), np.quantile(E_ref, q)
        assert abs(a - b) < 0.03 * max(b, 0.1), (q, a, b)


def test_wimp_chain_all_columns_match_reference(ctx, ref):
    """ref_chain with src_kind WIMP runs WIMP_prep_spectrum right after seeding, as execNEST does (:474-481)"""
    n, seed = 60_000, 9
    w = table_of_reference_seed(ref, seed, 50.0)
    r = ref_chain(ref, seed, n, capi.WIMP, capi.SRC_WIMP, 50.0, 1e-36)
    g, _ = gpu_chain(ctx, 4, n, capi.WIMP, capi.SRC_WIMP, 50.0, 1e-36, wimp=w)
    ks = ks_report(g, r)
    bad = {k: v for k, v in ks.items() if v < 1e-4}
    assert not bad, bad
    assert len(ks) == 36


def test_wimp_cli_matches_reference_binary():
    """`execNEST exposure WIMP mass xsec field pos seed`: the event count is Poisson(integral x exposure x xsec/1e-36)
    on both sides; rows statistically the same"""
    if not os.path.exists(CLI) or not os.path.exists(refprobe.REF_BIN_LUX):
        pytest.skip("command-line binaries not built")
    expo = 0.03  # kg-days: integral(50 GeV) ~ 1.013e6 /kg/day/pb -> ~30400 events
    counts_g, counts_r, G, R = [], [], None, None
    for s in (1, 2, 3, 4):
        args = (expo, "WIMP", "50", "1e-36", "180", "-1", s)
        rc_g, out_g, err_g = run(CLI, args)
        rc_r, out_r, err_r = run(refprobe.REF_BIN_LUX, args)
        assert rc_g == 0 and rc_r == 0, (err_g[-400:], err_r[-400:])
        g, r = rows_of(out_g), rows_of(out_r)
        counts_g.append(len(g))
        counts_r.append(len(r))
        G = g if G is None else np.vstack([G, g])
        R = r if R is None else np.vstack([R, r])
    mean = 1.0132567e6 * expo
    for c in counts_g + counts_r:   # each a Poisson draw about the same mean (isotope noise of the integral ~1e-4)
        assert abs(c - mean) < 5 * np.sqrt(mean), (c, mean)
    assert len(set(counts_g)) > 1   # seeds give different counts
    names = ["keV", "field", "dt", "x", "y", "z", "Nph", "Ne", "S1", "S1c", "spikeC", "Nee", "S2", "S2c"]
    for j, nme in enumerate(names):
        if np.all(G[:, j] == G[0, j]):
            assert np.all(R[:, j] == G[0, j]), nme
            continue
        p = stats.ks_2samp(G[:, j], R[:, j]).pvalue
        assert p > 1e-4, (nme, p, G[:, j].mean(), R[:, j].mean())
