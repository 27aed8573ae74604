"""Event sharding across ranks: events are i.i.d. given (seed, event id), so rank r of
W takes a contiguous id range and only the integer band accumulators are summed afterwards."""
import numpy as np

from . import capi


def event_range(rank, world, n_total, first_event=0):
    """[lo, hi) of global event ids for `rank`; ranges are contiguous, disjoint and cover
    [first_event, first_event + n_total)."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard request rank=%r world=%r n=%r" % (rank, world, n_total))
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return first_event + lo, first_event + hi


def sweep_points(rank, world, n_points):
    """grid points of a parameter sweep (config 4) that `rank` simulates: dealt round-robin, so that neighbouring
    points -- which cost about the same -- spread over the ranks.  A point is whole on one rank: no collective."""
    if world < 1 or not (0 <= rank < world) or n_points < 0:
        raise ValueError("bad shard request rank=%r world=%r points=%r" % (rank, world, n_points))
    return list(range(rank, int(n_points), int(world)))


def band_words_from_signals(ana, coin_level, signal1, signal2):
    """The integer band accumulator (same words and fixed-point scales as the device kernel,
    include/nest_b200.h) computed on the host from per-event signals: GetBand's binning and
    accept/reject rules (reference src/execNEST.cpp:1508-1580)."""
    s1 = np.asarray(signal1, np.float64)
    s2 = np.asarray(signal2, np.float64)
    nb, lb = ana.numBins, ana.logBins
    X, O = (s2, s1) if ana.useS2 == 2 else (s1, s2)
    if ana.useS2 == 2:
        width, border = (ana.maxS2 - ana.minS2) / nb, ana.minS2
    else:
        width, border = (ana.maxS1 - ana.minS1) / nb, ana.minS1
    words = np.zeros(nb * 6 + nb * lb, np.int64)
    aX = np.abs(X)
    req = -np.inf if coin_level == 0 else 0.0
    taken = np.zeros(X.size, bool)
    for j in range(nb):
        c = border + width / 2. + j * width
        if coin_level == 0:
            m = ~taken if j == 0 else np.zeros(X.size, bool)
        else:
            m = (aX > c - width / 2.) & (aX <= c + width / 2.) & ~taken
        taken |= m
        good = m & (X >= req) & (O >= 0.)
        with np.errstate(divide="ignore", invalid="ignore"):
            if ana.useS2 == 0:
                v = np.log10(O[good] / X[good])
            elif ana.useS2 == 1:
                v = np.log10(O[good])
            else:
                v = np.log10(X[good] / O[good])
        ok = (X[good] != 0) & (O[good] != 0) & (v > ana.logMin) & (v < ana.logMax)
        y = np.where(ok, v, 0.0)
        b = words[6 * j:6 * j + 6]
        b[0] = good.sum()
        b[1] = (m & ~good).sum()
        b[2] = np.rint(X[good] * capi.lib().nb200_band_scale_x(ana)).astype(np.int64).sum()
        b[3] = np.rint(y * capi.FX_Y).astype(np.int64).sum()
        b[4] = np.rint(y * y * capi.FX_Y2).astype(np.int64).sum()
        b[5] = np.rint(y * y * y * capi.FX_Y3).astype(np.int64).sum()
        if lb > 0:
            yy = y[y != 0]
            jb = np.clip(((yy - ana.logMin) / (ana.logMax - ana.logMin) * lb).astype(np.int64), 0, lb - 1)
            words[nb * 6 + j * lb:nb * 6 + (j + 1) * lb] += np.bincount(jb, minlength=lb)
    return words


def band_from_words(ana, words):
    """capi.Band filled from packed words (device layout: 6 words per bin, then the 2-D histogram)"""
    nb, lb = ana.numBins, ana.logBins
    b = capi.Band(ana)
    w = np.asarray(words, np.int64)
    per = w[:nb * 6].reshape(nb, 6)
    b.accepted[:] = per[:, 0].astype(np.uint64)
    b.rejected[:] = per[:, 1].astype(np.uint64)
    b.sumS1[:] = per[:, 2]
    b.sumY[:] = per[:, 3]
    b.sumY2[:] = per[:, 4]
    b.sumY3[:] = per[:, 5]
    if lb > 0:
        b.hist2d[:] = w[nb * 6:].astype(np.uint64)
    return b
