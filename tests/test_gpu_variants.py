"""The product library against its own plain build (nest_b200/csrc/libnest_b200_plain.so: per-lane BTRS
acceptance test, sequential box rejection of the spectra, erfc without table bounds).  The warp-cooperative
evaluations and the bounds are meant to be pure optimisations -- every event must come out bit for bit the same."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def hashes(lib, n):
    env = dict(os.environ, NB200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(HERE, "variant_hash.py"), str(n)], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_cooperative_paths_are_bit_identical_to_the_plain_algorithms():
    from nest_b200 import build as nb_build
    if not os.path.exists(nb_build.LIB_PLAIN):
        pytest.skip("libnest_b200_plain.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    n = 2_000_000
    a, b = hashes(nb_build.LIB, n), hashes(nb_build.LIB_PLAIN, n)
    assert a.keys() == b.keys() and len(a) == 9
    for k in a:
        assert a[k] == b[k], "configuration %s differs between the product and the plain build" % k
