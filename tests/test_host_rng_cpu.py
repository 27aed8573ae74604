"""Host-side checks of the counter-based generator (nest_b200/csrc/nb_rng.cuh compiled by g++, no GPU):
Random123 known-answer vectors for Philox4x32-10, the round-key / per-event variants, Stream word order."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_philox_known_answers_and_variants(tmp_path):
    exe = str(tmp_path / "host_checks")
    src = os.path.join(ROOT, "tests", "host_checks.cpp")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-x", "c++", src, "-o", exe], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "host checks ok" in r.stdout, r.stdout
