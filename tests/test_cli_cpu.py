"""The execNEST command line (nest_b200/host/execNEST_b200) where no GPU is needed: every argument
error the reference reports before it simulates anything (src/execNEST.cpp:65-95, 161-183, 468-610)
gives the same return code and the same stderr text as the unmodified reference binary run beside
it, and a valid command without a CUDA device fails loudly instead of falling back to the CPU."""
import os
import subprocess

import pytest

import refprobe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "nest_b200", "host", "execNEST_b200")


def run(binary, args):
    r = subprocess.run([binary] + list(args), stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=120)
    return r.returncode, r.stdout, [ln for ln in r.stderr.splitlines() if ln.strip()]


@pytest.fixture(scope="module")
def cli(built):
    if not os.path.exists(CLI):
        r = subprocess.run(["make", "-s", "-C", os.path.dirname(CLI)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True)
        assert r.returncode == 0, r.stdout
    return CLI


SAME_TEXT = [("0", "NR", "0", "100", "180", "-1"),           # :578-581 at least one event
             ("-5", "NR", "0", "100", "180", "-1"),
             ("100", "NR", "0", "0", "180", "-1"),           # :603 maximum energy 0
             ("100",),                                       # :65-95 usage
             ("100", "NR", "0"),
             ("100", "Kr83m", "20", "400", "180", "-1"),     # :583-601 not one of the three lines
             ("100", "Kr83m", "41.5", "-5", "180", "-1"),    # max time separation must be +
             ("100", "WIMP", "0.1", "1e-36", "180", "-1")]   # :484 WIMP mass too low


@pytest.mark.parametrize("args", SAME_TEXT, ids=[" ".join(a) for a in SAME_TEXT])
def test_argument_errors_read_like_the_reference(cli, args):
    rc_g, out_g, err_g = run(cli, args)
    assert rc_g == 1
    if not os.path.exists(refprobe.REF_BIN_LUX):
        pytest.skip("oracle/_ref/execNEST_ref_lux not built")
    rc_r, out_r, err_r = run(refprobe.REF_BIN_LUX, args)
    assert rc_r == 1
    assert err_g == err_r
    # the usage text (stdout, :65-95), up to the paragraph for muon tracks, which are not on the GPU path; else empty
    assert out_r.startswith(out_g) and (out_g == out_r or out_r[len(out_g):].startswith("For cosmic-ray muons"))
    assert not any(ln[:1].isdigit() for ln in out_g.splitlines())


def test_unknown_particle_type_lists_the_types(cli):
    rc, _, err = run(cli, ("100", "bogus", "0", "100", "180", "-1"))
    assert rc == 1 and "UNRECOGNIZED PARTICLE TYPE!! VALID OPTIONS ARE:" in err
    if os.path.exists(refprobe.REF_BIN_LUX):
        rc_r, _, err_r = run(refprobe.REF_BIN_LUX, ("100", "bogus", "0", "100", "180", "-1"))
        assert rc_r == 1
        # the same list line for line, up to the two types that are not on the GPU path (muon, fullGamma)
        k = err.index("UNRECOGNIZED PARTICLE TYPE!! VALID OPTIONS ARE:")
        kr = err_r.index("UNRECOGNIZED PARTICLE TYPE!! VALID OPTIONS ARE:")
        mine, theirs = err[k:], err_r[kr:]
        assert mine[:-1] == theirs[:len(mine) - 1] and theirs[len(mine) - 1] == mine[-1] + ","
        assert [t.split()[0] for t in theirs[len(mine):]] == ["muon", "fullGamma"]
    # out-of-scope types say so and point at the reference (no silent substitution)
    rc, _, err = run(cli, ("100", "fullGamma", "0", "400", "180", "-1"))
    assert rc == 1 and any("not part of the GPU path" in ln for ln in err)


def test_no_device_is_an_error_not_a_fallback(cli):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc, out, err = run(cli, ("100", "NR", "0", "100", "180", "-1", "1"))
    assert rc == 1
    assert any("no CUDA device" in ln and "no CPU fallback" in ln for ln in err)
    assert not any(ln[:1].isdigit() for ln in out.splitlines())       # not one event row was printed


def test_stage_api_compiles_against_the_reference_signatures(built, tmp_path):
    """tests/cpp/stage_calls.cpp calls GetYields -> GetQuanta -> GetS1 -> GetS2 and FullCalculation exactly as code
    written for the reference does (examples/bareNEST.cpp:183-212, NEST.hh:218-226, 308-311, 335-356): it must compile
    and link against the host mirror unchanged; without a device it must be refused loudly"""
    host = os.path.join(ROOT, "nest_b200", "host")
    r = subprocess.run(["make", "-s", "-C", host], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    exe = str(tmp_path / "stage_calls")
    csrc = os.path.join(ROOT, "nest_b200", "csrc")
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + host,
           os.path.join(ROOT, "tests", "cpp", "stage_calls.cpp"), "-L" + host, "-lnest_b200_host", "-L" + csrc, "-lnest_b200",
           "-Wl,-rpath," + host, "-Wl,-rpath," + csrc, "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().splitlines()
        assert len(lines[0].split("\t")) == 8 and len(lines[1].split("\t")) == 4
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr
        assert r.stdout.strip() == ""
