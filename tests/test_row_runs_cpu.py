"""The rasteriser's closed-form pixel runs (phase A/B of raster_fwd_kernel) restated on the CPU and checked
against the exact oracle-order edge-sign test on millions of random triangle rows, incl. slivers, (near-)horizontal
and vertical edges, sub-pixel to half-image sizes and a +-4 ulp error on the approximate division:
no accepted pixel may fall outside a run.  (The GPU tests check the same end to end; this one runs anywhere.)"""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def test_closed_form_runs_never_miss_a_pixel():
    exe = os.path.join(tempfile.mkdtemp(), "row_runs_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(HERE, "csrc", "row_runs_check.c"), "-lm"])
    out = subprocess.run([exe, "150000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == "TOTAL_MISSED 0"
    exact = sum(int(l.split("exact=")[1].split()[0]) for l in lines[:-1])
    cand = sum(int(l.split("candidates=")[1].split()[0]) for l in lines[:-1])
    assert exact > 5_000_000 and cand < 2 * exact          # the runs are tight, not just supersets
