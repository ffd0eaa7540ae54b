"""The closed form for repeated float32 addition used by the tile kernel (advance_bs / fill_run) against the plain
loop, on the CPU: same C arithmetic (SSE, no contraction) as the reference build."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_advance_closed_form_matches_loop():
    src = os.path.join(ROOT, "tests", "advance_check.c")
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "advance_check")
        subprocess.check_call(["gcc", "-O1", "-ffp-contract=off", "-msse2", "-mfpmath=sse", "-o", exe, src])
        out = subprocess.check_output([exe, "400000"], text=True)
    assert "0 mismatches" in out, out
