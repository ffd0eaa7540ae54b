"""Rows a10 - a15 on the CPU suite: the per-graph code the kernels compile (asm_device.cuh, frag_device.cuh: create_graph, the
read -> transfrag keys, process_transfrags with the reference's quicksort, find_transcripts / push_max_flow / store_transcript)
compiled for the HOST with one-lane warps (tools/asm_hostcheck.cu; warp_prims.cuh degenerates to the identity) and compared
with the dumps taken inside the reference's build_graphs (golden fixtures: a10_*, a11_*, a12_*, the prediction list).  It does
not replace the GPU parity tests (lane-parallel paths, the commit kernels): it pins the sequential logic where no GPU is."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostcheck_exe():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    out = os.path.join(ROOT, "tools", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "asm_hostcheck")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-ffp-contract=off",
                           "-o", exe, os.path.join(ROOT, "tools", "asm_hostcheck.cu")])
    return exe


@pytest.mark.parametrize("fixture", ["short_s11", "short_s12_deep", "short_s13_jitter"])
def test_per_graph_code_on_the_host_equals_the_reference_dumps(hostcheck_exe, fixture):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "asm_hostcheck.py"), fixture, "15"], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout.splitlines()[-1], r.stdout[-3000:] + r.stderr[-2000:]
