"""End-to-end drop-in check: the REFERENCE program with infer_transcripts() (rlink.h:380, call site stringtie.cpp:1477)
served by libstgpu through host/stgpu_adaptor.cpp (built by host/Makefile into host/_build/stringtie_stgpu) — i.e. coverage,
junctions, groups, splice graphs, flow and predictions all come from the GPU, the reference's own producer, printResults and
GTF writer run around them — must write the same GTF, byte for byte, as the unmodified reference.  The expected GTFs (tests/golden/gtf_*.gtf) were written by oracle/_ref/stringtie -p 1 on
the seeded SAM files tools/gen_sam.py regenerates here."""
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

BIN = os.path.join(ROOT, "host", "_build", "stringtie_stgpu")
CASES = {
    "gtf_s11": ["--loci", "24", "--pairs", "2500", "--seed", "11", "--long-intron-frac", "0.01"],
    "gtf_s13_jitter": ["--loci", "16", "--pairs", "4000", "--seed", "13", "--long-intron-frac", "0.02", "--nh-frac", "0.1",
                       "--jitter-frac", "0.06"],
    "gtf_s14_deep": ["--loci", "8", "--pairs", "12000", "--seed", "14", "--long-intron-frac", "0.0", "--nh-frac", "0.25",
                     "--jitter-frac", "0.03"],
}


def _body(path):
    return [l for l in open(path) if not l.startswith("#")]


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="host/_build/stringtie_stgpu not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("threads", [1, 4, 16])
def test_reference_host_with_gpu_stage_writes_identical_gtf(tmp_path, name, threads):
    sam, gtf = str(tmp_path / "in.sam"), str(tmp_path / "out.gtf")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_sam.py"), "--sam", sam] + CASES[name],
                          stderr=subprocess.DEVNULL)
    env = dict(os.environ, STGPU_VERBOSE="1")
    p = subprocess.run([BIN, sam, "-p", str(threads), "-o", gtf], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    import re as _re
    m = _re.search(r"libstgpu: (\d+) bundles in (\d+) device batches \(largest (\d+)\), (\d+) predictions", p.stderr)
    assert m, p.stderr[-2000:]
    nb, nbatch, largest, npred = (int(x) for x in m.groups())
    assert nb > 0 and npred > 0 and 1 <= nbatch <= nb and largest <= max(threads, 1)
    if threads == 1:
        assert nbatch == nb
    got, want = _body(gtf), _body(os.path.join(ROOT, "tests", "golden", name + ".gtf"))
    if threads == 1:
        assert got == want
    else:  # with -p N the reference numbers genes in completion order: compare the records up to the ids
        import re
        norm = lambda ls: sorted(re.sub(r'(gene_id|transcript_id) "STRG\.[0-9.]+";', r'\1 "";', l) for l in ls)
        assert norm(got) == norm(want)


def test_adaptor_source_declares_the_replaced_symbol():
    src = open(os.path.join(ROOT, "host", "stgpu_adaptor.cpp")).read()
    assert "int infer_transcripts(BundleData* bundle)" in src and "stg_infer_transcripts_batch" in src and "assemble = 1" in src
