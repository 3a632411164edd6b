"""The oracle's restatement of mecat2ref (oracle/oracle_ref.cpp) against the output of the UNMODIFIED `mecat2ref` binary
(tests/golden/refmap.*: 300 synthetic CLR reads against their own 100 kb genome, M4 and ref format).  mecat2ref is the
next row of the scope table: no CUDA path exists for it yet, this pins the checker it will be built against."""
import ctypes as C
import gzip
import hashlib
import json
import os

import pytest

import util

GOLD = json.load(open(os.path.join(util.GOLDEN, "golden.json")))


@pytest.fixture(scope="module")
def refmap_inputs(tmp_path_factory):
    c = GOLD["refmap"]
    d = tmp_path_factory.mktemp("refmap")
    fa, genome = str(d / "reads.fa"), str(d / "genome.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], genome_out=genome)
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == c["fasta_sha256"]
    assert hashlib.sha256(open(genome, "rb").read()).hexdigest() == c["genome_sha256"]
    return fa, genome


def run_oracle(fa, genome, fmt, tech=0):
    O = util.oracle()
    text, n = C.c_void_p(), C.c_size_t()
    assert O.orc_ref_map_x(genome.encode(), fa.encode(), 10, 10, fmt, tech, C.byref(text), C.byref(n)) == 0
    s = C.string_at(text.value, n.value).decode()
    O.orc_free(text)
    return s


def test_m4_output_matches_reference(refmap_inputs):
    fa, genome = refmap_inputs
    got = sorted(run_oracle(fa, genome, 1).splitlines())
    with gzip.open(os.path.join(util.GOLDEN, "refmap.m4.gz"), "rt") as f:
        want = f.read().splitlines()
    assert len(got) == len(want) == GOLD["refmap"]["num_m4"]
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, bad[:3]


def test_ref_format_output_matches_reference(refmap_inputs):
    fa, genome = refmap_inputs
    lines = run_oracle(fa, genome, 0).split("\n")
    got = sorted("\n".join(lines[i:i + 3]) for i in range(0, len(lines) - 1, 3))
    with gzip.open(os.path.join(util.GOLDEN, "refmap.ref.gz"), "rt") as f:
        want = f.read().rstrip("\n").split("\n")
    want = ["\n".join(want[i:i + 3]) for i in range(0, len(want), 3)]
    assert len(got) == len(want) == GOLD["refmap"]["num_ref"]
    bad = [g.split("\n")[0] for g, w in zip(got, want) if g != w]
    assert not bad, bad[:5]


def test_hard_inputs_match_reference(tmp_path):
    """Three contigs with a shared repeat and a run of N, chimeric reads (clipped alignments: rescue_clipped_align), very
    noisy reads (the second, more sensitive pass), reads with N and lower-case stretches, short reads."""
    fa, genome = str(tmp_path / "reads.fa"), str(tmp_path / "genome.fa")
    util.make_refmap_hard(fa, genome)
    c = GOLD["refmap_hard"]
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == c["fasta_sha256"]
    assert hashlib.sha256(open(genome, "rb").read()).hexdigest() == c["genome_sha256"]
    lines = run_oracle(fa, genome, 0).split("\n")
    got = sorted("\n".join(lines[i:i + 3]) for i in range(0, len(lines) - 1, 3))
    with gzip.open(os.path.join(util.GOLDEN, "refmap_hard.ref.gz"), "rt") as f:
        want = f.read().rstrip("\n").split("\n")
    want = ["\n".join(want[i:i + 3]) for i in range(0, len(want), 3)]
    gh, wh = [g.split("\n")[0] for g in got], [w.split("\n")[0] for w in want]
    assert gh == wh, (len(gh), len(wh), sorted(set(gh) - set(wh))[:5], sorted(set(wh) - set(gh))[:5])
    assert got == want


def test_cfg0_sized_reads_match_reference(tmp_path):
    """1 000 x 15 kb reads (BASELINE configs[0]'s read set) against their 1 Mb genome, M4 output."""
    c = GOLD["refmap_cfg0"]
    fa, genome = str(tmp_path / "reads.fa"), str(tmp_path / "genome.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], genome_out=genome)
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == c["fasta_sha256"]
    got = sorted(run_oracle(fa, genome, 1).splitlines())
    with gzip.open(os.path.join(util.GOLDEN, "refmap_cfg0.m4.gz"), "rt") as f:
        want = f.read().splitlines()
    assert len(got) == len(want) == c["num_m4"]
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, bad[:3]


def test_nanopore_outputs_match_reference(refmap_inputs):
    """mecat2ref -x 1: the same program with XdropAligner as its aligner (mecat2ref_impl_large.cpp:329-332); M4 and ref
    format of the unmodified binary."""
    fa, genome = refmap_inputs
    got = sorted(run_oracle(fa, genome, 1, tech=1).splitlines())
    with gzip.open(os.path.join(util.GOLDEN, "refmap.x1.m4.gz"), "rt") as f:
        want = f.read().splitlines()
    assert len(got) == len(want) == GOLD["x1"]["refmap_num_m4"]
    assert got == want
    lines = run_oracle(fa, genome, 0, tech=1).split("\n")
    got = sorted("\n".join(lines[i:i + 3]) for i in range(0, len(lines) - 1, 3))
    with gzip.open(os.path.join(util.GOLDEN, "refmap.x1.ref.gz"), "rt") as f:
        want = f.read().rstrip("\n").split("\n")
    want = ["\n".join(want[i:i + 3]) for i in range(0, len(want), 3)]
    assert len(got) == len(want) == GOLD["x1"]["refmap_num_ref"]
    bad = [g.split("\n")[0] for g, w in zip(got, want) if g != w]
    assert not bad, bad[:5]
