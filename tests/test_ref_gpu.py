"""GPU parity tests of mecat2ref (-m gpu): the CUDA path -- genome index, seeding / DDF scoring /
candidate walk, gapped extension on genome windows, clipped-end rescue -- through the C ABI (mecat_b200_ref_index_build,
mecat_b200_ref_map) and through the `mecat2ref` command-line driver, against the output of the UNMODIFIED reference binary
(tests/golden/refmap*) and the pinned oracle.

The same stage sequence and kernel bodies pass these fixtures on the host (tests/test_ref_host.py)."""
import ctypes as C
import gzip
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(util.GOLDEN, "golden.json")))


def groups(s):
    lines = s.rstrip("\n").split("\n") if s else []
    return sorted("\n".join(lines[i:i + 3]) for i in range(0, len(lines), 3))


def golden(name):
    with gzip.open(os.path.join(util.GOLDEN, name), "rt") as f:
        return f.read()


@pytest.fixture(scope="module")
def refmap_inputs(tmp_path_factory):
    c = GOLD["refmap"]
    d = tmp_path_factory.mktemp("refmap_gpu")
    fa, genome = str(d / "reads.fa"), str(d / "genome.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], genome_out=genome)
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == c["fasta_sha256"]
    return fa, genome


@pytest.fixture(scope="module")
def hard_inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("refmap_hard_gpu")
    fa, genome = str(d / "reads.fa"), str(d / "genome.fa")
    util.make_refmap_hard(fa, genome)
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == GOLD["refmap_hard"]["fasta_sha256"]
    return fa, genome


def map_through_abi(ctx, genome_path, reads_path, fmt, n=10, b=10):
    from mecat_b200 import api
    G = api.RefGenome.from_fasta(genome_path)
    seqs = util.read_fasta(reads_path)
    idx = ctx.ref_index_build(G)
    try:
        rec, q, s = ctx.ref_map(idx, api.RefReads(seqs), n, b, want_strings=fmt != 1)
    finally:
        ctx.release_ref_index(idx)
    return api.format_ref_results(G, list(range(len(seqs))), rec, q, s, fmt), rec


def run_cli(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([os.path.join(util.ROOT, "mecat_b200", "bin", "mecat2ref")] + args, env=e, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    return p


def run_oracle(genome, fa, n, b, fmt):
    O = util.oracle()
    text, nb = C.c_void_p(), C.c_size_t()
    assert O.orc_ref_map(genome.encode(), fa.encode(), n, b, fmt, C.byref(text), C.byref(nb)) == 0
    s = C.string_at(text.value, nb.value).decode()
    O.orc_free(text)
    return s


def test_index_and_first_pass_candidates_match_the_host_twin(gpu_ctx, refmap_inputs, hard_inputs, tmp_path):
    """Function-level parity below the end-to-end records: the genome's k-mer index (both CSR arrays) and the candidate
    list SeedFn leaves for every strand, against the same structures computed on the host by tests/ref_host_harness.cpp
    (whose end-to-end output is pinned to the unmodified binary by tests/test_ref_host.py)."""
    from mecat_b200 import api
    L = util.ref_harness()
    rep_fa, rep_genome = str(tmp_path / "reads.fa"), str(tmp_path / "genome.fa")
    util.make_refmap_repeats(rep_fa, rep_genome, seed=5, num_reads=150)
    for fa, genome, ncand in ((refmap_inputs[0], refmap_inputs[1], 10), (hard_inputs[0], hard_inputs[1], 10), (rep_fa, rep_genome, 40)):
        G = api.RefGenome.from_fasta(genome)
        R = api.RefReads(util.read_fasta(fa))
        gc = G.c()
        hidx = L.harness_ref_index_build(C.byref(gc))
        hbegin = np.zeros((1 << 26) + 1, dtype=np.uint32)
        hn = L.harness_ref_index_export(hidx, hbegin.ctypes.data_as(C.c_void_p), None)
        hpos = np.zeros(max(1, hn), dtype=np.int32)
        L.harness_ref_index_export(hidx, None, hpos.ctypes.data_as(C.c_void_p))
        idx = gpu_ctx.ref_index_build(G)
        try:
            begin, pos = gpu_ctx.ref_index_export(idx)
            assert len(pos) == hn and (begin == hbegin).all() and (pos == hpos[:hn]).all()
            rows, counts = gpu_ctx.ref_raw_candidates(idx, R, ncand)
        finally:
            gpu_ctx.release_ref_index(idx)
        rc, p = R.c(), api.RefParams(ncand, ncand, 0, 0)
        hrows, hcounts, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        assert L.harness_ref_raw_candidates(hidx, C.byref(rc), C.byref(p), C.byref(hrows), C.byref(hcounts), C.byref(n)) == 0
        want_counts = np.frombuffer(C.string_at(hcounts.value, 8 * len(R.read_len)), dtype="<i4")
        want_rows = np.frombuffer(C.string_at(hrows.value, 16 * n.value), dtype="<i4").reshape(-1, 4)
        for ptr in (hrows, hcounts):
            L.harness_free(ptr)
        L.harness_ref_index_release(hidx)
        assert n.value > len(R.read_len) // 2
        assert (counts == want_counts).all()
        assert rows.shape == want_rows.shape and (rows == want_rows).all()


def test_m4_records_match_reference(gpu_ctx, refmap_inputs):
    """300 CLR reads against their 100 kb genome, m4 records (coordinates, identity, score) of the unmodified binary."""
    fa, genome = refmap_inputs
    gpu_ctx.reset_stats()
    text, rec = map_through_abi(gpu_ctx, genome, fa, 1)
    st = gpu_ctx.stats()
    want = golden("refmap.m4.gz").splitlines()
    got = sorted(text.splitlines())
    assert len(got) == len(want) == GOLD["refmap"]["num_m4"]
    assert got == want
    for k in ("index_count", "index_fill", "ref_count", "ref_seed", "extend"):      # every stage was a kernel launch
        assert st["kernel_launches"][k] > 0, k
    assert (rec["str_offset"] == -1).all()                                           # no strings were asked for


def test_ref_format_matches_reference(gpu_ctx, refmap_inputs):
    """Same reads, ref format: both alignment strings of every record."""
    fa, genome = refmap_inputs
    text, _ = map_through_abi(gpu_ctx, genome, fa, 0)
    assert groups(text) == groups(golden("refmap.ref.gz"))


def test_hard_inputs_match_reference(gpu_ctx, hard_inputs):
    """Three contigs with a shared repeat and a run of N, chimeric reads (clipped ends: the rescue kernel), very noisy
    reads (second pass), reads with N and lower-case stretches (explicit reverse strands), short reads."""
    fa, genome = hard_inputs
    gpu_ctx.reset_stats()
    text, _ = map_through_abi(gpu_ctx, genome, fa, 0)
    st = gpu_ctx.stats()
    got, want = groups(text), groups(golden("refmap_hard.ref.gz"))
    gh, wh = [g.split("\n")[0] for g in got], [w.split("\n")[0] for w in want]
    assert gh == wh, (len(gh), len(wh), sorted(set(gh) - set(wh))[:5], sorted(set(wh) - set(gh))[:5])
    assert got == want
    assert st["kernel_launches"]["ref_seed"] >= 2 and st["kernel_launches"]["ref_rescue"] >= 1


@pytest.mark.parametrize("n,b", [(3, 2), (1, 1), (50, 4)])
def test_candidate_and_output_caps_match_oracle(gpu_ctx, hard_inputs, n, b):
    fa, genome = hard_inputs
    text, _ = map_through_abi(gpu_ctx, genome, fa, 0, n, b)
    assert groups(text) == groups(run_oracle(genome, fa, n, b, 0))


@pytest.mark.skipif(not os.path.exists(os.path.join(util.REF_DIR, "mecat2ref")), reason="needs the unmodified binary (oracle/_ref travels with the snapshot)")
@pytest.mark.parametrize("n,b", [(10, 10), (40, 5)])
def test_repeat_rich_inputs_match_the_unmodified_binary(gpu_ctx, tmp_path, n, b):
    """Differential run against oracle/_ref/mecat2ref on a repeat-rich genome (util.make_refmap_repeats): full candidate
    lists, block-consuming votes, insert_loc evictions, rescue between repeat copies, both passes."""
    fa, genome, out = str(tmp_path / "reads.fa"), str(tmp_path / "genome.fa"), str(tmp_path / "ref.out")
    util.make_refmap_repeats(fa, genome, seed=3, num_reads=200)
    subprocess.check_call([os.path.join(util.REF_DIR, "mecat2ref"), "-d", fa, "-r", genome, "-o", out, "-w", str(tmp_path / "w"), "-t", "4", "-m", "0",
                           "-n", str(n), "-b", str(b)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=str(tmp_path))
    text, _ = map_through_abi(gpu_ctx, genome, fa, 0, n, b)
    want = groups(open(out).read())
    assert len(want) > 3 * 200
    assert groups(text) == want


def _map_with_env(refmap_inputs, env):
    """Own context: both budgets are read when the context / the call starts."""
    import mecat_b200
    fa, genome = refmap_inputs
    os.environ.update(env)
    try:
        with mecat_b200.Context(0) as ctx:
            text, _ = map_through_abi(ctx, genome, fa, 0)
            st = ctx.stats()
    finally:
        for k in env:
            del os.environ[k]
    return text, st["kernel_launches"]


def test_small_table_batches_give_the_same_records(refmap_inputs):
    """A 1 MB budget for the block tables cuts the 300 reads into several table batches."""
    text, launches = _map_with_env(refmap_inputs, {"MECAT_B200_REF_TABLE_MB": "1"})
    assert groups(text) == groups(golden("refmap.ref.gz"))
    assert launches["ref_seed"] >= 4


def test_small_arena_batches_give_the_same_records(refmap_inputs):
    """A 1 MB column arena cuts the extension call of the (single) table batch into several launches: 300 extensions of
    6 kb reads on 2.2x windows need ~6 MB of columns."""
    text, launches = _map_with_env(refmap_inputs, {"MECAT_B200_ALIGN_ARENA_MB": "1"})
    assert groups(text) == groups(golden("refmap.ref.gz"))
    assert launches["extend"] >= 3


def test_small_table_and_arena_batches_together(refmap_inputs):
    text, launches = _map_with_env(refmap_inputs, {"MECAT_B200_REF_TABLE_MB": "1", "MECAT_B200_ALIGN_ARENA_MB": "1"})
    assert groups(text) == groups(golden("refmap.ref.gz"))
    assert launches["ref_seed"] >= 4 and launches["extend"] >= launches["ref_seed"]


def test_forward_only_extension_gives_the_same_m4_records(gpu_ctx, refmap_inputs, hard_inputs):
    """MECAT_B200_REF_EXTEND=forward: extensions whose strings are not wanted run through k_extend (forward pass only, the
    genome windows as a second offset table) instead of the kernel that also writes alignment strings; same records."""
    os.environ["MECAT_B200_REF_EXTEND"] = "forward"
    try:
        fa, genome = refmap_inputs
        gpu_ctx.reset_stats()
        text, _ = map_through_abi(gpu_ctx, genome, fa, 1)
        st = gpu_ctx.stats()
        assert sorted(text.splitlines()) == golden("refmap.m4.gz").splitlines()
        assert st["kernel_launches"]["extend"] > 0 and st["kernel_launches"]["finalize"] == 0      # no string assembly ran
        fa, genome = hard_inputs
        text, _ = map_through_abi(gpu_ctx, genome, fa, 1)
        assert sorted(text.splitlines()) == sorted(run_oracle(genome, fa, 10, 10, 1).splitlines())
        # with strings wanted the same switch extends forward-only first and asks for the strings of the printed records
        # only; the library itself checks that both kernels agree on every coordinate
        gpu_ctx.reset_stats()
        text, rec = map_through_abi(gpu_ctx, genome, fa, 0)
        st = gpu_ctx.stats()
        assert groups(text) == groups(golden("refmap_hard.ref.gz"))
        assert st["kernel_launches"]["finalize"] > 0 and (rec["str_offset"] >= 0).all()
    finally:
        del os.environ["MECAT_B200_REF_EXTEND"]


def test_command_line_driver_matches_reference(gpu_ctx, refmap_inputs, hard_inputs, tmp_path):
    """bin/mecat2ref with the reference's flags: ref, sam and m4 files equal the unmodified binary's, also with two devices."""
    import mecat_b200
    fa, genome = hard_inputs
    out = str(tmp_path / "hard.ref")
    run_cli(["-d", fa, "-r", genome, "-o", out, "-w", str(tmp_path / "w1"), "-t", "2", "-m", "0"])
    assert groups(open(out).read()) == groups(golden("refmap_hard.ref.gz"))
    out = str(tmp_path / "hard.sam")
    run_cli(["-d", fa, "-r", genome, "-o", out, "-w", str(tmp_path / "w1"), "-m", "2"])
    lines = open(out).read().splitlines()
    with gzip.open(os.path.join(util.GOLDEN, "refmap_hard.sam.gz"), "rt") as f:
        want = f.read().splitlines()
    assert [l for l in lines if l.startswith("@") and not l.startswith("@PG")] == [l for l in want if l.startswith("@")]
    assert sum(l.startswith("@PG\tID:0\tVN:0.0.1\tCL:") and l.endswith("\tPN:mecat2ref") for l in lines) == 1
    assert sorted(l for l in lines if not l.startswith("@")) == [l for l in want if not l.startswith("@")]
    fa, genome = refmap_inputs
    out = str(tmp_path / "refmap.m4")
    run_cli(["-d", fa, "-r", genome, "-o", out, "-w", str(tmp_path / "w2"), "-t", "2", "-m", "1"])
    assert sorted(open(out).read().splitlines()) == golden("refmap.m4.gz").splitlines()
    if mecat_b200.load_library().mecat_b200_device_count() >= 2:
        out2 = str(tmp_path / "refmap2.m4")
        run_cli(["-d", fa, "-r", genome, "-o", out2, "-w", str(tmp_path / "w3"), "-m", "1"], env={"MECAT_GPUS": "2"})
        assert open(out2).read() == open(out).read()


def test_cfg0_sized_reads_match_reference(gpu_ctx, tmp_path):
    """1 000 x 15 kb reads (BASELINE configs[0]'s read set) against their 1 Mb genome, m4 records."""
    c = GOLD["refmap_cfg0"]
    fa, genome = str(tmp_path / "reads.fa"), str(tmp_path / "genome.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], genome_out=genome)
    text, _ = map_through_abi(gpu_ctx, genome, fa, 1)
    got, want = sorted(text.splitlines()), golden("refmap_cfg0.m4.gz").splitlines()
    assert len(got) == len(want) == c["num_m4"]
    assert got == want


def test_degenerate_inputs(gpu_ctx, refmap_inputs):
    """No reads, reads shorter than a k-mer, a genome shorter than a k-mer: empty results, not errors; a read the
    reference's fixed buffers cannot hold is refused."""
    import mecat_b200
    from mecat_b200 import api
    fa, genome = refmap_inputs
    G = api.RefGenome.from_fasta(genome)
    idx = gpu_ctx.ref_index_build(G)
    try:
        for seqs in ([], [b"ACGT"], [b"ACGTACGTACGTA", b"N" * 50, b"acgt" * 30]):
            rec, q, s = gpu_ctx.ref_map(idx, api.RefReads(seqs))
            assert len(rec) == 0
        with pytest.raises(mecat_b200.MecatB200Error):
            gpu_ctx.ref_map(idx, api.RefReads([b"ACGT" * 25000]))
    finally:
        gpu_ctx.release_ref_index(idx)
    tiny = gpu_ctx.ref_index_build(api.RefGenome(["t"], [b"ACGTNNACGT"]))
    try:
        rec, _, _ = gpu_ctx.ref_map(tiny, api.RefReads(util.read_fasta(fa)[:5]))
        assert len(rec) == 0
    finally:
        gpu_ctx.release_ref_index(tiny)


def test_warp_per_strand_seeding_equals_thread_per_strand(gpu_ctx, hard_inputs, tmp_path, monkeypatch):
    """k_ref_seed_warp (SeedWarpFn: lanes share the prefetch, insert_loc, find_location and the neighbour votes; the
    default) against k_ref<SeedFn> (a thread per strand) on the device: same candidate lists per strand, same records, on the
    hard fixture and on a repeat-rich one (full candidate lists, evictions in most blocks, consumed neighbours)."""
    from mecat_b200 import api
    rep_fa, rep_genome = str(tmp_path / "reads.fa"), str(tmp_path / "genome.fa")
    util.make_refmap_repeats(rep_fa, rep_genome, seed=9, num_reads=200)
    for fa, genome, ncand in ((hard_inputs[0], hard_inputs[1], 10), (rep_fa, rep_genome, 40)):
        G = api.RefGenome.from_fasta(genome)
        R = api.RefReads(util.read_fasta(fa))
        idx = gpu_ctx.ref_index_build(G)
        try:
            out = {}
            for mode in ("warp", "thread"):
                monkeypatch.setenv("MECAT_B200_REF_SEED", mode)
                rows, counts = gpu_ctx.ref_raw_candidates(idx, R, ncand)
                rec, q, s = gpu_ctx.ref_map(idx, R, ncand, ncand, want_strings=False)
                out[mode] = (rows.tobytes(), counts.tobytes(), rec.tobytes())
            assert out["warp"] == out["thread"]
            assert len(out["warp"][2]) > 1000
        finally:
            gpu_ctx.release_ref_index(idx)
