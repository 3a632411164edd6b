"""mecat2ref: the product's stage sequence (mecat_b200/csrc/ref_pipeline.h), kernel bodies
(ref_core.cuh) and host I/O (host/refio.h) run on the host by tests/ref_host_harness.cpp -- a launch becomes a loop, the
gapped extension is the oracle's -- against the output of the UNMODIFIED `mecat2ref` binary (tests/golden/refmap*) and,
for other option values, against the pinned oracle."""
import ctypes as C
import gzip
import hashlib
import json
import os
import random

import pytest

import util

GOLD = json.load(open(os.path.join(util.GOLDEN, "golden.json")))


def run_harness(genome, fa, n=10, b=10, fmt=1, per_call=0, budget=0):
    L = util.ref_harness()
    text, nb = C.c_void_p(), C.c_size_t()
    st = (C.c_long * 4)()
    err = C.create_string_buffer(512)
    rc = L.harness_ref_map(genome.encode(), fa.encode(), n, b, fmt, per_call, budget, C.byref(text), C.byref(nb), st, err, 512)
    assert rc == 0, err.value
    s = C.string_at(text.value, nb.value).decode()
    L.harness_free(text)
    return s, list(st)


def run_oracle(genome, fa, n, b, fmt):
    O = util.oracle()
    text, nb = C.c_void_p(), C.c_size_t()
    assert O.orc_ref_map(genome.encode(), fa.encode(), n, b, fmt, C.byref(text), C.byref(nb)) == 0
    s = C.string_at(text.value, nb.value).decode()
    O.orc_free(text)
    return s


def groups(s):
    lines = s.rstrip("\n").split("\n") if s else []
    return sorted("\n".join(lines[i:i + 3]) for i in range(0, len(lines), 3))


def golden_groups(name):
    with gzip.open(os.path.join(util.GOLDEN, name), "rt") as f:
        return groups(f.read())


@pytest.fixture(scope="module")
def refmap_inputs(tmp_path_factory):
    c = GOLD["refmap"]
    d = tmp_path_factory.mktemp("refmap_host")
    fa, genome = str(d / "reads.fa"), str(d / "genome.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], genome_out=genome)
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == c["fasta_sha256"]
    return fa, genome


@pytest.fixture(scope="module")
def hard_inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("refmap_hard_host")
    fa, genome = str(d / "reads.fa"), str(d / "genome.fa")
    util.make_refmap_hard(fa, genome)
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == GOLD["refmap_hard"]["fasta_sha256"]
    return fa, genome


@pytest.fixture(scope="module")
def repeat_inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("refmap_repeats_host")
    fa, genome = str(d / "reads.fa"), str(d / "genome.fa")
    util.make_refmap_repeats(fa, genome, seed=2, num_reads=120)
    return fa, genome


@pytest.fixture(scope="module")
def repeat_small(repeat_inputs, tmp_path_factory):
    """The first 40 reads of the repeat-rich fixture (for the slower instrumented runs)."""
    fa, genome = repeat_inputs
    small = str(tmp_path_factory.mktemp("refmap_repeats_small") / "reads.fa")
    with open(fa, "rb") as f, open(small, "wb") as g:
        g.write(b"".join(f.readlines()[:80]))
    return small, genome


def test_m4_matches_reference(refmap_inputs):
    fa, genome = refmap_inputs
    s, st = run_harness(genome, fa, fmt=1)
    with gzip.open(os.path.join(util.GOLDEN, "refmap.m4.gz"), "rt") as f:
        want = f.read().splitlines()
    got = sorted(s.splitlines())
    assert len(got) == len(want) == GOLD["refmap"]["num_m4"]
    assert got == want
    assert st[0] >= len(got)        # every record is an extension that ran


def test_ref_format_matches_reference_in_small_batches(refmap_inputs):
    """Several ABI-sized calls of 70 reads and a table budget small enough for many table batches per call."""
    fa, genome = refmap_inputs
    s, st = run_harness(genome, fa, fmt=0, per_call=70, budget=200_000)
    assert st[1] > 10
    assert groups(s) == golden_groups("refmap.ref.gz")


def test_hard_inputs_match_reference(hard_inputs):
    """Three contigs with a shared repeat and a run of N, chimeric reads (clipped ends: RescueFn), very noisy reads (the
    second pass), reads with N and lower-case stretches (explicit reverse strands, k-mers that are not looked up), short
    reads."""
    fa, genome = hard_inputs
    s, st = run_harness(genome, fa, fmt=0)
    assert st[1] == 2               # both passes ran
    assert st[2] > 20               # clipped ends went through RescueFn
    assert groups(s) == golden_groups("refmap_hard.ref.gz")


@pytest.mark.skipif(not os.path.exists(os.path.join(util.REF_DIR, "mecat2ref")), reason="needs the unmodified binary (oracle/_ref, built where /root/reference exists)")
@pytest.mark.parametrize("n,b", [(10, 10), (40, 5)])
def test_repeat_rich_inputs_match_the_unmodified_binary(tmp_path, repeat_inputs, n, b):
    """Differential run against oracle/_ref/mecat2ref itself on a repeat-rich genome (util.make_refmap_repeats): full
    candidate lists, block-consuming votes, rescue between repeat copies, both passes; small calls and table batches."""
    import subprocess
    (fa, genome), out = repeat_inputs, str(tmp_path / "ref.out")
    subprocess.check_call([os.path.join(util.REF_DIR, "mecat2ref"), "-d", fa, "-r", genome, "-o", out, "-w", str(tmp_path / "w"), "-t", "4", "-m", "0",
                           "-n", str(n), "-b", str(b)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=str(tmp_path))
    got, st = run_harness(genome, fa, n, b, 0, per_call=50, budget=3_000_000)
    want = groups(open(out).read())
    assert len(want) > 3 * 120 and st[1] > 6
    assert groups(got) == want


def test_kernel_bodies_under_sanitizers(hard_inputs, repeat_small):
    """ASan + UBSan over the stage sequence and kernel bodies (small calls, tiny table budget, -n above the list sizes) on
    the hard fixture and on a repeat-rich one."""
    import subprocess
    exe = os.path.join(util.ROOT, "tests", "_build", "ref_sanitize")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    util.build_oracle()
    cmd = ["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-pthread", "-o", exe, os.path.join(util.ROOT, "tests", "ref_sanitize_main.cpp"), os.path.join(util.ROOT, "tests", "ref_host_harness.cpp"),
           "-L", util.ORACLE_DIR, "-loracle", "-Wl,-rpath," + util.ORACLE_DIR]
    if util.stale(exe, util.REF_HOST_SOURCES) and subprocess.run(cmd, capture_output=True).returncode != 0:
        pytest.skip("this toolchain has no sanitizer runtime")
    fa, genome = hard_inputs
    rfa, rgenome = repeat_small
    for args in ([genome, fa, "10", "10", "0", "50", "200000"], [genome, fa, "3", "2", "2", "0", "0"], [rgenome, rfa, "40", "5", "1", "25", "3000000"]):
        p = subprocess.run([exe] + args, capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
        if "Shadow memory range interleaves" in p.stderr or "ReserveShadowMemoryRange failed" in p.stderr:
            pytest.skip("AddressSanitizer cannot start in this environment")
        assert p.returncode == 0 and "rc=0" in p.stdout and "runtime error" not in p.stderr and "AddressSanitizer" not in p.stderr, p.stderr[-3000:]


def test_strings_for_printed_records_only(hard_inputs, repeat_small, monkeypatch):
    """The opt-in route that extends every candidate for its coordinates and computes alignment strings only for the
    records that are printed: same text, far fewer extensions with strings on the repeat-rich genome."""
    fa, genome = repeat_small
    base, st0 = run_harness(genome, fa, 40, 5, 0, per_call=25)
    monkeypatch.setenv("MECAT_HARNESS_STRINGS_FOR_PRINTED_ONLY", "1")
    got, st1 = run_harness(genome, fa, 40, 5, 0, per_call=25)
    assert got == base
    assert st1[3] == len(groups(got)) and st1[3] * 3 < st0[3]
    fa, genome = hard_inputs
    assert groups(run_harness(genome, fa, fmt=0)[0]) == golden_groups("refmap_hard.ref.gz")
    _, want = golden_sam()
    assert sorted(run_harness(genome, fa, fmt=2, per_call=40)[0].splitlines()) == want


def golden_sam():
    with gzip.open(os.path.join(util.GOLDEN, "refmap_hard.sam.gz"), "rt") as f:
        lines = f.read().splitlines()
    return [l for l in lines if l.startswith("@")], [l for l in lines if not l.startswith("@")]


def test_sam_records_match_reference(hard_inputs):
    """-m 2: flag, 1-based position, CIGAR with hard clips and SEQ of every record of the unmodified binary's SAM file."""
    fa, genome = hard_inputs
    _, want = golden_sam()
    got = sorted(run_harness(genome, fa, fmt=2)[0].splitlines())
    assert len(got) == len(want) == GOLD["refmap_hard"]["num_sam"]
    assert got == want
    assert sorted(packed_via_python(genome, fa, 2).splitlines()) == want


@pytest.mark.parametrize("n,b", [(3, 2), (1, 1), (50, 4)])
def test_candidate_and_output_caps_match_oracle(refmap_inputs, hard_inputs, n, b):
    fa, genome = hard_inputs
    assert groups(run_harness(genome, fa, n, b, 0)[0]) == groups(run_oracle(genome, fa, n, b, 0))
    fa, genome = refmap_inputs
    assert sorted(run_harness(genome, fa, n, b, 1, per_call=37)[0].splitlines()) == sorted(run_oracle(genome, fa, n, b, 1).splitlines())


def test_cfg0_sized_reads_match_reference(tmp_path):
    """1 000 x 15 kb reads (BASELINE configs[0]'s read set) against their 1 Mb genome, M4 output."""
    c = GOLD["refmap_cfg0"]
    fa, genome = str(tmp_path / "reads.fa"), str(tmp_path / "genome.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], genome_out=genome)
    got = sorted(run_harness(genome, fa, fmt=1)[0].splitlines())
    with gzip.open(os.path.join(util.GOLDEN, "refmap_cfg0.m4.gz"), "rt") as f:
        want = f.read().splitlines()
    assert len(got) == len(want) == c["num_m4"]
    assert got == want


def test_fastq_reads_and_empty_inputs(tmp_path, refmap_inputs):
    """FASTQ reads are numbered from 1 (chang_fastqfile); a read file without reads maps nothing."""
    fa, genome = refmap_inputs
    seqs = util.read_fasta(fa)[:40]
    fq = str(tmp_path / "reads.fq")
    with open(fq, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b"@r%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    got = run_harness(genome, fq, fmt=1)[0]
    assert got and sorted(got.splitlines()) == sorted(run_oracle(genome, fq, 10, 10, 1).splitlines())
    assert min(int(l.split("\t")[0]) for l in got.splitlines()) >= 1
    empty = str(tmp_path / "none.fa")
    open(empty, "w").close()
    assert run_harness(genome, empty, fmt=1)[0] == ""


def test_awkward_files_parse_like_the_reference(tmp_path, refmap_inputs):
    """Multi-line records with CRLF line ends, blank lines, header descriptions, a lower-case genome, a '>' inside a
    sequence line (chang_fastqfile opens a header there), FASTQ without a final newline and with an incomplete last
    record: the memchr-based readers of host/refio.h against the oracle's character-by-character ones."""
    fa, genome = refmap_inputs
    seqs = util.read_fasta(fa)[:30]
    g = util.read_fasta(genome)[0]
    genome2 = str(tmp_path / "genome.fa")
    with open(genome2, "wb") as f:
        f.write(b">chrA some description\there\r\n")
        half = len(g) // 2
        for i in range(0, half, 70):
            f.write(g[i:min(i + 70, half)].lower() + b"\r\n")
        f.write(b"\n>chrB\n" + g[half:] + b"\n\n")
    reads2 = str(tmp_path / "reads.fa")
    with open(reads2, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">read%d extra words\r\n" % i)
            for k in range(0, len(s), 80):
                f.write(s[k:k + 80] + (b"\r\n" if i % 2 else b"\n"))
            if i == 7:
                f.write(b"\n")
            if i == 11:
                f.write(s[:3000] + b">glued " + b"\n" + s[3000:9000] + b"\n")
    for fmt in (0, 1):
        got, want = run_harness(genome2, reads2, fmt=fmt)[0], run_oracle(genome2, reads2, 10, 10, fmt)
        assert want and got == want           # same records in the same order: one call, all reads found in the first pass
    fq = str(tmp_path / "reads.fq")
    with open(fq, "wb") as f:
        for i, s in enumerate(seqs[:12]):
            f.write(b"@r%d\r\n" % i + s + b"\r\n+\r\n" + b"I" * len(s) + b"\r\n")
        f.write(b"@last\n" + seqs[12] + b"\n+\n" + b"I" * len(seqs[12]))             # complete, no final newline
    got, want = run_harness(genome2, fq, fmt=1)[0], run_oracle(genome2, fq, 10, 10, 1)
    assert want and got == want
    with open(fq, "ab") as f:
        f.write(b"\n@cut\n" + seqs[13] + b"\n+\n")                                     # three lines: not a record
    assert run_harness(genome2, fq, fmt=1)[0] == want == run_oracle(genome2, fq, 10, 10, 1)


def numpy_index(G):
    """(codes, positions) of the kept 13-mers of an api.RefGenome, ordered by (code, position): an independent statement of
    creat_ref_index (every 13-mer inside a run of ACGT, lists of more than 128 dropped) in the product's conventions
    (A0 C1 G2 T3, first base most significant, 0-based starts)."""
    import numpy as np
    n = G.num_bases
    pac = G.pac
    codes = ((pac[np.arange(n) >> 2] >> (((~np.arange(n)) & 3) << 1)) & 3).astype(np.int64)
    good = np.zeros(n, dtype=bool)
    for s0, ln in G.runs:
        good[s0:s0 + ln] = True
    m = n - 12
    kmer = np.zeros(m, dtype=np.int64)
    for j in range(13):
        kmer = (kmer << 2) | codes[j:j + m]
    csum = np.concatenate(([0], np.cumsum(~good)))
    ok = (csum[13:13 + m] - csum[:m]) == 0
    pos = np.flatnonzero(ok)
    kmer = kmer[ok]
    order = np.lexsort((pos, kmer))
    kmer, pos = kmer[order], pos[order]
    uniq, inv, cnt = np.unique(kmer, return_inverse=True, return_counts=True)
    keep = cnt[inv] <= 128
    return kmer[keep], pos[keep]


def test_index_of_the_genome_against_numpy(hard_inputs, tmp_path):
    """The harness's index (the layout the device index has to have: GPU tests compare the two arrays) against numpy_index,
    on the three-contig genome with its N run plus a 200-copy tandem repeat (lists above the cutoff) and runs shorter
    than a k-mer."""
    import numpy as np
    from mecat_b200 import api
    L = util.ref_harness()
    fa, genome = hard_inputs
    G0 = api.RefGenome.from_fasta(genome)
    seqs = [open(genome, "rb").read().split(b"\n")[1], b"ACGTTGCAAGGCT" * 200 + b"NNACGTACGTACGNNNACGTACGTACGTTN" + b"GATTACA" * 30]
    G = api.RefGenome(["a", "b"], seqs)
    for g in (G0, G):
        gc = g.c()
        idx = L.harness_ref_index_build(C.byref(gc))
        begin = np.zeros((1 << 26) + 1, dtype=np.uint32)
        n = L.harness_ref_index_export(idx, begin.ctypes.data_as(C.c_void_p), None)
        pos = np.zeros(max(1, n), dtype=np.int32)
        L.harness_ref_index_export(idx, None, pos.ctypes.data_as(C.c_void_p))
        L.harness_ref_index_release(idx)
        kmer, want = numpy_index(g)
        assert n == len(want) == int(begin[-1])
        assert (pos[:n] == want).all()
        uniq, first = np.unique(kmer, return_index=True)
        assert (begin[uniq] == first).all() and (begin[uniq + 1] == np.append(first[1:], n)).all()
        assert np.count_nonzero(np.diff(begin.astype(np.int64))) == len(uniq)
    assert len(numpy_index(G)[0]) < G.num_bases - 12 - 13 * 150          # the tandem repeat's lists were dropped


def packed_via_python(genome_path, reads_path, fmt, n=10, b=10):
    """Python packing (mecat_b200.api RefGenome / RefReads) -> the host twin of the ABI call -> Python formatting."""
    import numpy as np
    from mecat_b200 import api
    L = util.ref_harness()
    G = api.RefGenome.from_fasta(genome_path)
    seqs = util.read_fasta(reads_path)
    R = api.RefReads(seqs)
    g, r, p = G.c(), R.c(), api.RefParams(n, b, 1 if fmt != 1 else 0, 0)
    res, cnt, qs, ss, nb = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_void_p(), C.c_size_t()
    assert L.harness_ref_map_packed(C.byref(g), C.byref(r), C.byref(p), C.byref(res), C.byref(cnt), C.byref(qs), C.byref(ss), C.byref(nb)) == 0
    rec = np.frombuffer(C.string_at(res.value, cnt.value * api.REF_RESULT_DTYPE.itemsize), dtype=api.REF_RESULT_DTYPE)
    q, s = C.string_at(qs.value, nb.value), C.string_at(ss.value, nb.value)
    for ptr in (res, qs, ss):
        L.harness_free(ptr)
    return api.format_ref_results(G, list(range(len(seqs))), rec, q, s, fmt)


def test_python_packing_and_formatting(refmap_inputs, hard_inputs):
    """The structures mecat_b200/api.py builds for mecat_b200_ref_index_build / mecat_b200_ref_map and the text it formats
    from the records, against the reference's golden output (the GPU tests use the same helpers)."""
    fa, genome = hard_inputs
    assert groups(packed_via_python(genome, fa, 0)) == golden_groups("refmap_hard.ref.gz")
    fa, genome = refmap_inputs
    with gzip.open(os.path.join(util.GOLDEN, "refmap.m4.gz"), "rt") as f:
        want = f.read().splitlines()
    assert sorted(packed_via_python(genome, fa, 1).splitlines()) == want


def test_ddf_integer_form_equals_the_float_forms():
    """|dloc / (dseed * BC) - 1| < 0.25: float32 in insert_loc / find_location, float64 in the neighbour votes, integers in
    the kernels (ref_core.cuh ddf_close).  Exhaustive over block-sized operands for every stride, random wide operands."""
    L = util.ref_harness()
    assert L.harness_ddf_sweep(4200, 420, 5, 20) == 0
    rng = random.Random(5)
    for _ in range(200_000):
        bc = rng.randint(5, 20)
        b = rng.randint(-20_000, 20_000)
        if b == 0:
            assert L.harness_ddf_forms(rng.randint(-10, 10), 0, bc) == 0
            continue
        centre = b * bc * rng.choice((0.75, 1.0, 1.25))
        a = int(centre) + rng.randint(-3, 3)
        assert L.harness_ddf_forms(a, b, bc) in (0, 15), (a, b, bc)


# ---- the command-line driver itself (mecat_b200/csrc/host/mecat2ref.cpp), linked against tests/ref_abi_shim.cpp
def run_driver(args, env=None, ok=True):
    import subprocess
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([util.ref_driver_on_host()] + args, env=e, capture_output=True, text=True)
    assert (p.returncode == 0) == ok, p.stderr[-2000:]
    return p


def test_driver_files_match_reference(tmp_path, refmap_inputs, hard_inputs):
    """ref, sam (with its header) and m4 files of the driver; many small batches through the packing / device / text
    pipeline and two (shim) devices give byte-identical files."""
    fa, genome = hard_inputs
    out = str(tmp_path / "hard.ref")
    run_driver(["-d", fa, "-r", genome, "-o", out, "-w", str(tmp_path / "w"), "-t", "3"])           # -m 0 is the default
    assert groups(open(out).read()) == golden_groups("refmap_hard.ref.gz")
    assert os.path.isdir(str(tmp_path / "w"))
    out = str(tmp_path / "hard.sam")
    run_driver(["-d", fa, "-r", genome, "-o", out, "-w", str(tmp_path / "w"), "-m", "2"])
    lines = open(out).read().splitlines()
    head, want = golden_sam()
    assert [l for l in lines if l.startswith("@") and not l.startswith("@PG")] == head
    assert lines[len(head)].startswith("@PG\tID:0\tVN:0.0.1\tCL:") and lines[len(head)].endswith(" -m 2 \tPN:mecat2ref")
    assert sorted(lines[len(head) + 1:]) == want
    fa, genome = refmap_inputs
    one = str(tmp_path / "one.m4")
    run_driver(["-d", fa, "-r", genome, "-o", one, "-w", str(tmp_path / "w"), "-m", "1"])
    with gzip.open(os.path.join(util.GOLDEN, "refmap.m4.gz"), "rt") as f:
        assert sorted(open(one).read().splitlines()) == f.read().splitlines()
    many = str(tmp_path / "many.m4")
    p = run_driver(["-d", fa, "-r", genome, "-o", many, "-w", str(tmp_path / "w"), "-m", "1"], env={"MECAT_B200_REF_BATCH_BASES": "100000"})
    two = str(tmp_path / "two.m4")
    run_driver(["-d", fa, "-r", genome, "-o", two, "-w", str(tmp_path / "w"), "-m", "1"],
               env={"MECAT_B200_REF_BATCH_BASES": "150000", "MECAT_GPUS": "2", "MECAT_SHIM_DEVICES": "2"})
    assert open(many).read() == open(one).read() == open(two).read()      # reads in input order whatever the batches


def test_driver_option_handling(tmp_path, refmap_inputs):
    """Missing arguments, -b above -n (reset with the reference's warning), -x 1, more devices than visible."""
    fa, genome = refmap_inputs
    out, w = str(tmp_path / "o"), str(tmp_path / "w")
    assert "reference must be specified" in run_driver(["-d", fa, "-o", out, "-w", w], ok=False).stderr
    assert "candidates must be > 0" in run_driver(["-d", fa, "-r", genome, "-o", out, "-w", w, "-n", "0"], ok=False).stderr
    run_driver(["-d", fa, "-r", genome, "-o", out, "-w", w, "-m", "1", "-x", "1"])           # nanopore: the other aligner
    assert sorted(open(out).read().splitlines()) == sorted(golden_lines("refmap.x1.m4.gz"))
    assert "CUDA device" in run_driver(["-d", fa, "-r", genome, "-o", out, "-w", w], env={"MECAT_GPUS": "3"}, ok=False).stderr
    assert "cannot open" in run_driver(["-d", str(tmp_path / "missing.fa"), "-r", genome, "-o", out, "-w", w], ok=False).stderr
    assert "for writing" in run_driver(["-d", fa, "-r", genome, "-o", str(tmp_path / "no" / "such" / "dir" / "o"), "-w", w], ok=False).stderr
    p = run_driver(["-d", fa, "-r", genome, "-o", out, "-w", w, "-m", "1", "-n", "2", "-b", "7"])
    assert "we reset it to 2" in p.stderr
    assert sorted(open(out).read().splitlines()) == sorted(run_oracle(genome, fa, 2, 2, 1).splitlines())


def test_driver_threads_under_thread_sanitizer(tmp_path, hard_inputs):
    """The driver's host threads -- one per device, packing threads, the packer / text tasks running beside the device
    call -- under ThreadSanitizer, two shim devices and many small batches; the file is still the reference's."""
    import subprocess
    exe = os.path.join(util.ROOT, "tests", "_build", "mecat2ref_tsan")
    util.build_oracle()
    cmd = ["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-I", os.path.join(util.ROOT, "include"), "-o", exe,
           os.path.join(util.ROOT, "mecat_b200", "csrc", "host", "mecat2ref.cpp"), os.path.join(util.ROOT, "tests", "ref_abi_shim.cpp"),
           os.path.join(util.ROOT, "tests", "ref_host_harness.cpp"), "-L", util.ORACLE_DIR, "-loracle", "-Wl,-rpath," + util.ORACLE_DIR]
    if util.stale(exe, util.REF_HOST_SOURCES) and subprocess.run(cmd, capture_output=True).returncode != 0:
        pytest.skip("this toolchain has no ThreadSanitizer runtime")
    fa, genome = hard_inputs
    out = str(tmp_path / "hard.ref")
    p = subprocess.run([exe, "-d", fa, "-r", genome, "-o", out, "-w", str(tmp_path / "w"), "-t", "4"], capture_output=True, text=True,
                       env=dict(os.environ, MECAT_B200_REF_BATCH_BASES="150000", MECAT_GPUS="2", MECAT_SHIM_DEVICES="2"))
    if "FATAL: ThreadSanitizer" in p.stderr:
        pytest.skip("ThreadSanitizer cannot start in this environment")
    assert p.returncode == 0 and "ThreadSanitizer" not in p.stderr, p.stderr[-3000:]
    assert groups(open(out).read()) == golden_groups("refmap_hard.ref.gz")


def test_nanopore_stage_sequence_matches_reference(refmap_inputs, hard_inputs):
    """-x 1: the same stages with the other aligner behind the extension hook (mecat2ref_impl_large.cpp:329-332); golden
    of the unmodified binary on the refmap fixture, the pinned oracle on the hard one (rescue between chimeric parts)."""
    L = util.ref_harness()
    L.harness_set_tech(1)
    try:
        fa, genome = refmap_inputs
        assert sorted(run_harness(genome, fa, fmt=1)[0].splitlines()) == sorted(golden_lines("refmap.x1.m4.gz"))
        assert groups(run_harness(genome, fa, fmt=0)[0]) == golden_groups("refmap.x1.ref.gz")
        fa, genome = hard_inputs
        O = util.oracle()
        text, nb = C.c_void_p(), C.c_size_t()
        assert O.orc_ref_map_x(genome.encode(), fa.encode(), 10, 10, 0, 1, C.byref(text), C.byref(nb)) == 0
        want = C.string_at(text.value, nb.value).decode()
        O.orc_free(text)
        assert groups(run_harness(genome, fa, fmt=0, per_call=7)[0]) == groups(want)
    finally:
        L.harness_set_tech(0)


def golden_lines(name):
    with gzip.open(os.path.join(util.GOLDEN, name), "rt") as f:
        return f.read().splitlines()


def test_warp_shaped_seeding_equals_the_scalar_form(refmap_inputs, hard_inputs, repeat_inputs, monkeypatch):
    """SeedWarpFn (a warp per strand: lane-parallel insert_loc / find_location / neighbour votes, the default) and SeedFn
    (a thread per strand) must leave the same output; the repeat-rich fixture fills candidate lists, consumes blocks by
    votes and evicts seeds in most blocks."""
    for (fa, genome), n, fmt in ((refmap_inputs, 10, 1), (hard_inputs, 10, 0), (repeat_inputs, 40, 1)):
        monkeypatch.delenv("MECAT_HARNESS_SEED_PER_THREAD", raising=False)
        a, sa = run_harness(genome, fa, n=n, fmt=fmt)
        monkeypatch.setenv("MECAT_HARNESS_SEED_PER_THREAD", "1")
        b, sb = run_harness(genome, fa, n=n, fmt=fmt)
        assert a == b and sa[0] == sb[0] and len(a) > 1000
