"""GPU parity tests (-m gpu) of the nanopore path, `-x 1`: the X-drop extension kernel
(xdrop.cu) against the oracle's restatement of XdropAligner, and whole mecat2pw tiles with tech = 1 against golden
outputs of the unmodified binary run with -x 1."""
import ctypes as C
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

import util
from util import PackedVolume

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(util.GOLDEN, "golden.json")))


def gold_lines(name, ext):
    with gzip.open(os.path.join(util.GOLDEN, "%s.%s.gz" % (name, ext)), "rt") as f:
        return f.read().splitlines()


def host_volume(v):
    import mecat_b200
    return mecat_b200.HostVolume(v.offset_size, v.pac, v.num_bases, v.start_read_id)


@pytest.fixture(scope="module")
def small_vol():
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f:
        seqs = [l for l in f.read().split(b"\n") if l and not l.startswith(b">")]
    return PackedVolume.from_seqs(seqs)


@pytest.fixture(scope="module")
def cfg0_vol(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg0")
    fa = str(d / "cfg0.fa")
    c = GOLD["cfg0"]
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == c["fasta_sha256"]
    return PackedVolume.from_seqs(util.read_fasta(fa))


def _tasks(small_vol, step=2):
    """Extension tasks: the oracle's own -x 1 candidates of every step-th read (some on a subject window), plus
    unrelated pairs and start points at the ends."""
    O = util.oracle()
    cv = small_vol.c()
    oidx = O.orc_index_build(C.byref(cv))
    p = util.pw_params(task=0, a=500, k=2, x=1)
    out = (C.c_int32 * (12 * 101))()
    tasks = []
    for rid in range(0, small_vol.num_reads, step):
        n = O.orc_pw_candidates(oidx, C.byref(cv), C.byref(cv), rid, C.byref(p), out)
        for i in range(n):
            c = out[12 * i:12 * i + 12]
            qstart, sstart = c[1], c[0]
            if qstart and sstart:
                qstart += 6; sstart += 6
            sidx = c[9]
            if i % 3 == 0:
                sl = int(small_vol.offset_size[sidx][1])
                lo = max(0, sstart - 2500); hi = min(sl, sstart + 3000)
                tasks.append((rid, c[11], qstart, sidx, sstart - lo, lo, hi - lo))
            else:
                tasks.append((rid, c[11], qstart, sidx, sstart, 0, 0))
    O.orc_index_free(oidx)
    rng = np.random.default_rng(8)
    nr = small_vol.num_reads
    for it in range(120):
        a, b = int(rng.integers(0, nr)), int(rng.integers(0, nr))
        la, lb = int(small_vol.offset_size[a][1]), int(small_vol.offset_size[b][1])
        qs, ss = [(0, 0), (la, lb), (la // 2, lb // 3), (int(rng.integers(0, la + 1)), int(rng.integers(0, lb + 1)))][it % 4]
        tasks.append((a, it & 1, qs, b, ss, 0, 0))
    return tasks


def _oracle_xdrop(vq, vs, t, min_aln):
    O = util.oracle()
    q = np.concatenate([[0], vq.codes(int(t[0]), int(t[1])), [0]]).astype(np.int8)
    full = vs.codes(int(t[3]), 0)
    if int(t[6]) > 0:
        full = full[int(t[5]):int(t[5]) + int(t[6])]
    s = np.concatenate([[0], full, [0]]).astype(np.int8)
    cap = len(q) + len(s) + 64
    qa, sa = C.create_string_buffer(cap), C.create_string_buffer(cap)
    out = (C.c_int32 * 8)()
    ident = C.c_double()
    O.orc_xdrop_go(C.cast(q.ctypes.data + 1, C.c_char_p), int(t[2]), len(q) - 2, C.cast(s.ctypes.data + 1, C.c_char_p), int(t[4]),
                   len(s) - 2, min_aln, out, C.byref(ident), qa, sa, cap)
    return tuple(out[:7]), ident.value, qa.value, sa.value


@pytest.mark.parametrize("min_aln", [500, 1])
def test_xdrop_with_strings_matches_oracle(gpu_ctx, small_vol, min_aln):
    """mecat_b200_align_batch policy 2 = XdropAligner::go + mapped strings (xdrop_gapalign.cpp:351-439)."""
    import mecat_b200
    tasks = np.array(_tasks(small_vol), dtype=mecat_b200.ALIGN_TASK_DTYPE)
    d = gpu_ctx.upload(host_volume(small_vol))
    res, qstr, sstr = gpu_ctx.align_batch(d, d, tasks, min_aln, policy=2, err=0.0)
    gpu_ctx.release_volume(d)
    bad, nok = [], 0
    for i, t in enumerate(tasks):
        w, wid, wq, ws = _oracle_xdrop(small_vol, small_vol, tuple(int(x) for x in t), min_aln)
        r = res[i]
        g = (int(r["ok"]), int(r["qstart"]), int(r["qend"]), int(r["sstart"]), int(r["send"]), int(r["columns"]), int(r["matches"]))
        if g != w:
            bad.append((i, tuple(int(x) for x in t), g, w))
            continue
        if g[0]:
            o, n = int(r["str_offset"]), g[5]
            if qstr[o:o + n] != wq[:n] or sstr[o:o + n] != ws[:n] or qstr[o + n] != 0 or float(r["ident"]) != wid:
                bad.append((i, tuple(int(x) for x in t), "strings", n))
            nok += 1
    assert nok > 100
    assert not bad, "%d of %d differ; first %r" % (len(bad), len(tasks), bad[0])


def test_xdrop_string_free_matches_oracle(gpu_ctx, small_vol):
    """mecat_b200_extend_batch policy 2: the accessors of XdropAligner::go without the strings (what mecat2pw -j 1 needs)."""
    import mecat_b200
    full = [t for t in _tasks(small_vol, 3) if t[6] == 0]
    tasks = np.array([t[:5] for t in full], dtype=mecat_b200.TASK_DTYPE)
    d = gpu_ctx.upload(host_volume(small_vol))
    got = gpu_ctx.extend_batch(d, d, tasks, 500, policy=2)
    gpu_ctx.release_volume(d)
    bad = []
    for i, t in enumerate(full):
        w, wid, _, _ = _oracle_xdrop(small_vol, small_vol, t, 500)
        g = got[i]
        gg = (int(g["ok"]), int(g["qstart"]), int(g["qend"]), int(g["sstart"]), int(g["send"]), int(g["columns"]), int(g["matches"]))
        if gg != w or float(g["ident"]) != wid:
            bad.append((i, t, gg, w))
    assert not bad, "%d of %d differ; first %r" % (len(bad), len(full), bad[0])


def test_small_nanopore_tiles_match_reference(gpu_ctx, small_vol):
    """mecat2pw -x 1 (-j 0 and -j 1) on the small fixture against the unmodified binary."""
    import mecat_b200
    hv = host_volume(small_vol)
    ec = gpu_ctx.pw_candidates(hv, hv, mecat_b200.pw_params(task=0, min_align_size=500, min_kmer_match=2, tech=1))
    assert util.ec_lines(ec) == gold_lines("small.x1", "can")
    m4 = gpu_ctx.pw_overlaps(hv, hv, mecat_b200.pw_params(task=1, min_align_size=500, min_kmer_match=2, tech=1))
    assert util.m4_lines(m4, gapped=True) == gold_lines("small.x1", "m4")


def test_cfg0_nanopore_m4_matches_reference(gpu_ctx, cfg0_vol):
    """BASELINE configs[0]-sized reads (1 000 x 15 kb) through mecat2pw -j 1 -x 1."""
    import mecat_b200
    hv = host_volume(cfg0_vol)
    m4 = gpu_ctx.pw_overlaps(hv, hv, mecat_b200.pw_params(task=1, min_align_size=500, min_kmer_match=2, tech=1))
    print("[x1 cfg0] stats", {k: v for k, v in gpu_ctx.stats().items() if "extend" in str(k) or "kernel" in str(k)}, "records", len(m4))
    assert util.m4_lines(m4, gapped=True) == gold_lines("cfg0.x1", "m4")


# ---------------------------------------------------------------- mecat2ref -x 1
def _groups(s):
    lines = s.rstrip("\n").split("\n") if s else []
    return sorted("\n".join(lines[i:i + 3]) for i in range(0, len(lines), 3))


def test_mecat2ref_nanopore_matches_reference(gpu_ctx, tmp_path):
    """mecat2ref -x 1 (XdropAligner behind extend_candidate, mecat2ref_impl_large.cpp:329-332) through the C ABI and
    through the command-line driver, M4 and ref format, against the unmodified binary."""
    import subprocess
    from mecat_b200 import api
    c = GOLD["refmap"]
    fa, genome = str(tmp_path / "reads.fa"), str(tmp_path / "genome.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], genome_out=genome)
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == c["fasta_sha256"]
    G = api.RefGenome.from_fasta(genome)
    seqs = util.read_fasta(fa)
    idx = gpu_ctx.ref_index_build(G)
    try:
        for fmt, gold in ((1, "refmap.x1.m4.gz"), (0, "refmap.x1.ref.gz")):
            rec, q, s = gpu_ctx.ref_map(idx, api.RefReads(seqs), 10, 10, want_strings=fmt != 1, tech=1)
            text = api.format_ref_results(G, list(range(len(seqs))), rec, q, s, fmt)
            with gzip.open(os.path.join(util.GOLDEN, gold), "rt") as f:
                want = f.read()
            if fmt == 1:
                assert sorted(text.splitlines()) == want.splitlines()
            else:
                assert _groups(text) == _groups(want)
    finally:
        gpu_ctx.release_ref_index(idx)
    out = str(tmp_path / "cli.m4")
    p = subprocess.run([os.path.join(util.ROOT, "mecat_b200", "bin", "mecat2ref"), "-d", fa, "-r", genome, "-o", out, "-w", str(tmp_path / "w"),
                        "-t", "4", "-m", "1", "-x", "1"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    with gzip.open(os.path.join(util.GOLDEN, "refmap.x1.m4.gz"), "rt") as f:
        assert sorted(open(out).read().splitlines()) == f.read().splitlines()


# ---------------------------------------------------------------- mecat2cns -x 1
def _gold_fasta(name, tag):
    with gzip.open(os.path.join(util.GOLDEN, "%s.%s.fa.gz" % (name, tag)), "rt") as f:
        lines = f.read().splitlines()
    return sorted(zip(lines[0::2], lines[1::2]))


def _gold_can(name):
    import io
    import mecat_b200
    with gzip.open(os.path.join(util.GOLDEN, "%s.can.gz" % name), "rt") as f:
        return mecat_b200.read_can(io.StringIO(f.read()))


def _cns_x1(gpu_ctx, vol, can):
    import mecat_b200
    d = gpu_ctx.upload(host_volume(vol))
    ec = mecat_b200.normalise_candidates(can, 2000)
    pieces = gpu_ctx.cns_reads(d, ec, 0.4, 400, 6, 2000, tech=1)      # the -x 1 defaults, options.cpp:21-29
    gpu_ctx.release_volume(d)
    return sorted((">%d_%d_%d_%d" % (i, b, e, len(s)), s.decode()) for i, b, e, s in pieces)


def test_nanopore_consensus_matches_reference(gpu_ctx, small_vol, tmp_path):
    """mecat2cns -x 1 -i 0 (consensus_one_read_can_nanopore, mecat_correction.cpp:453-512: error rate 0.20 -> the wide
    instance of the cns-flavour extension kernel, up to 100 alignments per read, the whole read as the effective range)
    against the corrected FASTA of the unmodified binary: the small fixture, the ~120x deep fixture, the command line."""
    import subprocess
    got = _cns_x1(gpu_ctx, small_vol, _gold_can("small.x1"))
    want = _gold_fasta("small.x1", "cns")
    assert len(got) == len(want) == GOLD["x1"]["small_num_cns"]
    assert got == want
    c = GOLD["deep"]
    fa = str(tmp_path / "deep.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
    vol = PackedVolume.from_seqs(util.read_fasta(fa))
    got = _cns_x1(gpu_ctx, vol, _gold_can("deep"))
    want = _gold_fasta("deep.x1", "cns")
    assert len(got) == len(want) == GOLD["x1"]["deep_num_cns"]
    bad = [g[0] for g, w in zip(got, want) if g != w]
    assert not bad, bad[:5]
    # command line: mecat2pw -j 0 -x 1 | mecat2cns -x 1 -i 0
    reads = str(tmp_path / "small.fa")
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f, open(reads, "wb") as g:
        g.write(f.read())
    can, out = str(tmp_path / "x1.can"), str(tmp_path / "x1.cns.fa")
    bindir = os.path.join(util.ROOT, "mecat_b200", "bin")
    p = subprocess.run([os.path.join(bindir, "mecat2pw"), "-j", "0", "-x", "1", "-d", reads, "-o", can, "-w", str(tmp_path / "wrk"), "-t", "4"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    assert sorted(open(can).read().splitlines()) == gold_lines("small.x1", "can")
    p = subprocess.run([os.path.join(bindir, "mecat2cns"), "-x", "1", "-i", "0", "-t", "4", can, reads, out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = open(out).read().splitlines()
    assert sorted(zip(lines[0::2], lines[1::2])) == _gold_fasta("small.x1", "cns")


def test_nanopore_m4_input_consensus_matches_reference(gpu_ctx, small_vol, tmp_path):
    """mecat2cns -x 1 (which means -i 1: consensus_one_read_m4_nanopore, mecat_correction.cpp:303-360 -- a partition ordered
    by std::sort on sid, the 100 largest overlaps of a read, every alignment that succeeds and passes the mapping-ratio
    test) on the -x 1 overlaps of the small fixture: golden of the unmodified binary with one OpenMP thread, through the
    C ABI and through the command line."""
    import subprocess
    import mecat_b200
    with gzip.open(os.path.join(util.GOLDEN, "small.x1.m4.gz"), "rt") as f:
        parts = mecat_b200.m4_partitions(f, 0.4, 2000)
    assert list(parts) == [0]
    d = gpu_ctx.upload(host_volume(small_vol))
    pieces = gpu_ctx.cns_reads(d, parts[0], 0.4, 400, 6, 2000, tech=1, input_type=1)
    gpu_ctx.release_volume(d)
    got = sorted((">%d_%d_%d_%d" % (i, b, e, len(s)), s.decode()) for i, b, e, s in pieces)
    want = _gold_fasta("small.x1i1", "cns")
    assert len(got) == len(want) == GOLD["i1"]["small_x1_num_cns"]
    bad = [g[0] for g, w in zip(got, want) if g != w]
    assert not bad, bad[:5]
    reads, m4, out = str(tmp_path / "small.fa"), str(tmp_path / "small.x1.m4"), str(tmp_path / "x1i1.cns.fa")
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f, open(reads, "wb") as g:
        g.write(f.read())
    with gzip.open(os.path.join(util.GOLDEN, "small.x1.m4.gz"), "rb") as f, open(m4, "wb") as g:
        g.write(f.read())
    p = subprocess.run([os.path.join(util.ROOT, "mecat_b200", "bin", "mecat2cns"), "-x", "1", "-t", "4", m4, reads, out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = open(out).read().splitlines()
    assert sorted(zip(lines[0::2], lines[1::2])) == want
