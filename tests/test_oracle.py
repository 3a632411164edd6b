"""CPU tests (no GPU): the oracle restatement against (a) golden outputs of the unmodified
reference binaries (tests/golden, made by tests/golden/make_golden.py) and (b) the unmodified
reference itself at function level through oracle/_ref/libmecatref.so when that was built."""
import ctypes as C
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

import util
from util import PackedVolume, pw_params

GOLD = json.load(open(os.path.join(util.GOLDEN, "golden.json")))


def gold_lines(name, ext):
    with gzip.open(os.path.join(util.GOLDEN, "%s.%s.gz" % (name, ext)), "rt") as f:
        return f.read().splitlines()


@pytest.fixture(scope="module")
def small_vol():
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f:
        tmp = f.read()
    seqs = [l for l in tmp.split(b"\n") if l and not l.startswith(b">")]
    return PackedVolume.from_seqs(seqs)


def vol_file_bytes(v):
    hdr = np.array([v.num_reads, v.num_bases, v.start_read_id], dtype="<i4").tobytes()
    return hdr + v.offset_size.astype("<i4").tobytes() + v.pac[:(v.num_bases + 3) // 4].tobytes()


def test_volume_layout_matches_reference_file(small_vol):
    # split_raw_dataset + dump_volume, split_database.cpp:136-153,222-266
    assert hashlib.sha256(vol_file_bytes(small_vol)).hexdigest() == GOLD["small"]["vol0_sha256"]


def test_oracle_can_matches_reference_binary(small_vol):
    ec = util.oracle_pw_tile(small_vol, small_vol, pw_params(task=0), threads=4)
    assert util.ec_lines(ec) == gold_lines("small", "can")


def test_oracle_m4_matches_reference_binary(small_vol):
    m4 = util.oracle_pw_tile(small_vol, small_vol, pw_params(task=1), threads=4)
    assert util.m4_lines(m4, gapped=True) == gold_lines("small", "m4")


def test_oracle_nanopore_matches_reference_binary(small_vol):
    """-x 1: min_kmer_dist 400 and -k 2 for the candidates, XdropAligner and -a 500 for the overlaps
    (pw_impl.cpp:638-642,843-849, pw_options.cpp:43-49); goldens of the unmodified binary."""
    ec = util.oracle_pw_tile(small_vol, small_vol, pw_params(task=0, a=500, k=2, x=1), threads=4)
    assert util.ec_lines(ec) == gold_lines("small.x1", "can")
    m4 = util.oracle_pw_tile(small_vol, small_vol, pw_params(task=1, a=500, k=2, x=1), threads=4)
    assert util.m4_lines(m4, gapped=True) == gold_lines("small.x1", "m4")


def test_oracle_num_candidates_cap(small_vol):
    # -n 3 keeps the first 3 lines per read of the -n 100 output
    full = util.oracle_pw_tile(small_vol, small_vol, pw_params(task=0), threads=4)
    cut = util.oracle_pw_tile(small_vol, small_vol, pw_params(task=0, n=3), threads=4)
    keep = []
    seen = {}
    for e in full:
        k = int(e["qid"])
        seen[k] = seen.get(k, 0) + 1
        if seen[k] <= 3:
            keep.append(e)
    assert util.ec_lines(cut) == util.ec_lines(np.array(keep, dtype=util.EC_DTYPE))


@pytest.fixture(scope="module")
def cfg0_vol(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg0")
    fa = str(d / "cfg0.fa")
    c = GOLD["cfg0"]
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
    h = hashlib.sha256(open(fa, "rb").read()).hexdigest()
    assert h == c["fasta_sha256"], "read generator is not reproducible on this machine"
    return PackedVolume.from_seqs(util.read_fasta(fa))


def test_cfg0_volume_and_can(cfg0_vol):
    # BASELINE.json configs[0]: mecat2pw -j 0 on 1k synthetic CLR reads
    assert hashlib.sha256(vol_file_bytes(cfg0_vol)).hexdigest() == GOLD["cfg0"]["vol0_sha256"]
    ec = util.oracle_pw_tile(cfg0_vol, cfg0_vol, pw_params(task=0), threads=8)
    assert util.ec_lines(ec) == gold_lines("cfg0", "can")


def test_cfg0_m4(cfg0_vol):
    m4 = util.oracle_pw_tile(cfg0_vol, cfg0_vol, pw_params(task=1), threads=8)
    assert util.m4_lines(m4, gapped=True) == gold_lines("cfg0", "m4")


def test_two_volume_tile_ids(small_vol):
    """Off-diagonal tile: index volume = first half, query volume = second half with its own
    start_read_id.  sid <= qid always holds, self hits do not occur."""
    n = small_vol.num_reads // 2
    seqs = [bytes(b"ACGT"[c] for c in small_vol.codes(i)) for i in range(small_vol.num_reads)]
    a = PackedVolume.from_seqs(seqs[:n], 0)
    b = PackedVolume.from_seqs(seqs[n:], n)
    ec = util.oracle_pw_tile(a, b, pw_params(task=0), threads=4)
    assert len(ec) > 0
    assert (ec["sid"] < n).all() and (ec["qid"] >= n).all()


# ---------------------------------------------------------------- integer DDF == float DDF
def test_ddf_integer_form_equals_float_forms():
    a = np.arange(-4200, 4200, dtype=np.int64)[:, None]
    b = np.arange(-2100, 2100, dtype=np.int64)[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        f32 = np.abs((a.astype(np.float32) / (b.astype(np.float32) * np.float32(10.0))).astype(np.float64) - 1.0) < 0.25
        f64 = np.abs(a.astype(np.float64) / (b.astype(np.float64) * 10.0) - 1.0) < 0.25
    a2 = 2 * a
    integer = np.where(b > 0, (15 * b < a2) & (a2 < 25 * b), np.where(b < 0, (25 * b < a2) & (a2 < 15 * b), False))
    assert (integer == f32).all()
    assert (integer == f64).all()
    # wide operands of the neighbour vote (float64 in the reference)
    rng = np.random.default_rng(1)
    a = rng.integers(-600000, 600000, size=2000000)
    b = rng.integers(-40000, 40000, size=2000000)
    with np.errstate(divide="ignore", invalid="ignore"):
        f64 = np.abs(a / (b * 10.0) - 1.0) < 0.25
    a2 = 2 * a
    integer = np.where(b > 0, (15 * b < a2) & (a2 < 25 * b), np.where(b < 0, (25 * b < a2) & (a2 < 15 * b), False))
    assert (integer == f64).all()


# ---------------------------------------------------------------- function level vs the real reference
needs_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/libmecatref.so not built (needs /root/reference)")


def ref_volume(seqs):
    R = util.ref()
    total = sum(len(s) + 1 for s in seqs)
    v = R.ref_volume_new(total + 16)
    for s in seqs:
        R.ref_volume_add(v, bytes(s), len(s))
    return v


@needs_ref
def test_oracle_functions_against_reference(small_vol):
    R, O = util.ref(), util.oracle()
    seqs = [bytes(b"ACGT"[c] for c in small_vol.codes(i)) for i in range(small_vol.num_reads)]
    rv = ref_volume(seqs)
    assert R.ref_volume_num_bases(rv) == small_vol.num_bases
    ridx = R.ref_index_create(rv, 1)
    cv = small_vol.c()
    oidx = O.orc_index_build(C.byref(cv))
    # A1: index lists for the k-mers of a few reads
    rng = np.random.default_rng(5)
    codes = set(int(x) for x in rng.integers(0, 1 << 26, size=2000))
    for rid in range(0, 20):
        c = small_vol.codes(rid).astype(np.int64)
        for p in range(0, len(c) - 13, 37):
            code = 0
            for j in range(13):
                code = (code << 2) | int(c[p + j])
            codes.add(code)
    lst = C.POINTER(C.c_int32)()
    for code in codes:
        n = R.ref_index_count(ridx, code)
        assert O.orc_index_lookup(oidx, code, C.byref(lst)) == n
        if n:
            rl = R.ref_index_list(ridx, code)
            assert [rl[i] for i in range(n)] == [lst[i] for i in range(n)]
    # A2-A4: seeding state of single strands ; A6: candidates
    R.ref_pw_set_options(100, 2000, 4, 0)
    ctx = R.ref_pw_ctx_new(rv, ridx)
    cap = 4096
    seg_r = (C.c_int32 * cap)(); sc_r = (C.c_int16 * cap)(); bk_r = (C.c_int16 * (cap * 84))()
    seg_o = (C.c_int32 * cap)(); sc_o = (C.c_int16 * cap)(); bk_o = (C.c_int16 * (cap * 82))()
    for rid in range(0, small_vol.num_reads, 9):
        for strand in (0, 1):
            nr = R.ref_pw_seeding_dump(ctx, rv, rid, strand, seg_r, sc_r, bk_r, cap)
            no = O.orc_seeding(oidx, C.byref(cv), C.byref(cv), rid, strand, seg_o, sc_o, bk_o, cap)
            assert nr == no
            assert list(seg_r[:nr]) == list(seg_o[:no])
            assert list(sc_r[:nr]) == list(sc_o[:no])
            for i in range(nr):
                rb = bk_r[i * 84:(i + 1) * 84]   # Back_List: score, loczhi[40], seedno[40], seednum, index
                ob = bk_o[i * 82:(i + 1) * 82]
                assert rb[0] == ob[0]
                m = min(rb[0], 40)
                assert rb[1:1 + m] == ob[1:1 + m] and rb[41:41 + m] == ob[41:41 + m]
    out_r = (C.c_int32 * (12 * 101))(); out_o = (C.c_int32 * (12 * 101))()
    p = pw_params(task=0)
    for rid in range(small_vol.num_reads):
        nr = R.ref_pw_candidates(ctx, rv, rid, 0, out_r)
        no = O.orc_pw_candidates(oidx, C.byref(cv), C.byref(cv), rid, C.byref(p), out_o)
        assert nr == no and list(out_r[:12 * nr]) == list(out_o[:12 * no])
    R.ref_pw_ctx_free(ctx)
    O.orc_index_free(oidx)
    R.ref_index_free(ridx)
    R.ref_volume_free(rv)


@needs_ref
@pytest.mark.parametrize("copies", [10, 18])
def test_oracle_on_repeat_rich_reads_against_reference(copies):
    """Long k-mer lists, the >128 cutoff, insert_loc far from self hits, full candidate lists."""
    R, O = util.ref(), util.oracle()
    reads = util.repeat_reads(copies=copies, n_reads=120)
    vol = PackedVolume.from_seqs(reads)
    rv = ref_volume(reads)
    ridx = R.ref_index_create(rv, 1)
    cv = vol.c()
    oidx = O.orc_index_build(C.byref(cv))
    R.ref_pw_set_options(100, 2000, 4, 0)
    ctx = R.ref_pw_ctx_new(rv, ridx)
    out_r = (C.c_int32 * (12 * 101))(); out_o = (C.c_int32 * (12 * 101))()
    p = pw_params(task=0)
    full = 0
    for rid in range(vol.num_reads):
        nr = R.ref_pw_candidates(ctx, rv, rid, 0, out_r)
        no = O.orc_pw_candidates(oidx, C.byref(cv), C.byref(cv), rid, C.byref(p), out_o)
        assert nr == no and list(out_r[:12 * nr]) == list(out_o[:12 * no])
        full += nr == 100
    assert full > 0   # the -n cap is exercised
    R.ref_pw_ctx_free(ctx)
    O.orc_index_free(oidx)
    R.ref_index_free(ridx)
    R.ref_volume_free(rv)


def mutate(rng, s, err):
    out = []
    for b in s:
        u = rng.random()
        if u < err * 0.3:
            continue
        if u < err * 0.4:
            out.append((b + 1 + rng.integers(0, 3)) & 3)
        else:
            out.append(b)
        while rng.random() < err * 0.6:
            out.append(rng.integers(0, 4))
    return np.array(out, dtype=np.int8)


@needs_ref
def test_oracle_extension_against_reference():
    """A8-A11 on random pairs: related sequences at several divergences, unrelated ones,
    tiny ones, start points at either end."""
    R, O = util.ref(), util.oracle()
    rng = np.random.default_rng(11)
    aligner = R.ref_diff_new()
    out_r = (C.c_int32 * 8)(); out_o = (C.c_int32 * 8)()
    id_r, id_o = C.c_double(), C.c_double()
    qs_r, ts_r = C.c_char_p(), C.c_char_p()
    cap = 60000
    qs_o, ts_o = C.create_string_buffer(cap), C.create_string_buffer(cap)
    cases = []
    for it in range(60):
        L = int(rng.integers(50, 6000))
        base = rng.integers(0, 4, size=L).astype(np.int8)
        err = [0.0, 0.05, 0.15, 0.3, 0.6][it % 5]
        q = mutate(rng, base, err)
        t = mutate(rng, base, err)
        if it % 7 == 0:
            t = rng.integers(0, 4, size=L).astype(np.int8)
        if len(q) < 20 or len(t) < 20:
            continue
        f = rng.random()
        qstart = int(f * len(q)); tstart = min(len(t), int(f * len(t)))
        if it % 11 == 0:
            qstart, tstart = 0, 0
        if it % 13 == 0:
            qstart, tstart = len(q), len(t)
        cases.append((q, qstart, t, tstart))
    for q, qstart, t, tstart in cases:
        # the reference reads query[-1] style offsets: keep one byte of slack in front
        qb = np.concatenate([[0], q, [0]]).astype(np.int8)
        tb = np.concatenate([[0], t, [0]]).astype(np.int8)
        qp = qb.ctypes.data + 1
        tp = tb.ctypes.data + 1
        for min_aln in (1, 2000):
            R.ref_diff_go(aligner, qp, qstart, len(q), tp, tstart, len(t), min_aln, out_r, C.byref(id_r), C.byref(qs_r), C.byref(ts_r))
            O.orc_diff_go(C.cast(qp, C.c_char_p), qstart, len(q), C.cast(tp, C.c_char_p), tstart, len(t), min_aln, out_o,
                          C.byref(id_o), qs_o, ts_o, cap)
            assert list(out_r[:6]) == list(out_o[:6]), (len(q), len(t), qstart, tstart)
            assert id_r.value == id_o.value
            n = out_r[5]
            assert qs_r.value[:n] == qs_o.value[:n] and ts_r.value[:n] == ts_o.value[:n]
    # single blocks incl. the unaligned fall-back
    blk_r = (C.c_int32 * 8)(); blk_o = (C.c_int32 * 8)()
    for it in range(200):
        ql = int(rng.integers(1, 720)); tl = int(rng.integers(1, 720))
        q = rng.integers(0, 4, size=ql).astype(np.int8)
        t = mutate(rng, q, [0.1, 0.3, 0.9][it % 3])[:tl] if it % 4 else rng.integers(0, 4, size=tl).astype(np.int8)
        if len(t) == 0:
            continue
        fwd = it % 2
        qb = np.concatenate([q[::-1], q]).astype(np.int8) if not fwd else q
        tb = np.concatenate([t[::-1], t]).astype(np.int8) if not fwd else t
        qp = qb.ctypes.data + (len(q) - 1 if not fwd else 0)
        tp = tb.ctypes.data + (len(t) - 1 if not fwd else 0)
        R.ref_diff_align_block(aligner, qp, len(q), tp, len(t), fwd, blk_r)
        O.orc_diff_align_block(qp, len(q), tp, len(t), fwd, blk_o)
        assert list(blk_r) == list(blk_o), (it, len(q), len(t), fwd)
    R.ref_diff_free(aligner)


def xdrop_cases(rng, n=70):
    """Pairs for the nanopore extension: related sequences at several divergences (substitutions included, which
    the diff aligner never emits), unrelated ones, tiny ones, start points at either end and at the ends."""
    cases = []
    for it in range(n):
        L = int(rng.integers(30, 5000))
        base = rng.integers(0, 4, size=L).astype(np.int8)
        err = [0.0, 0.05, 0.12, 0.2, 0.35, 0.6][it % 6]
        q = mutate(rng, base, err)
        t = mutate(rng, base, err)
        if it % 7 == 0:
            t = rng.integers(0, 4, size=L).astype(np.int8)
        if it % 9 == 0:      # low complexity: wide bands
            q = np.tile(np.array([0, 1], dtype=np.int8), L // 2)
            t = mutate(rng, q, 0.1)
        if len(q) < 10 or len(t) < 10:
            continue
        f = rng.random()
        qstart = int(f * len(q)); tstart = min(len(t), int(f * len(t)))
        if it % 11 == 0:
            qstart, tstart = 0, 0
        if it % 13 == 0:
            qstart, tstart = len(q), len(t)
        cases.append((q, qstart, t, tstart))
    return cases


@needs_ref
def test_oracle_xdrop_extension_against_reference():
    """XdropAligner::go (the -x 1 aligner, xdrop_gapalign.cpp) on random pairs: coordinates, identity and both
    alignment strings of the restatement equal the unmodified class."""
    R, O = util.ref(), util.oracle()
    rng = np.random.default_rng(23)
    aligner = R.ref_xdrop_new()
    out_r = (C.c_int32 * 8)(); out_o = (C.c_int32 * 8)()
    id_r, id_o = C.c_double(), C.c_double()
    qs_r, ts_r = C.c_char_p(), C.c_char_p()
    cap = 60000
    qs_o, ts_o = C.create_string_buffer(cap), C.create_string_buffer(cap)
    nonempty = 0
    for q, qstart, t, tstart in xdrop_cases(rng):
        qb = np.concatenate([[0], q, [0]]).astype(np.int8)
        tb = np.concatenate([[0], t, [0]]).astype(np.int8)
        qp = qb.ctypes.data + 1
        tp = tb.ctypes.data + 1
        for min_aln in (1, 500):
            R.ref_xdrop_go(aligner, qp, qstart, len(q), tp, tstart, len(t), min_aln, out_r, C.byref(id_r), C.byref(qs_r), C.byref(ts_r))
            O.orc_xdrop_go(C.cast(qp, C.c_char_p), qstart, len(q), C.cast(tp, C.c_char_p), tstart, len(t), min_aln, out_o,
                           C.byref(id_o), qs_o, ts_o, cap)
            assert list(out_r[:6]) == list(out_o[:6]), (len(q), len(t), qstart, tstart)
            assert id_r.value == id_o.value
            n = out_r[5]
            nonempty += n > 100
            assert qs_r.value[:n] == qs_o.value[:n] and ts_r.value[:n] == ts_o.value[:n]
    assert nonempty > 40
    R.ref_xdrop_free(aligner)


def test_xdrop_kernel_body_matches_oracle():
    """The statements the GPU executes for the -x 1 extension (csrc/xdrop_core.cuh, run on the host through
    tests/xdrop_host_harness.cpp over packed 2-bit words) against the oracle: coordinates, columns, matches, strings;
    with and without columns; with the score row in the 128-entry ring (unrelated and low-complexity pairs outgrow it and
    are redone with the global row) and with the global row only."""
    O, H = util.oracle(), util.xdrop_harness()
    rng = np.random.default_rng(29)
    out_o = (C.c_int32 * 8)(); out_h = (C.c_int32 * 8)(); out_n = (C.c_int32 * 8)()
    id_o = C.c_double()
    cap = 60000
    qs_o, ts_o = C.create_string_buffer(cap), C.create_string_buffer(cap)
    qs_h, ts_h = C.create_string_buffer(cap), C.create_string_buffer(cap)
    checked = 0
    for q, qstart, t, tstart in xdrop_cases(rng, 120):
        qb = np.concatenate([[0], q, [0]]).astype(np.int8)
        tb = np.concatenate([[0], t, [0]]).astype(np.int8)
        qp = qb.ctypes.data + 1
        tp = tb.ctypes.data + 1
        for min_aln in (1, 500):
            O.orc_xdrop_go(C.cast(qp, C.c_char_p), qstart, len(q), C.cast(tp, C.c_char_p), tstart, len(t), min_aln, out_o,
                           C.byref(id_o), qs_o, ts_o, cap)
            H.xh_go(qp, qstart, len(q), tp, tstart, len(t), min_aln, out_h, qs_h, ts_h, cap, 1)
            H.xh_go(qp, qstart, len(q), tp, tstart, len(t), min_aln, out_n, None, None, 0, 0)
            assert list(out_o[:7]) == list(out_h[:7]) == list(out_n[:7]), (len(q), len(t), qstart, tstart)
            H.xh_go(qp, qstart, len(q), tp, tstart, len(t), min_aln, out_n, None, None, 0, 2)      # score row in global memory only
            assert list(out_o[:7]) == list(out_n[:7])
            n = out_o[5]
            assert qs_o.value[:n] == qs_h.value[:n] and ts_o.value[:n] == ts_h.value[:n]
            checked += n > 100
    assert checked > 80


@needs_ref
def test_cns_alignment_against_reference():
    """C1-C2 (mecat2cns/dw.cpp GetAlignment) and C4 (normalize_gaps) on random related pairs."""
    R, O = util.ref(), util.oracle()
    rng = np.random.default_rng(17)
    drd = R.ref_cns_drd_new()
    cap = 120000
    out_r = (C.c_int32 * 8)(); out_o = (C.c_int32 * 8)()
    qa_r, sa_r = C.create_string_buffer(cap), C.create_string_buffer(cap)
    qa_o, sa_o = C.create_string_buffer(cap), C.create_string_buffer(cap)
    nq_r, nt_r = C.create_string_buffer(2 * cap), C.create_string_buffer(2 * cap)
    nq_o, nt_o = C.create_string_buffer(2 * cap), C.create_string_buffer(2 * cap)
    n_ok = 0
    for it in range(80):
        L = int(rng.integers(300, 9000))
        base = rng.integers(0, 4, size=L).astype(np.int8)
        err = [0.02, 0.08, 0.15, 0.22, 0.4][it % 5]
        q = mutate(rng, base, err)
        t = mutate(rng, base, err)
        if it % 9 == 0:
            t = rng.integers(0, 4, size=L).astype(np.int8)
        f = rng.random()
        qstart = int(f * len(q)); tstart = min(len(t), int(f * len(t)))
        if it % 11 == 0:
            qstart, tstart = 0, 0
        if it % 13 == 0:
            qstart, tstart = len(q), len(t)
        qb = np.concatenate([[0], q, [0]]).astype(np.int8)
        tb = np.concatenate([[0], t, [0]]).astype(np.int8)
        qp, tp = qb.ctypes.data + 1, tb.ctypes.data + 1
        for min_aln, e in ((1, 0.15), (2000, 0.15), (500, 0.20)):
            okr = R.ref_cns_get_alignment(drd, qp, qstart, len(q), tp, tstart, len(t), e, min_aln, out_r, qa_r, sa_r, cap)
            oko = O.orc_cns_get_alignment(C.cast(qp, C.c_char_p), qstart, len(q), C.cast(tp, C.c_char_p), tstart, len(t), e,
                                          min_aln, out_o, qa_o, sa_o, cap)
            assert bool(okr) == bool(oko), (it, min_aln, e)
            if okr:
                n_ok += 1
                assert list(out_r[:5]) == list(out_o[:5])
                assert qa_r.value == qa_o.value and sa_r.value == sa_o.value
                n = len(qa_r.value)
                a = R.ref_normalize_gaps(qa_r.value, sa_r.value, n, 1, nq_r, nt_r, 2 * cap)
                b = O.orc_normalize_gaps(qa_o.value, sa_o.value, n, 1, nq_o, nt_o, 2 * cap)
                assert a == b and nq_r.value == nq_o.value and nt_r.value == nt_o.value
    assert n_ok > 60
    R.ref_cns_drd_free(drd)
