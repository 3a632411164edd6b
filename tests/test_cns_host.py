"""CPU tests of the consensus half of mecat2cns (rows C3-C7), fed with GetAlignment results computed by the
oracle (pinned against the reference in test_oracle.py):

  * the oracle's restatement of C3-C7 (oracle/oracle_cns_consensus.cpp) must reproduce the corrected FASTA that
    the UNMODIFIED reference binary `mecat2cns -i 0` wrote for the same candidates (tests/golden) -- this pins it;
  * the product's consensus kernels -- stage sequence (csrc/cns_pipeline.h) and per-thread bodies
    (csrc/cns_core.cuh), compiled for the host by tests/cns_host_harness.cpp, where a launch is a loop -- must
    reproduce the same FASTA, all reads in ONE batch so that every arena offset and scan is exercised.
The GPU run of the same kernels is checked in test_gpu.py (test_cns_*)."""
import ctypes as C
import gzip
import io
import json
import os

import numpy as np
import pytest

import util
from util import PackedVolume

GOLD = json.load(open(os.path.join(util.GOLDEN, "golden.json")))


def gold_fasta(name, tag):
    with gzip.open(os.path.join(util.GOLDEN, "%s.%s.fa.gz" % (name, tag)), "rt") as f:
        lines = f.read().splitlines()
    return sorted(zip(lines[0::2], lines[1::2]))


def gold_can(name):
    import mecat_b200
    with gzip.open(os.path.join(util.GOLDEN, "%s.can.gz" % name), "rt") as f:
        return mecat_b200.read_can(io.StringIO(f.read()))


_aln_cache = {}


def _align_groups(vol, groups, min_aln, err):
    """The oracle's GetAlignment of every candidate of every group (candidates already in their final order)."""
    from mecat_b200.api import ALIGN_RESULT_DTYPE
    O = util.oracle()
    codes = {}

    def seq(rid, strand):
        k = (rid, strand)
        if k not in codes:
            codes[k] = np.concatenate([[0], vol.codes(rid, strand), [0]]).astype(np.int8)
        return codes[k]

    o5 = (C.c_int32 * 8)()
    first, results = [0], []
    qblob, sblob = bytearray(), bytearray()
    for grp in groups:
        res = np.zeros(len(grp), dtype=ALIGN_RESULT_DTYPE)
        t = seq(int(grp["sid"][0]), 0)
        cap = 2 * len(t) + 70000
        qa, sa = C.create_string_buffer(cap), C.create_string_buffer(cap)
        for k, e in enumerate(grp):
            q = seq(int(e["qid"]), int(e["qdir"]))
            qext = int(e["qsize"]) - 1 - int(e["qext"]) if e["qdir"] else int(e["qext"])
            ok = O.orc_cns_get_alignment(C.cast(q.ctypes.data + 1, C.c_char_p), qext, len(q) - 2,
                                         C.cast(t.ctypes.data + 1, C.c_char_p), int(e["sext"]), len(t) - 2, err, min_aln, o5, qa, sa, cap)
            if ok:
                res[k] = (1, o5[1], o5[2], o5[3], o5[4], len(qa.value), 0, 0, 0.0, len(qblob))
                qblob += qa.value + b"\0"
                sblob += sa.value + b"\0"
            else:
                res[k]["str_offset"] = -1
        results.append(res)
        first.append(first[-1] + len(grp))
    return (np.array(first, dtype=np.int32), np.concatenate(groups), np.concatenate(results), bytes(qblob) + b"\0", bytes(sblob) + b"\0")


def m4_groups(name, min_cov, min_size, ratio, cap, keep=None, batch_size=100000):
    """-i 1: the overlap file of the fixture -> per read the overlaps the reference works on, in its order (partition
    records in file order, std::sort by sid, the `cap` largest of a read by std::sort: oracle/orc_cns_m4_order)."""
    import mecat_b200
    O = util.oracle()
    with gzip.open(os.path.join(util.GOLDEN, "%s.m4.gz" % name), "rt") as f:
        parts = mecat_b200.m4_partitions(f, ratio, min_size, batch_size)
    assert batch_size < 100000 or list(parts) == [0]
    groups = []
    for _, ec in sorted(parts.items()):
        ec = np.ascontiguousarray(ec)
        O.orc_cns_m4_order(ec.ctypes.data_as(C.c_void_p), len(ec), cap)
        i = 0
        while i < len(ec):
            j = i + 1
            while j < len(ec) and ec["sid"][j] == ec["sid"][i]:
                j += 1
            if j - i >= min_cov and not ec["ssize"][i] < min_size * 0.95 and (keep is None or keep(int(ec["sid"][i]))):
                groups.append(np.ascontiguousarray(ec[i:min(j, i + cap)]))
            i = j
    return groups


def oracle_alignments(vol, can, min_aln, min_cov, min_size, keep=None, err=0.15, groups=None):
    """Per read to correct: candidates in trial order + the oracle's GetAlignment of each.  Returns
    (first[R+1], candidates[T], results[T], qblob, sblob).  keep: optional predicate on the read id."""
    key = (id(vol), min_aln, min_cov, min_size, keep, err, None if groups is None else id(groups))
    if key in _aln_cache:
        return _aln_cache[key]
    if groups is not None:
        return _aln_cache.setdefault(key, _align_groups(vol, groups, min_aln, err))
    import mecat_b200
    from mecat_b200.api import ALIGN_RESULT_DTYPE
    O = util.oracle()
    ec = mecat_b200.normalise_candidates(can, min_size)
    ec = ec[np.argsort(ec["sid"], kind="stable")]
    codes = {}

    def seq(rid, strand):
        k = (rid, strand)
        if k not in codes:
            codes[k] = np.concatenate([[0], vol.codes(rid, strand), [0]]).astype(np.int8)
        return codes[k]

    o5 = (C.c_int32 * 8)()
    first, cands, results = [0], [], []
    qblob, sblob = bytearray(), bytearray()
    i = 0
    while i < len(ec):
        j = i + 1
        while j < len(ec) and ec["sid"][j] == ec["sid"][i]:
            j += 1
        grp = np.ascontiguousarray(ec[i:j])
        i = j
        if keep is not None and not keep(int(grp["sid"][0])):
            continue
        if len(grp) < min_cov or grp["ssize"][0] < min_size * 0.95:
            continue
        O.orc_cns_sort_candidates(grp.ctypes.data_as(C.c_void_p), len(grp))
        grp = grp[:200]
        res = np.zeros(len(grp), dtype=ALIGN_RESULT_DTYPE)
        t = seq(int(grp["sid"][0]), 0)
        cap = 2 * len(t) + 70000
        qa, sa = C.create_string_buffer(cap), C.create_string_buffer(cap)
        for k, e in enumerate(grp):
            q = seq(int(e["qid"]), int(e["qdir"]))
            qext = int(e["qsize"]) - 1 - int(e["qext"]) if e["qdir"] else int(e["qext"])
            ok = O.orc_cns_get_alignment(C.cast(q.ctypes.data + 1, C.c_char_p), qext, len(q) - 2,
                                         C.cast(t.ctypes.data + 1, C.c_char_p), int(e["sext"]), len(t) - 2, err, min_aln, o5, qa, sa, cap)
            if ok:
                res[k] = (1, o5[1], o5[2], o5[3], o5[4], len(qa.value), 0, 0, 0.0, len(qblob))
                qblob += qa.value + b"\0"
                sblob += sa.value + b"\0"
            else:
                res[k]["str_offset"] = -1
        cands.append(grp)
        results.append(res)
        first.append(first[-1] + len(grp))
    out = (np.array(first, dtype=np.int32), np.concatenate(cands), np.concatenate(results), bytes(qblob) + b"\0", bytes(sblob) + b"\0")
    _aln_cache[key] = out
    return out


def _pieces(free, pieces, n, seqs, nb):
    from mecat_b200.api import CNS_PIECE_DTYPE
    pc = np.frombuffer(C.string_at(pieces.value, n.value * CNS_PIECE_DTYPE.itemsize), dtype=CNS_PIECE_DTYPE)
    blob = C.string_at(seqs.value, nb.value)
    out = []
    for x in pc:
        s = blob[int(x["seq_offset"]):int(x["seq_offset"]) + int(x["seq_len"])].decode()
        out.append((">%d_%d_%d_%d" % (x["id"], x["beg"], x["end"], len(s)), s))
    free(pieces)
    free(seqs)
    return out


def correct_with_oracle(vol, can, ratio, min_aln, min_cov, min_size, keep=None, tech=0, groups=None):
    from mecat_b200.api import CnsParams
    O = util.oracle()
    first, cand, res, qblob, sblob = oracle_alignments(vol, can, min_aln, min_cov, min_size, keep, 0.20 if tech else 0.15, groups)
    p = CnsParams(ratio, min_aln, min_cov, min_size, tech, 0 if groups is None else 1)
    out = []
    for r in range(len(first) - 1):
        a, b = int(first[r]), int(first[r + 1])
        c, x = np.ascontiguousarray(cand[a:b]), np.ascontiguousarray(res[a:b])
        pieces, n, seqs, nb = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        rc = O.orc_cns_consensus(c.ctypes.data_as(C.c_void_p), b - a, x.ctypes.data_as(C.c_void_p), qblob, sblob, C.byref(p),
                                 C.byref(pieces), C.byref(n), C.byref(seqs), C.byref(nb))
        assert rc == 0
        out += _pieces(O.orc_free, pieces, n, seqs, nb)
    return sorted(out)


def correct_with_kernel_bodies(vol, can, ratio, min_aln, min_cov, min_size, keep=None, tech=0, groups=None):
    from mecat_b200.api import CnsParams
    H = util.cns_harness()
    first, cand, res, qblob, sblob = oracle_alignments(vol, can, min_aln, min_cov, min_size, keep, 0.20 if tech else 0.15, groups)
    p = CnsParams(ratio, min_aln, min_cov, min_size, tech, 0 if groups is None else 1)
    pieces, n, seqs, nb = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
    err = C.create_string_buffer(256)
    rc = H.harness_cns_batch(len(first) - 1, first.ctypes.data_as(C.c_void_p), cand.ctypes.data_as(C.c_void_p),
                             res.ctypes.data_as(C.c_void_p), qblob, sblob, C.cast(C.byref(p), C.c_void_p), C.byref(pieces), C.byref(n),
                             C.byref(seqs), C.byref(nb), err, 256)
    assert rc == 0, err.value
    return sorted(_pieces(H.harness_free, pieces, n, seqs, nb))


@pytest.fixture(scope="module")
def small_vol():
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f:
        seqs = [l for l in f.read().split(b"\n") if l and not l.startswith(b">")]
    return PackedVolume.from_seqs(seqs)


def compare(got, want):
    if got == want:
        return
    gh, wh = dict(got), dict(want)
    only_g = sorted(set(gh) - set(wh)); only_w = sorted(set(wh) - set(gh))
    diff = [h for h in gh if h in wh and gh[h] != wh[h]]
    pytest.fail("corrected reads differ: %d vs %d records; only mine %s; only reference %s; same header other sequence %s"
                % (len(got), len(want), only_g[:5], only_w[:5], diff[:5]))


@pytest.mark.parametrize("index_bytes", ["2", "4"])
def test_kernel_bodies_with_wider_graph_indices(small_vol, monkeypatch, index_bytes):
    """The region graphs are templated on their index type: 8-bit for the common tiny graph, 16- and 32-bit for larger
    ones.  Forcing the wider types on every graph must give the same corrected reads."""
    monkeypatch.setenv("MECAT_HARNESS_GRAPH_INDEX_BYTES", index_bytes)
    compare(correct_with_kernel_bodies(small_vol, gold_can("small"), 0.9, 1000, 4, 2000), gold_fasta("small", "cns_relaxed"))


PARAMS = {"cns_default": (0.9, 2000, 6, 5000),
          # -l 2000 -c 4 -a 1000: many more reads qualify (148 corrected pieces, thousands of mini-POA regions)
          "cns_relaxed": (0.9, 1000, 4, 2000)}


@pytest.mark.parametrize("tag", sorted(PARAMS))
def test_oracle_consensus_matches_reference(small_vol, tag):
    compare(correct_with_oracle(small_vol, gold_can("small"), *PARAMS[tag]), gold_fasta("small", tag))


@pytest.mark.parametrize("tag", sorted(PARAMS))
def test_kernel_bodies_match_reference(small_vol, tag):
    compare(correct_with_kernel_bodies(small_vol, gold_can("small"), *PARAMS[tag]), gold_fasta("small", tag))


# ---------------------------------------------------------------- deep coverage (~120x): 60-alignment cap, 20x coverage gate
def _every_sixth(rid):
    return rid % 6 == 0


@pytest.fixture(scope="module")
def deep_vol(tmp_path_factory):
    import hashlib
    c = GOLD["deep"]
    fa = str(tmp_path_factory.mktemp("deep") / "deep.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == c["fasta_sha256"]
    return PackedVolume.from_seqs(util.read_fasta(fa))


def _deep_gold():
    return [(h, s) for h, s in gold_fasta("deep", "cns_relaxed") if _every_sixth(int(h[1:].split("_")[0]))]


def test_oracle_consensus_deep_coverage(deep_vol):
    want = _deep_gold()
    assert len(want) >= 40
    compare(correct_with_oracle(deep_vol, gold_can("deep"), *PARAMS["cns_relaxed"], keep=_every_sixth), want)


def test_kernel_bodies_deep_coverage(deep_vol):
    compare(correct_with_kernel_bodies(deep_vol, gold_can("deep"), *PARAMS["cns_relaxed"], keep=_every_sixth), _deep_gold())


# ---------------------------------------------------------------- nanopore consensus (-x 1)
NANOPORE = (0.4, 400, 6, 2000)      # the -x 1 defaults: -r 0.4 -a 400 -c 6 -l 2000 (options.cpp:21-29)


def gold_can_x1():
    import mecat_b200
    with gzip.open(os.path.join(util.GOLDEN, "small.x1.can.gz"), "rt") as f:
        return mecat_b200.read_can(io.StringIO(f.read()))


def test_nanopore_consensus_matches_reference(small_vol):
    """consensus_one_read_can_nanopore (mecat_correction.cpp:453-512): error rate 0.20, up to 100 alignments, the whole
    read as the one effective range -- oracle and kernel bodies against `mecat2cns -x 1 -i 0` of the unmodified binary."""
    want = gold_fasta("small.x1", "cns")
    assert len(want) == GOLD["x1"]["small_num_cns"]
    compare(correct_with_oracle(small_vol, gold_can_x1(), *NANOPORE, tech=1), want)
    compare(correct_with_kernel_bodies(small_vol, gold_can_x1(), *NANOPORE, tech=1), want)


def test_nanopore_consensus_deep_coverage(deep_vol):
    """~120x: more than 60 alignments are accepted per read (the nanopore cap is 100) before the 20x gate closes."""
    want = [(h, s) for h, s in gold_fasta("deep.x1", "cns") if _every_sixth(int(h[1:].split("_")[0]))]
    assert len(want) >= 40
    compare(correct_with_oracle(deep_vol, gold_can("deep"), *NANOPORE, keep=_every_sixth, tech=1), want)
    compare(correct_with_kernel_bodies(deep_vol, gold_can("deep"), *NANOPORE, keep=_every_sixth, tech=1), want)


# ---------------------------------------------------------------- M4 input (-i 1)
def test_m4_input_consensus_matches_reference(small_vol, deep_vol):
    """consensus_one_read_m4_pacbio (mecat_correction.cpp:242-300): the overlaps of `mecat2pw -j 1 -g 1` as input; a
    partition is ordered by std::sort on sid, a read with more than 60 overlaps keeps the 60 largest (std::sort again), every
    alignment that succeeds is used.  Goldens: the unmodified `mecat2cns -i 1` with one OpenMP thread (its parallel-mode sort
    is then the sequential introsort; tests/golden/make_golden.py i1)."""
    ratio, min_aln, min_cov, min_size = PARAMS["cns_relaxed"]
    g = m4_groups("small", min_cov, min_size, ratio, 60)
    want = gold_fasta("small.i1", "cns")
    assert len(want) == GOLD["i1"]["small_num_cns"]
    compare(correct_with_oracle(small_vol, None, ratio, min_aln, min_cov, min_size, groups=g), want)
    compare(correct_with_kernel_bodies(small_vol, None, ratio, min_aln, min_cov, min_size, groups=g), want)
    # partitions of 100 reads (-p 100): each partition is ordered on its own, equal keys fall differently than above
    g = m4_groups("small", min_cov, min_size, ratio, 60, batch_size=100)
    want_p = gold_fasta("small.i1p100", "cns")
    assert want_p != want
    compare(correct_with_kernel_bodies(small_vol, None, ratio, min_aln, min_cov, min_size, groups=g), want_p)
    g = m4_groups("deep", min_cov, min_size, ratio, 60, keep=_every_sixth)
    assert max(len(x) for x in g) == 60
    want = [(h, s) for h, s in gold_fasta("deep.i1", "cns") if _every_sixth(int(h[1:].split("_")[0]))]
    assert len(want) >= 40
    compare(correct_with_oracle(deep_vol, None, ratio, min_aln, min_cov, min_size, keep=_every_sixth, groups=g), want)
    compare(correct_with_kernel_bodies(deep_vol, None, ratio, min_aln, min_cov, min_size, keep=_every_sixth, groups=g), want)


def test_nanopore_m4_input_consensus_matches_reference(small_vol):
    """consensus_one_read_m4_nanopore (mecat_correction.cpp:303-360): `mecat2cns -x 1` alone means `-i 1`; the -x 1 overlaps
    of the small fixture as input, up to 100 overlaps per read, every alignment that also passes the mapping-ratio test.
    Golden: the unmodified `mecat2cns -x 1` with one OpenMP thread (tests/golden/make_golden.py i1)."""
    ratio, min_aln, min_cov, min_size = NANOPORE
    g = m4_groups("small.x1", min_cov, min_size, ratio, 100)
    want = gold_fasta("small.x1i1", "cns")
    assert len(want) == GOLD["i1"]["small_x1_num_cns"]
    compare(correct_with_oracle(small_vol, None, ratio, min_aln, min_cov, min_size, tech=1, groups=g), want)
    compare(correct_with_kernel_bodies(small_vol, None, ratio, min_aln, min_cov, min_size, tech=1, groups=g), want)


def test_fused_normalise_vote_kernel_body_matches_literal_restatement():
    """normalize_vote_index (one streaming pass, what the GPU thread runs) against normalize_gaps + add_votes +
    column_index written literally after the reference, on random gapped alignments rich in homopolymers, long gap
    runs, mismatches and adjacent opposite gaps (the cases where pushed gaps travel and collide)."""
    H = util.cns_harness()
    rng = np.random.default_rng(5)
    for trial in range(400):
        n = int(rng.integers(1, 400))
        alphabet = b"ACGT"[:int(rng.integers(1, 5))]          # small alphabets make pushes travel far
        q, t = bytearray(), bytearray()
        pgap = rng.uniform(0.05, 0.5)
        for _ in range(n):
            r = rng.random()
            a = alphabet[int(rng.integers(len(alphabet)))]
            b = alphabet[int(rng.integers(len(alphabet)))] if rng.random() < 0.3 else a
            if r < pgap / 2:
                q.append(ord("-")); t.append(b)
            elif r < pgap:
                q.append(a); t.append(ord("-"))
            else:
                q.append(a); t.append(b)
        if trial % 50 == 0:                                  # the reference never produces these, the code must still agree
            k = int(rng.integers(n))
            q[k] = t[k] = ord("-")
        positions = sum(1 for c in t if c != ord("-")) + 2
        rc = H.harness_normalize_compare(bytes(q), bytes(t), n, 1, positions)
        assert rc == 0, (trial, rc, bytes(q), bytes(t))


def test_lane_parallel_anchor_walk_matches_sequential_walk():
    """anchor_chunk (32 positions per step: previous anchor, refine interval, ordinals from bit masks and popcounts)
    against the literal walk of meap_consensus_one_segment on random flag arrays of every density."""
    H = util.cns_harness()
    rng = np.random.default_rng(11)
    FMAT, FDEL, FINS, UNDS = 1, 2, 4, 8
    for trial in range(600):
        n = int(rng.integers(1, 300)) if trial % 3 else int(rng.integers(1, 70))
        p_anchor = rng.choice([0.0, 0.02, 0.3, 0.8, 0.97, 1.0])
        p_prob = rng.choice([0.0, 0.05, 0.3, 0.9])
        f = np.zeros(n, dtype=np.uint8)
        for i in range(n):
            if rng.random() < p_anchor:
                f[i] = FMAT | (FDEL if rng.random() < p_prob else 0)
            else:
                f[i] = (UNDS if rng.random() < p_prob else FINS) | (FDEL if rng.random() < p_prob / 2 else 0)
        rc = H.harness_anchor_compare(f.ctypes.data_as(C.c_void_p), n, int(rng.integers(0, 1000)))
        assert rc == 0, (trial, rc, f.tolist())


def test_lane_parallel_segment_search_matches_sequential_search():
    H = util.cns_harness()
    rng = np.random.default_rng(12)
    for trial in range(300):
        n = int(rng.integers(1, 600))
        votes = np.zeros(n + 2, dtype=np.uint32)
        level, i = 0, 0
        while i < n:                                            # piecewise-constant coverage with short dips
            run = int(rng.integers(1, 80))
            level = int(rng.integers(0, 9))
            mat = int(rng.integers(0, level + 1))
            votes[i:i + run] = mat | ((level - mat) << 8) | (int(rng.integers(0, 5)) << 16)
            i += run
        cuts = sorted(set(int(x) for x in rng.integers(0, n + 1, size=int(rng.integers(2, 8)))))
        ranges = np.array([[a, b] for a, b in zip(cuts[:-1], cuts[1:])][::2], dtype=np.int32).reshape(-1, 2)
        if len(ranges) == 0:
            continue
        rc = H.harness_segments_compare(votes.ctypes.data_as(C.c_void_p), ranges.ctypes.data_as(C.c_void_p), len(ranges),
                                        int(rng.integers(1, 8)), float(rng.choice([0.5, 1.0, 3.0, 20.0, 57.0])))
        assert rc == 0, (trial, rc)


def test_region_graph_matches_oracle_graph_on_random_pileups():
    """The flat-array graph of the kernels (intrusive adjacency lists, explicit merge stack, exact-size arena; 8-, 16- and
    32-bit indices) against the oracle's restatement of AlnGraphBoost on random pile-ups: up to 60 alignments over a short
    backbone, tiny alphabets (so that many sibling nodes merge, recursively), long insertion runs, double gaps, partial
    coverage.  Every run must finish without touching its capacity limits and give the oracle's consensus; where
    oracle/_ref is built, the oracle itself is compared with the UNMODIFIED reference class on the same pile-ups."""
    H, O = util.cns_harness(), util.oracle()
    R = util.ref() if util.have_ref() else None
    rng = np.random.default_rng(17)
    for trial in range(1500):
        blen = int(rng.integers(2, 16))
        alphabet = "ACGT"[:int(rng.integers(1, 5))]
        backbone = "".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=blen))
        naln = int(rng.integers(1, 61)) if trial % 4 else int(rng.integers(1, 6))
        p_ins, p_del = rng.uniform(0, 0.5), rng.uniform(0, 0.3)
        qs, ts, starts = [], [], []
        for _ in range(naln):
            pos = int(rng.integers(1, blen + 1))
            starts.append(pos)
            q, t = [], []
            last = int(rng.integers(pos, blen + 1))
            while pos <= last:
                r = rng.random()
                if r < p_ins:
                    q.append(alphabet[int(rng.integers(len(alphabet)))]); t.append("-")
                elif r < p_ins + p_del:
                    q.append("-"); t.append(backbone[pos - 1]); pos += 1
                elif r < p_ins + p_del + 0.02:
                    q.append("-"); t.append("-")
                else:
                    q.append(backbone[pos - 1]); t.append(backbone[pos - 1]); pos += 1
            if not q:
                q, t = [backbone[starts[-1] - 1]], [backbone[starts[-1] - 1]]
            qs.append("".join(q).encode()); ts.append("".join(t).encode())
        qa = (C.c_char_p * naln)(*qs); ta = (C.c_char_p * naln)(*ts); sa = (C.c_int * naln)(*starts)
        min_weight = int(rng.integers(0, max(2, naln // 2)))
        cap = 4096
        want = C.create_string_buffer(cap)
        nw = O.orc_poa_consensus(blen, naln, qa, ta, sa, min_weight, want, cap)
        assert nw >= 0
        if R is not None:                                       # the unmodified AlnGraphBoost (Boost.Graph) pins the oracle here
            ref = C.create_string_buffer(cap)
            nr = R.ref_poa_consensus(blen, naln, qa, ta, sa, min_weight, ref, cap)
            assert nr == nw and ref.raw[:nr] == want.raw[:nw], (trial, "oracle differs from the reference", backbone, qs, ts, starts, min_weight)
        for index_bytes in (1, 2, 4):
            nodes = blen + 2 + sum(1 for q, t in zip(qs, ts) for a, b in zip(q, t) if a != 45 and a != b)
            e0 = blen + 1 + sum(1 for q in qs for a in q if a != 45) + naln
            if index_bytes == 1 and max(nodes, e0 + nodes + 2) >= 120:
                continue                                        # the kernels use wider indices for such a graph
            got = C.create_string_buffer(cap)
            ng = H.harness_poa_consensus(blen, naln, qa, ta, sa, min_weight, index_bytes, got, cap)
            assert ng == nw and got.raw[:ng] == want.raw[:nw], (trial, index_bytes, ng, nw, backbone, qs, ts, starts, min_weight)


needs_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/libmecatref.so not built (needs /root/reference)")


@needs_ref
def test_effective_ranges_kernel_body_against_the_reference_function():
    """get_effective_ranges (mecat_correction.cpp:118-153) of the UNMODIFIED reference against the kernels' version on random
    mapping ranges: nested, chained, touching, duplicated, with and without a read-spanning one."""
    H, R = util.cns_harness(), util.ref()
    rng = np.random.default_rng(23)
    for trial in range(3000):
        read_size = int(rng.integers(2000, 40000))
        n = int(rng.integers(1, 61))
        starts = rng.integers(0, read_size - 1, size=n)
        if trial % 3 == 0:
            starts = (starts // 700) * 700                       # many equal starts
        ends = np.minimum(read_size, starts + rng.integers(1, read_size, size=n))
        if trial % 7 == 0:
            starts[0], ends[0] = int(rng.integers(0, 520)), read_size - int(rng.integers(0, 520))
        pairs = np.ascontiguousarray(np.stack([starts, ends], axis=1).astype(np.int32))
        min_size = int(rng.choice([1, 500, 2000, 5000, 12000]))
        a, b = np.zeros(2 * 64, dtype=np.int32), np.zeros(2 * 64, dtype=np.int32)
        na = R.ref_effective_ranges(pairs.ctypes.data_as(C.c_void_p), n, read_size, min_size, a.ctypes.data_as(C.c_void_p), 64)
        nb = H.harness_effective_ranges(pairs.ctypes.data_as(C.c_void_p), n, read_size, min_size, b.ctypes.data_as(C.c_void_p))
        assert na == nb and a[:2 * na].tolist() == b[:2 * nb].tolist(), (trial, pairs.tolist(), read_size, min_size)


@needs_ref
def test_one_pass_normalise_vote_against_the_reference_functions():
    """normalize_gaps + meap_add_one_aln of the UNMODIFIED reference against the kernels' single streaming pass: same
    normalised strings, same vote table (base, match, insert, delete counts per template position)."""
    H, R = util.cns_harness(), util.ref()
    rng = np.random.default_rng(29)
    for trial in range(600):
        n = int(rng.integers(1, 500))
        alphabet = b"ACGT"[:int(rng.integers(1, 5))]
        pgap = rng.uniform(0.05, 0.5)
        q, t = bytearray(), bytearray()
        for k in range(n):
            a = alphabet[int(rng.integers(len(alphabet)))]
            b = alphabet[int(rng.integers(len(alphabet)))] if rng.random() < 0.3 else a
            r = rng.random()
            if k == 0 or k == n - 1 or r >= pgap:                 # GetAlignment's output starts and ends on a column with both bases
                q.append(a); t.append(b if 0 < k < n - 1 else a)
            elif r < pgap / 2:
                q.append(ord("-")); t.append(b)
            else:
                q.append(a); t.append(ord("-"))
        positions = sum(1 for c in t if c != ord("-")) + 3
        cap = 2 * n + 16
        out_r, out_h = np.zeros(4 * positions, dtype=np.uint8), np.zeros(4 * positions, dtype=np.uint8)
        qr, tr, qh, th = (C.create_string_buffer(cap) for _ in range(4))
        lr = R.ref_normalize_and_vote(bytes(q), bytes(t), n, 1, positions, out_r.ctypes.data_as(C.c_void_p), qr, tr, cap)
        lh = H.harness_normalize_and_vote(bytes(q), bytes(t), n, 1, positions, out_h.ctypes.data_as(C.c_void_p), qh, th)
        assert lr == lh and qr.value == qh.value and tr.value == th.value, (trial, bytes(q), bytes(t))
        assert out_r.tolist() == out_h.tolist(), (trial, bytes(q), bytes(t))
