"""CPU test of the host half of mecat2cns (rows C3-C7): the product's per-read consensus
(mecat_b200/csrc/cns.cpp, reached through its test hook) is fed GetAlignment results computed by the
oracle (pinned against the reference in test_oracle.py) and must reproduce the corrected FASTA that the
UNMODIFIED reference binary `mecat2cns -i 0` wrote for the same candidates (tests/golden)."""
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


def correct_with_oracle_alignments(vol, can, ratio, min_aln, min_cov, min_size):
    """mecat2cns -i 0 with the alignments taken from the oracle and everything else from the product."""
    import mecat_b200
    from mecat_b200.api import ALIGN_RESULT_DTYPE, CnsParams, CNS_PIECE_DTYPE, EC_DTYPE
    L = mecat_b200.load_library()
    O = util.oracle()
    ec = mecat_b200.normalise_candidates(can, min_size)
    ec = ec[np.argsort(ec["sid"], kind="stable")]
    p = CnsParams(ratio, min_aln, min_cov, min_size)
    out = []
    codes = {}

    def seq(rid, strand):
        k = (rid, strand)
        if k not in codes:
            codes[k] = np.concatenate([[0], vol.codes(rid, strand), [0]]).astype(np.int8)
        return codes[k]

    o5 = (C.c_int32 * 8)()
    i = 0
    while i < len(ec):
        j = i + 1
        while j < len(ec) and ec["sid"][j] == ec["sid"][i]:
            j += 1
        grp = np.ascontiguousarray(ec[i:j])
        i = j
        if len(grp) < min_cov or grp["ssize"][0] < min_size * 0.95:
            continue
        L.mecat_b200_cns_sort_candidates(grp.ctypes.data_as(C.c_void_p), len(grp))
        grp = grp[:200]
        res = np.zeros(len(grp), dtype=ALIGN_RESULT_DTYPE)
        qblob, sblob = bytearray(), bytearray()
        t = seq(int(grp["sid"][0]), 0)
        cap = 2 * len(t) + 70000
        qa, sa = C.create_string_buffer(cap), C.create_string_buffer(cap)
        for k, e in enumerate(grp):
            q = seq(int(e["qid"]), int(e["qdir"]))
            qext = int(e["qsize"]) - 1 - int(e["qext"]) if e["qdir"] else int(e["qext"])
            ok = O.orc_cns_get_alignment(C.cast(q.ctypes.data + 1, C.c_char_p), qext, len(q) - 2,
                                         C.cast(t.ctypes.data + 1, C.c_char_p), int(e["sext"]), len(t) - 2, 0.15, min_aln, o5, qa, sa, cap)
            if ok:
                res[k] = (1, o5[1], o5[2], o5[3], o5[4], len(qa.value), 0, 0, 0.0, len(qblob))
                qblob += qa.value + b"\0"
                sblob += sa.value + b"\0"
            else:
                res[k]["str_offset"] = -1
        pieces, n, seqs, nb = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        rc = L.mecat_b200_cns_consensus_host(grp.ctypes.data_as(C.c_void_p), len(grp), res.ctypes.data_as(C.c_void_p),
                                             bytes(qblob) + b"\0", bytes(sblob) + b"\0", C.byref(p), C.byref(pieces), C.byref(n),
                                             C.byref(seqs), C.byref(nb))
        assert rc == 0
        pc = np.frombuffer(C.string_at(pieces.value, n.value * CNS_PIECE_DTYPE.itemsize), dtype=CNS_PIECE_DTYPE)
        blob = C.string_at(seqs.value, nb.value)
        for x in pc:
            s = blob[int(x["seq_offset"]):int(x["seq_offset"]) + int(x["seq_len"])].decode()
            out.append((">%d_%d_%d_%d" % (x["id"], x["beg"], x["end"], len(s)), s))
        L.mecat_b200_host_free(pieces)
        L.mecat_b200_host_free(seqs)
    return sorted(out)


@pytest.fixture(scope="module")
def small_vol():
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f:
        seqs = [l for l in f.read().split(b"\n") if l and not l.startswith(b">")]
    return PackedVolume.from_seqs(seqs)


@pytest.fixture(scope="module", autouse=True)
def built():
    from mecat_b200 import build
    build.build()


def compare(got, want):
    if got == want:
        return
    gh, wh = dict(got), dict(want)
    only_g = sorted(set(gh) - set(wh)); only_w = sorted(set(wh) - set(gh))
    diff = [h for h in gh if h in wh and gh[h] != wh[h]]
    pytest.fail("corrected reads differ: %d vs %d records; only mine %s; only reference %s; same header other sequence %s"
                % (len(got), len(want), only_g[:5], only_w[:5], diff[:5]))


def test_consensus_default_parameters(small_vol):
    got = correct_with_oracle_alignments(small_vol, gold_can("small"), 0.9, 2000, 6, 5000)
    compare(got, gold_fasta("small", "cns_default"))


def test_consensus_relaxed_parameters(small_vol):
    # -l 2000 -c 4 -a 1000: many more reads qualify (148 corrected pieces, thousands of mini-POA regions)
    got = correct_with_oracle_alignments(small_vol, gold_can("small"), 0.9, 1000, 4, 2000)
    compare(got, gold_fasta("small", "cns_relaxed"))
