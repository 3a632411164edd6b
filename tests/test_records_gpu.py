"""GPU parity tests (-m gpu) of the record assembly (row A12) and result text on the device:
fill_m4record / std::sort / containment filter in kernels (records.cu) against the golden `.m4` of the unmodified binary
and against the earlier host-thread form, the lines written on the device against the golden files byte for byte."""
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


@pytest.fixture(scope="module")
def deep_vol(tmp_path_factory):
    c = GOLD["deep"]
    fa = str(tmp_path_factory.mktemp("deep") / "deep.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
    return PackedVolume.from_seqs(util.read_fasta(fa))


def test_device_assembly_equals_host_assembly_record_for_record(gpu_ctx, cfg0_vol, deep_vol, monkeypatch):
    """Same records in the same order from the kernels and from the host threads (std::sort, containment filter); the
    ~120x fixture gives every read 100 candidates, many of them of the same pair (ties, partitions beyond 16 records)."""
    for vol in (cfg0_vol, deep_vol):
        hv = host_volume(vol)
        monkeypatch.setenv("MECAT_B200_M4", "host")
        a = gpu_ctx.pw_overlaps(hv, hv)
        monkeypatch.setenv("MECAT_B200_M4", "device")
        b = gpu_ctx.pw_overlaps(hv, hv)
        assert len(a) == len(b) > 1000
        assert a.tobytes() == b.tobytes()


def test_tile_text_is_the_reference_file(gpu_ctx, small_vol, cfg0_vol):
    """mecat_b200_pw_tile_text: the lines of `.can` and `.m4 -g 1` written on the device, against the unmodified binary's
    files (sorted: the reference's threads interleave reads), and against the host formatter line for line in order."""
    import mecat_b200
    from mecat_b200 import api
    for name, vol in (("small", small_vol), ("cfg0", cfg0_vol)):
        hv = host_volume(vol)
        d = gpu_ctx.upload(hv)
        idx = gpu_ctx.index_build(d)
        try:
            for task, ext, gapped in ((0, "can", False), (1, "m4", True), (1, None, False)):
                p = mecat_b200.pw_params(task=task)
                text, n = gpu_ctx.pw_tile_text(idx, d, d, p, gapped=gapped)
                rec = gpu_ctx.pw_tile(idx, d, d, p)
                assert n == len(rec)
                lines = text.decode().splitlines()
                want = util.ec_lines(rec) if task == 0 else util.m4_lines(rec, gapped=gapped)
                assert sorted(lines) == want                     # util's formatter (printf %g) on the records
                if ext:
                    assert sorted(lines) == gold_lines(name, ext)
                assert text.endswith(b"\n") and text.count(b"\n") == n
                # record order is kept: line i is record i
                first = lines[0].split("\t")
                assert int(first[0]) == int(rec[0]["qid"]) and int(first[1]) == int(rec[0]["sid"])
                assert gpu_ctx.records_text(rec, gapped=gapped) == text
        finally:
            gpu_ctx.release_index(idx)
            gpu_ctx.release_volume(d)


def test_records_text_of_extreme_values(gpu_ctx):
    """Device formatter on field values the fixtures do not reach (negative scores, 40-bit ids, identities across the
    whole printable range) against printf."""
    import mecat_b200
    rng = np.random.default_rng(5)
    m4 = np.zeros(20000, dtype=mecat_b200.M4_DTYPE)
    for name in m4.dtype.names:
        if name == "ident":
            m4[name] = 100.0 * rng.integers(0, 30000, len(m4)) / rng.integers(30000, 60000, len(m4))
        elif name not in ("pad", "pad_"):
            hi = 2 ** 31 - 1 if m4.dtype[name].itemsize == 4 else 2 ** 40
            m4[name] = rng.integers(0, hi, len(m4))
    m4["vscore"][:10] = -5
    m4["ident"][:8] = [0.0, 100.0, 99.99995, 9.999995, 1e-4, 3.0517578125e-05, 12.5, 99.999949999]
    got = gpu_ctx.records_text(m4, gapped=True).decode().splitlines()
    assert sorted(got) == util.m4_lines(m4, gapped=True)
    assert got[0].split("\t")[0] == str(int(m4[0]["qid"])) and got[-1].split("\t")[1] == str(int(m4[-1]["sid"]))
    assert gpu_ctx.records_text(m4[:0]) == b""


def test_device_packing_equals_host_packing(gpu_ctx, small_vol, tmp_path):
    """mecat_b200_volume_from_text (2-bit packing on the device) against the product's host packer, which is pinned to the
    unmodified split_raw_dataset on the same kind of input (tests/test_host_io.py): reads with N runs, lower case, IUPAC
    letters, '-' and characters outside the table (codes above 3 spill inside their byte like PackedDB::set_char)."""
    import mecat_b200
    rng = np.random.default_rng(12)
    seqs = []
    for i in range(small_vol.num_reads):
        s = bytearray(b"ACGT"[c] for c in small_vol.codes(i))
        if i % 5 == 1:
            for _ in range(6):
                p = int(rng.integers(0, max(1, len(s) - 40))); L = min(int(rng.integers(1, 30)), len(s) - p); s[p:p + L] = b"N" * L
            for _ in range(20):
                p = int(rng.integers(0, len(s))); s[p] = b"NnRYKMSWBDHVacgt-"[int(rng.integers(0, 17))]
        if i % 7 == 2:
            s = bytearray(bytes(s).lower())
        if i % 11 == 3:
            s = s[:int(rng.integers(1, 9))]             # tiny reads: several reads inside one packed word
        seqs.append(bytes(s))
    fa = str(tmp_path / "awkward.fa")
    with open(fa, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">%d\n%s\n" % (i, s))
    hv = mecat_b200.volume_from_fasta(fa)
    text = open(fa, "rb").read()
    src, osz, at, curr = [], [], 0, 0
    for i, s in enumerate(seqs):
        at = text.index(b"\n", at) + 1          # past the header line
        src.append(at); osz += [curr, len(s)]
        at += len(s) + 1; curr += len(s) + 1
    assert curr == hv.num_bases and np.array_equal(np.array(osz, dtype=np.int32).reshape(-1, 2), np.asarray(hv.offset_size).reshape(-1, 2))
    d, pac = gpu_ctx.volume_from_text(text, src, osz, curr)
    try:
        want = np.frombuffer(bytes(hv.pac[:(curr + 3) // 4]), dtype=np.uint8)
        bad = np.nonzero(pac != want)[0]
        assert len(bad) == 0, (len(bad), bad[:5], pac[bad[:5]], want[bad[:5]])
        # the resident volume is the one an upload of the host volume gives: same index, same candidates
        d2 = gpu_ctx.upload(hv)
        i1, i2 = gpu_ctx.index_build(d), gpu_ctx.index_build(d2)
        b1, p1 = gpu_ctx.index_export(i1); b2, p2 = gpu_ctx.index_export(i2)
        assert (b1 == b2).all() and (p1 == p2).all()
        p = mecat_b200.pw_params(task=0)
        assert gpu_ctx.pw_tile(i1, d, d, p).tobytes() == gpu_ctx.pw_tile(i2, d2, d2, p).tobytes()
        gpu_ctx.release_index(i1); gpu_ctx.release_index(i2); gpu_ctx.release_volume(d2)
    finally:
        gpu_ctx.release_volume(d)


# ---------------------------------------------------------------- mecat2cns -i 1 (M4 input)
def _gold_fasta(name, tag):
    with gzip.open(os.path.join(util.GOLDEN, "%s.%s.fa.gz" % (name, tag)), "rt") as f:
        lines = f.read().splitlines()
    return sorted(zip(lines[0::2], lines[1::2]))


def test_m4_input_consensus_matches_reference(gpu_ctx, small_vol, deep_vol, tmp_path):
    """mecat2cns -i 1: the library orders a partition like the reference (std::sort by sid, the 60 largest overlaps of a
    read by std::sort) and uses every alignment that succeeds (consensus_one_read_m4_pacbio, mecat_correction.cpp:242-300);
    goldens of the unmodified binary (one OpenMP thread) on the sorted overlap files of the small and the ~120x fixture,
    through the C ABI and through the command line."""
    import subprocess
    import mecat_b200
    for name, vol in (("small", small_vol), ("deep", deep_vol)):
        with gzip.open(os.path.join(util.GOLDEN, "%s.m4.gz" % name), "rt") as f:
            parts = mecat_b200.m4_partitions(f, 0.9, 2000)
        assert list(parts) == [0]
        d = gpu_ctx.upload(host_volume(vol))
        pieces = gpu_ctx.cns_reads(d, parts[0], 0.9, 1000, 4, 2000, input_type=1)
        gpu_ctx.release_volume(d)
        got = sorted((">%d_%d_%d_%d" % (i, b, e, len(s)), s.decode()) for i, b, e, s in pieces)
        want = _gold_fasta(name + ".i1", "cns")
        assert len(got) == len(want) == GOLD["i1"][name + "_num_cns"]
        bad = [g[0] for g, w in zip(got, want) if g != w]
        assert not bad, bad[:5]
    reads, m4, out = str(tmp_path / "small.fa"), str(tmp_path / "small.m4"), str(tmp_path / "cns.fa")
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f, open(reads, "wb") as g:
        g.write(f.read())
    with gzip.open(os.path.join(util.GOLDEN, "small.m4.gz"), "rb") as f, open(m4, "wb") as g:
        g.write(f.read())
    p = subprocess.run([os.path.join(util.ROOT, "mecat_b200", "bin", "mecat2cns"), "-i", "1", "-t", "4", "-l", "2000", "-c", "4", "-a", "1000", m4, reads, out],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = open(out).read().splitlines()
    assert sorted(zip(lines[0::2], lines[1::2])) == _gold_fasta("small.i1", "cns")
    # partitions of 100 reads: three library calls, each partition ordered on its own like the reference's partition files
    p = subprocess.run([os.path.join(util.ROOT, "mecat_b200", "bin", "mecat2cns"), "-i", "1", "-t", "4", "-p", "100", "-l", "2000", "-c", "4", "-a", "1000", m4, reads, out],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = open(out).read().splitlines()
    assert sorted(zip(lines[0::2], lines[1::2])) == _gold_fasta("small.i1p100", "cns")
