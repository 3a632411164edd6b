"""CPU tests of the host-side data formats (mecat_b200/csrc/host_io.cpp): the library's FASTA/FASTQ reader and 2-bit
packer must write the volume files the reference's split_raw_dataset writes, byte for byte.

  * against the committed sha256 of `volN` written by the UNMODIFIED reference for the golden read sets;
  * against the unmodified reference binary itself on awkward inputs (multi-line records, lower case, IUPAC codes that
    spill into neighbouring bases like PackedDB::set_char does, CRLF / lone CR line ends, FASTQ records, comments,
    `;` tails) -- only where oracle/_ref has been built (this container)."""
import gzip
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import util

GOLD = json.load(open(os.path.join(util.GOLDEN, "golden.json")))
REF_PW = os.path.join(util.REF_DIR, "mecat2pw")
needs_ref_binary = pytest.mark.skipif(not os.path.exists(REF_PW), reason="oracle/_ref/mecat2pw not built (needs /root/reference)")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def test_split_matches_reference_volume_of_the_golden_reads(tmp_path):
    import mecat_b200
    fa = str(tmp_path / "small.fa")
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f, open(fa, "wb") as g:
        g.write(f.read())
    vols = mecat_b200.split_dataset(fa, str(tmp_path / "wrk"))
    assert [os.path.basename(v) for v in vols] == ["vol0"]
    assert sha(vols[0]) == GOLD["small"]["vol0_sha256"]
    c = GOLD["cfg0"]
    fa = str(tmp_path / "cfg0.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
    vols = mecat_b200.split_dataset(fa, str(tmp_path / "wrk0"))
    assert sha(vols[0]) == c["vol0_sha256"]


def test_volume_from_fasta_equals_the_split_volume(tmp_path):
    """The in-memory loader mecat2cns uses gives exactly the records and bytes of vol0."""
    import mecat_b200
    fa = str(tmp_path / "small.fa")
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f, open(fa, "wb") as g:
        g.write(f.read())
    hv = mecat_b200.volume_from_fasta(fa)
    ref = mecat_b200.HostVolume.load(mecat_b200.split_dataset(fa, str(tmp_path / "wrk"))[0])
    assert hv.num_bases == ref.num_bases and hv.start_read_id == 0
    assert np.array_equal(hv.offset_size, ref.offset_size) and np.array_equal(hv.pac, ref.pac)
    with pytest.raises(mecat_b200.MecatB200Error):
        mecat_b200.volume_from_fasta(str(tmp_path / "missing.fa"))


def awkward_inputs():
    rng = np.random.default_rng(3)

    def seq(n, alphabet="ACGT"):
        return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=n))

    cases = {}
    cases["multiline_lowercase"] = "".join(">r%d some description\n%s\n" % (i, "\n".join(
        seq(int(rng.integers(1, 90)), "ACGTacgt") for _ in range(int(rng.integers(1, 6))))) for i in range(40))
    cases["iupac_codes"] = "".join(">r%d\n%s\n" % (i, seq(int(rng.integers(1, 200)), "ACGTNRYKMSWBDHVnryk-")) for i in range(60))
    cases["crlf"] = "".join(">r%d\r\n%s\r\n" % (i, seq(int(rng.integers(1, 150)))) for i in range(30))
    cases["lone_cr"] = "".join(">r%d\r%s\r" % (i, seq(int(rng.integers(1, 150)))) for i in range(30))
    cases["fastq"] = "".join("@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s)) for i, s in enumerate(seq(int(rng.integers(1, 120))) for _ in range(30)))
    # (no blank lines: the reference's BufferLineReader reports an EMPTY line as end of input once its last 16 MB buffer
    # is loaded, buffer_line_iterator.cpp:31,44 -- a quirk of its buffering that the library does not reproduce)
    cases["comments_semicolons"] = "# a comment\n" + "".join(
        ">r%d\n! note\n%s ;trailing words\n%s\n" % (i, seq(int(rng.integers(1, 80))), seq(int(rng.integers(1, 80)))) for i in range(25))
    cases["spaces_inside"] = "".join(">r%d\n%s %s\t%s\n" % (i, seq(9), seq(17), seq(30)) for i in range(20))
    cases["no_final_newline"] = ">a\nACGTACGTAC\n>b\nGGGTTTAAACCC"
    cases["one_base_reads"] = "".join(">r%d\n%s\n" % (i, "ACGT"[i % 4]) for i in range(37))
    cases["long_acgt"] = "".join(">r%d\n%s\n" % (i, seq(int(rng.integers(3000, 9000)))) for i in range(12))
    return cases


@needs_ref_binary
@pytest.mark.parametrize("name", sorted(awkward_inputs()))
def test_split_matches_reference_binary_on_awkward_inputs(tmp_path, name):
    import mecat_b200
    text = awkward_inputs()[name]
    fa = str(tmp_path / "in.fa")
    with open(fa, "w", newline="") as f:
        f.write(text)
    ref_wrk = str(tmp_path / "ref_wrk")
    p = subprocess.run([REF_PW, "-j", "0", "-d", fa, "-o", str(tmp_path / "ref.can"), "-w", ref_wrk, "-t", "1"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-500:]
    vols = mecat_b200.split_dataset(fa, str(tmp_path / "wrk"))
    ref_vols = sorted(f for f in os.listdir(ref_wrk) if f.startswith("vol"))
    assert [os.path.basename(v) for v in vols] == ref_vols
    for v in vols:
        assert open(v, "rb").read() == open(os.path.join(ref_wrk, os.path.basename(v)), "rb").read(), os.path.basename(v)


def test_split_into_several_volumes_round_trips(tmp_path):
    """A small volume cap (the reference's own debug knob) cuts the reads into several volumes; loading them back gives
    every read once, in order, with the right bases."""
    import mecat_b200
    rng = np.random.default_rng(9)
    reads = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=int(rng.integers(50, 400)))) for _ in range(200)]
    fa = str(tmp_path / "r.fa")
    with open(fa, "w") as f:
        for i, s in enumerate(reads):
            f.write(">%d\n%s\n" % (i, s))
    vols = mecat_b200.split_dataset(fa, str(tmp_path / "wrk"), max_volume_bases=5000)
    assert len(vols) > 3
    got, next_id = [], 0
    for v in vols:
        hv = mecat_b200.HostVolume.load(v)
        assert hv.start_read_id == next_id
        for off, size in hv.offset_size:
            codes = [(hv.pac[(off + i) >> 2] >> (2 * (3 - ((off + i) & 3)))) & 3 for i in range(size)]
            got.append("".join("ACGT"[c] for c in codes))
        next_id += len(hv.offset_size)
    assert got == reads


@pytest.mark.parametrize("text,msg", [
    ("ACGT\n>a\nACGT\n", "defline"),
    (">a\n>b\nACGT\n", "missing"),
    (">a\nAC!GT\n", "invalid residue"),
    ("@a\nACGT\n+\n", "quality"),
])
def test_split_reports_malformed_input_like_the_reference_reader(tmp_path, text, msg):
    import mecat_b200
    fa = str(tmp_path / "bad.fa")
    open(fa, "w").write(text)
    with pytest.raises(mecat_b200.MecatB200Error) as e:
        mecat_b200.split_dataset(fa, str(tmp_path / "wrk"))
    assert msg in str(e.value)


def _format_harness():
    import ctypes as C
    out_dir = os.path.join(util.ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libformat_harness.so")
    src = [os.path.join(util.ROOT, "tests", "format_harness.cpp"), os.path.join(util.ROOT, "mecat_b200", "csrc", "host", "format.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src[0]])
    L = C.CDLL(so)
    for f in (L.harness_format_m4, L.harness_ostream_m4):
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t]
    for f in (L.harness_format_can, L.harness_ostream_can):
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
    return L


def test_result_text_equals_the_reference_stream_formatting():
    """The drivers format .can / .m4 lines without iostreams; the characters must be the ones `operator<<` of the
    reference's records produces (integers of every sign and width, the identity as a default-formatted double)."""
    import ctypes as C
    L = _format_harness()
    rng = np.random.default_rng(21)
    n = 4000
    m4 = np.zeros(n, dtype=util.M4_DTYPE)
    for f in ("qid", "sid", "qoff", "qend", "qsize", "soff", "send", "ssize", "qext", "sext"):
        m4[f] = rng.integers(-5, 1 << 40, size=n) >> rng.integers(0, 40, size=n)
    m4["vscore"] = rng.integers(-2 ** 31, 2 ** 31 - 1, size=n)
    m4["qdir"] = rng.integers(0, 2, size=n); m4["sdir"] = rng.integers(0, 2, size=n)
    m4["ident"] = np.where(rng.random(n) < 0.9, rng.uniform(60, 100, size=n), rng.choice([0.0, 100.0, 7e-05, 1e+20, 99.99995, 123456.7, 85.5], size=n))
    m4["ident"][:200] = np.round(m4["ident"][:200], 2)
    m4[0] = tuple([-(2 ** 63)] * 2 + [75.0, -2 ** 31, 1] + [2 ** 63 - 1] * 3 + [0, 0] + [0] * 5)
    cap = 400 * n
    for gapped in (0, 1):
        a, b = C.create_string_buffer(cap), C.create_string_buffer(cap)
        la = L.harness_format_m4(m4.ctypes.data_as(C.c_void_p), n, gapped, a, cap)
        lb = L.harness_ostream_m4(m4.ctypes.data_as(C.c_void_p), n, gapped, b, cap)
        assert la == lb and la <= cap and a.raw[:la] == b.raw[:lb]
    ec = np.zeros(n, dtype=util.EC_DTYPE)
    for f in ec.dtype.names:
        ec[f] = rng.integers(-2 ** 31, 2 ** 31 - 1, size=n) >> rng.integers(0, 31, size=n)
    a, b = C.create_string_buffer(cap), C.create_string_buffer(cap)
    la = L.harness_format_can(ec.ctypes.data_as(C.c_void_p), n, a, cap)
    lb = L.harness_ostream_can(ec.ctypes.data_as(C.c_void_p), n, b, cap)
    assert la == lb and a.raw[:la] == b.raw[:lb]


@pytest.mark.parametrize("threads", [2, 5, 16])
def test_threaded_split_equals_the_sequential_split(tmp_path, threads, monkeypatch):
    """mecat_b200_split_dataset cuts plain FASTA at '>' lines and parses / packs it on several threads; every awkward input
    (and a multi-volume cap) must give the very bytes of the one-thread path -- FASTQ and malformed pieces fall back to it."""
    import mecat_b200
    cases = dict(awkward_inputs())
    rng = np.random.default_rng(17)
    cases["many_reads_small_cap"] = "".join(">%d\n%s\n" % (i, "".join("ACGT"[c] for c in rng.integers(0, 4, size=int(rng.integers(1, 700)))))
                                            for i in range(600))
    for name, text in sorted(cases.items()):
        fa = str(tmp_path / (name + ".fa"))
        with open(fa, "w", newline="") as f:
            f.write(text)
        cap = 9000 if name == "many_reads_small_cap" else 0
        monkeypatch.setenv("MECAT_B200_SPLIT_THREADS", "1")
        seq = mecat_b200.split_dataset(fa, str(tmp_path / (name + "_w1")), max_volume_bases=cap)
        monkeypatch.setenv("MECAT_B200_SPLIT_THREADS", str(threads))
        par = mecat_b200.split_dataset(fa, str(tmp_path / (name + "_wn")), max_volume_bases=cap)
        assert [os.path.basename(v) for v in par] == [os.path.basename(v) for v in seq], name
        for a, b in zip(seq, par):
            assert open(a, "rb").read() == open(b, "rb").read(), (name, os.path.basename(a))


@pytest.mark.parametrize("threads", [1, 4])
def test_volumes_from_fasta_equal_the_split_files(tmp_path, threads, monkeypatch):
    """The in-memory loader of mecat2cns for read sets beyond one volume gives the very volumes the splitter writes."""
    import mecat_b200
    rng = np.random.default_rng(23)
    fa = str(tmp_path / "r.fa")
    with open(fa, "w") as f:
        for i in range(300):
            f.write(">%d\n%s\n" % (i, "".join("ACGTN"[c] for c in rng.integers(0, 5 if i % 7 == 0 else 4, size=int(rng.integers(1, 900))))))
    monkeypatch.setenv("MECAT_B200_SPLIT_THREADS", str(threads))
    files = mecat_b200.split_dataset(fa, str(tmp_path / "wrk"), max_volume_bases=20000)
    vols = mecat_b200.volumes_from_fasta(fa, max_volume_bases=20000)
    assert len(vols) == len(files) > 4
    for path, v in zip(files, vols):
        w = mecat_b200.HostVolume.load(path)
        assert (v.num_reads, v.num_bases, v.start_read_id) == (w.num_reads, w.num_bases, w.start_read_id)
        assert np.array_equal(v.offset_size, w.offset_size) and v.pac.tobytes() == w.pac.tobytes()
    one = mecat_b200.volumes_from_fasta(fa)
    assert len(one) == 1 and one[0].pac.tobytes() == mecat_b200.volume_from_fasta(fa).pac.tobytes()


def test_fixed3_equals_printf():
    """mecat2asmpw prints its score with fprintf("%.3f") of a float (mecat2asmpw.c:944); the driver's integer-arithmetic
    form (format.h fixed3) must give the same characters: random values, the scores both programs can print, exact
    ties at the fourth decimal (rounded half to even on the exact value), denormals, large values, zeros of both signs."""
    import ctypes as C
    L = _format_harness()
    L.fh_fixed3.restype = C.c_int
    L.fh_fixed3.argtypes = [C.c_float, C.c_char_p, C.c_int]
    rng = np.random.default_rng(9)
    vals = [rng.random(20000, dtype=np.float32) * np.float32(300.0), rng.random(5000, dtype=np.float32) * np.float32(0.3),
            (np.arange(0, 4000, dtype=np.float32) + np.float32(0.5)) / np.float32(1000.0),          # near ties
            np.arange(1, 2049, dtype=np.float32) / np.float32(16.0) / np.float32(1000.0) * np.float32(8.0),
            np.array([0.0, -0.0, 0.0005, 0.0015, 0.0025, 0.0625, 0.1875, 0.3125, 1e-45, 1e-38, 1e-10, 65535.9995, 32768.0, 1e7, 3e38, -2.5, -0.0004],
                     dtype=np.float32)]
    n = np.arange(500, 3000, 7, dtype=np.int64)
    for m in (0, 1, 13, 57, 211):
        js = (2 * n - m).astype(np.float32)
        vals.append(js * np.float32(30) * np.float32(4) / n.astype(np.float32))                     # mecat2asmpw.c:942-943
        vals.append(np.full(len(n), m, dtype=np.float32) / (4 * n).astype(np.float32))             # mecat2trimpw.c:942-943
    buf = C.create_string_buffer(80)
    bad = []
    for arr in vals:
        for v in arr:
            L.fh_fixed3(float(v), buf, 80)
            want = "%.3f" % float(v)
            if buf.value.decode() != want:
                bad.append((float(v), buf.value.decode(), want))
    assert not bad, bad[:5]
