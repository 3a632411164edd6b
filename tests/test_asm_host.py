"""CPU tests of mecat2asmpw / mecat2trimpw (mecat2canu/src/mecat2asmpw/*.c): the oracle
restatement (oracle/oracle_asmpw.cpp) against goldens of the unmodified binaries, and the product's stage sequence and
kernel bodies (mecat_b200/csrc/asm_pipeline.h, asm_core.cuh) run on the host against both."""
import gzip
import json
import os

import numpy as np
import pytest

import util

GOLD = json.load(open(os.path.join(util.GOLDEN, "golden.json")))["asm"]


def gold(name):
    with gzip.open(os.path.join(util.GOLDEN, name + ".r.gz"), "rt") as f:
        return f.read().splitlines()


@pytest.fixture(scope="module")
def asm_files(tmp_path_factory):
    return util.asm_workdir("asm", str(tmp_path_factory.mktemp("asm")))


@pytest.fixture(scope="module")
def deep_files(tmp_path_factory):
    return util.asm_workdir("asmdeep", str(tmp_path_factory.mktemp("asmdeep")))


def all_pairs(fn, files, first_file=0, **kw):
    """-S<first_file+1> -E<n>: the index of file first_file, the reads of that file and of the following ones (main :1112-1150)."""
    sfirst, sub = files[first_file]
    lines = []
    for qfirst, qry in files[first_file:]:
        lines += util.asm_lines(fn(sub, sfirst, qry, qfirst, **kw))
    return sorted(lines)


def test_oracle_matches_the_unmodified_binaries(asm_files, deep_files):
    """Two files, -S1 -E2 and -S2 -E2, both programs; on this fixture the binary's output does not depend on what earlier
    reads left in the block array, so both conventions of the oracle reproduce it."""
    for history in (0, 1):
        assert all_pairs(util.asm_oracle_overlaps, asm_files, history=history) == gold("asm.asmpw")
    assert len(gold("asm.asmpw")) == GOLD["num_asm_asmpw"]
    assert all_pairs(util.asm_oracle_overlaps, asm_files, first_file=1) == gold("asm.asmpw.s2")
    assert all_pairs(util.asm_oracle_overlaps, asm_files, variant=1) == gold("asm.trimpw")
    # the deep file (MAXC = 50 cuts the candidate lists, block scores pass SM): with one thread's memory carried from read to
    # read the restatement is the binary; with zeroed blocks a handful of candidates score differently at the cut
    for variant, name in ((0, "asmdeep.asmpw50"), (1, "asmdeep.trimpw50")):
        want = gold(name)
        assert all_pairs(util.asm_oracle_overlaps, deep_files, variant=variant, maxc=50, history=1) == want
        got = all_pairs(util.asm_oracle_overlaps, deep_files, variant=variant, maxc=50, history=0)
        assert got != want
        assert len(set(got) ^ set(want)) <= 16 and abs(len(got) - len(want)) <= 4


def test_kernel_bodies_match_the_unmodified_binaries(asm_files):
    got = all_pairs(lambda *a, **k: util.asm_harness_overlaps(*a, **k)[0], asm_files)
    assert got == gold("asm.asmpw")
    got = all_pairs(lambda *a, **k: util.asm_harness_overlaps(*a, **k)[0], asm_files, first_file=1, variant=1, maxc=50)
    assert got == all_pairs(util.asm_oracle_overlaps, asm_files, first_file=1, variant=1, maxc=50)


def test_kernel_bodies_match_the_oracle_on_the_deep_file(deep_files):
    """More candidates than MAXC, block scores beyond SM (the neighbour votes read past a block's 60 entries: seedno[],
    seednum, index, the next block), N letters, lower-case reads, a 300-letter and a 12-letter read."""
    sfirst, sub = deep_files[0]
    for variant, maxc in ((0, 50), (1, 100)):
        want = util.asm_oracle_overlaps(sub, sfirst, sub, sfirst, variant=variant, maxc=maxc)
        got, stats = util.asm_harness_overlaps(sub, sfirst, sub, sfirst, variant=variant, maxc=maxc)
        assert stats[0] >= 1 and stats[3] == 1          # a genome this small fills every block: the first table estimate runs out and the range is split
        assert util.asm_lines(got) == util.asm_lines(want)          # same records in the same order
    # tables cut into many batches, a record pool that runs out and splits its batch: same records
    got, stats = util.asm_harness_overlaps(sub, sfirst, sub, sfirst, variant=1, maxc=100, budget=200000, divisor=64)
    assert stats[0] > 400
    assert util.asm_lines(got) == util.asm_lines(want)


def test_integer_consistency_test_equals_the_float_forms():
    """ddf_close / ddf_close_d (asm_core.cuh) against |a / (b * 10.0f) - 1| < 0.10 with a float quotient (find_location,
    mecat2asmpw.c:340) and |a / (b * 10 * 1.0) - 1.0| < 0.10 with a double one (the neighbour votes, :702)."""
    H = util.asm_harness()
    a = np.arange(-2500, 25000, dtype=np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        for b in list(range(-40, 41)) + [97, 250, 999, 1500, 2000]:
            qf = a.astype(np.float32) / (np.float32(b) * np.float32(10.0))
            want_f = np.abs((qf - np.float32(1)).astype(np.float64)) < 0.10
            qd = a.astype(np.float64) / (b * 10 * 1.0)
            want_d = np.abs(qd - 1.0) < 0.10
            got_f = np.array([H.ah_ddf_close(int(x), b, 0) for x in a], dtype=bool)
            got_d = np.array([H.ah_ddf_close(int(x), b, 1) for x in a], dtype=bool)
            assert (got_f == want_f).all(), b
            assert (got_d == want_d).all(), b
    assert H.ah_ddf_close(90, 10, 1) == 1 and H.ah_ddf_close(90, 10, 0) == 0       # the quotient 0.9 is close in double only


def _run_driver(bindir, prog, wrk, first, last, threads=3, devices=1, part_reads=None):
    import subprocess
    for f in os.listdir(wrk):
        if f.endswith(".r"):
            os.remove(os.path.join(wrk, f))
    env = dict(os.environ, MECAT_GPUS=str(devices), MECAT_SHIM_DEVICES=str(devices))
    if part_reads:
        env["MECAT_B200_ASM_PART_READS"] = str(part_reads)
    p = subprocess.run([os.path.join(bindir, prog), "-P" + wrk, "-T%d" % threads, "-S%d" % first, "-E%d" % last], capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    names = sorted(f for f in os.listdir(wrk) if f.endswith(".r"))
    assert names == ["%d_%d.r" % (first, t) for t in range(threads)]          # what the pipeline's `cat <S>_*.r` expects
    lines = []
    for f in names:
        lines += open(os.path.join(wrk, f)).read().splitlines()
    return sorted(lines)


def test_command_line_driver_on_the_host(tmp_path):
    """mecat_b200/csrc/host/mecat2asmpw.cpp (linked against the host harness instead of the library): the four program
    names, -P -T -S -E, ovlprep and the block files in, <S>_<t>.r out -- the files of the unmodified binaries."""
    import subprocess
    bindir = util.asm_driver_on_host()
    wrk = str(tmp_path / "blocks")
    util.asm_workdir("asm", wrk)
    assert _run_driver(bindir, "mecat2asmpw", wrk, 1, 2) == gold("asm.asmpw")
    assert _run_driver(bindir, "mecat2asmpw", wrk, 2, 2, threads=1) == gold("asm.asmpw.s2")
    assert _run_driver(bindir, "mecat2trimpw", wrk, 1, 2) == gold("asm.trimpw")
    # three devices (MECAT_GPUS): each maps a third of every query file's reads against its own index; more devices than -T files
    assert _run_driver(bindir, "mecat2asmpw", wrk, 1, 2, threads=2, devices=3) == gold("asm.asmpw")
    # query files taken in parts (load_fastq's SVM / MAXSTR batches; here 70 reads at a time), each part split between two devices
    assert _run_driver(bindir, "mecat2asmpw", wrk, 1, 2, threads=1, devices=2, part_reads=70) == gold("asm.asmpw")
    p = subprocess.run([os.path.join(bindir, "mecat2asmpw"), "-P" + wrk, "-T2", "-S1", "-E1"], capture_output=True, text=True, env=dict(os.environ, MECAT_GPUS="2"))
    assert p.returncode != 0 and "devices" in p.stderr
    p = subprocess.run([os.path.join(bindir, "mecat2asmpw"), "-P" + wrk, "-T2", "-S1", "-E3"], capture_output=True, text=True)
    assert p.returncode != 0 and "ovlprep" in p.stderr
    p = subprocess.run([os.path.join(bindir, "mecat2asmpw"), "-P" + wrk, "-T2"], capture_output=True, text=True)
    assert p.returncode != 0 and "usage" in p.stderr
    # the *50 names keep 50 candidates per read: on the deep file that cuts the lists (the oracle with the same convention)
    wrk = str(tmp_path / "deep")
    files = util.asm_workdir("asmdeep", wrk)
    want = all_pairs(util.asm_oracle_overlaps, files, variant=1, maxc=50)
    assert _run_driver(bindir, "mecat2trimpw50", wrk, 1, 1) == want
    assert len(want) != len(all_pairs(util.asm_oracle_overlaps, files, variant=1, maxc=100))


def test_index_kernel_bodies_match_a_numpy_restatement(deep_files):
    """creat_ref_index (mecat2asmpw.c:397-497) through the product's count / scan / fill / sort bodies, on the deep file
    (N letters) with a poly-A read appended so that one list passes 256 entries and is dropped."""
    import ctypes as C
    sfirst, sub = deep_files[0]
    text, starts, lens = util.asm_text(list(sub) + ["A" * 400])
    kept, cnt, want_pos, dropped = util.asm_index_numpy(text)
    assert dropped
    H = util.asm_harness()
    begin, pos = np.zeros((1 << 26) + 1, dtype=np.uint32), np.zeros(len(want_pos) + 16, dtype=np.int32)
    n = H.ah_index(text, len(text), starts.ctypes.data, lens.ctypes.data, len(lens), begin.ctypes.data, pos.ctypes.data, len(pos))
    assert n == len(want_pos)
    assert (pos[:n] == want_pos).all()
    assert (np.diff(begin.astype(np.int64))[kept] == cnt).all() and int(begin[-1]) == n


def test_oracle_reproduces_both_schedules_of_the_binary(tmp_path):
    """1 500 reads at ~130x through mecat2asmpw50 (three chunks of PLL = 500 reads): the binary prints different overlaps
    with one thread than with a thread per chunk (what its vote loops find beyond a block's entries is what the thread mapped
    before).  The restatement with one thread's memory carried from read to read equals -T1, with fresh memory per chunk -T4
    (digests of the sorted lines, tests/golden/make_golden.py asm).  Zeroed blocks per strand -- the product's convention --
    differ from both in a fraction of a percent of the lines (tools/asm_reference_schedule.py)."""
    import hashlib
    assert GOLD["asmsched_T1_sha256"] != GOLD["asmsched_T4_sha256"] and 0 < GOLD["asmsched_T1_vs_T4_lines"] < 100
    first, reads = util.asm_workdir("asmsched", str(tmp_path / "sched"))[0]
    digest = {}
    for history in (1, 2):
        lines = sorted(util.asm_lines(util.asm_oracle_overlaps(reads, first, reads, first, variant=0, maxc=50, history=history)))
        assert len(lines) == GOLD["asmsched_lines"]
        digest[history] = hashlib.sha256("\n".join(lines).encode()).hexdigest()
    assert digest[1] == GOLD["asmsched_T1_sha256"]
    assert digest[2] == GOLD["asmsched_T4_sha256"]


def test_awkward_reads_match_the_unmodified_binaries(tmp_path):
    """A duplicated read, tandem repeats (every k-mer of the unit ~180 times), poly-A (its list is dropped), ACGT x 600, N every
    50 letters, a read of N only, other IUPAC letters, mixed case, reads of 1 / 13 / 14 letters, a contained read: the oracle in
    both conventions and the product's bodies print what the binaries print."""
    first, reads = util.asm_workdir("asmodd", str(tmp_path / "odd"))[0]
    for variant, maxc, name in ((0, 100, "asmodd.asmpw"), (1, 50, "asmodd.trimpw50")):
        want = gold(name)
        assert len(want) == GOLD["num_" + name.replace(".", "_")]
        for history in (0, 1):
            assert sorted(util.asm_lines(util.asm_oracle_overlaps(reads, first, reads, first, variant=variant, maxc=maxc, history=history))) == want
        got, _ = util.asm_harness_overlaps(reads, first, reads, first, variant=variant, maxc=maxc)
        assert sorted(util.asm_lines(got)) == want
