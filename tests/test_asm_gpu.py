"""GPU parity tests (-m gpu) of mecat2asmpw / mecat2trimpw: the CUDA path through the C ABI
(mecat_b200_asm_index_build / mecat_b200_asm_overlaps) and through the command line, against goldens of the unmodified
binaries (tests/golden/asm*.r.gz) and against the oracle where the binary's result depends on its thread history."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

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


def gpu_pairs(ctx, files, first_file=0, variant=0, maxc=100):
    import mecat_b200
    sfirst, sub = files[first_file]
    idx = ctx.asm_index_build(mecat_b200.AsmReads(sub, sfirst))
    lines = []
    for qfirst, qry in files[first_file:]:
        lines += mecat_b200.asm_lines(ctx.asm_overlaps(idx, mecat_b200.AsmReads(qry, qfirst), variant, maxc))
    ctx.asm_index_release(idx)
    return lines


def test_index_matches_a_numpy_restatement(gpu_ctx, deep_files):
    """creat_ref_index (mecat2asmpw.c:397-497): 13-mers in A0 T1 C2 G3, a letter other than ACGT (N, the NUL between reads)
    ends a k-mer, lists of more than 256 dropped, 1-based starts ascending inside a list."""
    import mecat_b200
    sfirst, sub = deep_files[0]
    reads = mecat_b200.AsmReads(sub, sfirst)
    idx = gpu_ctx.asm_index_build(reads)
    begin, pos = gpu_ctx.asm_index_export(idx)
    gpu_ctx.asm_index_release(idx)
    kept, cnt, want_pos, _ = util.asm_index_numpy(reads.text.upper())
    assert len(pos) == len(want_pos) and (pos == want_pos).all()
    assert (np.diff(begin.astype(np.int64))[kept] == cnt).all()
    assert int(begin[-1]) == len(want_pos)


def test_overlaps_match_the_unmodified_binaries(gpu_ctx, asm_files):
    """Two block files: the index of file 1 with the reads of files 1 and 2 (-S1 -E2), the index of file 2 with its own reads
    (-S2 -E2), both programs."""
    got = gpu_pairs(gpu_ctx, asm_files)
    print("[asm] stats", {k: v for k, v in gpu_ctx.stats()["kernel_ms"].items() if k.startswith("asm")}, "records", len(got))
    assert sorted(got) == gold("asm.asmpw") and len(got) == GOLD["num_asm_asmpw"]
    assert sorted(gpu_pairs(gpu_ctx, asm_files, first_file=1)) == gold("asm.asmpw.s2")
    assert sorted(gpu_pairs(gpu_ctx, asm_files, variant=1)) == gold("asm.trimpw")


def test_deep_file_matches_the_oracle(gpu_ctx, deep_files, monkeypatch):
    """More candidates than MAXC, block scores beyond SM, N letters, lower-case reads, stubs of 300 and 12 letters: the
    records and their order are the oracle's (zeroed blocks per strand; the binary's own result depends on what its thread
    mapped before, tests/test_asm_host.py::test_oracle_matches_the_unmodified_binaries); the binary's golden differs from
    it in a handful of lines."""
    sfirst, sub = deep_files[0]
    for variant, maxc, name in ((0, 50, "asmdeep.asmpw50"), (1, 50, "asmdeep.trimpw50"), (1, 100, None)):
        want = util.asm_lines(util.asm_oracle_overlaps(sub, sfirst, sub, sfirst, variant=variant, maxc=maxc))
        got = gpu_pairs(gpu_ctx, deep_files, variant=variant, maxc=maxc)
        assert got == want
        if name:
            g = gold(name)
            assert len(set(got) ^ set(g)) <= 16
    # many table batches and a record pool that runs out and splits its batch: same records
    monkeypatch.setenv("MECAT_B200_ASM_TABLE_MB", "1")
    monkeypatch.setenv("MECAT_B200_ASM_POOL_DIV", "64")
    before = gpu_ctx.stats()["kernel_launches"]["asm_seed"]
    assert gpu_pairs(gpu_ctx, deep_files, variant=1, maxc=100) == want
    assert gpu_ctx.stats()["kernel_launches"]["asm_seed"] - before > 20


def test_awkward_reads_match_the_unmodified_binaries(gpu_ctx, tmp_path):
    """The `asmodd` fixture (tests/util.py: a duplicate, tandem repeats, poly-A, N runs, other letters, mixed case, reads of
    1 / 13 / 14 letters, a contained read) through the C ABI against the goldens of both programs."""
    files = util.asm_workdir("asmodd", str(tmp_path / "odd"))
    assert sorted(gpu_pairs(gpu_ctx, files)) == gold("asmodd.asmpw")
    assert sorted(gpu_pairs(gpu_ctx, files, variant=1, maxc=50)) == gold("asmodd.trimpw50")


def test_bad_inputs_fail_loudly(gpu_ctx):
    import mecat_b200
    idx = gpu_ctx.asm_index_build(mecat_b200.AsmReads(["ACGT" * 100, "TTGCA" * 90], 1))
    assert len(gpu_ctx.asm_overlaps(idx, mecat_b200.AsmReads(["ACGT" * 100, "", "ACGTAC"], 1))) == 0
    with pytest.raises(mecat_b200.MecatB200Error):
        gpu_ctx.asm_overlaps(idx, mecat_b200.AsmReads(["A" * 100000], 1))
    with pytest.raises(mecat_b200.MecatB200Error):
        gpu_ctx.asm_overlaps(idx, mecat_b200.AsmReads(["ACGT" * 100], 1), variant=2)
    with pytest.raises(mecat_b200.MecatB200Error):
        gpu_ctx.asm_overlaps(idx, mecat_b200.AsmReads(["ACGT" * 100], 1), max_candidates=101)
    gpu_ctx.asm_index_release(idx)
    with pytest.raises(mecat_b200.MecatB200Error):
        gpu_ctx.asm_index_build(mecat_b200.AsmReads([], 1))


def test_command_line_drivers_match_the_unmodified_binaries(tmp_path):
    """bin/mecat2asmpw, mecat2trimpw and the *50 names with the pipeline's call (Overlapmecat2asmpw.pm:483-496)."""
    bindir = os.path.join(util.ROOT, "mecat_b200", "bin")
    wrk = str(tmp_path / "blocks")
    util.asm_workdir("asm", wrk)

    def run(prog, first, last, threads, devices=1):
        for f in os.listdir(wrk):
            if f.endswith(".r"):
                os.remove(os.path.join(wrk, f))
        p = subprocess.run([os.path.join(bindir, prog), "-P" + wrk, "-T%d" % threads, "-S%d" % first, "-E%d" % last], capture_output=True, text=True,
                           env=dict(os.environ, MECAT_GPUS=str(devices)))
        assert p.returncode == 0, p.stderr[-2000:]
        names = sorted(f for f in os.listdir(wrk) if f.endswith(".r"))
        assert names == ["%d_%d.r" % (first, t) for t in range(threads)]
        lines = []
        for f in names:
            lines += open(os.path.join(wrk, f)).read().splitlines()
        return sorted(lines)

    assert run("mecat2asmpw", 1, 2, 4) == gold("asm.asmpw")
    assert run("mecat2asmpw50", 2, 2, 1) == gold("asm.asmpw.s2")       # fewer than 50 candidates per read here
    assert run("mecat2trimpw", 1, 2, 2) == gold("asm.trimpw")
    import mecat_b200
    if mecat_b200.load_library().mecat_b200_device_count() >= 2:      # an index replica per device, the reads of every file split between them
        assert run("mecat2asmpw", 1, 2, 4, devices=2) == gold("asm.asmpw")
