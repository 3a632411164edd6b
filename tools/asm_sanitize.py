"""Small mecat2asmpw run for compute-sanitizer (memcheck / racecheck): the second block file of the `asm` fixture against
itself (300 reads) and the first reads of the deep fixture with forced table batches; checks the records against the golden
of the unmodified binary / the oracle's count."""
import gzip
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mecat_b200  # noqa: E402
import util  # noqa: E402


def main():
    nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    tmp = tempfile.mkdtemp(prefix="asm_sanitize_")
    files = util.asm_workdir("asm", os.path.join(tmp, "asm"))
    first, reads = files[1]
    ctx = mecat_b200.Context(0)
    if nreads >= len(reads):
        idx = ctx.asm_index_build(mecat_b200.AsmReads(reads, first))
        got = sorted(mecat_b200.asm_lines(ctx.asm_overlaps(idx, mecat_b200.AsmReads(reads, first), 0, 100)))
        ctx.asm_index_release(idx)
        with gzip.open(os.path.join(util.GOLDEN, "asm.asmpw.s2.r.gz"), "rt") as f:
            want = f.read().splitlines()
        print("asm file 2:", len(got), "records, equal to the unmodified binary's:", got == want)
        assert got == want
    deep = util.asm_workdir("asmdeep", os.path.join(tmp, "deep"))[0][1][:nreads]
    os.environ["MECAT_B200_ASM_TABLE_MB"] = "1"
    idx = ctx.asm_index_build(mecat_b200.AsmReads(deep, 1))
    got = ctx.asm_overlaps(idx, mecat_b200.AsmReads(deep, 1), 1, 50)
    ctx.asm_index_release(idx)
    want = util.asm_oracle_overlaps(deep, 1, deep, 1, variant=1, maxc=50)
    print("deep, first %d reads, forced batches:" % len(deep), len(got), "records, oracle", len(want), "equal:", mecat_b200.asm_lines(got) == util.asm_lines(want))
    assert mecat_b200.asm_lines(got) == util.asm_lines(want)
    ctx.close()


if __name__ == "__main__":
    main()
