"""Does the unmodified mecat2asmpw50 print the same overlaps whatever its thread count?  (It does not: DESIGN.md section 4.11.)

  python tools/asm_reference_schedule.py > profiles/r2_asm_reference_schedule.json      (CPU only, needs oracle/_ref)

1 500 corrected-read like reads at ~130x (three chunks of PLL = 500 reads, mecat2asmpw.c:26), the *50 program (the candidate
list of a read is cut at 50, so a score that differs by one changes which overlaps are printed).  Runs the binary with 1, 4
and 8 threads and the oracle in both conventions: `history` = one thread's block memory carried from read to read (must
equal -T1), zeroed blocks per strand (what the product implements)."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402


def main():
    tmp = tempfile.mkdtemp(prefix="asm_schedule_")
    fa = os.path.join(tmp, "000001.fasta")
    n = 1500
    util.gen_reads(fa, n, 40000, 123, 3500, 900, err=0.01)
    with open(os.path.join(tmp, "ovlprep"), "w") as f:
        f.write("-allreads -allbases -b 1 -e %d\n" % n)
    runs = {}
    for t in (1, 4, 8):
        for f in os.listdir(tmp):
            if f.endswith(".r"):
                os.remove(os.path.join(tmp, f))
        subprocess.check_call([os.path.join(util.REF_DIR, "mecat2asmpw50"), "-P" + tmp, "-T%d" % t, "-S1", "-E1"])
        lines = []
        for f in os.listdir(tmp):
            if f.endswith(".r"):
                lines += open(os.path.join(tmp, f)).read().splitlines()
        runs[t] = sorted(lines)
    seqs = [s for _, s in util.read_fasta_raw(fa)]
    hist = sorted(util.asm_lines(util.asm_oracle_overlaps(seqs, 1, seqs, 1, variant=0, maxc=50, history=1)))
    zero = sorted(util.asm_lines(util.asm_oracle_overlaps(seqs, 1, seqs, 1, variant=0, maxc=50, history=0)))
    d = lambda a, b: len(set(a) ^ set(b))
    print(json.dumps({
        "input": "%d reads x 3.5 kb, 1 %% error, 40 kb genome (~130x), mecat2asmpw50 -S1 -E1" % n,
        "lines": {"T1": len(runs[1]), "T4": len(runs[4]), "T8": len(runs[8]), "oracle_history": len(hist), "oracle_zeroed": len(zero)},
        "lines_differing": {"T1_vs_T4": d(runs[1], runs[4]), "T1_vs_T8": d(runs[1], runs[8]), "T4_vs_T8": d(runs[4], runs[8]),
                            "oracle_history_vs_T1": d(hist, runs[1]), "oracle_zeroed_vs_T1": d(zero, runs[1]), "oracle_zeroed_vs_T4": d(zero, runs[4])},
        "reading": "the binary's output depends on its thread count (which reads a thread mapped before decides what the vote loops find beyond a "
                   "block's entries); the restatement with one thread's memory equals -T1; zeroed blocks change which overlaps survive the cut "
                   "at 50 candidates for a fraction of a percent of the lines, and nothing on the 16x / 32x sets (tests, profiles/r2_bench_asm.json)"}))


if __name__ == "__main__":
    main()
