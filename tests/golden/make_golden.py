"""Regenerates tests/golden/* from the UNMODIFIED reference binaries (oracle/_ref, built by
`make -C oracle ref` in the build container where /root/reference exists).

  python tests/golden/make_golden.py

Cases (reads come from oracle/gen_reads, a seeded deterministic generator):
  small : 250 reads, mean 6 kb, genome 100 kb, seed 3   -> FASTA committed (gzip)
  cfg0  : 1000 reads, mean 15 kb, genome 1 Mb, seed 7   -> BASELINE.json configs[0]; FASTA
          regenerated on demand, its sha256 is committed
  deep  : 300 reads, mean 6 kb, genome 15 kb, seed 5    -> ~120x coverage: every read has the full 100 candidates,
          so mecat2cns reaches its 60-alignment cap and its 20x coverage gate (check_cov_stats); only the
          candidates and the corrected FASTA (-l 2000 -c 4 -a 1000) are kept, FASTA regenerated on demand
  refmap: 300 reads, mean 6 kb, genome 100 kb, seed 5  -> `mecat2ref -m 1` (M4) and `-m 0` (ref format with alignment
          strings) of the reads against their own genome: fixtures for the mecat2ref driver,
          which is not built yet; reads and genome are regenerated on demand (sha256 committed)
  x1    : the small / cfg0 / deep / refmap inputs with `-x 1` (nanopore: XdropAligner, min_kmer_dist 400, the nanopore
          consensus variant) -- SURVEY.md section 8(f) item 2
  i1    : `mecat2cns -i 1` (M4 input) on the sorted small / deep overlap files, one OpenMP thread
  asm   : mecat2asmpw / mecat2trimpw (and the *50 programs) of mecat2canu on corrected-read like fixtures (tests/util.py
          ASM_CASES): `asm` two files (-S1 -E2 and -S2 -E2), `asmdeep` one deep file through mecat2asmpw50 with one thread
  python tests/golden/make_golden.py [case ...]   regenerates only the named cases
For each: vol0 sha256 (split_raw_dataset), sorted `mecat2pw -j 0` lines, sorted
`mecat2pw -j 1 -g 1` lines.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import gen_reads, make_refmap_hard, REF_DIR, ASM_CASES, asm_workdir  # noqa: E402

CASES = {
    "small": dict(n=250, genome=100000, seed=3, mean=6000, sd=1500),
    "cfg0": dict(n=1000, genome=1000000, seed=7, mean=15000, sd=1500),
    "deep": dict(n=300, genome=15000, seed=5, mean=6000, sd=1000),
}
REFMAP = dict(n=300, genome=100000, seed=5, mean=6000, sd=1500)


def make_refmap(meta):
    c = REFMAP
    tmp = tempfile.mkdtemp(prefix="golden_ref_")
    fa, genome = os.path.join(tmp, "reads.fa"), os.path.join(tmp, "genome.fa")
    gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], genome_out=genome)
    m = dict(c)
    m["fasta_sha256"] = sha(fa)
    m["genome_sha256"] = sha(genome)
    for fmt, ext in ((1, "m4"), (0, "ref")):
        out = os.path.join(tmp, "out." + ext)
        subprocess.check_call([os.path.join(REF_DIR, "mecat2ref"), "-d", fa, "-r", genome, "-o", out, "-w", os.path.join(tmp, "w" + ext),
                               "-t", "4", "-m", str(fmt)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
        text = open(out).read()
        if fmt == 1:
            lines = sorted(text.splitlines())
            with gzip.open(os.path.join(HERE, "refmap.m4.gz"), "wt") as f:
                f.write("\n".join(lines) + "\n")
            m["num_m4"] = len(lines)
        else:
            recs = text.split("\n")
            groups = sorted("\n".join(recs[i:i + 3]) for i in range(0, len(recs) - 1, 3))     # header, query string, subject string
            with gzip.open(os.path.join(HERE, "refmap.ref.gz"), "wt") as f:
                f.write("\n".join(groups) + "\n")
            m["num_ref"] = len(groups)
    meta["refmap"] = m
    shutil.rmtree(tmp)
    # BASELINE configs[0]-sized reads (1 000 x 15 kb) against their 1 Mb genome: M4 only
    c = dict(n=1000, genome=1000000, seed=7, mean=15000, sd=1500)
    tmp = tempfile.mkdtemp(prefix="golden_ref1_")
    fa, genome = os.path.join(tmp, "reads.fa"), os.path.join(tmp, "genome.fa")
    gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], genome_out=genome)
    m = dict(c)
    m["fasta_sha256"] = sha(fa); m["genome_sha256"] = sha(genome)
    out = os.path.join(tmp, "out.m4")
    subprocess.check_call([os.path.join(REF_DIR, "mecat2ref"), "-d", fa, "-r", genome, "-o", out, "-w", os.path.join(tmp, "w"), "-t", "8", "-m", "1"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
    lines = sorted(open(out).read().splitlines())
    with gzip.open(os.path.join(HERE, "refmap_cfg0.m4.gz"), "wt") as f:
        f.write("\n".join(lines) + "\n")
    m["num_m4"] = len(lines)
    meta["refmap_cfg0"] = m
    shutil.rmtree(tmp)
    # second fixture: inputs that leave the main path (util.make_refmap_hard)
    tmp = tempfile.mkdtemp(prefix="golden_ref2_")
    fa, genome = os.path.join(tmp, "reads.fa"), os.path.join(tmp, "genome.fa")
    make_refmap_hard(fa, genome)
    m = {"fasta_sha256": sha(fa), "genome_sha256": sha(genome)}
    out = os.path.join(tmp, "out.ref")
    subprocess.check_call([os.path.join(REF_DIR, "mecat2ref"), "-d", fa, "-r", genome, "-o", out, "-w", os.path.join(tmp, "w"), "-t", "3", "-m", "0"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)      # the binary leaves ./config.txt behind
    recs = open(out).read().split("\n")
    groups = sorted("\n".join(recs[i:i + 3]) for i in range(0, len(recs) - 1, 3))
    with gzip.open(os.path.join(HERE, "refmap_hard.ref.gz"), "wt") as f:
        f.write("\n".join(groups) + "\n")
    m["num_ref"] = len(groups)
    # the same inputs as SAM (-m 2): header lines except @PG (it holds the command line), then the records, sorted
    out = os.path.join(tmp, "out.sam")
    subprocess.check_call([os.path.join(REF_DIR, "mecat2ref"), "-d", fa, "-r", genome, "-o", out, "-w", os.path.join(tmp, "ws"), "-t", "3", "-m", "2"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
    lines = open(out).read().splitlines()
    head = [l for l in lines if l.startswith("@") and not l.startswith("@PG")]
    recs = sorted(l for l in lines if not l.startswith("@"))
    with gzip.open(os.path.join(HERE, "refmap_hard.sam.gz"), "wt") as f:
        f.write("\n".join(head + recs) + "\n")
    m["num_sam"] = len(recs)
    meta["refmap_hard"] = m
    shutil.rmtree(tmp)


def make_nanopore(meta):
    """`-x 1` (nanopore) goldens: the same reads through the reference's other aligner and parameter set --
    mecat2pw -x 1 (-j 0: min_kmer_dist 400, -k 2; -j 1: XdropAligner, -a 500), mecat2cns -x 1 -i 0 (error rate 0.20, up to
    100 alignments per read, the whole read as the only effective range), mecat2ref -x 1 (XdropAligner)."""
    m = {}
    for name in ("small", "cfg0"):
        c = CASES[name]
        tmp = tempfile.mkdtemp(prefix="golden_x1_")
        fa = os.path.join(tmp, "reads.fa")
        gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
        for job, ext, extra in ((0, "can", []), (1, "m4", ["-g", "1"])):
            if name == "cfg0" and job == 0:
                continue
            out = os.path.join(tmp, "out." + ext)
            subprocess.check_call([os.path.join(REF_DIR, "mecat2pw"), "-j", str(job), "-x", "1", "-d", fa, "-o", out, "-w", os.path.join(tmp, "wrk%d" % job),
                                   "-t", "8"] + extra, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            lines = sorted(open(out).read().splitlines())
            with gzip.open(os.path.join(HERE, "%s.x1.%s.gz" % (name, ext)), "wt") as f:
                f.write("\n".join(lines) + "\n")
            m["%s_num_%s" % (name, ext)] = len(lines)
        if name == "small":
            can = os.path.join(tmp, "in.can")
            shutil.copy(os.path.join(tmp, "out.can"), can)
            m["small_num_cns"] = run_cns_x1(can, fa, os.path.join(HERE, "small.x1.cns.fa.gz"), tmp)
        shutil.rmtree(tmp)
    # the deep fixture's own (pacbio) candidates through the nanopore consensus: 100 candidates per read at ~120x
    c = CASES["deep"]
    tmp = tempfile.mkdtemp(prefix="golden_x1d_")
    fa = os.path.join(tmp, "reads.fa")
    gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
    can = os.path.join(tmp, "in.can")
    with gzip.open(os.path.join(HERE, "deep.can.gz"), "rt") as f, open(can, "w") as g:
        g.write(f.read())
    m["deep_num_cns"] = run_cns_x1(can, fa, os.path.join(HERE, "deep.x1.cns.fa.gz"), tmp)
    shutil.rmtree(tmp)
    # mecat2ref -x 1 on the refmap fixture
    c = REFMAP
    tmp = tempfile.mkdtemp(prefix="golden_x1r_")
    fa, genome = os.path.join(tmp, "reads.fa"), os.path.join(tmp, "genome.fa")
    gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], genome_out=genome)
    for fmt, ext in ((1, "m4"), (0, "ref")):
        out = os.path.join(tmp, "out." + ext)
        subprocess.check_call([os.path.join(REF_DIR, "mecat2ref"), "-d", fa, "-r", genome, "-o", out, "-w", os.path.join(tmp, "w" + ext),
                               "-t", "4", "-m", str(fmt), "-x", "1"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
        text = open(out).read()
        if fmt == 1:
            lines = sorted(text.splitlines())
            with gzip.open(os.path.join(HERE, "refmap.x1.m4.gz"), "wt") as f:
                f.write("\n".join(lines) + "\n")
            m["refmap_num_m4"] = len(lines)
        else:
            recs = text.split("\n")
            groups = sorted("\n".join(recs[i:i + 3]) for i in range(0, len(recs) - 1, 3))
            with gzip.open(os.path.join(HERE, "refmap.x1.ref.gz"), "wt") as f:
                f.write("\n".join(groups) + "\n")
            m["refmap_num_ref"] = len(groups)
    shutil.rmtree(tmp)
    meta["x1"] = m


def make_m4_input(meta):
    """`mecat2cns -i 1` (M4 input, the reference's default input type): the overlaps of `mecat2pw -j 1 -g 1` as SORTED lines
    (the order of the lines in the file decides how equal keys fall in the reference's std::sort, so the file is part of the
    fixture), corrected by the unmodified binary with ONE OpenMP thread -- the parallel-mode sort of the partition
    (reads_correction_aux.cpp:102) is then the sequential introsort, which is the order this repository reproduces.
    small: the committed small.m4.gz, relaxed options; deep (~120x): more than 60 overlaps per read, so the 60 largest are chosen."""
    m = {}
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for name, args in (("small", ["-l", "2000", "-c", "4", "-a", "1000"]), ("deep", ["-l", "2000", "-c", "4", "-a", "1000"])):
        c = CASES[name]
        tmp = tempfile.mkdtemp(prefix="golden_i1_")
        fa = os.path.join(tmp, "reads.fa")
        gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
        m4 = os.path.join(tmp, "in.m4")
        if name == "small":
            with gzip.open(os.path.join(HERE, "small.m4.gz"), "rt") as f, open(m4, "w") as g:
                g.write(f.read())
        else:
            raw = os.path.join(tmp, "raw.m4")
            subprocess.check_call([os.path.join(REF_DIR, "mecat2pw"), "-j", "1", "-g", "1", "-d", fa, "-o", raw, "-w", os.path.join(tmp, "wrk"), "-t", "8"],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            lines = sorted(open(raw).read().splitlines())
            with open(m4, "w") as g:
                g.write("\n".join(lines) + "\n")
            with gzip.open(os.path.join(HERE, "deep.m4.gz"), "wt") as g:
                g.write("\n".join(lines) + "\n")
            m["deep_num_m4"] = len(lines)
        # small also with partitions of 100 reads (-p 100: three partition files, each ordered on its own)
        for tag, extra in (("i1", []), ("i1p100", ["-p", "100"])) if name == "small" else (("i1", []),):
            out = os.path.join(tmp, "cns_%s.fa" % tag)
            subprocess.check_call([os.path.join(REF_DIR, "mecat2cns"), "-i", "1", "-t", "1"] + extra + args + [m4, fa, out], env=env,
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            lines = open(out).read().splitlines()
            recs = sorted(zip(lines[0::2], lines[1::2]))
            with gzip.open(os.path.join(HERE, "%s.%s.cns.fa.gz" % (name, tag)), "wt") as f:
                for h, q in recs:
                    f.write(h + "\n" + q + "\n")
            m["%s_num_cns%s" % (name, "" if tag == "i1" else "_p100")] = len(recs)
        shutil.rmtree(tmp)
    # nanopore defaults: -x 1 alone means -i 1 (consensus_one_read_m4_nanopore: every alignment that also passes the
    # mapping-ratio test, up to 100 per read), on the -x 1 overlaps of the small fixture
    c = CASES["small"]
    tmp = tempfile.mkdtemp(prefix="golden_x1i1_")
    fa = os.path.join(tmp, "reads.fa")
    gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
    m4 = os.path.join(tmp, "in.m4")
    with gzip.open(os.path.join(HERE, "small.x1.m4.gz"), "rt") as f, open(m4, "w") as g:
        g.write(f.read())
    out = os.path.join(tmp, "cns.fa")
    subprocess.check_call([os.path.join(REF_DIR, "mecat2cns"), "-x", "1", "-t", "1", m4, fa, out], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lines = open(out).read().splitlines()
    recs = sorted(zip(lines[0::2], lines[1::2]))
    with gzip.open(os.path.join(HERE, "small.x1i1.cns.fa.gz"), "wt") as f:
        for h, q in recs:
            f.write(h + "\n" + q + "\n")
    m["small_x1_num_cns"] = len(recs)
    shutil.rmtree(tmp)
    meta["i1"] = m


def make_asm(meta):
    """The overlappers mecat2canu runs on corrected reads, unmodified, compiled without
    CFLAGS like the reference's build.  The programs write one file per thread; the goldens are the sorted lines."""
    m = {k: dict(v) for k, v in ASM_CASES.items()}
    def run(prog, wrk, s, e, threads):
        for f in os.listdir(wrk):
            if f.endswith(".r"):
                os.remove(os.path.join(wrk, f))
        subprocess.check_call([os.path.join(REF_DIR, prog), "-P" + wrk, "-T%d" % threads, "-S%d" % s, "-E%d" % e])
        lines = []
        for f in sorted(os.listdir(wrk)):
            if f.endswith(".r"):
                lines += open(os.path.join(wrk, f)).read().splitlines()
        return sorted(lines)
    def keep(name, lines):
        with gzip.open(os.path.join(HERE, name + ".r.gz"), "wt") as f:
            f.write("\n".join(lines) + "\n")
        m["num_" + name.replace(".", "_")] = len(lines)
    tmp = tempfile.mkdtemp(prefix="golden_asm_")
    wrk = os.path.join(tmp, "asm")
    asm_workdir("asm", wrk)
    keep("asm.asmpw", run("mecat2asmpw", wrk, 1, 2, 4))
    keep("asm.asmpw.s2", run("mecat2asmpw", wrk, 2, 2, 4))
    keep("asm.trimpw", run("mecat2trimpw", wrk, 1, 2, 4))
    wrk = os.path.join(tmp, "asmdeep")
    asm_workdir("asmdeep", wrk)
    keep("asmdeep.asmpw50", run("mecat2asmpw50", wrk, 1, 1, 1))
    keep("asmdeep.trimpw50", run("mecat2trimpw50", wrk, 1, 1, 1))
    wrk = os.path.join(tmp, "asmodd")
    asm_workdir("asmodd", wrk)
    keep("asmodd.asmpw", run("mecat2asmpw", wrk, 1, 1, 4))
    keep("asmodd.trimpw50", run("mecat2trimpw50", wrk, 1, 1, 1))
    # the schedule fixture: only digests of the sorted lines, one thread and one thread per chunk
    wrk = os.path.join(tmp, "asmsched")
    asm_workdir("asmsched", wrk)
    t1, t4 = run("mecat2asmpw50", wrk, 1, 1, 1), run("mecat2asmpw50", wrk, 1, 1, 4)
    m["asmsched_lines"] = len(t1)
    m["asmsched_T1_sha256"] = hashlib.sha256("\n".join(t1).encode()).hexdigest()
    m["asmsched_T4_sha256"] = hashlib.sha256("\n".join(t4).encode()).hexdigest()
    m["asmsched_T1_vs_T4_lines"] = len(set(t1) ^ set(t4))
    meta["asm"] = m
    shutil.rmtree(tmp)


def run_cns_x1(can, fa, dest, tmp):
    out = os.path.join(tmp, "cns_x1.fa")
    subprocess.check_call([os.path.join(REF_DIR, "mecat2cns"), "-x", "1", "-i", "0", "-t", "1", can, fa, out],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lines = open(out).read().splitlines()
    recs = sorted(zip(lines[0::2], lines[1::2]))
    with gzip.open(dest, "wt") as f:
        for h, q in recs:
            f.write(h + "\n" + q + "\n")
    return len(recs)


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for b in iter(lambda: f.read(1 << 20), b""):
            h.update(b)
    return h.hexdigest()


def main():
    meta = {}
    if os.path.exists(os.path.join(HERE, "golden.json")):
        meta = json.load(open(os.path.join(HERE, "golden.json")))
    for name, c in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        if name in ("x1", "i1"):
            continue
        tmp = tempfile.mkdtemp(prefix="golden_")
        fa = os.path.join(tmp, "reads.fa")
        gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
        m = dict(c)
        m["fasta_sha256"] = sha(fa)
        for job, ext, extra in ((0, "can", []), (1, "m4", ["-g", "1"])):
            if name == "deep" and job == 1:
                continue
            wrk = os.path.join(tmp, "wrk%d" % job)
            out = os.path.join(tmp, "out." + ext)
            subprocess.check_call([os.path.join(REF_DIR, "mecat2pw"), "-j", str(job), "-d", fa, "-o", out, "-w", wrk,
                                   "-t", "8"] + extra, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            lines = sorted(open(out).read().splitlines())
            with gzip.open(os.path.join(HERE, "%s.%s.gz" % (name, ext)), "wt") as f:
                f.write("\n".join(lines) + "\n")
            m["num_" + ext] = len(lines)
            if job == 0:
                m["vol0_sha256"] = sha(os.path.join(wrk, "vol0"))
        # mecat2cns -i 0 on the .can above (written next to a copy: it drops partition files beside its input)
        can = os.path.join(tmp, "in.can")
        for tag, args in (("cns_default", []), ("cns_relaxed", ["-l", "2000", "-c", "4", "-a", "1000"])):
            if (name == "cfg0" and tag == "cns_relaxed") or (name == "deep" and tag == "cns_default"):
                continue
            with open(can, "w") as f:
                f.write(open(os.path.join(tmp, "out.can")).read())
            out = os.path.join(tmp, tag + ".fa")
            subprocess.check_call([os.path.join(REF_DIR, "mecat2cns"), "-i", "0", "-t", "1"] + args + [can, fa, out],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            lines = open(out).read().splitlines()
            recs = sorted(zip(lines[0::2], lines[1::2]))
            with gzip.open(os.path.join(HERE, "%s.%s.fa.gz" % (name, tag)), "wt") as f:
                for h, q in recs:
                    f.write(h + "\n" + q + "\n")
            m["num_" + tag] = len(recs)
        if name == "small":
            with open(fa, "rb") as f, gzip.open(os.path.join(HERE, "small.fa.gz"), "wb") as g:
                shutil.copyfileobj(f, g)
        meta[name] = m
        shutil.rmtree(tmp)
    if len(sys.argv) == 1 or "refmap" in sys.argv[1:]:
        make_refmap(meta)
    if len(sys.argv) == 1 or "x1" in sys.argv[1:]:
        make_nanopore(meta)
    if len(sys.argv) == 1 or "i1" in sys.argv[1:]:
        make_m4_input(meta)
    if len(sys.argv) == 1 or "asm" in sys.argv[1:]:
        make_asm(meta)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
