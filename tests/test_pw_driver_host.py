"""The mecat2pw command-line driver (mecat_b200/csrc/host/mecat2pw.cpp) run on the host: the product's own volume split and
host I/O, with the device-side ABI calls played by the CPU oracle (tests/pw_abi_shim.cpp).  What is under test is the driver:
flags, the split into volumes, the tile schedule over several devices, the wrk/r_N resume protocol, the merged output.
(The same driver against the CUDA library is part of `pytest -m gpu`.)"""
import gzip
import os
import re
import shutil
import subprocess

import pytest

import util


def run_driver(args, env=None, ok=True):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([util.pw_driver_on_host()] + args, env=e, capture_output=True, text=True)
    assert (p.returncode == 0) == ok, p.stderr[-2000:]
    return p


def gold(name):
    with gzip.open(os.path.join(util.GOLDEN, name), "rt") as f:
        return f.read().splitlines()


@pytest.fixture(scope="module")
def small_fa(tmp_path_factory):
    fa = str(tmp_path_factory.mktemp("pwdrv") / "small.fa")
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f, open(fa, "wb") as g:
        g.write(f.read())
    return fa


def test_single_volume_files_match_reference(small_fa, tmp_path):
    """-j 0 and -j 1 -g 1 on the 250-read fixture: the unmodified binary's candidate and overlap lines."""
    can, m4 = str(tmp_path / "o.can"), str(tmp_path / "o.m4")
    run_driver(["-j", "0", "-d", small_fa, "-o", can, "-w", str(tmp_path / "w0"), "-t", "2"])
    assert sorted(open(can).read().splitlines()) == gold("small.can.gz")
    run_driver(["-j", "1", "-g", "1", "-d", small_fa, "-o", m4, "-w", str(tmp_path / "w1")])
    assert sorted(open(m4).read().splitlines()) == gold("small.m4.gz")
    assert os.path.exists(str(tmp_path / "w1" / "r_0")) and not os.path.exists(str(tmp_path / "w1" / "r_0.working"))


def test_tiles_are_shared_between_devices_and_rows_resume(small_fa, tmp_path):
    """The first 130 reads in volumes of 250 kbase: four or more volumes, 10+ tiles (the oracle behind the shim spends half a
    second per tile on its 2^26-entry index, whatever the tile holds).  One device and three devices write byte-identical
    files (rows in volume order, tiles in order inside a row); the records are the oracle's, tile by tile; a finished row
    (r_N present) is not recomputed."""
    env = {"MECAT_VOLUME_BASES": "250000", "MECAT_SHIM_REPORT": "1", "MECAT_B200_FAST_EXIT": "0"}   # the shim reports when device 0 is released
    part = str(tmp_path / "part.fa")
    with open(small_fa, "rb") as f, open(part, "wb") as g:
        g.write(b"".join(f.readlines()[:260]))
    small_fa = part
    one, w1 = str(tmp_path / "one.m4"), str(tmp_path / "w1")
    p = run_driver(["-j", "1", "-d", small_fa, "-o", one, "-w", w1], env=env)
    nv = len(open(os.path.join(w1, "fileindex.txt")).read().split())
    assert nv >= 4
    ntiles = nv * (nv + 1) // 2
    assert re.search(r"\[shim\] tiles=%d index_builds=%d\b" % (ntiles, nv), p.stderr), p.stderr[-400:]     # one device: one index per row
    vols = [util.PackedVolume.load(os.path.join(w1, "vol%d" % i)) for i in range(nv)]
    want = []
    for s in range(nv):
        for v in range(s, nv):
            want += util.m4_lines(util.oracle_pw_tile(vols[s], vols[v], util.pw_params(task=1), threads=4))
    assert sorted(open(one).read().splitlines()) == sorted(want)
    three, w3 = str(tmp_path / "three.m4"), str(tmp_path / "w3")
    p = run_driver(["-j", "1", "-d", small_fa, "-o", three, "-w", w3], env=dict(env, MECAT_GPUS="3", MECAT_SHIM_DEVICES="3"))
    assert open(three).read() == open(one).read()
    m = re.search(r"\[shim\] tiles=(\d+) index_builds=(\d+)", p.stderr)
    assert int(m.group(1)) == ntiles and nv <= int(m.group(2)) <= ntiles
    for i in range(nv):
        assert open(os.path.join(w3, "r_%d" % i)).read() == open(os.path.join(w1, "r_%d" % i)).read()
    # resume: rows 0 and 2 are there already, the others are recomputed
    w4 = str(tmp_path / "w4")
    shutil.copytree(w1, w4)
    for i in range(nv):
        if i not in (0, 2):
            os.remove(os.path.join(w4, "r_%d" % i))
    again = str(tmp_path / "again.m4")
    p = run_driver(["-j", "1", "-d", small_fa, "-o", again, "-w", w4], env=dict(env, MECAT_GPUS="2", MECAT_SHIM_DEVICES="2"))
    assert "volume 0 has been finished" in p.stderr and "volume 2 has been finished" in p.stderr
    m = re.search(r"\[shim\] tiles=(\d+)", p.stderr)
    assert int(m.group(1)) == ntiles - nv - (nv - 2)
    assert open(again).read() == open(one).read()


def test_option_handling(small_fa, tmp_path):
    w = str(tmp_path / "w")
    assert "output must be specified" in run_driver(["-d", small_fa, "-w", w], ok=False).stderr
    assert "task (-j) must be 0 or 1" in run_driver(["-j", "3", "-d", small_fa, "-o", str(tmp_path / "o"), "-w", w], ok=False).stderr
    assert "invalid argument" in run_driver(["-x", "2", "-d", small_fa, "-o", str(tmp_path / "o"), "-w", w], ok=False).stderr


def test_nanopore_option_sets_the_other_defaults(small_fa, tmp_path):
    """-x 1: -a 500 -k 2 unless given, min_kmer_dist 400 and the X-drop aligner behind the tile call; the unmodified
    binary's candidate and overlap lines for `-x 1`."""
    can, m4 = str(tmp_path / "o.can"), str(tmp_path / "o.m4")
    run_driver(["-j", "0", "-x", "1", "-d", small_fa, "-o", can, "-w", str(tmp_path / "w0"), "-t", "2"])
    assert sorted(open(can).read().splitlines()) == gold("small.x1.can.gz")
    run_driver(["-j", "1", "-g", "1", "-x", "1", "-d", small_fa, "-o", m4, "-w", str(tmp_path / "w1")])
    assert sorted(open(m4).read().splitlines()) == gold("small.x1.m4.gz")
