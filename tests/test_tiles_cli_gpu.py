"""`mecat2pw` on a multi-volume read set through the command-line driver on the GPU (-m gpu): six volumes = 21 tiles handed
out tile by tile (mecat_b200/csrc/host/mecat2pw.cpp), records equal to the oracle's tile by tile, one device and -- where
the box has them -- two devices byte-identical.  The same schedule runs in the CPU suite against an ABI shim
(tests/test_pw_driver_host.py); this file adds the real library underneath.  Sorts last: first hardware run pending."""
import gzip
import os
import subprocess

import pytest

import util

pytestmark = pytest.mark.gpu


def run_cli(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([os.path.join(util.ROOT, "mecat_b200", "bin", "mecat2pw")] + args, env=e, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    return p


def test_multi_volume_tiles_match_oracle(gpu_ctx, tmp_path):
    import mecat_b200
    fa = str(tmp_path / "small.fa")
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f, open(fa, "wb") as g:
        g.write(f.read())
    env = {"MECAT_VOLUME_BASES": "320000"}
    one, w1 = str(tmp_path / "one.m4"), str(tmp_path / "w1")
    run_cli(["-j", "1", "-d", fa, "-o", one, "-w", w1], env=env)
    nv = len(open(os.path.join(w1, "fileindex.txt")).read().split())
    assert nv >= 5
    vols = [util.PackedVolume.load(os.path.join(w1, "vol%d" % i)) for i in range(nv)]
    want = []
    for s in range(nv):
        for v in range(s, nv):
            want += util.m4_lines(util.oracle_pw_tile(vols[s], vols[v], util.pw_params(task=1), threads=4))
    assert sorted(open(one).read().splitlines()) == sorted(want)
    if mecat_b200.load_library().mecat_b200_device_count() >= 2:
        two = str(tmp_path / "two.m4")
        run_cli(["-j", "1", "-d", fa, "-o", two, "-w", str(tmp_path / "w2")], env=dict(env, MECAT_GPUS="2"))
        assert open(two).read() == open(one).read()
