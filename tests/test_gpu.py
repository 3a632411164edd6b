"""GPU parity tests (-m gpu): every result of the CUDA path, obtained through the C ABI, is
compared bit for bit with the CPU oracle and with golden outputs of the unmodified reference."""
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


def report(tag, mine, want, limit=8):
    if mine == want:
        return
    sm, sw = set(mine), set(want)
    msg = ["%s: %d lines vs %d expected; %d only-mine, %d only-expected" % (tag, len(mine), len(want), len(sm - sw), len(sw - sm))]
    msg += ["  mine: " + x for x in sorted(sm - sw)[:limit]]
    msg += ["  want: " + x for x in sorted(sw - sm)[:limit]]
    os.makedirs(os.path.join(util.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(util.ROOT, "gpurun_out", "mismatch_%s.txt" % tag), "w") as f:
        f.write("\n".join(msg) + "\n")
        f.write("\n".join("M " + x for x in sorted(sm - sw)) + "\n")
        f.write("\n".join("W " + x for x in sorted(sw - sm)) + "\n")
    pytest.fail("\n".join(msg))


# ---------------------------------------------------------------- A1
def test_index_matches_oracle(gpu_ctx, small_vol):
    O = util.oracle()
    cv = small_vol.c()
    oidx = O.orc_index_build(C.byref(cv))
    d = gpu_ctx.upload(host_volume(small_vol))
    idx = gpu_ctx.index_build(d)
    begin, pos = gpu_ctx.index_export(idx)
    assert len(pos) == O.orc_index_num_kmers(oidx)
    lst = C.POINTER(C.c_int32)()
    lens = np.diff(begin.astype(np.int64))
    assert lens.max() <= 128
    nz = np.nonzero(lens)[0]
    rng = np.random.default_rng(2)
    for code in list(rng.choice(nz, size=3000, replace=False)) + list(rng.integers(0, 1 << 26, size=500)):
        n = O.orc_index_lookup(oidx, int(code), C.byref(lst))
        assert n == lens[code]
        assert [lst[i] for i in range(n)] == list(pos[begin[code]:begin[code] + n])
    # whole-array property: every list ascending, every position a valid 13-mer start of the right code
    starts = pos.astype(np.int64)
    same = np.repeat(np.arange(len(lens)), lens)
    asc = (np.diff(starts) > 0) | (np.diff(same) != 0)
    assert asc.all()
    gpu_ctx.release_index(idx)
    gpu_ctx.release_volume(d)
    O.orc_index_free(oidx)


def test_index_built_in_code_slices_equals_one_shot(gpu_ctx, small_vol):
    """The two-stage, sliced build used by N GPUs sharing a tile (count_part / exchange / finish_part),
    replayed on one GPU for 2 and 3 slices, gives the same CSR arrays as index_build."""
    import torch
    from mecat_b200 import multi
    d = gpu_ctx.upload(host_volume(small_vol))
    ref = gpu_ctx.index_build(d)
    rbegin, rpos = gpu_ctx.index_export(ref)
    gpu_ctx.release_index(ref)
    NC = 1 << 26
    for world in (2, 3):
        sl = multi.code_slices(world)
        parts = [gpu_ctx.index_count_part(d, lo, hi) for lo, hi in sl]
        views = [multi.device_view(gpu_ctx.index_device_arrays(p)[0], NC, torch.int32, 4) for p in parts]
        for q, (lo, hi) in enumerate(sl):           # "all-gather" of the histogram slices
            for r in range(world):
                if r != q:
                    views[r][lo:hi].copy_(views[q][lo:hi])
        torch.cuda.synchronize()
        for p, (lo, hi) in zip(parts, sl):
            gpu_ctx.index_finish_part(d, p, lo, hi)
        arrs = [gpu_ctx.index_device_arrays(p) for p in parts]
        begin0 = multi.device_view(arrs[0][1], NC + 1, torch.int32, 4)
        pos = [multi.device_view(a[2], a[3], torch.int32, 4) for a in arrs]
        for q, (lo, hi) in enumerate(sl):           # "broadcast" of every position slice
            b0, b1 = int(begin0[lo].item()) & 0xFFFFFFFF, int(begin0[hi].item()) & 0xFFFFFFFF
            for r in range(world):
                if r != q:
                    pos[r][b0:b1].copy_(pos[q][b0:b1])
        torch.cuda.synchronize()
        for p in parts:
            b, ps = gpu_ctx.index_export(p)
            assert (b == rbegin).all() and len(ps) == len(rpos) and (ps == rpos).all()
            gpu_ctx.release_index(p)
    gpu_ctx.release_volume(d)


# ---------------------------------------------------------------- A8-A11
def oracle_extend(vq, vs, tasks, min_aln):
    O = util.oracle()
    out = (C.c_int32 * 8)()
    ident = C.c_double()
    res = []
    cache = {}
    for t in tasks:
        kq = (int(t["qread"]), int(t["qstrand"]))
        if kq not in cache:
            cache[kq] = np.concatenate([[0], vq.codes(kq[0], kq[1]), [0]]).astype(np.int8)
        ks = ("s", int(t["sread"]))
        if ks not in cache:
            cache[ks] = np.concatenate([[0], vs.codes(ks[1], 0), [0]]).astype(np.int8)
        q, s = cache[kq], cache[ks]
        O.orc_diff_go(C.cast(q.ctypes.data + 1, C.c_char_p), int(t["qstart"]), len(q) - 2,
                      C.cast(s.ctypes.data + 1, C.c_char_p), int(t["sstart"]), len(s) - 2, min_aln, out, C.byref(ident), None, None, 0)
        res.append((out[0], out[1], out[2], out[3], out[4], out[5], out[6], ident.value))
    return res


def test_extend_matches_oracle(gpu_ctx, small_vol):
    import mecat_b200
    O = util.oracle()
    cv = small_vol.c()
    oidx = O.orc_index_build(C.byref(cv))
    p = util.pw_params(task=0)
    out = (C.c_int32 * (12 * 101))()
    tasks = []
    for rid in range(small_vol.num_reads):
        n = O.orc_pw_candidates(oidx, C.byref(cv), C.byref(cv), rid, C.byref(p), out)
        for i in range(n):
            c = out[12 * i:12 * i + 12]
            qstart, sstart = c[1], c[0]
            if qstart and sstart:
                qstart += 6; sstart += 6
            tasks.append((rid, c[11], qstart, c[9], sstart))
    O.orc_index_free(oidx)
    rng = np.random.default_rng(4)
    n = small_vol.num_reads
    # edge cases: start points at the very ends, unrelated pairs, random interior points
    for it in range(300):
        a, b = int(rng.integers(0, n)), int(rng.integers(0, n))
        la, lb = int(small_vol.offset_size[a][1]), int(small_vol.offset_size[b][1])
        mode = it % 4
        if mode == 0: qs, ss = 0, 0
        elif mode == 1: qs, ss = la, lb
        elif mode == 2: qs, ss = int(rng.integers(0, la + 1)), int(rng.integers(0, lb + 1))
        else: qs, ss = la // 2, 0
        tasks.append((a, it & 1, qs, b, ss))
    tasks = np.array(tasks, dtype=mecat_b200.TASK_DTYPE)
    d = gpu_ctx.upload(host_volume(small_vol))
    got = gpu_ctx.extend_batch(d, d, tasks, 2000)
    gpu_ctx.release_volume(d)
    want = oracle_extend(small_vol, small_vol, tasks, 2000)
    bad = []
    for i, (g, w) in enumerate(zip(got, want)):
        gg = (int(g["ok"]), int(g["qstart"]), int(g["qend"]), int(g["sstart"]), int(g["send"]), int(g["columns"]), int(g["matches"]), float(g["ident"]))
        if gg != tuple(w):
            bad.append((i, tuple(int(x) for x in tasks[i]), gg, tuple(w)))
    if bad:
        os.makedirs(os.path.join(util.ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(util.ROOT, "gpurun_out", "mismatch_extend.txt"), "w") as f:
            for b in bad:
                f.write(repr(b) + "\n")
    assert not bad, "%d of %d extension results differ; first: %r" % (len(bad), len(tasks), bad[0])


def _oracle_strings(vq, vs, t, policy, min_aln, err=0.15):
    """Oracle result of one align task: (ok, qs, qe, ss, se, qstr, sstr)."""
    O = util.oracle()
    q = np.concatenate([[0], vq.codes(int(t["qread"]), int(t["qstrand"])), [0]]).astype(np.int8)
    full = vs.codes(int(t["sread"]), 0)
    if int(t["swin_len"]) > 0:
        full = full[int(t["swin_off"]):int(t["swin_off"]) + int(t["swin_len"])]
    s = np.concatenate([[0], full, [0]]).astype(np.int8)
    cap = len(q) + len(s) + 64
    qa, sa = C.create_string_buffer(cap), C.create_string_buffer(cap)
    out = (C.c_int32 * 8)()
    ident = C.c_double()
    qp, sp = C.cast(q.ctypes.data + 1, C.c_char_p), C.cast(s.ctypes.data + 1, C.c_char_p)
    if policy == 0:
        O.orc_diff_go(qp, int(t["qstart"]), len(q) - 2, sp, int(t["sstart"]), len(s) - 2, min_aln, out, C.byref(ident), qa, sa, cap)
        return (out[0], out[1], out[2], out[3], out[4], qa.value if out[0] else b"", sa.value if out[0] else b"")
    ok = O.orc_cns_get_alignment(qp, int(t["qstart"]), len(q) - 2, sp, int(t["sstart"]), len(s) - 2, err, min_aln, out, qa, sa, cap)
    return (int(ok), out[1], out[2], out[3], out[4], qa.value if ok else b"", sa.value if ok else b"")


@pytest.mark.parametrize("policy,min_aln", [(0, 1000), (0, 1), (1, 2000), (1, 1)])
def test_align_with_strings_matches_oracle(gpu_ctx, small_vol, policy, min_aln):
    """R1 (pw/ref flavour with mapped strings, incl. subject windows) and C1-C2 (cns GetAlignment)."""
    import mecat_b200
    O = util.oracle()
    cv = small_vol.c()
    oidx = O.orc_index_build(C.byref(cv))
    p = util.pw_params(task=0)
    out = (C.c_int32 * (12 * 101))()
    tasks = []
    for rid in range(0, small_vol.num_reads, 2):
        n = O.orc_pw_candidates(oidx, C.byref(cv), C.byref(cv), rid, C.byref(p), out)
        for i in range(n):
            c = out[12 * i:12 * i + 12]
            qstart, sstart = c[1], c[0]
            if qstart and sstart:
                qstart += 6; sstart += 6
            sidx = c[9]
            if policy == 0 and i % 3 == 0:
                # a window on the subject like mecat2ref's extract_sequences
                sl = int(small_vol.offset_size[sidx][1])
                lo = max(0, sstart - 2500); hi = min(sl, sstart + 3000)
                tasks.append((rid, c[11], qstart, sidx, sstart - lo, lo, hi - lo))
            else:
                tasks.append((rid, c[11], qstart, sidx, sstart, 0, 0))
    O.orc_index_free(oidx)
    rng = np.random.default_rng(8)
    nr = small_vol.num_reads
    for it in range(120):     # unrelated pairs, end points
        a, b = int(rng.integers(0, nr)), int(rng.integers(0, nr))
        la, lb = int(small_vol.offset_size[a][1]), int(small_vol.offset_size[b][1])
        qs, ss = [(0, 0), (la, lb), (la // 2, lb // 3), (int(rng.integers(0, la + 1)), int(rng.integers(0, lb + 1)))][it % 4]
        tasks.append((a, it & 1, qs, b, ss, 0, 0))
    tasks = np.array(tasks, dtype=mecat_b200.ALIGN_TASK_DTYPE)
    d = gpu_ctx.upload(host_volume(small_vol))
    res, qstr, sstr = gpu_ctx.align_batch(d, d, tasks, min_aln, policy=policy, err=0.15)
    gpu_ctx.release_volume(d)
    bad = []
    nok = 0
    for i, t in enumerate(tasks):
        w = _oracle_strings(small_vol, small_vol, t, policy, min_aln)
        r = res[i]
        if int(r["ok"]):
            o = int(r["str_offset"]); n = int(r["columns"])
            g = (1, int(r["qstart"]), int(r["qend"]), int(r["sstart"]), int(r["send"]), qstr[o:o + n], sstr[o:o + n])
            assert qstr[o + n] == 0 and sstr[o + n] == 0
            # matches / ident describe the returned strings (for the cns flavour: after its trimming to 4-match runs)
            same = sum(1 for x, y in zip(qstr[o:o + n], sstr[o:o + n]) if x == y)
            assert int(r["matches"]) == same and abs(float(r["ident"]) - 100.0 * same / n) < 1e-9 and float(r["ident"]) <= 100.0
            nok += 1
        else:
            g = (0,) + tuple(w[1:5]) + (b"", b"") if not w[0] else (0, 0, 0, 0, 0, b"", b"")
        if g != w:
            bad.append((i, tuple(int(x) for x in t), g[:5], w[:5], len(g[5]), len(w[5])))
    if bad:
        os.makedirs(os.path.join(util.ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(util.ROOT, "gpurun_out", "mismatch_align_p%d_%d.txt" % (policy, min_aln)), "w") as f:
            for b in bad:
                f.write(repr(b) + "\n")
    assert nok > 100
    assert not bad, "%d of %d differ; first %r" % (len(bad), len(tasks), bad[0])


# ---------------------------------------------------------------- A2-A6
def test_raw_candidates_match_oracle(gpu_ctx, small_vol):
    O = util.oracle()
    cv = small_vol.c()
    oidx = O.orc_index_build(C.byref(cv))
    p = util.pw_params(task=0)
    d = gpu_ctx.upload(host_volume(small_vol))
    idx = gpu_ctx.index_build(d)
    import mecat_b200
    rows, counts = gpu_ctx.pw_raw_candidates(idx, d, d, mecat_b200.pw_params(task=0), small_vol.num_reads)
    gpu_ctx.release_index(idx)
    gpu_ctx.release_volume(d)
    out = (C.c_int32 * (12 * 101))()
    k = 0
    bad = []
    for rid in range(small_vol.num_reads):
        n = O.orc_pw_candidates(oidx, C.byref(cv), C.byref(cv), rid, C.byref(p), out)
        want = [tuple(out[12 * i:12 * i + 12]) for i in range(n)]
        got = [tuple(int(x) for x in rows[k + i]) for i in range(int(counts[rid]))]
        k += int(counts[rid])
        if got != want:
            bad.append((rid, got, want))
    O.orc_index_free(oidx)
    if bad:
        os.makedirs(os.path.join(util.ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(util.ROOT, "gpurun_out", "mismatch_rawcand.txt"), "w") as f:
            for rid, got, want in bad:
                f.write("read %d\n" % rid)
                for g in got: f.write("  G %r\n" % (g,))
                for w in want: f.write("  W %r\n" % (w,))
    assert not bad, "%d reads differ; first read %d" % (len(bad), bad[0][0])


# ---------------------------------------------------------------- repeat-rich stress (long lists, overflow buckets)
@pytest.fixture(scope="module", params=[10, 18])
def repeat_vol(request):
    return PackedVolume.from_seqs(util.repeat_reads(copies=request.param))


def test_repeat_index_every_list(gpu_ctx, repeat_vol):
    O = util.oracle()
    cv = repeat_vol.c()
    oidx = O.orc_index_build(C.byref(cv))
    d = gpu_ctx.upload(host_volume(repeat_vol))
    idx = gpu_ctx.index_build(d)
    begin, pos = gpu_ctx.index_export(idx)
    gpu_ctx.release_index(idx)
    gpu_ctx.release_volume(d)
    assert len(pos) == O.orc_index_num_kmers(oidx)
    lens = np.diff(begin.astype(np.int64))
    lst = C.POINTER(C.c_int32)()
    nz = np.nonzero(lens)[0]
    assert lens.max() > 64, "fixture no longer exercises the 4-register sort"
    want = np.empty(len(pos), dtype=np.int32)
    k = 0
    for code in nz:
        n = O.orc_index_lookup(oidx, int(code), C.byref(lst))
        assert n == lens[code]
        want[k:k + n] = np.ctypeslib.as_array(lst, shape=(n,))
        k += n
    O.orc_index_free(oidx)
    assert k == len(pos)
    assert (want == pos).all()


def test_repeat_raw_candidates(gpu_ctx, repeat_vol):
    import mecat_b200
    O = util.oracle()
    cv = repeat_vol.c()
    oidx = O.orc_index_build(C.byref(cv))
    p = util.pw_params(task=0)
    d = gpu_ctx.upload(host_volume(repeat_vol))
    idx = gpu_ctx.index_build(d)
    rows, counts = gpu_ctx.pw_raw_candidates(idx, d, d, mecat_b200.pw_params(task=0), repeat_vol.num_reads)
    gpu_ctx.release_index(idx)
    gpu_ctx.release_volume(d)
    out = (C.c_int32 * (12 * 101))()
    k, bad = 0, []
    for rid in range(repeat_vol.num_reads):
        n = O.orc_pw_candidates(oidx, C.byref(cv), C.byref(cv), rid, C.byref(p), out)
        want = [tuple(out[12 * i:12 * i + 12]) for i in range(n)]
        got = [tuple(int(x) for x in rows[k + i]) for i in range(int(counts[rid]))]
        k += int(counts[rid])
        if got != want:
            bad.append((rid, got, want))
    O.orc_index_free(oidx)
    if bad:
        os.makedirs(os.path.join(util.ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(util.ROOT, "gpurun_out", "mismatch_repeat_rawcand.txt"), "w") as f:
            for rid, got, want in bad:
                f.write("read %d  got %d want %d\n" % (rid, len(got), len(want)))
                for i in range(max(len(got), len(want))):
                    g = got[i] if i < len(got) else None
                    w = want[i] if i < len(want) else None
                    if g != w:
                        f.write("  [%d] G %r\n      W %r\n" % (i, g, w))
    assert not bad, "%d reads differ; first read %d" % (len(bad), bad[0][0])


def test_repeat_tiles(gpu_ctx, repeat_vol):
    import mecat_b200
    hv = host_volume(repeat_vol)
    want = util.oracle_pw_tile(repeat_vol, repeat_vol, util.pw_params(task=0), threads=8)
    got = gpu_ctx.pw_candidates(hv, hv)
    report("repeat_can", util.ec_lines(got), util.ec_lines(want))
    assert [tuple(x) for x in got.tolist()] == [tuple(x) for x in want.tolist()]
    want = util.oracle_pw_tile(repeat_vol, repeat_vol, util.pw_params(task=1), threads=8)
    got = gpu_ctx.pw_overlaps(hv, hv)
    report("repeat_m4", util.m4_lines(got, True), util.m4_lines(want, True))


def test_repeat_heavy_long_reads(gpu_ctx):
    """15 kb reads over a genome of 12 near-identical copies at ~11x: a sampled k-mer has ~80 index hits, a strand
    collects > 100 000 hits of buckets that pass the gate -- more than the 65 536 a strand's scratch used to hold (the
    tile failed with `more candidate seeds than the per-strand capacity`; the reference maps such reads like any other)."""
    import mecat_b200
    vol = PackedVolume.from_seqs(util.repeat_reads(seed=33, unit=4000, copies=12, n_reads=48, mean=15000, err=0.02, div=0.005))
    hv = host_volume(vol)
    want = util.oracle_pw_tile(vol, vol, util.pw_params(task=0), threads=8)
    gpu_ctx.reset_stats()
    got = gpu_ctx.pw_candidates(hv, hv)
    assert gpu_ctx.stats()["num_hits"] > 2 * 48 * 65536 // 2          # the strands really are that heavy
    report("repeat_heavy_can", util.ec_lines(got), util.ec_lines(want))
    want = util.oracle_pw_tile(vol, vol, util.pw_params(task=1), threads=8)
    got = gpu_ctx.pw_overlaps(hv, hv)
    report("repeat_heavy_m4", util.m4_lines(got, True), util.m4_lines(want, True))


def test_deterministic_across_runs(gpu_ctx, cfg0_vol):
    hv = host_volume(cfg0_vol)
    a = gpu_ctx.pw_overlaps(hv, hv)
    b = gpu_ctx.pw_overlaps(hv, hv)
    assert a.tobytes() == b.tobytes()


# ---------------------------------------------------------------- whole tiles vs the reference binaries
def test_small_can_matches_reference(gpu_ctx, small_vol):
    hv = host_volume(small_vol)
    ec = gpu_ctx.pw_candidates(hv, hv)
    report("small_can", util.ec_lines(ec), gold_lines("small", "can"))


def test_small_m4_matches_reference(gpu_ctx, small_vol):
    hv = host_volume(small_vol)
    m4 = gpu_ctx.pw_overlaps(hv, hv)
    report("small_m4", util.m4_lines(m4, gapped=True), gold_lines("small", "m4"))


def test_cfg0_can_matches_reference(gpu_ctx, cfg0_vol):
    hv = host_volume(cfg0_vol)
    ec = gpu_ctx.pw_candidates(hv, hv)
    report("cfg0_can", util.ec_lines(ec), gold_lines("cfg0", "can"))


def test_cfg0_m4_matches_reference(gpu_ctx, cfg0_vol):
    hv = host_volume(cfg0_vol)
    m4 = gpu_ctx.pw_overlaps(hv, hv)
    report("cfg0_m4", util.m4_lines(m4, gapped=True), gold_lines("cfg0", "m4"))


def test_off_diagonal_tile_matches_oracle(gpu_ctx, small_vol):
    n = small_vol.num_reads // 2
    seqs = [bytes(b"ACGT"[c] for c in small_vol.codes(i)) for i in range(small_vol.num_reads)]
    a = PackedVolume.from_seqs(seqs[:n], 0)
    b = PackedVolume.from_seqs(seqs[n:], n)
    for task in (0, 1):
        want = util.oracle_pw_tile(a, b, util.pw_params(task=task), threads=4)
        import mecat_b200
        if task == 0:
            got = gpu_ctx.pw_candidates(host_volume(a), host_volume(b), mecat_b200.pw_params(task=0))
            report("offdiag_can", util.ec_lines(got), util.ec_lines(want))
        else:
            got = gpu_ctx.pw_overlaps(host_volume(a), host_volume(b), mecat_b200.pw_params(task=1))
            report("offdiag_m4", util.m4_lines(got, True), util.m4_lines(want, True))


def test_tile_range_halves_equal_whole_tile(gpu_ctx, small_vol):
    """mecat_b200_pw_tile_range: two GPUs splitting a query volume produce, concatenated, exactly the
    whole-tile records (the multi-GPU schedule relies on it); also the device-resident constructor."""
    import mecat_b200
    import torch
    hv = host_volume(small_vol)
    d = gpu_ctx.upload(hv)
    idx = gpu_ctx.index_build(d)
    pac = torch.zeros((len(small_vol.pac) + 3) // 4 * 4, dtype=torch.uint8)
    pac.numpy()[:len(small_vol.pac)] = small_vol.pac
    pac_d = pac.cuda()
    torch.cuda.synchronize()
    d2 = gpu_ctx.volume_from_device(small_vol.num_reads, small_vol.num_bases, 0, small_vol.offset_size, pac_d.data_ptr())
    n = small_vol.num_reads
    for task in (0, 1):
        p = mecat_b200.pw_params(task=task)
        whole = gpu_ctx.pw_tile(idx, d, d, p)
        a = gpu_ctx.pw_tile_range(idx, d, d2, p, 0, n // 2)
        b = gpu_ctx.pw_tile_range(idx, d, d2, p, n // 2, n)
        assert len(a) and len(b)
        assert whole.tobytes() == a.tobytes() + b.tobytes()
    gpu_ctx.release_volume(d2)
    gpu_ctx.release_index(idx)
    gpu_ctx.release_volume(d)


@pytest.mark.parametrize("world", [2, 3])
def test_multi_gpu_schedule_reproduces_every_tile(gpu_ctx, small_vol, world):
    """The N-rank schedule (mecat_b200/multi.py) replayed rank by rank on one GPU: the union of all
    ranks' records equals the oracle's records of every (index volume, query volume) tile."""
    import mecat_b200
    from mecat_b200 import multi
    seqs = [bytes(b"ACGT"[c] for c in small_vol.codes(i)) for i in range(small_vol.num_reads)]
    n = small_vol.num_reads
    cuts = [n * i // world for i in range(world + 1)]
    vols = [PackedVolume.from_seqs(seqs[cuts[i]:cuts[i + 1]], cuts[i]) for i in range(world)]
    dv = [gpu_ctx.upload(host_volume(v)) for v in vols]
    idx = [gpu_ctx.index_build(d) for d in dv]
    reads = [v.num_reads for v in vols]
    for task in (0, 1):
        p = mecat_b200.pw_params(task=task)
        got = {}
        for rank in range(world):
            for step in range(world):
                for s, v, rb, re in multi.tile_work(world, rank, step, reads):
                    rec = gpu_ctx.pw_tile_range(idx[s], dv[s], dv[v], p, rb, re)
                    got.setdefault((s, v), []).append((rb, rec))
        for s in range(world):
            for v in range(s, world):
                want = util.oracle_pw_tile(vols[s], vols[v], util.pw_params(task=task), threads=4)
                parts = [r for _, r in sorted(got[(s, v)], key=lambda x: x[0])]
                mine = np.concatenate(parts) if parts else want[:0]
                assert mine.tobytes() == want.tobytes(), (task, s, v, len(mine), len(want))
    for i in idx:
        gpu_ctx.release_index(i)
    for d in dv:
        gpu_ctx.release_volume(d)


def _gold_fasta(name, tag):
    with gzip.open(os.path.join(util.GOLDEN, "%s.%s.fa.gz" % (name, tag)), "rt") as f:
        lines = f.read().splitlines()
    return sorted(zip(lines[0::2], lines[1::2]))


def _gold_can(name):
    import io
    import mecat_b200
    with gzip.open(os.path.join(util.GOLDEN, "%s.can.gz" % name), "rt") as f:
        return mecat_b200.read_can(io.StringIO(f.read()))


def _cns(gpu_ctx, vol, can, ratio, a, c, l):
    import mecat_b200
    d = gpu_ctx.upload(host_volume(vol))
    ec = mecat_b200.normalise_candidates(can, l)
    pieces = gpu_ctx.cns_reads(d, ec, ratio, a, c, l)
    gpu_ctx.release_volume(d)
    return sorted((">%d_%d_%d_%d" % (i, b, e, len(s)), s.decode()) for i, b, e, s in pieces)


def test_cns_small_matches_reference(gpu_ctx, small_vol):
    """mecat2cns -i 0 (rows C1-C7) against the corrected FASTA of the unmodified reference binary."""
    can = _gold_can("small")
    assert _cns(gpu_ctx, small_vol, can, 0.9, 2000, 6, 5000) == _gold_fasta("small", "cns_default")
    got = _cns(gpu_ctx, small_vol, can, 0.9, 1000, 4, 2000)
    want = _gold_fasta("small", "cns_relaxed")
    assert len(got) == len(want)
    assert got == want


def test_cns_cfg0_matches_reference(gpu_ctx, cfg0_vol):
    got = _cns(gpu_ctx, cfg0_vol, _gold_can("cfg0"), 0.9, 2000, 6, 5000)
    want = _gold_fasta("cfg0", "cns_default")
    assert len(got) == len(want)
    bad = [g[0] for g, w in zip(got, want) if g != w]
    assert not bad, bad[:5]


def test_cns_deep_coverage_matches_reference(gpu_ctx, tmp_path):
    """~120x coverage: every read has the full candidate list, so the accept loop reaches its 60-alignment cap and
    the 20x coverage gate, and the region graphs carry up to 60 paths (golden: unmodified mecat2cns -l 2000 -c 4 -a 1000)."""
    c = GOLD["deep"]
    fa = str(tmp_path / "deep.fa")
    util.gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"])
    assert hashlib.sha256(open(fa, "rb").read()).hexdigest() == c["fasta_sha256"]
    vol = PackedVolume.from_seqs(util.read_fasta(fa))
    got = _cns(gpu_ctx, vol, _gold_can("deep"), 0.9, 1000, 4, 2000)
    want = _gold_fasta("deep", "cns_relaxed")
    assert len(got) == len(want) == c["num_cns_relaxed"]
    bad = [g[0] for g, w in zip(got, want) if g != w]
    assert not bad, bad[:5]


def test_cns_small_batches_and_graph_waves(small_vol, monkeypatch):
    """The same output when the extension arena only holds a few reads per batch and the region graphs have to run in
    many scratch waves (both limits are normally sized from the device memory)."""
    import mecat_b200
    monkeypatch.setenv("MECAT_B200_ALIGN_ARENA_MB", "8")
    monkeypatch.setenv("MECAT_B200_POA_BUDGET_MB", "2")
    with mecat_b200.Context(0) as ctx:
        ctx.reset_stats()
        got = _cns(ctx, small_vol, _gold_can("small"), 0.9, 1000, 4, 2000)
        st = ctx.stats()
    assert got == _gold_fasta("small", "cns_relaxed")
    assert st["kernel_launches"]["extend"] >= 3        # several extension batches ...
    assert st["kernel_launches"]["cns_poa"] >= 3 * st["kernel_launches"]["cns_normvote"]   # ... and several graph waves in each


def _run_cli(exe, args, env=None):
    import subprocess
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([os.path.join(util.ROOT, "mecat_b200", "bin", exe)] + args, env=e, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    return p


def test_command_line_drivers_match_reference(gpu_ctx, tmp_path):
    """mecat2pw -j 0 | mecat2cns -i 0 through the C++ drivers (reference flags and file formats): candidate lines and
    corrected FASTA equal the unmodified reference binaries' output; with two devices the read-sharded run is identical."""
    import mecat_b200
    fa = str(tmp_path / "small.fa")
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f, open(fa, "wb") as g:
        g.write(f.read())
    can = str(tmp_path / "small.can")
    _run_cli("mecat2pw", ["-j", "0", "-d", fa, "-o", can, "-w", str(tmp_path / "wrk"), "-t", "2"])
    assert sorted(open(can).read().splitlines()) == gold_lines("small", "can")
    want = _gold_fasta("small", "cns_relaxed")

    def corrected(path):
        lines = open(path).read().splitlines()
        return sorted(zip(lines[0::2], lines[1::2]))

    out1 = str(tmp_path / "cns1.fa")
    _run_cli("mecat2cns", ["-i", "0", "-t", "2", "-l", "2000", "-c", "4", "-a", "1000", can, fa, out1])
    assert corrected(out1) == want
    if mecat_b200.load_library().mecat_b200_device_count() >= 2:
        out2 = str(tmp_path / "cns2.fa")
        _run_cli("mecat2cns", ["-i", "0", "-t", "2", "-l", "2000", "-c", "4", "-a", "1000", can, fa, out2], env={"MECAT_GPUS": "2"})
        assert open(out2).read() == open(out1).read()


def test_cns_on_several_volumes_matches_reference(gpu_ctx, tmp_path, monkeypatch):
    """The read set cut into 5+ volumes (as a read set beyond 2.14 Gbase is): mecat_b200_cns_reads_multi gathers the reads
    a run of templates needs into a working volume on the device; the corrected FASTA is the unmodified reference's, also
    when a small working-volume cap forces many runs, and through the command-line driver (one and two devices)."""
    import mecat_b200
    fa = str(tmp_path / "small.fa")
    with gzip.open(os.path.join(util.GOLDEN, "small.fa.gz"), "rb") as f, open(fa, "wb") as g:
        g.write(f.read())
    vols = mecat_b200.volumes_from_fasta(fa, max_volume_bases=320000)
    assert len(vols) >= 5 and sum(v.num_reads for v in vols) == 250
    dv = [gpu_ctx.upload(v) for v in vols]
    want = _gold_fasta("small", "cns_relaxed")
    ec = mecat_b200.normalise_candidates(_gold_can("small"), 2000)

    def run():
        pieces = gpu_ctx.cns_reads_multi(dv, ec, 0.9, 1000, 4, 2000)
        return sorted((">%d_%d_%d_%d" % (i, b, e, len(s)), s.decode()) for i, b, e, s in pieces)

    assert run() == want
    monkeypatch.setenv("MECAT_B200_CNS_WORK_BASES", "150000")          # ~25 reads per working volume: dozens of runs
    assert run() == want
    monkeypatch.delenv("MECAT_B200_CNS_WORK_BASES")
    for d in dv:
        gpu_ctx.release_volume(d)
    can = str(tmp_path / "small.can")
    with gzip.open(os.path.join(util.GOLDEN, "small.can.gz"), "rb") as f, open(can, "wb") as g:
        g.write(f.read())

    def corrected(path):
        lines = open(path).read().splitlines()
        return sorted(zip(lines[0::2], lines[1::2]))

    out1 = str(tmp_path / "cns1.fa")
    _run_cli("mecat2cns", ["-i", "0", "-l", "2000", "-c", "4", "-a", "1000", can, fa, out1], env={"MECAT_VOLUME_BASES": "320000"})
    assert corrected(out1) == want
    if mecat_b200.load_library().mecat_b200_device_count() >= 2:
        out2 = str(tmp_path / "cns2.fa")
        _run_cli("mecat2cns", ["-i", "0", "-l", "2000", "-c", "4", "-a", "1000", can, fa, out2], env={"MECAT_VOLUME_BASES": "320000", "MECAT_GPUS": "2"})
        assert open(out2).read() == open(out1).read()


def test_cns_degenerate_inputs(gpu_ctx, small_vol):
    """No candidates, and candidates that no read can use (coverage gate above every read's candidate count / no read
    long enough): an empty result, not an error."""
    import mecat_b200
    d = gpu_ctx.upload(host_volume(small_vol))
    can = _gold_can("small")
    empty = mecat_b200.normalise_candidates(can[:0], 2000)
    assert gpu_ctx.cns_reads(d, empty, 0.9, 1000, 4, 2000) == []
    ec = mecat_b200.normalise_candidates(can, 2000)
    assert gpu_ctx.cns_reads(d, ec, 0.9, 1000, 1000, 2000) == []          # -c 1000: no read has that many candidates
    assert gpu_ctx.cns_reads(d, ec, 0.9, 1000, 4, 10 ** 6) == []          # -l 1e6: no read is long enough
    assert gpu_ctx.cns_reads(d, ec, 0.9, 10 ** 6, 4, 2000) == []          # -a 1e6: no alignment is accepted
    gpu_ctx.release_volume(d)


def test_cns_consensus_runs_on_the_gpu(gpu_ctx, small_vol):
    """The consensus stages (accept, normalise/vote, segments, regions, graphs, assembly) are kernel launches."""
    gpu_ctx.reset_stats()
    _cns(gpu_ctx, small_vol, _gold_can("small"), 0.9, 1000, 4, 2000)
    st = gpu_ctx.stats()
    for k in ("cns_accept", "cns_normvote", "cns_segment", "cns_region", "cns_poa", "cns_assemble"):
        assert st["kernel_launches"][k] > 0, k


def test_candidate_cap_and_order(gpu_ctx, small_vol):
    """-n 3: per read the first 3 candidates of the -n 100 list, in the same order."""
    import mecat_b200
    hv = host_volume(small_vol)
    full = gpu_ctx.pw_candidates(hv, hv, mecat_b200.pw_params(task=0))
    cut = gpu_ctx.pw_candidates(hv, hv, mecat_b200.pw_params(task=0, num_candidates=3))
    want = util.oracle_pw_tile(small_vol, small_vol, util.pw_params(task=0, n=3), threads=4)
    assert [tuple(x) for x in cut.tolist()] == [tuple(x) for x in want.tolist()]   # same order, not only same set
    assert len(full) >= len(cut)


def test_empty_and_tiny_inputs(gpu_ctx):
    import mecat_b200
    tiny = PackedVolume.from_seqs([b"ACGTACGTAC", b"A", b"ACGTTGCATGCATGCATGCAAGCTTAGC" * 3])
    hv = host_volume(tiny)
    assert len(gpu_ctx.pw_candidates(hv, hv)) == 0
    assert len(gpu_ctx.pw_overlaps(hv, hv)) == 0
    with pytest.raises(mecat_b200.MecatB200Error):
        gpu_ctx.pw_candidates(hv, hv, mecat_b200.pw_params(task=0, num_candidates=0))
