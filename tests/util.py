"""Shared helpers for the test-suite: ctypes views of the oracle (oracle/liboracle.so), of the
unmodified reference (oracle/_ref/libmecatref.so, only when it has been built) and small
numpy utilities for packed volumes.  Test infrastructure only."""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")


class Volume(C.Structure):
    """orc_volume / mecat_volume: same layout (oracle.h, include/mecat_b200.h)."""
    _fields_ = [("num_reads", C.c_int32), ("num_bases", C.c_int32), ("start_read_id", C.c_int32),
                ("offset_size", C.POINTER(C.c_int32)), ("pac", C.POINTER(C.c_uint8))]


class PwParams(C.Structure):
    _fields_ = [("task", C.c_int32), ("num_candidates", C.c_int32), ("min_align_size", C.c_int32),
                ("min_kmer_match", C.c_int32), ("tech", C.c_int32)]


def pw_params(task=1, n=100, a=2000, k=4, x=0):
    return PwParams(task, n, a, k, x)


EC_DTYPE = np.dtype([(n, "<i4") for n in
                     ("qdir", "qid", "qext", "qsize", "qoff", "qend", "sdir", "sid", "sext", "ssize", "soff",
                      "send", "score")])
M4_DTYPE = np.dtype([("qid", "<i8"), ("sid", "<i8"), ("ident", "<f8"), ("vscore", "<i4"), ("qdir", "<i4"),
                     ("qoff", "<i8"), ("qend", "<i8"), ("qsize", "<i8"), ("sdir", "<i4"), ("pad", "<i4"),
                     ("soff", "<i8"), ("send", "<i8"), ("ssize", "<i8"), ("qext", "<i8"), ("sext", "<i8")])
assert EC_DTYPE.itemsize == 52 and M4_DTYPE.itemsize == 104


class PackedVolume:
    """A 2-bit packed volume held in numpy arrays, in the on-disk layout of `wrk/volN`
    (reference split_database.cpp:136-153)."""

    def __init__(self, offset_size, pac, num_bases, start_read_id=0):
        self.offset_size = np.ascontiguousarray(offset_size, dtype=np.int32).reshape(-1, 2)
        self.pac = np.ascontiguousarray(pac, dtype=np.uint8)
        self.num_reads = self.offset_size.shape[0]
        self.num_bases = int(num_bases)
        self.start_read_id = int(start_read_id)

    def c(self):
        return Volume(self.num_reads, self.num_bases, self.start_read_id,
                      self.offset_size.ctypes.data_as(C.POINTER(C.c_int32)),
                      self.pac.ctypes.data_as(C.POINTER(C.c_uint8)))

    @staticmethod
    def from_seqs(seqs, start_read_id=0):
        """seqs: list of ASCII strings/bytes of pure ACGT.  One pad base after every read."""
        lut = np.zeros(256, dtype=np.uint8)
        for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
            lut[ch] = v
        total = sum(len(s) + 1 for s in seqs)
        codes = np.zeros(((total + 3) // 4) * 4, dtype=np.uint8)
        os_ = np.zeros((len(seqs), 2), dtype=np.int32)
        cur = 0
        for i, s in enumerate(seqs):
            b = s.encode() if isinstance(s, str) else bytes(s)
            a = np.frombuffer(b, dtype=np.uint8)
            codes[cur:cur + len(a)] = lut[a]
            os_[i] = (cur, len(a))
            cur += len(a) + 1
        q = codes.reshape(-1, 4)
        pac = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
        return PackedVolume(os_, pac.astype(np.uint8), total, start_read_id)

    @staticmethod
    def load(path):
        """Read a reference `volN` file."""
        with open(path, "rb") as f:
            hdr = np.frombuffer(f.read(12), dtype="<i4")
            n, nb, sid = int(hdr[0]), int(hdr[1]), int(hdr[2])
            os_ = np.frombuffer(f.read(8 * n), dtype="<i4").reshape(-1, 2).copy()
            pac = np.frombuffer(f.read((nb + 3) // 4), dtype=np.uint8).copy()
        return PackedVolume(os_, pac, nb, sid)

    def codes(self, rid, strand=0):
        """Unpacked base codes (0..3) of one read; strand 1 = reverse complement."""
        off, sz = self.offset_size[rid]
        idx = np.arange(off, off + sz, dtype=np.int64)
        c = (self.pac[idx >> 2] >> (((~idx) & 3) << 1)) & 3
        if strand:
            c = (3 - c)[::-1]
        return np.ascontiguousarray(c, dtype=np.int8)


def read_fasta(path):
    seqs, cur = [], []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if cur:
                    seqs.append(b"".join(cur))
                    cur = []
            else:
                cur.append(line.strip())
    if cur:
        seqs.append(b"".join(cur))
    return seqs


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


_harness = None


def cns_harness():
    """Host build of the product's consensus stage sequence and kernel bodies (tests/cns_host_harness.cpp)."""
    global _harness
    if _harness is not None:
        return _harness
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcns_harness.so")
    src = [os.path.join(ROOT, "tests", "cns_host_harness.cpp"), os.path.join(ROOT, "mecat_b200", "csrc", "cns_pipeline.h"),
           os.path.join(ROOT, "mecat_b200", "csrc", "cns_core.cuh"), os.path.join(ROOT, "include", "mecat_b200.h"),
           os.path.join(ROOT, "tests", "cns_literal.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src[0]])
    L = C.CDLL(so)
    vp = C.c_void_p
    L.harness_cns_batch.restype = C.c_int
    L.harness_cns_batch.argtypes = [C.c_int, vp, vp, vp, C.c_char_p, C.c_char_p, vp, C.POINTER(vp), C.POINTER(C.c_size_t),
                                    C.POINTER(vp), C.POINTER(C.c_size_t), C.c_char_p, C.c_int]
    L.harness_free.argtypes = [vp]
    L.harness_effective_ranges.restype = C.c_int
    L.harness_effective_ranges.argtypes = [vp, C.c_int, C.c_int, C.c_longlong, vp]
    L.harness_normalize_and_vote.restype = C.c_int
    L.harness_normalize_and_vote.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, vp, C.c_char_p, C.c_char_p]
    L.harness_poa_consensus.restype = C.c_int
    L.harness_poa_consensus.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.harness_anchor_compare.restype = C.c_int
    L.harness_anchor_compare.argtypes = [vp, C.c_int, C.c_int]
    L.harness_segments_compare.restype = C.c_int
    L.harness_segments_compare.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double]
    L.harness_normalize_compare.restype = C.c_int
    L.harness_normalize_compare.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    _harness = L
    return L


_m4_harness = None


def m4_harness():
    """Host build of the product's record assembly / text bodies (tests/m4_host_harness.cpp over csrc/m4_core.cuh)."""
    global _m4_harness
    if _m4_harness is not None:
        return _m4_harness
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libm4_harness.so")
    src = [os.path.join(ROOT, "tests", "m4_host_harness.cpp"), os.path.join(ROOT, "mecat_b200", "csrc", "m4_core.cuh"),
           os.path.join(ROOT, "mecat_b200", "csrc", "host", "format.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", so, src[0]])
    L = C.CDLL(so)
    L.mh_sort_random.restype = C.c_long
    L.mh_sort_random.argtypes = [C.c_int, C.c_int, C.c_uint]
    L.mh_sort_adversary.restype = C.c_long
    L.mh_sort_adversary.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.mh_fmt_ratios.restype = C.c_long
    L.mh_fmt_ratios.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.mh_fmt_values.restype = C.c_long
    L.mh_fmt_values.argtypes = [C.c_void_p, C.c_long, C.c_char_p, C.c_int]
    L.mh_lines.restype = C.c_long
    L.mh_lines.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long]
    _m4_harness = L
    return L


_xdrop_harness = None


def xdrop_harness():
    """Host build of the product's nanopore extension body (tests/xdrop_host_harness.cpp over csrc/xdrop_core.cuh)."""
    global _xdrop_harness
    if _xdrop_harness is not None:
        return _xdrop_harness
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libxdrop_harness.so")
    src = [os.path.join(ROOT, "tests", "xdrop_host_harness.cpp"), os.path.join(ROOT, "mecat_b200", "csrc", "xdrop_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", so, src[0]])
    L = C.CDLL(so)
    L.xh_go.restype = C.c_int
    L.xh_go.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32),
                        C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    _xdrop_harness = L
    return L


_ref_harness = None


def ref_harness():
    """Host build of the product's mecat2ref stage sequence, kernel bodies and host I/O (tests/ref_host_harness.cpp); the
    gapped extension inside it is the oracle's."""
    global _ref_harness
    if _ref_harness is not None:
        return _ref_harness
    build_oracle()
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libref_harness.so")
    src = [os.path.join(ROOT, "tests", "ref_host_harness.cpp"), os.path.join(ROOT, "mecat_b200", "csrc", "ref_pipeline.h"),
           os.path.join(ROOT, "mecat_b200", "csrc", "ref_core.cuh"), os.path.join(ROOT, "mecat_b200", "csrc", "host", "refio.h"),
           os.path.join(ROOT, "include", "mecat_b200.h"), os.path.join(ORACLE_DIR, "liboracle.so")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", so, src[0], "-L", ORACLE_DIR, "-loracle",
                               "-Wl,-rpath," + ORACLE_DIR])
    L = C.CDLL(so)
    vp = C.c_void_p
    L.harness_ref_map.restype = C.c_int
    L.harness_ref_map.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, C.POINTER(vp), C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_long), C.c_char_p, C.c_int]
    L.harness_set_tech.argtypes = [C.c_int]
    L.harness_ref_index_build.restype = vp
    L.harness_ref_index_build.argtypes = [vp]
    L.harness_ref_index_release.argtypes = [vp]
    L.harness_ref_index_export.restype = C.c_int64
    L.harness_ref_index_export.argtypes = [vp, vp, vp]
    L.harness_ref_raw_candidates.restype = C.c_int
    L.harness_ref_raw_candidates.argtypes = [vp, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.harness_ref_map_packed.restype = C.c_int
    L.harness_ref_map_packed.argtypes = [vp, vp, vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.harness_ddf_forms.restype = C.c_int
    L.harness_ddf_forms.argtypes = [C.c_int, C.c_int, C.c_int]
    L.harness_ddf_sweep.restype = C.c_long
    L.harness_ddf_sweep.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.harness_free.argtypes = [vp]
    _ref_harness = L
    return L


def gen_reads(path, n, genome_len, seed, mean=15000, sd=1500, err=0.15, genome_out=None):
    exe = os.path.join(ROOT, "mecat_b200", "bin", "gen_reads")
    if not os.path.exists(exe):
        sys.path.insert(0, ROOT)
        from mecat_b200 import build as _b
        _b.build()
    cmd = [exe, path, str(n), str(genome_len), str(seed), str(mean), str(sd),
           str(err)]
    if genome_out:
        cmd.append(genome_out)
    subprocess.check_call(cmd)


_oracle = None


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    build_oracle()
    L = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
    VP, PP = C.POINTER(Volume), C.POINTER(PwParams)
    i32p, i16p = C.POINTER(C.c_int32), C.POINTER(C.c_int16)
    L.orc_index_build.restype = C.c_void_p
    L.orc_index_build.argtypes = [VP]
    L.orc_index_free.argtypes = [C.c_void_p]
    L.orc_index_lookup.restype = C.c_int
    L.orc_index_lookup.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(i32p)]
    L.orc_index_num_kmers.restype = C.c_int64
    L.orc_index_num_kmers.argtypes = [C.c_void_p]
    L.orc_seeding.restype = C.c_int
    L.orc_seeding.argtypes = [C.c_void_p, VP, VP, C.c_int, C.c_int, i32p, i16p, i16p, C.c_int]
    L.orc_insert_loc.argtypes = [i16p, i16p, i16p, C.c_int, C.c_int]
    L.orc_find_location.restype = C.c_int
    L.orc_find_location.argtypes = [i32p, i32p, i32p, i32p, C.c_int, i32p, C.c_int]
    L.orc_pw_candidates.restype = C.c_int
    L.orc_pw_candidates.argtypes = [C.c_void_p, VP, VP, C.c_int, PP, i32p]
    L.orc_diff_go.restype = C.c_int
    L.orc_diff_go.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, i32p,
                              C.POINTER(C.c_double), C.c_char_p, C.c_char_p, C.c_int]
    L.orc_xdrop_go.restype = C.c_int
    L.orc_xdrop_go.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, i32p,
                              C.POINTER(C.c_double), C.c_char_p, C.c_char_p, C.c_int]
    L.orc_diff_align_block.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, i32p]
    L.orc_pw_tile.restype = C.c_int
    L.orc_pw_tile.argtypes = [VP, VP, PP, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_cns_sort_candidates.argtypes = [C.c_void_p, C.c_int]
    L.orc_cns_m4_order.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_ref_map.restype = C.c_int
    L.orc_ref_map.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.orc_ref_map_x.restype = C.c_int
    L.orc_ref_map_x.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.orc_poa_consensus.restype = C.c_int
    L.orc_poa_consensus.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    L.orc_cns_consensus.restype = C.c_int
    L.orc_cns_consensus.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.orc_cns_get_alignment.restype = C.c_int
    L.orc_cns_get_alignment.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_double,
                                        C.c_int, i32p, C.c_char_p, C.c_char_p, C.c_int]
    L.orc_normalize_gaps.restype = C.c_int
    L.orc_normalize_gaps.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    _oracle = L
    return L


def oracle_pw_tile(ref, reads, params, threads=8):
    L = oracle()
    out, n = C.c_void_p(), C.c_size_t()
    rv, qv = ref.c(), reads.c()
    rc = L.orc_pw_tile(C.byref(rv), C.byref(qv), C.byref(params), threads, C.byref(out), C.byref(n))
    assert rc == 0
    dt = EC_DTYPE if params.task == 0 else M4_DTYPE
    arr = np.frombuffer(C.string_at(out.value, n.value * dt.itemsize), dtype=dt).copy()
    L.orc_free(out)
    return arr


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libmecatref.so"))


_ref = None


def ref():
    """ctypes handle on the unmodified reference (function-level shim)."""
    global _ref
    if _ref is not None:
        return _ref
    L = C.CDLL(os.path.join(REF_DIR, "libmecatref.so"))
    i32p, i16p = C.POINTER(C.c_int32), C.POINTER(C.c_int16)
    L.ref_volume_new.restype = C.c_void_p
    L.ref_volume_new.argtypes = [C.c_int]
    L.ref_volume_add.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.ref_volume_set_start_id.argtypes = [C.c_void_p, C.c_int]
    L.ref_volume_load.restype = C.c_void_p
    L.ref_volume_load.argtypes = [C.c_char_p]
    L.ref_volume_dump.argtypes = [C.c_void_p, C.c_char_p]
    L.ref_volume_free.argtypes = [C.c_void_p]
    for f in ("ref_volume_num_reads", "ref_volume_num_bases", "ref_volume_start_id"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.ref_volume_pac.restype = C.POINTER(C.c_uint8)
    L.ref_volume_pac.argtypes = [C.c_void_p]
    L.ref_volume_offsets.restype = i32p
    L.ref_volume_offsets.argtypes = [C.c_void_p]
    L.ref_read_id_from_offset.argtypes = [C.c_void_p, C.c_int]
    L.ref_index_create.restype = C.c_void_p
    L.ref_index_create.argtypes = [C.c_void_p, C.c_int]
    L.ref_index_free.argtypes = [C.c_void_p]
    L.ref_index_count.argtypes = [C.c_void_p, C.c_uint32]
    L.ref_index_list.restype = i32p
    L.ref_index_list.argtypes = [C.c_void_p, C.c_uint32]
    L.ref_pw_set_options.argtypes = [C.c_int] * 4
    L.ref_insert_loc.argtypes = [i16p, C.c_int, C.c_int]
    L.ref_find_location.argtypes = [i32p, i32p, i32p, i32p, C.c_int, i32p, C.c_int]
    L.ref_pw_ctx_new.restype = C.c_void_p
    L.ref_pw_ctx_new.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_pw_ctx_free.argtypes = [C.c_void_p]
    L.ref_pw_seeding_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, i32p, i16p, i16p, C.c_int]
    L.ref_pw_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, i32p]
    L.ref_diff_new.restype = C.c_void_p
    L.ref_diff_free.argtypes = [C.c_void_p]
    L.ref_diff_go.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, i32p,
                              C.POINTER(C.c_double), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    L.ref_diff_align_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, i32p]
    L.ref_xdrop_new.restype = C.c_void_p
    L.ref_xdrop_free.argtypes = [C.c_void_p]
    L.ref_xdrop_go.argtypes = L.ref_diff_go.argtypes
    L.ref_effective_ranges.restype = C.c_int
    L.ref_effective_ranges.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.ref_normalize_and_vote.restype = C.c_int
    L.ref_normalize_and_vote.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    L.ref_poa_consensus.restype = C.c_int
    L.ref_poa_consensus.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    L.ref_cns_drd_new.restype = C.c_void_p
    L.ref_cns_drd_free.argtypes = [C.c_void_p]
    L.ref_cns_get_alignment.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_double, C.c_int, i32p, C.c_char_p, C.c_char_p, C.c_int]
    L.ref_normalize_gaps.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    _ref = L
    return L


def ec_lines(ec):
    """ExtensionCandidate records -> sorted `.can` lines (alignment.cpp:18-32)."""
    out = ["%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d" % (e["qid"], e["sid"], e["qdir"], e["sdir"], e["qext"], e["sext"],
                                                   e["score"], e["qsize"], e["ssize"]) for e in ec]
    return sorted(out)


def fmt_g6(x):
    """default ostream << double : %g with 6 significant digits."""
    return "%g" % x


def m4_lines(m4, gapped=False):
    """M4 records -> sorted `.m4` lines (pw_impl.cpp:509-531)."""
    out = []
    for m in m4:
        s = "%d\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d" % (
            m["qid"], m["sid"], fmt_g6(m["ident"]), m["vscore"], m["qdir"], m["qoff"], m["qend"], m["qsize"],
            m["sdir"], m["soff"], m["send"], m["ssize"])
        if gapped:
            s += "\t%d\t%d" % (m["qext"], m["sext"])
        out.append(s)
    return sorted(out)


def repeat_reads(seed=21, unit=4000, copies=10, n_reads=260, mean=5000, err=0.05, div=0.01):
    """Repeat-rich synthetic reads: a genome made of `copies` diverged copies of one random unit plus
    unique flanks.  Exact 13-mers then occur 30-200 times in the volume, which exercises what uniform
    genomes never do: long index lists, the >128 cutoff, 40-seed bucket overflow (insert_loc) far
    from self hits, ties in DDF scoring and full candidate lists."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=unit)
    parts = [rng.integers(0, 4, size=3000)]
    for _ in range(copies):
        u = base.copy()
        m = rng.random(unit) < div
        u[m] = rng.integers(0, 4, size=int(m.sum()))
        parts.append(u)
        parts.append(rng.integers(0, 4, size=int(rng.integers(200, 1500))))
    genome = np.concatenate(parts)
    G = len(genome)
    reads = []
    for _ in range(n_reads):
        L = int(max(2500, rng.normal(mean, 800)))
        L = min(L, G)
        s = int(rng.integers(0, G - L + 1))
        t = genome[s:s + L]
        u = rng.random(L)
        keep = u >= err * 0.3
        sub = (u >= err * 0.3) & (u < err * 0.4)
        t = t.copy()
        t[sub] = (t[sub] + 1 + rng.integers(0, 3, size=int(sub.sum()))) & 3
        t = t[keep]
        ins = rng.random(len(t)) < err * 0.6
        out = np.empty(len(t) + int(ins.sum()), dtype=np.int64)
        idx = np.arange(len(t)) + np.cumsum(ins) - ins
        out[:] = -1
        out[idx] = t
        gaps = out < 0
        out[gaps] = rng.integers(0, 4, size=int(gaps.sum()))
        if rng.random() < 0.5:
            out = (3 - out)[::-1]
        reads.append(bytes(b"ACGT"[int(c)] for c in out))
    return reads


def stale(target, sources):
    return not os.path.exists(target) or any(os.path.getmtime(f) > os.path.getmtime(target) for f in sources)


REF_HOST_SOURCES = [os.path.join(ROOT, "tests", "ref_host_harness.cpp"), os.path.join(ROOT, "tests", "ref_abi_shim.cpp"),
                    os.path.join(ROOT, "tests", "ref_sanitize_main.cpp"), os.path.join(ROOT, "mecat_b200", "csrc", "ref_pipeline.h"),
                    os.path.join(ROOT, "mecat_b200", "csrc", "ref_core.cuh"), os.path.join(ROOT, "mecat_b200", "csrc", "host", "refio.h"),
                    os.path.join(ROOT, "mecat_b200", "csrc", "host", "mecat2ref.cpp"), os.path.join(ROOT, "include", "mecat_b200.h"),
                    os.path.join(ORACLE_DIR, "liboracle.so")]


def ref_driver_on_host():
    """mecat_b200/csrc/host/mecat2ref.cpp linked against tests/ref_abi_shim.cpp + the host harness instead of the product
    library: the command-line driver as the CPU suite can run it.  Returns the path of the executable."""
    ref_harness()
    out_dir = os.path.join(ROOT, "tests", "_build")
    exe = os.path.join(out_dir, "mecat2ref_host")
    src = [os.path.join(ROOT, "mecat_b200", "csrc", "host", "mecat2ref.cpp"), os.path.join(ROOT, "tests", "ref_abi_shim.cpp"),
           os.path.join(ROOT, "mecat_b200", "csrc", "host", "refio.h"), os.path.join(ROOT, "include", "mecat_b200.h"),
           os.path.join(out_dir, "libref_harness.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(f) > os.path.getmtime(exe) for f in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe, src[0], src[1],
                               "-L", out_dir, "-lref_harness", "-Wl,-rpath," + out_dir, "-Wl,-rpath," + ORACLE_DIR])
    return exe


def pw_driver_on_host():
    """mecat_b200/csrc/host/mecat2pw.cpp + the product's host I/O linked against tests/pw_abi_shim.cpp (the oracle plays the
    device): the command-line driver as the CPU suite can run it.  Returns the path of the executable."""
    build_oracle()
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "mecat2pw_host")
    src = [os.path.join(ROOT, "mecat_b200", "csrc", "host", "mecat2pw.cpp"), os.path.join(ROOT, "mecat_b200", "csrc", "host_io.cpp"),
           os.path.join(ROOT, "tests", "pw_abi_shim.cpp"), os.path.join(ROOT, "mecat_b200", "csrc", "host", "format.h"),
           os.path.join(ROOT, "include", "mecat_b200.h"), os.path.join(ORACLE_DIR, "liboracle.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(f) > os.path.getmtime(exe) for f in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe] + src[:3] +
                              ["-L", ORACLE_DIR, "-loracle", "-Wl,-rpath," + ORACLE_DIR])
    return exe


def make_refmap_repeats(reads_path, genome_path, seed, num_reads, copies=40):
    """Repeat-rich inputs for mecat2ref: a 6 kb segment copied `copies` times (3 % diverged) between short unique stretches,
    a 150-copy tandem 40-mer, a 3-mer run, a second contig with a run of N; reads of 3-45 kb with 10-30 % errors, half of
    them reverse strand, 15 % chimeric.  Candidate lists overflow -n, neighbour votes consume blocks, rescue hops between
    copies.  (Reads near the reference's 100 000-base buffers make the unmodified binary corrupt its heap, so none here.)"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return acgt[rng.integers(0, 4, n)]

    parts, seg = [], rnd(6000)
    for _ in range(copies):
        parts.append(rnd(int(rng.integers(2000, 9000))))
        s = seg.copy()
        m = rng.random(len(s)) < 0.03
        s[m] = rnd(int(m.sum()))
        parts.append(s)
    parts += [np.tile(rnd(40), 150), rnd(30000), np.tile(rnd(3), 400), rnd(50000)]
    g1, g2 = np.concatenate(parts), rnd(80000)
    with open(genome_path, "wb") as f:
        f.write(b">c1\n" + g1.tobytes() + b"\n>c2 x\n" + g2.tobytes()[:40000] + b"NNNNNNNNNN" + g2.tobytes()[40000:] + b"\n")
    G = np.concatenate([g1, g2])
    comp = np.zeros(256, np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")

    def mutate(s, err):
        out, r = [], rng.random(len(s))
        for b, x in zip(s, r):
            if x < 0.3 * err:
                continue
            out.append(acgt[rng.integers(0, 4)] if x < 0.4 * err else b)
            while rng.random() < 0.6 * err:
                out.append(acgt[rng.integers(0, 4)])
        return np.array(out, dtype=np.uint8)

    with open(reads_path, "wb") as f:
        for i in range(num_reads):
            n = min(int(rng.choice([3000, 8000, 15000, 30000, 45000], p=[.2, .3, .3, .15, .05])), len(G) - 1)
            st = int(rng.integers(0, len(G) - n))
            s = mutate(G[st:st + n], float(rng.choice([0.1, 0.15, 0.22, 0.3])))
            if rng.random() < 0.5:
                s = comp[s[::-1]]
            if rng.random() < 0.15:
                st2 = int(rng.integers(0, len(G) - 5000))
                s = np.concatenate([s, mutate(G[st2:st2 + 5000], 0.12)])
            f.write(b">%d\n" % i + s[:99000].tobytes() + b"\n")


def make_refmap_hard(reads_path, genome_path, seed=19):
    """Deterministic inputs that push mecat2ref off its main path: three contigs (one holding a 3 kb repeat of another),
    noisy reads (15 % errors), chimeric reads glued from two places (clipped alignments -> rescue_clipped_align), reads with
    a stretch of N, short reads, very noisy reads (second, more sensitive pass) and lower-case letters."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rand_seq(n):
        return acgt[rng.integers(0, 4, size=n)].tobytes().decode()

    contigs = [rand_seq(60000), rand_seq(35000), rand_seq(20000)]
    contigs[1] = contigs[1][:10000] + contigs[0][20000:23000] + contigs[1][13000:]
    contigs[2] = contigs[2][:5000] + "N" * 300 + contigs[2][5300:]
    with open(genome_path, "w") as f:
        for i, c in enumerate(contigs):
            f.write(">chr%d some description\n" % (i + 1))
            for k in range(0, len(c), 70):
                f.write(c[k:k + 70] + "\n")

    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}

    def noisy(s, err):
        out = []
        for ch in s:
            r = rng.random()
            if r < err * 0.3:
                continue
            if r < err * 0.4:
                out.append("ACGT"[int(rng.integers(4))])
                continue
            out.append(ch)
            if rng.random() < err * 0.6:
                out.append("ACGT"[int(rng.integers(4))])
        return "".join(out)

    def piece(length):
        c = contigs[int(rng.integers(len(contigs)))]
        length = min(length, len(c) - 1)
        b = int(rng.integers(0, len(c) - length))
        s = c[b:b + length]
        if rng.random() < 0.5:
            s = "".join(comp[x] for x in reversed(s))
        return s

    reads = []
    for i in range(160):
        kind = i % 8
        if kind < 4:
            reads.append(noisy(piece(int(rng.integers(3000, 12000))), 0.15))
        elif kind == 4:
            reads.append(noisy(piece(int(rng.integers(4000, 7000))), 0.15) + noisy(piece(int(rng.integers(4000, 7000))), 0.15))
        elif kind == 5:
            reads.append(noisy(piece(int(rng.integers(2500, 5000))), 0.28))
        elif kind == 6:
            s = noisy(piece(int(rng.integers(5000, 9000))), 0.12)
            reads.append(s[:2000] + "N" * 50 + s[2050:])
        else:
            s = noisy(piece(int(rng.integers(1200, 2500))), 0.1)
            reads.append(s[:600] + s[600:900].lower() + s[900:])
    with open(reads_path, "w") as f:
        for i, r in enumerate(reads):
            f.write(">read_%d\n%s\n" % (i, r))


# ---------------------------------------------------------------- mecat2asmpw / mecat2trimpw fixtures
ASM_CASES = {
    # corrected-read like: ~16x of a 150 kb genome at 1.5 % error, two files of 300 reads
    "asm": dict(n=600, genome=150000, seed=77, mean=4000, sd=800, err=0.015, files=2),
    # ~56x of a 25 kb genome: more candidates per read than the *50 programs keep, blocks whose score passes SM = 60
    # (the neighbour votes then read beyond a block's 60 entries), a few N letters, lower-case reads, two stubs
    "asmdeep": dict(n=400, genome=25000, seed=91, mean=3500, sd=900, err=0.01, files=1),
    # three chunks of PLL = 500 reads at ~130x: the binary's output depends on its thread count here (only digests are kept)
    "asmsched": dict(n=1500, genome=40000, seed=123, mean=3500, sd=900, err=0.01, files=1),
    # awkward reads among ordinary ones: a duplicate, tandem repeats, poly-A, N runs, other letters, reads of 1 / 13 / 14 letters
    "asmodd": dict(n=200, genome=30000, seed=17, mean=3000, sd=600, err=0.02, files=1),
}


def asm_reads(name, tmp_dir):
    """The reads of an ASM_CASES fixture, in file order, as written to the FASTA files (case kept, N letters in)."""
    c = ASM_CASES[name]
    fa = os.path.join(tmp_dir, name + ".all.fa")
    gen_reads(fa, c["n"], c["genome"], c["seed"], c["mean"], c["sd"], err=c["err"])
    seqs = [s for _, s in read_fasta_raw(fa)]
    if name == "asmodd":
        unit = seqs[3][100:137]
        seqs[11] = seqs[10]                                   # the same read twice
        seqs[20] = unit * 100                                 # tandem repeat: every k-mer of the unit ~180 times in the file
        seqs[21] = unit * 80
        seqs[30] = "A" * 2000                                 # one k-mer 1 988 times: its list is dropped (> 256)
        seqs[31] = "ACGT" * 600
        seqs[40] = "".join("N" if i % 50 == 49 else c for i, c in enumerate(seqs[40]))
        seqs[41] = "N" * 2000
        seqs[50], seqs[51], seqs[52] = seqs[50][:13], seqs[51][:14], seqs[52][:1]
        seqs[60] = "".join(c.lower() if i % 3 else c for i, c in enumerate(seqs[60]))
        seqs[61] = "".join("RYKM"[i % 4] if i % 97 == 5 else c for i, c in enumerate(seqs[61]))
        seqs[70] = seqs[10][500:2500]                         # contained in reads 10 and 11
    if name == "asmdeep":
        for r in range(len(seqs)):
            s = seqs[r]
            if r % 37 == 5:
                s = list(s)
                for t in range(3):
                    s[(r * 7919 + t * 104729) % len(s)] = "N"
                s = "".join(s)
            if r % 41 == 7:
                s = s.lower()
            if r == 100:
                s = s[:300]
            if r == 200:
                s = s[:12]
            seqs[r] = s
    return seqs


def read_fasta_raw(path):
    out, name = [], None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            name = line[1:]
        elif line:
            out.append((name, line))
    return out


def asm_workdir(name, wrk):
    """The working directory mecat2canu hands the overlapper: NNNNNN.fasta files and `ovlprep` with the read ranges
    (mecat2asmpw.c:1083-1090 parses `-allreads -allbases -b <first> -e <last>`).  Returns [(first_id, [reads])] per file."""
    c = ASM_CASES[name]
    os.makedirs(wrk, exist_ok=True)
    seqs = asm_reads(name, wrk)
    per = len(seqs) // c["files"]
    files = []
    with open(os.path.join(wrk, "ovlprep"), "w") as prep:
        for i in range(c["files"]):
            part = seqs[i * per:(i + 1) * per]
            with open(os.path.join(wrk, "%06d.fasta" % (i + 1)), "w") as f:
                for j, s in enumerate(part):
                    f.write(">%d\n%s\n" % (i * per + j, s))
            prep.write("-allreads -allbases -b %d -e %d\n" % (i * per + 1, (i + 1) * per))
            files.append((i * per + 1, part))
    return files


def asm_text(reads):
    """(text, starts, lengths) of a file's reads the way load_read keeps them: upper-cased, a NUL behind each."""
    up = [s.upper() for s in reads]
    starts, o = [], 0
    for s in up:
        starts.append(o)
        o += len(s) + 1
    return ("\0".join(up) + "\0").encode(), np.array(starts, dtype=np.int32), np.array([len(s) for s in up], dtype=np.int32)


ASM_OVERLAP_DTYPE = np.dtype([("sread", "<i4"), ("qread", "<i4"), ("score", "<f4"), ("sbeg", "<i4"), ("send", "<i4"), ("slen", "<i4"),
                              ("strand", "<i4"), ("qbeg", "<i4"), ("qend", "<i4"), ("qlen", "<i4")])


def asm_lines(recs):
    """The line the reference prints per overlap (mecat2asmpw.c:944-945)."""
    return ["%d %d %.3f 100 0 %d %d %d %d %d %d %d" % (r["sread"], r["qread"], float(r["score"]), r["sbeg"], r["send"], r["slen"], r["strand"],
                                                      r["qbeg"], r["qend"], r["qlen"]) for r in recs]


def asm_oracle_overlaps(sub, sub_first, qry, qry_first, variant=0, maxc=100, history=0):
    """oracle/oracle_asmpw.cpp: the reads of one query file against the index of one subject file."""
    L = oracle()
    L.orc_asm_overlaps.restype = C.c_int
    L.orc_asm_overlaps.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    t, st, ln = asm_text(sub)
    qt, qs, ql = asm_text(qry)
    out, n = C.c_void_p(), C.c_size_t()
    rc = L.orc_asm_overlaps(t, len(t), st.ctypes.data, ln.ctypes.data, len(sub), sub_first, qt, qs.ctypes.data, ql.ctypes.data, len(qry),
                            qry_first, variant, maxc, history, C.byref(out), C.byref(n))
    assert rc == 0
    arr = np.frombuffer(C.string_at(out.value, n.value * ASM_OVERLAP_DTYPE.itemsize), dtype=ASM_OVERLAP_DTYPE).copy()
    L.orc_free(out)
    return arr


_asm_harness = None


def asm_harness():
    """Host build of the product's mecat2asmpw stage sequence and kernel bodies (tests/asm_host_harness.cpp)."""
    global _asm_harness
    if _asm_harness is not None:
        return _asm_harness
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libasm_harness.so")
    src = [os.path.join(ROOT, "tests", "asm_host_harness.cpp"), os.path.join(ROOT, "mecat_b200", "csrc", "asm_pipeline.h"),
           os.path.join(ROOT, "mecat_b200", "csrc", "asm_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", so, src[0]])
    L = C.CDLL(so)
    L.ah_overlaps.restype = C.c_int
    L.ah_overlaps.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p,
                              C.c_int32, C.c_int32, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                              C.c_void_p, C.c_char_p, C.c_int]
    L.ah_free.argtypes = [C.c_void_p]
    L.ah_index.restype = C.c_int64
    L.ah_index.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
    L.ah_ddf_close.restype = C.c_int
    L.ah_ddf_close.argtypes = [C.c_int, C.c_int, C.c_int]
    _asm_harness = L
    return L


def asm_harness_overlaps(sub, sub_first, qry, qry_first, variant=0, maxc=100, budget=0, divisor=0):
    """The product's stage sequence on the host.  Returns (records, [seed batches, candidates, hits, extension passes])."""
    H = asm_harness()
    t, st, ln = asm_text(sub)
    qt, qs, ql = asm_text(qry)
    out, n, err, stats = C.c_void_p(), C.c_size_t(), C.create_string_buffer(256), (C.c_int64 * 4)()
    rc = H.ah_overlaps(t, len(t), st.ctypes.data, ln.ctypes.data, len(sub), sub_first, qt, len(qt), qs.ctypes.data, ql.ctypes.data, len(qry),
                       qry_first, variant, maxc, budget, divisor, C.byref(out), C.byref(n), stats, err, 256)
    assert rc == 0, err.value
    arr = np.frombuffer(C.string_at(out.value, n.value * ASM_OVERLAP_DTYPE.itemsize), dtype=ASM_OVERLAP_DTYPE).copy()
    H.ah_free(out)
    return arr, list(stats)


def asm_driver_on_host():
    """mecat_b200/csrc/host/mecat2asmpw.cpp linked against tests/asm_abi_shim.cpp + the host harness instead of the product
    library.  Returns the directory holding the four program names."""
    out_dir = os.path.join(ROOT, "tests", "_build", "asm_driver")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "mecat2asmpw")
    src = [os.path.join(ROOT, "mecat_b200", "csrc", "host", "mecat2asmpw.cpp"), os.path.join(ROOT, "tests", "asm_abi_shim.cpp"),
           os.path.join(ROOT, "tests", "asm_host_harness.cpp")]
    deps = src + [os.path.join(ROOT, "mecat_b200", "csrc", "asm_pipeline.h"), os.path.join(ROOT, "mecat_b200", "csrc", "asm_core.cuh"),
                  os.path.join(ROOT, "include", "mecat_b200.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(f) > os.path.getmtime(exe) for f in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", exe] + src)
    for twin in ("mecat2asmpw50", "mecat2trimpw", "mecat2trimpw50"):
        dst = os.path.join(out_dir, twin)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(exe):
            shutil.copy2(exe, dst)
    return out_dir


def asm_index_numpy(text):
    """creat_ref_index (mecat2asmpw.c:397-497) in numpy: (codes kept, their list lengths, all positions in list order).
    13-mers in A0 T1 C2 G3; a letter other than ACGT ends a k-mer; lists of more than 256 dropped; 1-based starts."""
    t = np.frombuffer(text, dtype=np.uint8)
    code = np.full(256, 4, dtype=np.int64)
    for i, c in enumerate(b"ATCG"):
        code[c] = i
    v = code[t]
    n = len(v) - 12
    kmer = np.zeros(n, dtype=np.int64)
    ok = np.ones(n, dtype=bool)
    for j in range(13):
        kmer = kmer * 4 + (v[j:j + n] & 3)
        ok &= v[j:j + n] < 4
    starts = np.nonzero(ok)[0]
    codes = kmer[starts]
    order = np.lexsort((starts, codes))
    codes, starts = codes[order], starts[order]
    uniq, cnt = np.unique(codes, return_counts=True)
    keep = np.repeat(cnt <= 256, cnt)
    return uniq[cnt <= 256], cnt[cnt <= 256], (starts[keep] + 1).astype(np.int32), bool((cnt > 256).any())
