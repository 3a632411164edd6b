"""ctypes binding of the C ABI in include/mecat_b200.h.

This is the only way Python reaches the product: every call goes through
mecat_b200/libmecat_b200.so (hand-written sm_100a kernels).  There is no CPU fallback --
a missing library or a machine without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmecat_b200.so")


class MecatB200Error(RuntimeError):
    pass


class Volume(C.Structure):
    _fields_ = [("num_reads", C.c_int32), ("num_bases", C.c_int32), ("start_read_id", C.c_int32),
                ("offset_size", C.POINTER(C.c_int32)), ("pac", C.POINTER(C.c_uint8))]


class PwParams(C.Structure):
    _fields_ = [("task", C.c_int32), ("num_candidates", C.c_int32), ("min_align_size", C.c_int32),
                ("min_kmer_match", C.c_int32), ("tech", C.c_int32)]


KERNEL_NAMES = ["orient", "index_count", "index_scan", "index_fill", "index_sort", "seed", "walk", "merge", "extend",
                "finalize", "cns_accept", "cns_normvote", "cns_segment", "cns_region", "cns_poa", "cns_assemble",
                "ref_count", "ref_seed", "ref_rescue", "asm_index", "asm_seed", "asm_extend"]
K_NUM = len(KERNEL_NAMES)      # MECAT_K_NUM


class CnsParams(C.Structure):
    _fields_ = [("min_mapping_ratio", C.c_double), ("min_align_size", C.c_int32), ("min_cov", C.c_int32),
                ("min_size", C.c_int64), ("tech", C.c_int32), ("input_type", C.c_int32)]


CNS_PIECE_DTYPE = np.dtype([("id", "<i8"), ("beg", "<i8"), ("end", "<i8"), ("seq_offset", "<i8"), ("seq_len", "<i8")])


class AsmReadsC(C.Structure):      # mecat_asm_reads
    _fields_ = [("text", C.c_char_p), ("num_letters", C.c_int64), ("num_reads", C.c_int32), ("first_read_id", C.c_int32),
                ("read_start", C.POINTER(C.c_int32)), ("read_len", C.POINTER(C.c_int32))]


class AsmParams(C.Structure):      # mecat_asm_params
    _fields_ = [("variant", C.c_int32), ("max_candidates", C.c_int32)]


ASM_OVERLAP_DTYPE = np.dtype([("sread", "<i4"), ("qread", "<i4"), ("score", "<f4"), ("sbeg", "<i4"), ("send", "<i4"), ("slen", "<i4"),
                              ("strand", "<i4"), ("qbeg", "<i4"), ("qend", "<i4"), ("qlen", "<i4")])


class AsmReads:
    """A file of reads the way mecat2asmpw keeps it (load_read, mecat2asmpw.c:345-372): one text with a NUL behind every
    read; read r is numbered first_read_id + r."""

    def __init__(self, seqs, first_read_id):
        self.first_read_id = int(first_read_id)
        self.start = np.zeros(len(seqs), dtype=np.int32)
        self.len = np.array([len(s) for s in seqs], dtype=np.int32)
        if len(seqs):
            self.start[1:] = np.cumsum(self.len[:-1] + 1)
        self.text = ("\0".join(seqs) + "\0").encode() if len(seqs) else b""

    def c(self):
        return AsmReadsC(self.text, len(self.text), len(self.len), self.first_read_id, self.start.ctypes.data_as(C.POINTER(C.c_int32)),
                         self.len.ctypes.data_as(C.POINTER(C.c_int32)))


def asm_lines(recs):
    """The line mecat2asmpw / mecat2trimpw print per overlap (mecat2asmpw.c:944-945)."""
    return ["%d %d %.3f 100 0 %d %d %d %d %d %d %d" % (r["sread"], r["qread"], float(r["score"]), r["sbeg"], r["send"], r["slen"], r["strand"],
                                                      r["qbeg"], r["qend"], r["qlen"]) for r in recs]


class Stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_float * K_NUM), ("kernel_launches", C.c_int64 * K_NUM),
                ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("host_ms", C.c_float), ("total_ms", C.c_float),
                ("wall_index_ms", C.c_float), ("wall_seed_ms", C.c_float), ("wall_extend_ms", C.c_float), ("wall_other_ms", C.c_float),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("num_hits", C.c_int64), ("num_candidates", C.c_int64), ("num_extend_blocks", C.c_int64),
                ("index_kmers", C.c_int64), ("index_bases", C.c_int64), ("num_records", C.c_int64),
                ("num_extend_cells", C.c_int64), ("num_extend_spills", C.c_int64)]


class RefGenomeC(C.Structure):      # mecat_ref_genome
    _fields_ = [("num_bases", C.c_int64), ("pac", C.POINTER(C.c_uint8)), ("num_runs", C.c_int32),
                ("run_start_len", C.POINTER(C.c_int64))]


class RefReadsC(C.Structure):       # mecat_ref_reads
    _fields_ = [("num_reads", C.c_int32), ("vol", C.POINTER(Volume)), ("read_len", C.POINTER(C.c_int32)),
                ("fwd_read", C.POINTER(C.c_int32)), ("rev_read", C.POINTER(C.c_int32)), ("rev_is_rc", C.POINTER(C.c_int32)),
                ("num_bad", C.c_int64), ("bad", C.POINTER(C.c_int64))]


class RefParams(C.Structure):       # mecat_ref_params
    _fields_ = [("num_candidates", C.c_int32), ("num_output", C.c_int32), ("want_strings", C.c_int32), ("tech", C.c_int32)]


REF_RESULT_DTYPE = np.dtype([("read", "<i4"), ("dir", "<i4"), ("vscore", "<i4"), ("qb", "<i4"), ("qe", "<i4"), ("qs", "<i4"),
                             ("sb", "<i8"), ("se", "<i8"), ("columns", "<i4"), ("matches", "<i4"), ("str_offset", "<i8")])

EC_DTYPE = np.dtype([(n, "<i4") for n in
                     ("qdir", "qid", "qext", "qsize", "qoff", "qend", "sdir", "sid", "sext", "ssize", "soff",
                      "send", "score")])
M4_DTYPE = np.dtype([("qid", "<i8"), ("sid", "<i8"), ("ident", "<f8"), ("vscore", "<i4"), ("qdir", "<i4"),
                     ("qoff", "<i8"), ("qend", "<i8"), ("qsize", "<i8"), ("sdir", "<i4"), ("pad", "<i4"),
                     ("soff", "<i8"), ("send", "<i8"), ("ssize", "<i8"), ("qext", "<i8"), ("sext", "<i8")])
TASK_DTYPE = np.dtype([(n, "<i4") for n in ("qread", "qstrand", "qstart", "sread", "sstart")])
RESULT_DTYPE = np.dtype([("ok", "<i4"), ("qstart", "<i4"), ("qend", "<i4"), ("sstart", "<i4"), ("send", "<i4"),
                         ("columns", "<i4"), ("matches", "<i4"), ("pad", "<i4"), ("ident", "<f8")])

ALIGN_TASK_DTYPE = np.dtype([(n, "<i4") for n in ("qread", "qstrand", "qstart", "sread", "sstart", "swin_off", "swin_len")])
ALIGN_RESULT_DTYPE = np.dtype([("ok", "<i4"), ("qstart", "<i4"), ("qend", "<i4"), ("sstart", "<i4"), ("send", "<i4"),
                               ("columns", "<i4"), ("matches", "<i4"), ("pad", "<i4"), ("ident", "<f8"), ("str_offset", "<i8")])

EXPORTS = [
    "mecat_b200_abi_version", "mecat_b200_device_count", "mecat_b200_init", "mecat_b200_destroy",
    "mecat_b200_last_error", "mecat_b200_free", "mecat_b200_get_stats", "mecat_b200_reset_stats",
    "mecat_b200_volume_upload",
    "mecat_b200_volume_release", "mecat_b200_index_build", "mecat_b200_index_count_part", "mecat_b200_index_finish_part",
    "mecat_b200_index_device_arrays", "mecat_b200_index_release", "mecat_b200_index_export",
    "mecat_b200_pw_tile", "mecat_b200_pw_candidates", "mecat_b200_pw_overlaps", "mecat_b200_pw_raw_candidates",
    "mecat_b200_extend_batch", "mecat_b200_align_batch", "mecat_b200_cns_reads", "mecat_b200_cns_sort_candidates",
    "mecat_b200_host_free", "mecat_b200_pw_tile_range", "mecat_b200_volume_from_device", "mecat_b200_split_dataset", "mecat_b200_volume_load", "mecat_b200_volume_unload", "mecat_b200_volume_from_fasta",
    "mecat_b200_ref_index_build", "mecat_b200_ref_index_release", "mecat_b200_ref_map",
    "mecat_b200_ref_index_export", "mecat_b200_ref_raw_candidates",
    "mecat_b200_cns_reads_multi", "mecat_b200_volumes_from_fasta", "mecat_b200_volumes_unload",
    "mecat_b200_pw_tile_text", "mecat_b200_records_text", "mecat_b200_volume_from_text",
    "mecat_b200_asm_index_build", "mecat_b200_asm_index_release", "mecat_b200_asm_overlaps", "mecat_b200_asm_index_export",
]

_lib = None


def _adopt(lib, addr, nbytes, dtype):
    """numpy view of a library-owned host buffer without copying; the buffer is free()d when the last
    array referring to it is collected."""
    import weakref
    mem = (C.c_char * nbytes).from_address(addr)
    weakref.finalize(mem, lib.mecat_b200_host_free, C.c_void_p(addr))
    return np.frombuffer(mem, dtype=dtype)


def load_library():
    """dlopen the product library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MecatB200Error("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    VP, PP, vp = C.POINTER(Volume), C.POINTER(PwParams), C.c_void_p
    L.mecat_b200_init.argtypes = [C.POINTER(vp), C.c_int, vp]
    L.mecat_b200_destroy.argtypes = [vp]
    L.mecat_b200_last_error.restype = C.c_char_p
    L.mecat_b200_last_error.argtypes = [vp]
    L.mecat_b200_free.argtypes = [vp, vp]
    L.mecat_b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.mecat_b200_reset_stats.argtypes = [vp]
    L.mecat_b200_volume_upload.argtypes = [vp, VP, C.POINTER(vp)]
    L.mecat_b200_volume_release.argtypes = [vp, vp]
    L.mecat_b200_index_build.argtypes = [vp, vp, C.POINTER(vp)]
    L.mecat_b200_index_release.argtypes = [vp, vp]
    L.mecat_b200_index_count_part.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.mecat_b200_index_finish_part.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32]
    L.mecat_b200_index_device_arrays.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int64)]
    L.mecat_b200_index_export.argtypes = [vp, vp, C.POINTER(C.c_int64), vp, vp]
    L.mecat_b200_pw_tile.argtypes = [vp, vp, vp, vp, PP, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_volume_from_text.argtypes = [vp, vp, C.c_size_t, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, C.POINTER(vp)]
    L.mecat_b200_pw_tile_text.argtypes = [vp, vp, vp, vp, PP, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.mecat_b200_records_text.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_pw_candidates.argtypes = [vp, VP, VP, PP, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_pw_overlaps.argtypes = [vp, VP, VP, PP, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_pw_raw_candidates.argtypes = [vp, vp, vp, vp, PP, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_extend_batch.argtypes = [vp, C.c_int, vp, vp, vp, C.c_size_t, C.c_int, C.POINTER(vp)]
    L.mecat_b200_align_batch.argtypes = [vp, C.c_int, C.c_double, vp, vp, vp, C.c_size_t, C.c_int, C.POINTER(vp),
                                         C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_cns_reads.argtypes = [vp, vp, vp, C.c_size_t, C.POINTER(CnsParams), C.POINTER(vp), C.POINTER(C.c_size_t),
                                       C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_cns_reads_multi.argtypes = [vp, C.POINTER(vp), C.c_int, vp, C.c_size_t, C.POINTER(CnsParams), C.POINTER(vp),
                                             C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_volumes_from_fasta.argtypes = [C.c_char_p, C.c_int64, C.POINTER(VP), C.POINTER(C.c_int), C.c_char_p, C.c_int]
    L.mecat_b200_volumes_unload.argtypes = [VP, C.c_int]
    L.mecat_b200_cns_sort_candidates.argtypes = [vp, C.c_int]
    L.mecat_b200_host_free.argtypes = [vp]
    L.mecat_b200_pw_tile_range.argtypes = [vp, vp, vp, vp, PP, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_volume_from_device.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), vp, C.POINTER(vp)]
    L.mecat_b200_split_dataset.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(C.c_int), C.c_char_p, C.c_int]
    L.mecat_b200_volume_load.argtypes = [C.c_char_p, VP]
    L.mecat_b200_volume_from_fasta.argtypes = [C.c_char_p, VP, C.c_char_p, C.c_int]
    L.mecat_b200_volume_unload.argtypes = [VP]
    L.mecat_b200_ref_index_build.argtypes = [vp, C.POINTER(RefGenomeC), C.POINTER(vp)]
    L.mecat_b200_ref_index_release.argtypes = [vp, vp]
    L.mecat_b200_ref_map.argtypes = [vp, vp, C.POINTER(RefReadsC), C.POINTER(RefParams), C.POINTER(vp), C.POINTER(C.c_size_t),
                                     C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_ref_index_export.argtypes = [vp, vp, C.POINTER(C.c_int64), vp, vp]
    L.mecat_b200_ref_raw_candidates.argtypes = [vp, vp, C.POINTER(RefReadsC), C.POINTER(RefParams), C.POINTER(vp), C.POINTER(vp),
                                                C.POINTER(C.c_size_t)]
    L.mecat_b200_asm_index_build.argtypes = [vp, C.POINTER(AsmReadsC), C.POINTER(vp)]
    L.mecat_b200_asm_index_release.argtypes = [vp, vp]
    L.mecat_b200_asm_overlaps.argtypes = [vp, vp, C.POINTER(AsmReadsC), C.POINTER(AsmParams), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mecat_b200_asm_index_export.argtypes = [vp, vp, C.POINTER(C.c_int64), vp, vp]
    _lib = L
    return L


class HostVolume:
    """A packed volume in host memory in the reference's `volN` layout."""

    def __init__(self, offset_size, pac, num_bases, start_read_id=0):
        self.offset_size = np.ascontiguousarray(offset_size, dtype=np.int32).reshape(-1, 2)
        self.pac = np.ascontiguousarray(pac, dtype=np.uint8)
        self.num_reads = int(self.offset_size.shape[0])
        self.num_bases = int(num_bases)
        self.start_read_id = int(start_read_id)

    def c(self):
        return Volume(self.num_reads, self.num_bases, self.start_read_id,
                      self.offset_size.ctypes.data_as(C.POINTER(C.c_int32)),
                      self.pac.ctypes.data_as(C.POINTER(C.c_uint8)))

    @staticmethod
    def load(path):
        with open(path, "rb") as f:
            hdr = np.frombuffer(f.read(12), dtype="<i4")
            n, nb, sid = int(hdr[0]), int(hdr[1]), int(hdr[2])
            os_ = np.frombuffer(f.read(8 * n), dtype="<i4").reshape(-1, 2).copy()
            pac = np.frombuffer(f.read((nb + 3) // 4), dtype=np.uint8).copy()
        return HostVolume(os_, pac, nb, sid)


_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate(b"ACGT"):
    _CODE[_ch] = _i
    _CODE[_ch + 32] = _i            # lower case aligns like upper case (extract_sequences)
_UPPER = np.zeros(256, dtype=bool)
_UPPER[list(b"ACGT")] = True


def pack_bases(codes):
    """2 bits per base in the reference's volume layout: base i in byte i >> 2 at shift ((~i) & 3) << 1."""
    n = len(codes)
    c = np.zeros((n + 3) // 4 * 4, dtype=np.uint8)
    c[:n] = codes
    c = c.reshape(-1, 4)
    return (c[:, 0] << 6 | c[:, 1] << 4 | c[:, 2] << 2 | c[:, 3]).astype(np.uint8)


class RefGenome:
    """The genome as mecat2ref's creat_ref_index keeps it (all sequences concatenated, upper-cased), packed for
    mecat_b200_ref_index_build.  `from_fasta` reads plain FASTA (a header is a line starting with '>')."""

    def __init__(self, names, seqs):
        self.names = list(names)
        self.starts, self.sizes = [], []
        at = 0
        for s in seqs:
            self.starts.append(at)
            self.sizes.append(len(s))
            at += len(s)
        raw = np.frombuffer(b"".join(seqs).upper(), dtype=np.uint8)
        good = _UPPER[raw]
        self.num_bases = int(len(raw))
        self.pac = np.ascontiguousarray(pack_bases(np.where(good, _CODE[raw], 0).astype(np.uint8)))
        edge = np.flatnonzero(np.diff(np.concatenate(([0], good.view(np.int8), [0]))))
        self.runs = np.ascontiguousarray(np.stack([edge[0::2], edge[1::2] - edge[0::2]], axis=1).astype(np.int64))

    @staticmethod
    def from_fasta(path):
        names, seqs, cur = [], [], []
        with open(path, "rb") as f:
            for line in f:
                if line.startswith(b">"):
                    if names:
                        seqs.append(b"".join(cur))
                    names.append(line[1:].split()[0].decode() if line[1:].split() else "")
                    cur = []
                else:
                    cur.append(line.rstrip(b"\r\n"))
        if names:
            seqs.append(b"".join(cur))
        return RefGenome(names, seqs)

    def c(self):
        return RefGenomeC(self.num_bases, self.pac.ctypes.data_as(C.POINTER(C.c_uint8)), len(self.runs),
                          self.runs.ctypes.data_as(C.POINTER(C.c_int64)))

    def contig_of(self, offset):
        """(index, start, size) of the sequence holding a concatenated offset (get_chr_id)."""
        k = int(np.searchsorted(np.asarray(self.starts), offset, side="right")) - 1
        return k, self.starts[k], self.sizes[k]


class RefReads:
    """A batch of reads packed for mecat_b200_ref_map: reads of upper-case ACGT are packed once; any other read also
    carries its reverse strand, built the way the reference builds it (complement of upper-case ACGT only)."""

    _COMP = bytes.maketrans(b"ACGT", b"TGCA")

    def __init__(self, seqs):
        strands, self.read_len, self.fwd_read, self.rev_read, self.rev_is_rc = [], [], [], [], []
        for s in seqs:
            s = bytes(s)
            self.read_len.append(len(s))
            self.fwd_read.append(len(strands))
            strands.append(s)
            if _UPPER[np.frombuffer(s, dtype=np.uint8)].all():
                self.rev_read.append(self.fwd_read[-1])
                self.rev_is_rc.append(1)
            else:
                self.rev_read.append(len(strands))
                self.rev_is_rc.append(0)
                strands.append(s[::-1].translate(self._COMP))
        offsz, at = [], 0
        for s in strands:
            offsz.append((at, len(s)))
            at += len(s) + 1                                   # one pad base between reads
        raw = np.frombuffer(b"A".join(strands) + b"A", dtype=np.uint8) if strands else np.zeros(0, np.uint8)
        good = _UPPER[raw]
        pad = np.zeros(len(raw), dtype=bool)
        if strands:
            pad[np.asarray([o + n for o, n in offsz])] = True
        self.bad = np.ascontiguousarray(np.flatnonzero(~good & ~pad).astype(np.int64))
        code = _CODE[raw]
        self.volume = HostVolume(np.asarray(offsz, dtype=np.int32).reshape(-1, 2), pack_bases(np.where(code > 3, 0, code).astype(np.uint8)),
                                 len(raw))
        self._arrays = [np.ascontiguousarray(a, dtype=np.int32) for a in (self.read_len, self.fwd_read, self.rev_read, self.rev_is_rc)]

    def c(self):
        self._vol = self.volume.c()
        p = [a.ctypes.data_as(C.POINTER(C.c_int32)) for a in self._arrays]
        return RefReadsC(len(self.read_len), C.pointer(self._vol), p[0], p[1], p[2], p[3], len(self.bad),
                         self.bad.ctypes.data_as(C.POINTER(C.c_int64)))


def format_ref_results(genome, read_names, records, qstrings=b"", sstrings=b"", fmt=1):
    """The text mecat2ref writes for `records` of Context.ref_map (print_ref_result / print_m4_result,
    src/mecat2ref/output.cpp:8-191): fmt 0 = ref (header + both alignment strings), 1 = m4, 2 = sam records."""
    out = []
    for r in records:
        k, start, size = genome.contig_of(int(r["sb"]))
        qb, qe, qs = int(r["qb"]), int(r["qe"]), int(r["qs"])
        if r["dir"]:
            qb, qe = qs - qe, qs - qb
        name, sb, se = genome.names[k], int(r["sb"]) - start, int(r["se"]) - start
        if fmt == 0:
            o, n = int(r["str_offset"]), int(r["columns"])
            out.append("%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n%s\n%s\n" % (
                read_names[int(r["read"])], name, "R" if r["dir"] else "F", int(r["vscore"]), qb, qe, qs, sb, se, size,
                qstrings[o:o + n].decode(), sstrings[o:o + n].decode()))
        elif fmt == 2:
            o, n = int(r["str_offset"]), int(r["columns"])
            qm, sm = qstrings[o:o + n], sstrings[o:o + n]
            cigar = ["%dH" % int(r["qb"])] if r["qb"] else []
            i = 0
            while i < n:
                j = i + 1
                if qm[i] == 45:
                    while j < n and qm[j] == 45:
                        j += 1
                    cigar.append("%dD" % (j - i))
                elif sm[i] == 45:
                    while j < n and sm[j] == 45:
                        j += 1
                    cigar.append("%dI" % (j - i))
                else:
                    while j < n and qm[j] != 45 and sm[j] != 45:
                        j += 1
                    cigar.append("%dM" % (j - i))
                i = j
            if int(r["qe"]) != qs:
                cigar.append("%dH" % (qs - int(r["qe"])))
            out.append("%d\t%d\t%s\t%d\t255\t%s\t*\t0\t0\t%s\t*\n" % (
                read_names[int(r["read"])], 16 if r["dir"] else 0, name, sb + 1, "".join(cigar), qm.replace(b"-", b"").decode()))
        else:
            ident = float(int(r["matches"])) / float(int(r["columns"]))
            ident *= 100.0
            out.append("%d\t%s\t%.4f\t%d\t%d\t%d\t%d\t%d\t0\t%d\t%d\t%d\n" % (
                read_names[int(r["read"])], name, ident, int(r["vscore"]), 1 if r["dir"] else 0, qb, qe, qs, sb, se, size))
    return "".join(out)


def volume_from_fasta(reads_path):
    """All reads of a FASTA/FASTQ file as one HostVolume (the reference's PackedDB::load_fasta_db), no files written."""
    L = load_library()
    v = Volume()
    err = C.create_string_buffer(512)
    if L.mecat_b200_volume_from_fasta(reads_path.encode(), C.byref(v), err, 512) != 0:
        raise MecatB200Error(err.value.decode())
    try:
        os_ = np.ctypeslib.as_array(v.offset_size, shape=(2 * v.num_reads,)).reshape(-1, 2).copy() if v.num_reads else np.zeros((0, 2), np.int32)
        pac = np.ctypeslib.as_array(v.pac, shape=((v.num_bases + 3) // 4,)).copy()
    finally:
        L.mecat_b200_volume_unload(C.byref(v))
    return HostVolume(os_, pac, v.num_bases, v.start_read_id)


def volumes_from_fasta(reads_path, max_volume_bases=0):
    """All reads of a FASTA/FASTQ file as a list of HostVolumes cut like split_dataset cuts its files, no files written."""
    L = load_library()
    arr = C.POINTER(Volume)()
    n = C.c_int()
    err = C.create_string_buffer(512)
    if L.mecat_b200_volumes_from_fasta(reads_path.encode(), max_volume_bases, C.byref(arr), C.byref(n), err, 512) != 0:
        raise MecatB200Error(err.value.decode())
    out = []
    try:
        for i in range(n.value):
            v = arr[i]
            os_ = np.ctypeslib.as_array(v.offset_size, shape=(2 * v.num_reads,)).reshape(-1, 2).copy() if v.num_reads else np.zeros((0, 2), np.int32)
            pac = np.ctypeslib.as_array(v.pac, shape=((v.num_bases + 3) // 4,)).copy()
            out.append(HostVolume(os_, pac, v.num_bases, v.start_read_id))
    finally:
        L.mecat_b200_volumes_unload(arr, n.value)
    return out


def split_dataset(reads_path, wrk_dir, max_volume_bases=0):
    """FASTA/FASTQ -> wrk_dir/volN + fileindex.txt (the reference's split_raw_dataset). Returns the volume paths."""
    L = load_library()
    os.makedirs(wrk_dir, exist_ok=True)
    n = C.c_int()
    err = C.create_string_buffer(512)
    if L.mecat_b200_split_dataset(reads_path.encode(), wrk_dir.encode(), max_volume_bases, C.byref(n), err, 512) != 0:
        raise MecatB200Error(err.value.decode())
    with open(os.path.join(wrk_dir, "fileindex.txt")) as f:
        names = [l.strip() for l in f if l.strip()]
    assert len(names) == n.value
    return names


def normalise_candidates(can, min_read_size):
    """The two partition-file records of every `.can` line (reference partition_candidates +
    normalise_candidate, src/mecat2cns/overlaps_partition.cpp:141-165,176-224): one with each read as the
    read to correct (`sid`), strands flipped so that sdir == 0.  `can`: EC_DTYPE array as read from a .can."""
    can = can[(can["qsize"] >= min_read_size) & (can["ssize"] >= min_read_size)]
    a = np.zeros(len(can), dtype=EC_DTYPE)                 # query becomes the target
    a["qdir"], a["qid"], a["qext"], a["qsize"] = can["sdir"], can["sid"], can["sext"], can["ssize"]
    a["sdir"], a["sid"], a["sext"], a["ssize"] = can["qdir"], can["qid"], can["qext"], can["qsize"]
    a["score"] = can["score"]
    b = can.copy()                                         # subject stays the target
    out = np.empty(2 * len(can), dtype=EC_DTYPE)
    out[0::2], out[1::2] = a, b
    rev = out["sdir"] == 1
    out["qdir"][rev] = 1 - out["qdir"][rev]
    out["sdir"][rev] = 0
    for f in ("qoff", "qend", "soff", "send"):
        out[f] = 0
    return out


def m4_partitions(m4_file, min_mapping_ratio=0.9, min_read_size=5000, batch_size=100000):
    """The records partition_m4records writes for `mecat2cns -i 1` (src/mecat2cns/overlaps_partition.cpp:345-410): per
    partition of batch_size reads an EC_DTYPE array, records in file order -- size and mapping-range filters, then for
    either read as the one to correct m4_to_candidate(normalize_m4record(...)) (src/common/alignment.h:71-103,170-186).
    m4_file: path or file object of `mecat2pw -j 1 -g 1` lines.  Feed each array to cns_reads(..., input_type=1)."""
    f = open(m4_file) if isinstance(m4_file, str) else m4_file
    ratio = min_mapping_ratio - 0.02
    parts = {}
    for line in f:
        t = line.split()
        if len(t) < 12:
            continue
        if len(t) < 14:
            raise MecatB200Error("no gapped start position is provided (mecat2pw -g 1)")
        qid, sid = int(t[0]), int(t[1])
        vscore, qdir, qoff, qend, qsize, sdir, soff, send, ssize, qext, sext = (int(x) for x in t[3:14])
        if qsize < min_read_size or ssize < min_read_size:
            continue
        if not (qend - qoff >= int(qsize * ratio) or send - soff >= int(ssize * ratio)):
            continue
        for subject_is_target in (False, True):
            if subject_is_target:
                e = [qdir, qid, qext, qsize, qoff, qend, sdir, sid, sext, ssize, soff, send, vscore]
            else:
                e = [sdir, sid, sext, ssize, soff, send, qdir, qid, qext, qsize, qoff, qend, vscore]
            if e[6] == 1:
                e[6] = 0
                e[0] = 1 - e[0]
            parts.setdefault(e[7] // batch_size, []).append(tuple(e))
    if isinstance(m4_file, str):
        f.close()
    return {k: np.array(v, dtype=EC_DTYPE) for k, v in sorted(parts.items())}


def read_can(path):
    """`.can` text (qid sid qdir sdir qext sext score qsize ssize) -> EC_DTYPE array."""
    raw = np.loadtxt(path, dtype=np.int64, ndmin=2)
    ec = np.zeros(len(raw), dtype=EC_DTYPE)
    for i, f in enumerate(("qid", "sid", "qdir", "sdir", "qext", "sext", "score", "qsize", "ssize")):
        ec[f] = raw[:, i]
    return ec


def pw_params(task=1, num_candidates=100, min_align_size=2000, min_kmer_match=4, tech=0):
    """Defaults of mecat2pw for PacBio (pw_options.cpp:8-13,30-50)."""
    return PwParams(task, num_candidates, min_align_size, min_kmer_match, tech)


class Context:
    """One context per device (mecat_b200_init)."""

    def __init__(self, device=0):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.mecat_b200_init(C.byref(self.h), device, None)
        if rc != 0:
            raise MecatB200Error("mecat_b200_init(device=%d) failed with code %d (no CUDA device? there is no CPU fallback)"
                                 % (device, rc))

    def close(self):
        if self.h:
            self.L.mecat_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise MecatB200Error("%s: %s" % (what, self.L.mecat_b200_last_error(self.h).decode()))

    def _take(self, ptr, n, dtype):
        if not ptr.value or n == 0:
            if ptr.value:
                self.L.mecat_b200_free(self.h, ptr)
            return np.zeros(0, dtype=dtype)
        if n * dtype.itemsize < (1 << 20):
            buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr.value)
            arr = np.frombuffer(buf, dtype=dtype).copy()
            del buf
            self.L.mecat_b200_free(self.h, ptr)
            return arr
        # large results: wrap the library-owned buffer without copying; it is released with the array
        return _adopt(self.L, ptr.value, n * dtype.itemsize, dtype)

    def stats(self):
        s = Stats()
        self._check(self.L.mecat_b200_get_stats(self.h, C.byref(s)), "get_stats")
        d = {f[0]: getattr(s, f[0]) for f in Stats._fields_[2:]}
        d["kernel_ms"] = {n: float(s.kernel_ms[i]) for i, n in enumerate(KERNEL_NAMES)}
        d["kernel_launches"] = {n: int(s.kernel_launches[i]) for i, n in enumerate(KERNEL_NAMES)}
        d["gpu_launches"] = int(sum(s.kernel_launches))
        return d

    def reset_stats(self):
        self.L.mecat_b200_reset_stats(self.h)

    # ---- device-resident objects
    def upload(self, vol):
        d = C.c_void_p()
        cv = vol.c()
        self._check(self.L.mecat_b200_volume_upload(self.h, C.byref(cv), C.byref(d)), "volume_upload")
        return d

    def release_volume(self, d):
        self.L.mecat_b200_volume_release(self.h, d)

    def index_build(self, dvol):
        i = C.c_void_p()
        self._check(self.L.mecat_b200_index_build(self.h, dvol, C.byref(i)), "index_build")
        return i

    def index_count_part(self, dvol, code_lo, code_hi):
        i = C.c_void_p()
        self._check(self.L.mecat_b200_index_count_part(self.h, dvol, code_lo, code_hi, C.byref(i)), "index_count_part")
        return i

    def index_finish_part(self, dvol, index, code_lo, code_hi):
        self._check(self.L.mecat_b200_index_finish_part(self.h, dvol, index, code_lo, code_hi), "index_finish_part")

    def index_device_arrays(self, index):
        """(counts_ptr, begin_ptr, positions_ptr, num_kmers): raw device addresses (for NCCL exchanges)."""
        a, b, p, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        self._check(self.L.mecat_b200_index_device_arrays(self.h, index, C.byref(a), C.byref(b), C.byref(p), C.byref(n)),
                    "index_device_arrays")
        return a.value, b.value, p.value, n.value

    def release_index(self, i):
        self.L.mecat_b200_index_release(self.h, i)

    def index_export(self, index):
        n = C.c_int64()
        self._check(self.L.mecat_b200_index_export(self.h, index, C.byref(n), None, None), "index_export")
        begin = np.zeros((1 << 26) + 1, dtype=np.uint32)
        pos = np.zeros(max(1, n.value), dtype=np.int32)
        self._check(self.L.mecat_b200_index_export(self.h, index, C.byref(n), begin.ctypes.data_as(C.c_void_p),
                                                   pos.ctypes.data_as(C.c_void_p)), "index_export")
        return begin, pos[:n.value]

    def pw_tile(self, index, dref, dreads, params):
        out, n = C.c_void_p(), C.c_size_t()
        self._check(self.L.mecat_b200_pw_tile(self.h, index, dref, dreads, C.byref(params), C.byref(out), C.byref(n)),
                    "pw_tile")
        return self._take(out, n.value, EC_DTYPE if params.task == 0 else M4_DTYPE)

    def volume_from_text(self, text, src_offset, offset_size, num_bases, start_read_id=0, want_pac=True):
        """Pack a volume on the device from the letters of its reads.  Returns (device volume, packed bytes or None)."""
        src = np.ascontiguousarray(src_offset, dtype=np.int64)
        osz = np.ascontiguousarray(offset_size, dtype=np.int32).reshape(-1)
        n = len(src)
        pac = np.zeros((num_bases + 3) // 4, dtype=np.uint8) if want_pac else None
        d = C.c_void_p()
        # (a bytes object is handed over as a pointer to its buffer: no copy of the letters on the Python side)
        self._check(self.L.mecat_b200_volume_from_text(self.h, text if len(text) else None, len(text),
                                                       src.ctypes.data_as(C.c_void_p), osz.ctypes.data_as(C.c_void_p), n, num_bases, start_read_id,
                                                       pac.ctypes.data_as(C.c_void_p) if want_pac else None, C.byref(d)), "volume_from_text")
        return d, pac

    def pw_tile_text(self, index, dref, dreads, params, gapped=False):
        """One tile as the text of the reference's output file (written on the device).  Returns (bytes, number of records)."""
        text, nb, n = C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._check(self.L.mecat_b200_pw_tile_text(self.h, index, dref, dreads, C.byref(params), 1 if gapped else 0, C.byref(text),
                                                   C.byref(nb), C.byref(n)), "pw_tile_text")
        out = C.string_at(text.value, nb.value) if text.value else b""
        if text.value:
            self.L.mecat_b200_free(self.h, text)
        return out, n.value

    def records_text(self, records, gapped=False):
        """`.can` / `.m4` lines of EC_DTYPE / M4_DTYPE records, formatted on the device."""
        kind = 0 if records.dtype == EC_DTYPE else 1
        rec = np.ascontiguousarray(records)
        text, nb = C.c_void_p(), C.c_size_t()
        self._check(self.L.mecat_b200_records_text(self.h, kind, 1 if gapped else 0, rec.ctypes.data_as(C.c_void_p), len(rec), C.byref(text),
                                                   C.byref(nb)), "records_text")
        out = C.string_at(text.value, nb.value) if text.value else b""
        if text.value:
            self.L.mecat_b200_free(self.h, text)
        return out

    def align_batch(self, dquery, dsubject, tasks, min_align_size, policy=0, err=0.15):
        """Returns (results, qstrings, sstrings): result['str_offset'] indexes the two NUL-separated byte blobs."""
        tasks = np.ascontiguousarray(tasks, dtype=ALIGN_TASK_DTYPE)
        res, qs, ss, nb = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._check(self.L.mecat_b200_align_batch(self.h, policy, err, dquery, dsubject, tasks.ctypes.data_as(C.c_void_p),
                                                  len(tasks), min_align_size, C.byref(res), C.byref(qs), C.byref(ss),
                                                  C.byref(nb)), "align_batch")
        r = self._take(res, len(tasks), ALIGN_RESULT_DTYPE)
        q = C.string_at(qs.value, nb.value) if qs.value else b""
        s = C.string_at(ss.value, nb.value) if ss.value else b""
        if qs.value:
            self.L.mecat_b200_free(self.h, qs)
        if ss.value:
            self.L.mecat_b200_free(self.h, ss)
        return r, q, s

    def cns_reads(self, dvol, candidates, min_mapping_ratio=0.9, min_align_size=2000, min_cov=6, min_size=5000, tech=0, input_type=0):
        """mecat2cns -i 0 on normalised candidates (EC_DTYPE array).  Returns [(id, beg, end, seq bytes), ...]."""
        ec = np.ascontiguousarray(candidates, dtype=EC_DTYPE)
        p = CnsParams(min_mapping_ratio, min_align_size, min_cov, min_size, tech, input_type)
        pieces, n, seqs, nb = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        self._check(self.L.mecat_b200_cns_reads(self.h, dvol, ec.ctypes.data_as(C.c_void_p), len(ec), C.byref(p), C.byref(pieces),
                                                C.byref(n), C.byref(seqs), C.byref(nb)), "cns_reads")
        pc = self._take(pieces, n.value, CNS_PIECE_DTYPE)
        blob = C.string_at(seqs.value, nb.value) if seqs.value else b""
        if seqs.value:
            self.L.mecat_b200_free(self.h, seqs)
        return [(int(x["id"]), int(x["beg"]), int(x["end"]), blob[int(x["seq_offset"]):int(x["seq_offset"]) + int(x["seq_len"])])
                for x in pc]

    # ---- mecat2asmpw / mecat2trimpw
    def asm_index_build(self, reads):
        """Index of a subject file (AsmReads): replaces load_read + creat_ref_index."""
        d, cr = C.c_void_p(), reads.c()
        self._check(self.L.mecat_b200_asm_index_build(self.h, C.byref(cr), C.byref(d)), "asm_index_build")
        return d

    def asm_index_release(self, idx):
        self._check(self.L.mecat_b200_asm_index_release(self.h, idx), "asm_index_release")

    def asm_index_export(self, idx):
        n = C.c_int64()
        self._check(self.L.mecat_b200_asm_index_export(self.h, idx, C.byref(n), None, None), "asm_index_export")
        begin, pos = np.zeros((1 << 26) + 1, dtype=np.uint32), np.zeros(max(1, n.value), dtype=np.int32)
        self._check(self.L.mecat_b200_asm_index_export(self.h, idx, C.byref(n), begin.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p)),
                    "asm_index_export")
        return begin, pos[:n.value]

    def asm_overlaps(self, idx, reads, variant=0, max_candidates=100):
        """pairwise_mapping for the reads of one query file (AsmReads) against a subject index.  Returns ASM_OVERLAP_DTYPE records."""
        out, n, cr, p = C.c_void_p(), C.c_size_t(), reads.c(), AsmParams(variant, max_candidates)
        self._check(self.L.mecat_b200_asm_overlaps(self.h, idx, C.byref(cr), C.byref(p), C.byref(out), C.byref(n)), "asm_overlaps")
        return self._take(out, n.value, ASM_OVERLAP_DTYPE)

    def cns_reads_multi(self, dvols, candidates, min_mapping_ratio=0.9, min_align_size=2000, min_cov=6, min_size=5000, tech=0, input_type=0):
        """cns_reads for a read set that spans several resident volumes (consecutive read ids)."""
        ec = np.ascontiguousarray(candidates, dtype=EC_DTYPE)
        p = CnsParams(min_mapping_ratio, min_align_size, min_cov, min_size, tech, input_type)
        arr = (C.c_void_p * len(dvols))(*[d.value if hasattr(d, "value") else d for d in dvols])
        pieces, n, seqs, nb = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        self._check(self.L.mecat_b200_cns_reads_multi(self.h, arr, len(dvols), ec.ctypes.data_as(C.c_void_p), len(ec), C.byref(p),
                                                      C.byref(pieces), C.byref(n), C.byref(seqs), C.byref(nb)), "cns_reads_multi")
        pc = self._take(pieces, n.value, CNS_PIECE_DTYPE)
        blob = C.string_at(seqs.value, nb.value) if seqs.value else b""
        if seqs.value:
            self.L.mecat_b200_free(self.h, seqs)
        return [(int(x["id"]), int(x["beg"]), int(x["end"]), blob[int(x["seq_offset"]):int(x["seq_offset"]) + int(x["seq_len"])])
                for x in pc]

    def pw_tile_range(self, index, dref, dreads, params, read_begin, read_end):
        out, n = C.c_void_p(), C.c_size_t()
        self._check(self.L.mecat_b200_pw_tile_range(self.h, index, dref, dreads, C.byref(params), read_begin, read_end,
                                                    C.byref(out), C.byref(n)), "pw_tile_range")
        return self._take(out, n.value, EC_DTYPE if params.task == 0 else M4_DTYPE)

    def volume_from_device(self, num_reads, num_bases, start_read_id, host_offset_size, device_ptr):
        """host_offset_size: int32 numpy [num_reads, 2]; device_ptr: address of the packed bytes in device memory."""
        d = C.c_void_p()
        osz = np.ascontiguousarray(host_offset_size, dtype=np.int32)
        self._check(self.L.mecat_b200_volume_from_device(self.h, num_reads, num_bases, start_read_id,
                                                         osz.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(device_ptr),
                                                         C.byref(d)), "volume_from_device")
        return d

    def pw_raw_candidates(self, index, dref, dreads, params, num_reads):
        """Test hook: (rows[n,12], counts[num_reads]) = the candidate_save lists of every read."""
        rows, counts, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._check(self.L.mecat_b200_pw_raw_candidates(self.h, index, dref, dreads, C.byref(params), C.byref(rows),
                                                        C.byref(counts), C.byref(n)), "pw_raw_candidates")
        i4 = np.dtype("<i4")
        cnt = self._take(counts, num_reads, i4)
        r = self._take(rows, max(1, n.value) * 12, i4).reshape(-1, 12)[:n.value]
        return r, cnt

    # ---- mecat2ref
    def ref_index_build(self, genome):
        i = C.c_void_p()
        g = genome.c()
        self._check(self.L.mecat_b200_ref_index_build(self.h, C.byref(g), C.byref(i)), "ref_index_build")
        return i

    def release_ref_index(self, i):
        self.L.mecat_b200_ref_index_release(self.h, i)

    def ref_map(self, refidx, reads, num_candidates=10, num_output=10, want_strings=True, tech=0):
        """mecat2ref on a RefReads batch.  Returns (records, qstrings, sstrings): REF_RESULT_DTYPE records in output order
        (a read's records adjacent); record['str_offset'] indexes the two NUL-separated byte blobs."""
        p = RefParams(num_candidates, num_output, 1 if want_strings else 0, tech)
        r = reads.c()
        res, n, qs, ss, nb = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._check(self.L.mecat_b200_ref_map(self.h, refidx, C.byref(r), C.byref(p), C.byref(res), C.byref(n), C.byref(qs),
                                              C.byref(ss), C.byref(nb)), "ref_map")
        rec = self._take(res, n.value, REF_RESULT_DTYPE)
        q = C.string_at(qs.value, nb.value) if qs.value else b""
        s = C.string_at(ss.value, nb.value) if ss.value else b""
        if qs.value:
            self.L.mecat_b200_free(self.h, qs)
        if ss.value:
            self.L.mecat_b200_free(self.h, ss)
        return rec, q, s

    def ref_index_export(self, refidx):
        """Test hook: (begin[2^26 + 1], positions) of the genome's k-mer index (0-based k-mer starts)."""
        n = C.c_int64()
        self._check(self.L.mecat_b200_ref_index_export(self.h, refidx, C.byref(n), None, None), "ref_index_export")
        begin = np.zeros((1 << 26) + 1, dtype=np.uint32)
        pos = np.zeros(max(1, n.value), dtype=np.int32)
        self._check(self.L.mecat_b200_ref_index_export(self.h, refidx, C.byref(n), begin.ctypes.data_as(C.c_void_p),
                                                       pos.ctypes.data_as(C.c_void_p)), "ref_index_export")
        return begin, pos[:n.value]

    def ref_raw_candidates(self, refidx, reads, num_candidates=10):
        """Test hook: (rows[n, 4] = loc1 loc2 score chain, counts[2 * reads]) of the first seeding pass, strand by strand."""
        p = RefParams(num_candidates, num_candidates, 0, 0)
        r = reads.c()
        rows, counts, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._check(self.L.mecat_b200_ref_raw_candidates(self.h, refidx, C.byref(r), C.byref(p), C.byref(rows), C.byref(counts),
                                                         C.byref(n)), "ref_raw_candidates")
        i4 = np.dtype("<i4")
        cnt = self._take(counts, 2 * len(reads.read_len), i4)
        row = self._take(rows, max(1, n.value) * 4, i4).reshape(-1, 4)[:n.value]
        return row, cnt

    # ---- host-buffer entry points (the end-to-end calls)
    def pw_candidates(self, ref, reads, params=None):
        p = params or pw_params(task=0)
        out, n = C.c_void_p(), C.c_size_t()
        rv, qv = ref.c(), reads.c()
        self._check(self.L.mecat_b200_pw_candidates(self.h, C.byref(rv), C.byref(qv), C.byref(p), C.byref(out),
                                                    C.byref(n)), "pw_candidates")
        return self._take(out, n.value, EC_DTYPE)

    def pw_overlaps(self, ref, reads, params=None):
        p = params or pw_params(task=1)
        out, n = C.c_void_p(), C.c_size_t()
        rv, qv = ref.c(), reads.c()
        self._check(self.L.mecat_b200_pw_overlaps(self.h, C.byref(rv), C.byref(qv), C.byref(p), C.byref(out),
                                                  C.byref(n)), "pw_overlaps")
        return self._take(out, n.value, M4_DTYPE)

    def extend_batch(self, dquery, dsubject, tasks, min_align_size=2000, policy=0):
        tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
        out = C.c_void_p()
        self._check(self.L.mecat_b200_extend_batch(self.h, policy, dquery, dsubject, tasks.ctypes.data_as(C.c_void_p),
                                                   len(tasks), min_align_size, C.byref(out)), "extend_batch")
        return self._take(out, len(tasks), RESULT_DTYPE)
