"""CPU tests of the record-assembly and result-text bodies the GPU runs (mecat_b200/csrc/m4_core.cuh, row A12 and
SURVEY.md 8(f) item 3), compiled for the host by tests/m4_host_harness.cpp: the sort must leave equal keys exactly where
the library's std::sort leaves them (append_m4v's containment filter keeps the first of two identical alignments), the
identity must print like `out << double`."""
import ctypes as C
import gzip
import io
import os

import numpy as np

import util


def test_sort_permutation_is_the_librarys():
    H = util.m4_harness()
    # every size 0..200 (the threshold of 16, partitions, several levels), 30 key distributions each, ties everywhere
    assert H.mh_sort_random(200, 30, 7) == 0
    assert H.mh_sort_random(1000, 2, 11) == 0


def test_sort_follows_the_library_into_its_heap_sort():
    """Adversarial keys (McIlroy's construction, built against the library's own std::sort) exhaust the depth limit
    2 floor(log2 n); the heap sort behind it must be the library's too."""
    H = util.m4_harness()
    reached = 0
    for n in (40, 64, 100, 129, 200, 500, 1000):
        used = C.c_int(0)
        assert H.mh_sort_adversary(n, C.byref(used)) == 0, n
        reached += used.value
    assert reached >= 3


def test_identity_prints_like_printf_g():
    H = util.m4_harness()
    msg = C.create_string_buffer(256)
    assert H.mh_fmt_ratios(3000, 1, msg, 256) == 0, msg.value          # every 100 m / n up to n = 3000
    assert H.mh_fmt_ratios(600000, 7919, msg, 256) == 0, msg.value
    rng = np.random.default_rng(3)
    vals = np.concatenate([
        10.0 ** rng.uniform(-9, 6, 200000) * 0.999999,
        np.array([0.0, 1e-9, 9.999995e-5, 1e-4, 0.00099999949, 0.5, 1.0, 9.9999949, 9.999995, 99.99995, 99.999949999, 100.0, 999999.4, 123456.5, 0.1, 0.3,
                  2.5e-5, 1.25, 1234565e-1 / 10]),
        np.round(rng.uniform(0, 100, 100000), 4),            # values that sit exactly on or next to a rounding tie
        np.arange(0, 1000000, 1237) / 1e4 + 0.00005,
    ])
    vals = np.ascontiguousarray(vals[(vals == 0) | ((vals >= 1e-9) & (vals < 1e6))])
    assert H.mh_fmt_values(vals.ctypes.data_as(C.c_void_p), len(vals), msg, 256) == 0, msg.value


def test_lines_equal_the_host_formatter():
    """Whole `.can` / `.m4` lines on the golden records of the small fixture and on extreme field values."""
    import mecat_b200
    H = util.m4_harness()
    with gzip.open(os.path.join(util.GOLDEN, "small.can.gz"), "rt") as f:
        ec = mecat_b200.read_can(io.StringIO(f.read()))
    ec = np.ascontiguousarray(ec)
    rng = np.random.default_rng(5)
    m4 = np.zeros(5000, dtype=mecat_b200.M4_DTYPE)
    for name in m4.dtype.names:
        if name == "ident":
            m4[name] = 100.0 * rng.integers(0, 30000, len(m4)) / rng.integers(30000, 60000, len(m4))
        elif name not in ("pad", "pad_"):
            hi = 2 ** 31 - 1 if m4.dtype[name].itemsize == 4 else 2 ** 40
            m4[name] = rng.integers(0, hi, len(m4))
    m4["vscore"][:10] = -5
    assert H.mh_lines(ec.ctypes.data_as(C.c_void_p), len(ec), m4.ctypes.data_as(C.c_void_p), len(m4)) == 0
