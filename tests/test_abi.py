"""CPU tests of the drop-in boundary: the C-ABI library builds (nvcc cross-compiles without a
GPU), loads, and exports every symbol include/mecat_b200.h declares.  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import mecat_b200
    from mecat_b200 import build
    build.build()
    return mecat_b200.load_library()


def header_functions():
    src = open(os.path.join(ROOT, "include", "mecat_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mecat_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    import mecat_b200
    assert header_functions() == sorted(mecat_b200.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.mecat_b200_abi_version() == 1


def test_struct_sizes_match_reference_records():
    import mecat_b200
    # ExtensionCandidate = 13 x int32 (alignment.h:8-13); M4Record = 104 bytes (alignment.h:21-37, idx_t = int64)
    assert mecat_b200.EC_DTYPE.itemsize == 52
    assert mecat_b200.M4_DTYPE.itemsize == 104
    # one printed line of mecat2asmpw: two ids, the float score, seven ints (mecat2asmpw.c:944-945)
    assert mecat_b200.ASM_OVERLAP_DTYPE.itemsize == 40
    import ctypes
    assert ctypes.sizeof(mecat_b200.api.AsmReadsC) == 40 and ctypes.sizeof(mecat_b200.AsmParams) == 8


def test_no_cpu_fallback(lib):
    """Without a CUDA device the library refuses to initialise; it never computes on the CPU."""
    if lib.mecat_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.mecat_b200_init(C.byref(h), 0, None) != 0
    assert not h.value
    import mecat_b200
    with pytest.raises(mecat_b200.MecatB200Error):
        mecat_b200.Context(0)


def test_product_does_not_reference_oracle():
    """Nothing under mecat_b200/ may import, link or name oracle/."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "mecat_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                s = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"liboracle|oracle/|orc_|libmecatref", s) and f != "build.py":
                    bad.append(os.path.join(d, f))
                elif f == "build.py" and re.search(r"liboracle|orc_|libmecatref", s):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
