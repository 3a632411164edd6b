"""Builds the in-tree native artefacts:

  mecat_b200/libmecat_b200.so   CUDA kernels + C ABI (include/mecat_b200.h), sm_100a only
  mecat_b200/bin/mecat2pw       C++ host drivers with the reference's CLIs (link the library)
  mecat_b200/bin/mecat2cns
  mecat_b200/bin/mecat2ref
  mecat_b200/bin/mecat2asmpw    (+ mecat2asmpw50, mecat2trimpw, mecat2trimpw50: copies, the name selects the variant)
  mecat_b200/bin/gen_reads      seeded synthetic CLR read generator (tools/gen_reads.cpp; test/bench tooling)

nvcc cross-compiles without a GPU.  Nothing here touches oracle/.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libmecat_b200.so")
BIN = os.path.join(HERE, "bin")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOSTCXX = "/usr/bin/g++"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
              "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC,-O2,-pthread", "-Xptxas", "-v"]

CU = ["volume.cu", "index.cu", "seed.cu", "extend.cu", "align.cu", "xdrop.cu", "records.cu", "cns.cu", "refmap.cu", "asmpw.cu", "capi.cu"]


def _newer(src_list, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in src_list)


def _compile(cu):
    src = os.path.join(CSRC, cu)
    obj = os.path.join(OBJ, cu.replace(".cu", ".o"))
    deps = [src, os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "cns_pipeline.h"), os.path.join(CSRC, "cns_core.cuh"), os.path.join(CSRC, "ref_pipeline.h"), os.path.join(CSRC, "ref_core.cuh"), os.path.join(CSRC, "dev_backend.cuh"), os.path.join(CSRC, "xdrop_core.cuh"), os.path.join(CSRC, "m4_core.cuh"), os.path.join(CSRC, "asm_core.cuh"), os.path.join(CSRC, "asm_pipeline.h"),
            os.path.join(ROOT, "include", "mecat_b200.h")]
    if not _newer(deps, obj):
        return obj, ""
    p = subprocess.run([NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (cu, p.stdout, p.stderr))
    return obj, p.stderr


def build(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(BIN, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=len(CU)) as ex:
        res = list(ex.map(_compile, CU))
    objs = [r[0] for r in res]
    for name in ("host_io",):
        src, obj = os.path.join(CSRC, name + ".cpp"), os.path.join(OBJ, name + ".o")
        if _newer([src, os.path.join(ROOT, "include", "mecat_b200.h"), os.path.join(CSRC, "cns_pipeline.h"), os.path.join(CSRC, "cns_core.cuh")], obj):
            subprocess.check_call([HOSTCXX, "-O2", "-std=c++17", "-fPIC", "-pthread", "-c", src, "-o", obj])
        objs.append(obj)
    log = "\n".join(r[1] for r in res if r[1])
    if log:
        with open(os.path.join(ROOT, "build", "ptxas.log"), "w") as f:
            f.write(log)
        if verbose:
            print(log)
    if _newer(objs, LIB):
        subprocess.check_call([NVCC, "-shared", "-ccbin", HOSTCXX, "-gencode", "arch=compute_100a,code=sm_100a",
                               "-o", LIB] + objs)
    host = os.path.join(CSRC, "host")
    for name in ("mecat2pw", "mecat2cns", "mecat2ref", "mecat2asmpw"):
        src, exe = os.path.join(host, name + ".cpp"), os.path.join(BIN, name)
        if os.path.exists(src) and _newer([src, LIB, os.path.join(ROOT, "include", "mecat_b200.h"), os.path.join(host, "format.h"), os.path.join(host, "refio.h")], exe):
            subprocess.check_call([HOSTCXX, "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                                   "-L", HERE, "-lmecat_b200", "-Wl,-rpath,$ORIGIN/.."])
    # one source, four programs: the name selects the variant (mecat2asmpw.cpp)
    for twin in ("mecat2asmpw50", "mecat2trimpw", "mecat2trimpw50"):
        src, dst = os.path.join(BIN, "mecat2asmpw"), os.path.join(BIN, twin)
        if _newer([src], dst):
            shutil.copy2(src, dst)
    gen_src, gen_exe = os.path.join(ROOT, "tools", "gen_reads.cpp"), os.path.join(BIN, "gen_reads")
    if _newer([gen_src], gen_exe):
        subprocess.check_call([HOSTCXX, "-O2", "-std=c++17", "-pthread", "-o", gen_exe, gen_src])
    return LIB


if __name__ == "__main__":
    build(verbose="-v" in sys.argv)
    print(LIB)
