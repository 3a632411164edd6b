"""mecat_b200 -- B200-native (sm_100a) implementation of MECAT's all-vs-all overlap hot path.

Layout:
  csrc/        hand-written CUDA kernels + the C ABI (include/mecat_b200.h) + C++ host driver
  api.py       ctypes binding of the C ABI (the only route from Python to the product)
  build.py     in-tree build of libmecat_b200.so / bin/mecat2pw, mecat2cns, mecat2ref, mecat2asmpw (+ mecat2trimpw, *50)
"""
from .api import (CnsParams, CNS_PIECE_DTYPE, normalise_candidates, m4_partitions, read_can, Context, HostVolume, MecatB200Error, PwParams, pw_params, split_dataset, volume_from_fasta, volumes_from_fasta, load_library, LIB_PATH, EXPORTS,
                  EC_DTYPE, M4_DTYPE, TASK_DTYPE, RESULT_DTYPE, ALIGN_TASK_DTYPE, ALIGN_RESULT_DTYPE,
                  RefGenome, RefReads, RefParams, REF_RESULT_DTYPE, format_ref_results,
                  AsmReads, AsmParams, ASM_OVERLAP_DTYPE, asm_lines)

__all__ = ["CnsParams", "CNS_PIECE_DTYPE", "normalise_candidates", "m4_partitions", "read_can", "Context", "HostVolume", "MecatB200Error", "PwParams", "pw_params", "split_dataset", "volume_from_fasta", "volumes_from_fasta", "load_library", "LIB_PATH", "EXPORTS",
           "EC_DTYPE", "M4_DTYPE", "TASK_DTYPE", "RESULT_DTYPE", "ALIGN_TASK_DTYPE", "ALIGN_RESULT_DTYPE",
           "RefGenome", "RefReads", "RefParams", "REF_RESULT_DTYPE", "format_ref_results",
           "AsmReads", "AsmParams", "ASM_OVERLAP_DTYPE", "asm_lines"]
