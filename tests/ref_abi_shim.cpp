// tests/ref_abi_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The handful of C-ABI entry points the mecat2ref command-line driver calls (include/mecat_b200.h), played by the host
// harness (tests/ref_host_harness.cpp).  tests/util.py links mecat_b200/csrc/host/mecat2ref.cpp against this file instead
// of the product library, so the CPU test-suite can run the driver itself -- flags, batching, the packing / device / text
// pipeline, output files -- without a GPU.  Never part of the product: bin/mecat2ref links libmecat_b200.so only.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../include/mecat_b200.h"

extern "C" {
void* harness_ref_index_build(const mecat_ref_genome* g);
void harness_ref_index_release(void* idx);
int harness_ref_map_indexed(void* idx, const mecat_ref_reads* reads, const mecat_ref_params* p, mecat_ref_result** results, size_t* n,
                            char** qstrings, char** sstrings, size_t* string_bytes, char* errbuf, int errcap);
}

extern "C" void harness_set_tech(int tech);

struct mecat_b200_ctx { std::string err; int device; };

extern "C" {

int mecat_b200_device_count(void) { return getenv("MECAT_SHIM_DEVICES") ? atoi(getenv("MECAT_SHIM_DEVICES")) : 1; }
int mecat_b200_init(mecat_b200_ctx** ctx, int device, void*) { *ctx = new mecat_b200_ctx; (*ctx)->device = device; return 0; }
void mecat_b200_destroy(mecat_b200_ctx* ctx) { delete ctx; }
const char* mecat_b200_last_error(mecat_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void mecat_b200_free(mecat_b200_ctx*, void* p) { free(p); }
void mecat_b200_host_free(void* p) { free(p); }
int mecat_b200_get_stats(mecat_b200_ctx*, mecat_b200_stats* out) { memset(out, 0, sizeof *out); return 0; }

int mecat_b200_ref_index_build(mecat_b200_ctx*, const mecat_ref_genome* g, void** refidx) { *refidx = harness_ref_index_build(g); return 0; }
int mecat_b200_ref_index_release(mecat_b200_ctx*, void* refidx) { harness_ref_index_release(refidx); return 0; }
int mecat_b200_ref_map(mecat_b200_ctx* ctx, void* refidx, const mecat_ref_reads* reads, const mecat_ref_params* p, mecat_ref_result** results, size_t* n,
                       char** qstrings, char** sstrings, size_t* string_bytes)
{
	char err[512];
	err[0] = 0;
	harness_set_tech(p->tech);       // which oracle aligner plays the extension kernel (one run uses one technology)
	const int rc = harness_ref_map_indexed(refidx, reads, p, results, n, qstrings, sstrings, string_bytes, err, (int)sizeof err);
	if (rc) ctx->err = err;
	return rc;
}

}  // extern "C"
