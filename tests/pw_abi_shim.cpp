// tests/pw_abi_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The device-side C-ABI entry points the mecat2pw command-line driver calls (include/mecat_b200.h), played by the CPU
// oracle (oracle/oracle.h: orc_pw_tile is one (index volume, query volume) tile of the reference).  tests/util.py links
// mecat_b200/csrc/host/mecat2pw.cpp and the product's own host I/O (mecat_b200/csrc/host_io.cpp) against this file
// instead of the CUDA library, so the CPU test-suite can run the driver itself -- volume split, the tile schedule over
// several "devices", the wrk/r_N resume protocol, the merged output -- without a GPU.  Never part of the product.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../include/mecat_b200.h"
#include "../mecat_b200/csrc/host/format.h"
#include "../oracle/oracle.h"

struct mecat_b200_ctx { std::string err; int device; };

static std::atomic<long> g_tiles(0), g_index_builds(0);

extern "C" {

int mecat_b200_device_count(void) { return getenv("MECAT_SHIM_DEVICES") ? atoi(getenv("MECAT_SHIM_DEVICES")) : 1; }
int mecat_b200_init(mecat_b200_ctx** ctx, int device, void*) { *ctx = new mecat_b200_ctx; (*ctx)->device = device; return 0; }
void mecat_b200_destroy(mecat_b200_ctx* ctx)
{
	if (ctx && ctx->device == 0 && getenv("MECAT_SHIM_REPORT")) fprintf(stderr, "[shim] tiles=%ld index_builds=%ld\n", g_tiles.load(), g_index_builds.load());
	delete ctx;
}
const char* mecat_b200_last_error(mecat_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void mecat_b200_free(mecat_b200_ctx*, void* p) { free(p); }

int mecat_b200_volume_upload(mecat_b200_ctx*, const mecat_volume* v, void** dvol)
{
	mecat_volume* c = new mecat_volume(*v);        // the driver keeps the host volume loaded while its "device" copy is in use
	*dvol = c;
	return 0;
}
int mecat_b200_volume_release(mecat_b200_ctx*, void* dvol) { delete (mecat_volume*)dvol; return 0; }
int mecat_b200_index_build(mecat_b200_ctx*, void* dvol_ref, void** index) { ++g_index_builds; *index = dvol_ref; return 0; }
int mecat_b200_index_release(mecat_b200_ctx*, void*) { return 0; }

int mecat_b200_pw_tile(mecat_b200_ctx* ctx, void* index, void* dvol_ref, void* dvol_reads, const mecat_pw_params* p, void** records, size_t* n)
{
	static_assert(sizeof(orc_volume) == sizeof(mecat_volume) && sizeof(orc_pw_params) == sizeof(mecat_pw_params), "same plain-data views");
	if (index != dvol_ref) { ctx->err = "tile run against the index of another volume"; return 1; }
	++g_tiles;
	return orc_pw_tile((const orc_volume*)dvol_ref, (const orc_volume*)dvol_reads, (const orc_pw_params*)p, 2, records, n);
}

// the tile as text: the oracle's records through the drivers' host formatter (the device formatter of the product is
// compared with it line by line in tests/test_records_host.py and on the GPU)
int mecat_b200_pw_tile_text(mecat_b200_ctx* ctx, void* index, void* dvol_ref, void* dvol_reads, const mecat_pw_params* p, int gapped,
                            char** text, size_t* bytes, size_t* num_records)
{
	void* rec = NULL;
	size_t n = 0;
	if (mecat_b200_pw_tile(ctx, index, dvol_ref, dvol_reads, p, &rec, &n)) return 1;
	mbfmt::TextBuf b;
	if (p->task == 0) mbfmt::format_candidates(b, (const mecat_candidate*)rec, n);
	else mbfmt::format_m4(b, (const mecat_m4*)rec, n, gapped != 0);
	free(rec);
	char* out = (char*)malloc(b.s.size() + 1);
	memcpy(out, b.s.data(), b.s.size());
	out[b.s.size()] = 0;
	*text = out; *bytes = b.s.size(); *num_records = n;
	return 0;
}

}  // extern "C"
