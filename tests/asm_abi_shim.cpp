// tests/asm_abi_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The C-ABI entry points the mecat2asmpw command-line driver calls (include/mecat_b200.h), played by the host harness
// (tests/asm_host_harness.cpp: the product's stage sequence and kernel bodies on the host).  tests/util.py links
// mecat_b200/csrc/host/mecat2asmpw.cpp against this file instead of the product library, so the CPU test-suite can run the
// driver itself -- program names, flags, ovlprep, block files, result files -- without a GPU.  Never part of the product.
#include <ctype.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../include/mecat_b200.h"

extern "C" int ah_overlaps(const char* text, int64_t n, const int32_t* starts, const int32_t* lens, int32_t nreads, int32_t first_id,
                           const char* qtext, int64_t qn, const int32_t* qstarts, const int32_t* qlens, int32_t nq, int32_t qfirst,
                           int variant, int maxc, int64_t budget, int divisor, void** out, size_t* nout, int64_t* stats, char* err, int errcap);

struct mecat_b200_ctx { std::string err; };

namespace {
struct Subject { std::string text; std::vector<int32_t> start, len; int32_t first; };
std::string upper(const char* t, int64_t n) { return std::string(t, (size_t)n); }      // the pipeline upper-cases the letters itself
}  // namespace

extern "C" {

int mecat_b200_device_count(void) { return getenv("MECAT_SHIM_DEVICES") ? atoi(getenv("MECAT_SHIM_DEVICES")) : 1; }
int mecat_b200_init(mecat_b200_ctx** ctx, int, void*) { *ctx = new mecat_b200_ctx; return 0; }
void mecat_b200_destroy(mecat_b200_ctx* ctx) { delete ctx; }
const char* mecat_b200_last_error(mecat_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void mecat_b200_free(mecat_b200_ctx*, void* p) { free(p); }
int mecat_b200_get_stats(mecat_b200_ctx*, mecat_b200_stats* out) { memset(out, 0, sizeof *out); return 0; }

int mecat_b200_asm_index_build(mecat_b200_ctx*, const mecat_asm_reads* s, void** asmidx)
{
	Subject* S = new Subject;
	S->text = upper(s->text, s->num_letters);
	S->start.assign(s->read_start, s->read_start + s->num_reads);
	S->len.assign(s->read_len, s->read_len + s->num_reads);
	S->first = s->first_read_id;
	*asmidx = S;
	return 0;
}

int mecat_b200_asm_index_release(mecat_b200_ctx*, void* asmidx) { delete (Subject*)asmidx; return 0; }

int mecat_b200_asm_overlaps(mecat_b200_ctx* ctx, void* asmidx, const mecat_asm_reads* q, const mecat_asm_params* p, mecat_asm_overlap** overlaps, size_t* n)
{
	const Subject* S = (const Subject*)asmidx;
	const std::string qt = upper(q->text, q->num_letters);
	char err[256] = "";
	void* out = NULL;
	const int rc = ah_overlaps(S->text.data(), (int64_t)S->text.size(), S->start.data(), S->len.data(), (int32_t)S->len.size(), S->first, qt.data(),
	                           q->num_letters, q->read_start, q->read_len, q->num_reads, q->first_read_id, p->variant, p->max_candidates, 0, 0, &out, n,
	                           NULL, err, sizeof err);
	if (rc) { ctx->err = err; return rc; }
	*overlaps = (mecat_asm_overlap*)out;
	return 0;
}

}
