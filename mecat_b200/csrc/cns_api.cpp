// mecat_b200/csrc/cns_api.cpp -- test hooks around the host-side consensus (no device code).
// They let the CPU test-suite feed alignments computed elsewhere (the oracle) through exactly the
// code mecat_b200_cns_reads runs after its GPU extensions.  They compute no alignments themselves.
#include <stdlib.h>
#include <string.h>

#include "cns.h"

extern "C" {

void mecat_b200_cns_sort_candidates(mecat_candidate* c, int n) { mbcns::sort_candidates(c, n); }

int mecat_b200_cns_consensus_host(const mecat_candidate* cand, int ncand, const mecat_align_result* res, const char* qstr,
                                  const char* sstr, const mecat_cns_params* p, mecat_cns_piece** pieces, size_t* npieces,
                                  char** seqs, size_t* seq_bytes)
{
	if (!cand || ncand <= 0 || !res || !p || !pieces || !npieces || !seqs || !seq_bytes) return 1;
	mbcns::Params P;
	P.min_mapping_ratio = p->min_mapping_ratio; P.min_align_size = p->min_align_size; P.min_cov = p->min_cov; P.min_size = p->min_size;
	mbcns::Scratch scratch;
	std::vector<mbcns::Piece> out;
	mbcns::consensus_one_read(cand[0].sid, cand[0].ssize, cand, ncand, res, qstr, sstr, P, scratch, out);
	size_t bytes = 0;
	for (auto& pc : out) bytes += pc.seq.size();
	mecat_cns_piece* o = (mecat_cns_piece*)malloc(sizeof(mecat_cns_piece) * (out.size() ? out.size() : 1));
	char* sq = (char*)malloc(bytes + 1);
	if (!o || !sq) { free(o); free(sq); return 1; }
	size_t at = 0;
	for (size_t i = 0; i < out.size(); ++i) {
		o[i].id = out[i].id; o[i].beg = out[i].beg; o[i].end = out[i].end; o[i].seq_offset = (int64_t)at; o[i].seq_len = (int64_t)out[i].seq.size();
		memcpy(sq + at, out[i].seq.data(), out[i].seq.size());
		at += out[i].seq.size();
	}
	sq[bytes] = 0;
	*pieces = o; *npieces = out.size(); *seqs = sq; *seq_bytes = bytes;
	return 0;
}

void mecat_b200_host_free(void* p) { free(p); }

}  // extern "C"
