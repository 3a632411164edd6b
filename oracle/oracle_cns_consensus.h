// oracle/oracle_cns_consensus.h -- TEST INFRASTRUCTURE ONLY: CPU consensus of one read (oracle_cns_consensus.cpp).
// Uses the record layouts of the public C header (candidates, alignment results, pieces); nothing else of the product.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../include/mecat_b200.h"

namespace orccns {

struct Params { double min_mapping_ratio; int min_align_size; int min_cov; int64_t min_size; int tech = 0; int input_type = 0; };
struct Piece { int64_t id, beg, end; std::string seq; };   // CnsResult, src/common/alignment.h

struct Scratch
{
	struct Impl;
	Impl* p;
	Scratch();
	~Scratch();
	Scratch(const Scratch&) = delete;
	Scratch& operator=(const Scratch&) = delete;
};

// order in which a read's candidates are tried (mecat_correction.cpp:362-370,409)
void sort_candidates(mecat_candidate* c, int n);

// cand[0..ncand): the read's candidates in trial order (at most 200 are looked at), res[i] / strings: the
// GPU's GetAlignment result of candidate i.  Appends the corrected pieces of the read to `out`.
void consensus_one_read(int64_t read_id, int read_size, const mecat_candidate* cand, int ncand, const mecat_align_result* res,
                        const char* qstr, const char* sstr, const Params& P, Scratch& scratch, std::vector<Piece>& out);

}  // namespace orccns
