// SIMD tokenizer of the state-label columns of one matrix row (csrc/label_simd.cpp; host code, x86-64 SSSE3 with a run-time
// check, compiled by the host compiler alone).  Used by the row parser of csrc/hostio.cu in front of its scalar loops.
#pragma once
#include <stdint.h>

namespace epi {

// True when the CPU has what parse_labels_simd needs (else the caller keeps to its scalar loops).
bool label_simd_available();

// Labels `d` or `dd`, each followed by a tab, starting at p; p[-3..-1] must be readable (the end of the coordinate fields
// in front of the labels).  Consumes whole 16-byte windows of [p, e) while at least 16 of the `want` label slots of dst are
// still free, writes label-1 for every tab-terminated label in them and returns how many it wrote, with *resume = the
// position behind the last tab it consumed.  Returns -1 if anything in those windows is not of that shape (another
// character, an empty or three-digit field) or a label is outside 1..num_states: the caller then parses the row from the
// start with its careful loop, which knows how to report what is wrong.
int parse_labels_simd(const char* p, const char* e, int want, int num_states, int8_t* dst, const char** resume);

}  // namespace epi
