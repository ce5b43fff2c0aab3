// alignment_io.h -- text rendering of an alignment pair.
// Format of the reference's print_alignment (src/alignment_io.cpp:13-38): the
// score on its own line, then for every block of `width` columns the query
// slice, a line with '|' wherever the two rows hold the same byte (blank and
// gap columns included, as the reference does), the subject slice and an empty line.
#pragma once

#include <cstdint>
#include <iosfwd>
#include <string>

namespace anyseq_host {

void print_alignment(std::ostream& os, std::int64_t score, const std::string& q, const std::string& s,
                     std::size_t width = 80);

}  // namespace anyseq_host
