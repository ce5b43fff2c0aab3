// alignment_io.cpp -- see alignment_io.h
#include "alignment_io.h"

#include <ostream>

namespace anyseq_host {

namespace {
// symbol k of a row, NUL beyond its end (rows of unequal length are a caller error the reference does not check either)
inline char at(const std::string& row, std::size_t k) { return k < row.size() ? row[k] : '\0'; }
}  // namespace

void print_alignment(std::ostream& os, std::int64_t score, const std::string& q, const std::string& s,
                     std::size_t width)
{
    os << score << '\n';
    if (width == 0) width = 80;
    const std::size_t total = q.size();
    std::string block;                       // three lines + the blank line of one slice, written in one go
    for (std::size_t first = 0; first < total; first += width) {
        const std::size_t len = (total - first < width) ? total - first : width;
        block.assign(3 * (len + 1) + 1, '\n');
        char* top = &block[0];
        char* mid = top + len + 1;
        char* bot = mid + len + 1;
        for (std::size_t c = 0; c < len; ++c) {
            const char a = q[first + c], b = at(s, first + c);
            top[c] = a;
            mid[c] = (a == b) ? '|' : ' ';   // also "blank == blank", like the reference
            bot[c] = b;
        }
        os.write(block.data(), static_cast<std::streamsize>(block.size()));
    }
}

}  // namespace anyseq_host
