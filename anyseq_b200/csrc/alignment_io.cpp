// alignment_io.cpp -- see alignment_io.h
#include "alignment_io.h"

#include <algorithm>
#include <ostream>

namespace anyseq_host {

void print_alignment(std::ostream& os, std::int64_t score, const std::string& q, const std::string& s,
                     std::size_t width)
{
    os << score << '\n';
    const std::size_t n = q.size();
    for (std::size_t b = 0; b < n; b += width) {
        const std::size_t e = std::min(n, b + width);
        os.write(q.data() + b, static_cast<std::streamsize>(e - b));
        os << '\n';
        std::string bars(e - b, ' ');
        for (std::size_t k = b; k < e; ++k)
            if (q[k] == s[k]) bars[k - b] = '|';
        os << bars << '\n';
        os.write(s.data() + b, static_cast<std::streamsize>(e - b));
        os << "\n\n";
    }
}

}  // namespace anyseq_host
