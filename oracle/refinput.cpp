// refinput.cpp -- TEST INFRASTRUCTURE ONLY (part of the oracle, see anyseq_oracle.h).
//
// Regenerates the inputs the reference CLI produces for `align -r [min [max]]`:
// a default-seeded std::mt19937_64, the query drawn first and then the subject,
// each with a uniformly drawn length in [min,max] followed by one draw in {0..3}
// per symbol mapped to A,C,G,T (reference: src/main.cpp:90-120 and :207-209).
// The draws go through libstdc++'s uniform_int_distribution, instantiated for
// std::size_t (length) and for char (symbols) exactly as the reference does,
// because the produced bytes depend on the distribution's algorithm and on its
// result type.  Built with the same g++/libstdc++ 13.3
// the reference host code compiles with in this image.
#include <cstdint>
#include <random>
#include "anyseq_oracle.h"

namespace {
const char kAlphabet[4] = {'A', 'C', 'G', 'T'};

template <class Rng>
int draw_sequence(std::int64_t lo, std::int64_t hi, Rng& rng, std::uint8_t* out)
{
    std::uniform_int_distribution<std::size_t> len_dist(static_cast<std::size_t>(lo),
                                                        static_cast<std::size_t>(hi));
    const std::size_t len = len_dist(rng);
    std::uniform_int_distribution<char> sym_dist(0, 3);
    for (std::size_t k = 0; k < len; ++k) {
        const char v = sym_dist(rng);
        out[k] = (v >= 0 && v < 4) ? static_cast<std::uint8_t>(kAlphabet[static_cast<int>(v)])
                                   : static_cast<std::uint8_t>('_');
    }
    return static_cast<int>(len);
}
}  // namespace

extern "C" void oracle_reference_random_pair(int64_t minlen, int64_t maxlen,
                                             uint8_t* q, int* m, uint8_t* s, int* n)
{
    if (maxlen < minlen) { int64_t t = minlen; minlen = maxlen; maxlen = t; }  // src/main.cpp:205
    std::mt19937_64 rng;   // default seed, as the reference
    *m = draw_sequence(minlen, maxlen, rng, q);
    *n = draw_sequence(minlen, maxlen, rng, s);
}
