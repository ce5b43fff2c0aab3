// test driver: the inputs `align -r [min [max]]` of the REFERENCE generates, produced by the reference's own generator
// (random_string / uniform_ACGT_distribution of /root/reference/src/main.cpp, included where it lies with its main()
// renamed).  Prints "len_q fnv_q len_s fnv_s" for each "min max" pair on the command line.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define main anyseq_reference_main
#include "main.cpp"
#undef main

// the six entry points main.cpp references are never called here
extern "C" {
score_t global_alignment_score(const char*, int, const char*, int) { return 0; }
score_t semiglobal_alignment_score(const char*, int, const char*, int) { return 0; }
score_t local_alignment_score(const char*, int, const char*, int) { return 0; }
score_t construct_global_alignment(const char*, int, const char*, int, char*, char*) { return 0; }
score_t construct_semiglobal_alignment(const char*, int, const char*, int, char*, char*) { return 0; }
score_t construct_local_alignment(const char*, int, const char*, int, char*, char*) { return 0; }
}

static std::uint64_t fnv(const std::string& s)
{
    std::uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv)
{
    for (int k = 1; k + 1 < argc; k += 2) {
        std::size_t lo = std::strtoull(argv[k], nullptr, 10), hi = std::strtoull(argv[k + 1], nullptr, 10);
        if (hi < lo) std::swap(lo, hi);                       // src/main.cpp:205
        std::mt19937_64 urng;                                 // src/main.cpp:207: default seed
        const std::string q = random_string(lo, hi, urng);
        const std::string s = random_string(lo, hi, urng);
        std::printf("%zu %016llx %zu %016llx\n", q.size(), (unsigned long long)fnv(q), s.size(), (unsigned long long)fnv(s));
    }
    return 0;
}
