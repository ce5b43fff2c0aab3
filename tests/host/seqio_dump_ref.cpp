// test driver: dump the records the REFERENCE reader (/root/reference/src/sequence_io.*, compiled where it lies) produces for a file
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include "sequence_io.h"

static std::uint64_t fnv(const std::string& s)
{
    std::uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    const unsigned long skip = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 0;
    try {
        auto r = anyseq::make_sequence_reader(argv[1]);
        if (skip) r->skip(skip);
        while (r->has_next()) {
            auto rec = r->next();
            std::cout << rec.index << '\t' << rec.header << '\t' << rec.data.size() << '\t' << fnv(rec.data) << '\t'
                      << rec.qualities.size() << '\t' << fnv(rec.qualities) << '\n';
        }
    } catch (std::exception& e) {
        std::cout << "EXCEPTION\n";
        return 1;
    }
    return 0;
}
