// test driver: print_alignment of the REFERENCE (/root/reference/src/alignment_io.*, compiled where it lies) for "score|q|s|width" lines on stdin
#include <iostream>
#include <sstream>
#include <string>
#include "alignment_io.h"

int main()
{
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line);
        std::string score, q, s, w;
        std::getline(is, score, '|'); std::getline(is, q, '|'); std::getline(is, s, '|'); std::getline(is, w, '|');
        std::cout << "<<<\n";
        anyseq::print_alignment(std::cout, std::stoll(score), q, s, static_cast<std::size_t>(std::stoul(w)));
        std::cout << ">>>\n";
    }
    return 0;
}
