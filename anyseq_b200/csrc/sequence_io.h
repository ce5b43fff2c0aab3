// sequence_io.h -- minimal FASTA / FASTQ ingestion for the `align` CLI.
//
// Behavioural mirror of the reference reader as far as the hot path's caller
// needs it (src/sequence_io.cpp:62-110,131-163,207-241): records are read one at
// a time; a FASTA record's lines are concatenated verbatim (no case folding, no
// stripping of '\r' -- symbols are compared as raw bytes downstream); the file
// type is taken from the extension (.fa/.fna/.fasta, .fq/.fnq/.fastq) and
// otherwise sniffed from the first character ('>' or '@').
#pragma once

#include <cstdint>
#include <fstream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>

namespace anyseq_host {

struct io_error : std::runtime_error { using std::runtime_error::runtime_error; };

struct SequenceRecord {
    std::uint64_t index = 0;   // 1-based position in the file
    std::string header;
    std::string data;
    std::string qualities;     // FASTQ only
};

class SequenceReader {
public:
    enum class Format { fasta, fastq };
    SequenceReader(const std::string& path, Format fmt);
    bool has_next() const { return valid_; }
    SequenceRecord next();               // thread-safe, like the reference's next()
    void skip(std::uint64_t n);

private:
    void read_fasta(SequenceRecord& rec);
    void read_fastq(SequenceRecord& rec);
    std::ifstream file_;
    Format fmt_;
    std::string pending_header_;
    std::uint64_t index_ = 0;
    bool valid_ = true;
    std::mutex mu_;
};

std::unique_ptr<SequenceReader> make_sequence_reader(const std::string& path);

}  // namespace anyseq_host
