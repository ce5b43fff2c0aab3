// sequence_io.cpp -- see sequence_io.h
#include "sequence_io.h"

namespace anyseq_host {

static bool ends_with(const std::string& s, const char* suf)
{
    const std::string t(suf);
    return s.size() >= t.size() && s.compare(s.size() - t.size(), t.size(), t) == 0;
}

SequenceReader::SequenceReader(const std::string& path, Format fmt) : file_(path.c_str()), fmt_(fmt)
{
    if (!file_.is_open()) {
        valid_ = false;
        throw io_error("can't open file " + path);
    }
}

SequenceRecord SequenceReader::next()
{
    SequenceRecord rec;
    if (!valid_) return rec;
    std::lock_guard<std::mutex> lock(mu_);
    rec.index = ++index_;
    if (fmt_ == Format::fasta) read_fasta(rec); else read_fastq(rec);
    return rec;
}

void SequenceReader::skip(std::uint64_t n)
{
    std::lock_guard<std::mutex> lock(mu_);
    SequenceRecord rec;
    for (; n > 0 && valid_; --n) {
        rec.index = ++index_;
        if (fmt_ == Format::fasta) read_fasta(rec); else read_fastq(rec);
    }
}

void SequenceReader::read_fasta(SequenceRecord& rec)
{
    if (!file_.good()) { valid_ = false; return; }
    std::string line;
    if (pending_header_.empty()) std::getline(file_, line);
    else line.swap(pending_header_);
    if (line.empty() || line[0] != '>')
        throw io_error("malformed fasta file - expected header char > not found");
    rec.header = line.substr(1);
    rec.data.clear();
    while (file_.good()) {
        std::getline(file_, line);
        if (!line.empty() && line[0] == '>') { pending_header_ = line; break; }
        rec.data += line;                        // verbatim: '\r', case, IUPAC codes all kept
    }
    if (rec.data.empty()) throw io_error("malformed fasta file - zero-length sequence: " + rec.header);
    if (!file_.good()) valid_ = false;
}

void SequenceReader::read_fastq(SequenceRecord& rec)
{
    if (!file_.good()) { valid_ = false; return; }
    std::string line;
    std::getline(file_, line);
    if (line.empty()) { valid_ = false; return; }
    if (line[0] != '@') {
        if (line[0] != '\r') throw io_error("malformed fastq file - sequence header: " + line);
        valid_ = false;
        return;
    }
    rec.header = line.substr(1);
    std::getline(file_, rec.data);
    std::getline(file_, line);
    if (line.empty() || line[0] != '+') {
        if (line.empty() || line[0] != '\r') throw io_error("malformed fastq file - quality header: " + line);
        valid_ = false;
        return;
    }
    std::getline(file_, rec.qualities);
}

std::unique_ptr<SequenceReader> make_sequence_reader(const std::string& path)
{
    using F = SequenceReader::Format;
    if (ends_with(path, ".fq") || ends_with(path, ".fnq") || ends_with(path, ".fastq"))
        return std::make_unique<SequenceReader>(path, F::fastq);
    if (ends_with(path, ".fa") || ends_with(path, ".fna") || ends_with(path, ".fasta"))
        return std::make_unique<SequenceReader>(path, F::fasta);
    std::ifstream probe(path.c_str());
    if (!probe.good()) throw io_error("file not accessible");
    std::string line;
    std::getline(probe, line);
    if (!line.empty() && line[0] == '>') return std::make_unique<SequenceReader>(path, F::fasta);
    if (!line.empty() && line[0] == '@') return std::make_unique<SequenceReader>(path, F::fastq);
    throw io_error("file format not recognized");
}

}  // namespace anyseq_host
