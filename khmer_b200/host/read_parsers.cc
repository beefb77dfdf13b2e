// FASTA/FASTQ reader for the device feed.  Stands in for the reference's FastxReader over seqan
// (src/oxli/read_parsers.cc:259-372): same record semantics and error behaviour, but block-buffered and able
// to hand out whole batches of reads under one lock instead of one read per spin-lock round trip.
// Plain and gzip input go through zlib's gz* API (it reads uncompressed files transparently); bzip2 input
// through libbz2 loaded at run time.
#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "oxli_b200.hh"

namespace oxli_b200 {
namespace read_parsers {

void Read::reset()
{
    name.clear();
    description.clear();
    sequence.clear();
    quality.clear();
    cleaned_seq.clear();
}

// _to_valid_dna (src/oxli/read_parsers.cc:53-69): ACGT kept, acgt upper-cased, everything else 'A'
void Read::set_clean_seq()
{
    cleaned_seq = sequence;
    for (char& c : cleaned_seq) {
        switch (c) {
        case 'A': case 'C': case 'G': case 'T':
            break;
        case 'a': case 'c': case 'g': case 't':
            c = (char)(c - 32);
            break;
        default:
            c = 'A';
        }
    }
}

namespace {

// minimal libbz2 binding (no bzlib.h in this image; the shared object is there)
struct Bz2 {
    void* lib = nullptr;
    void* (*open)(const char*, const char*) = nullptr;
    int (*read)(void*, void*, int) = nullptr;
    void (*close)(void*) = nullptr;
    const char* (*error)(void*, int*) = nullptr;
    Bz2()
    {
        for (const char* n : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"}) {
            lib = dlopen(n, RTLD_NOW);
            if (lib) break;
        }
        if (!lib) return;
        open = (void* (*)(const char*, const char*))dlsym(lib, "BZ2_bzopen");
        read = (int (*)(void*, void*, int))dlsym(lib, "BZ2_bzread");
        close = (void (*)(void*))dlsym(lib, "BZ2_bzclose");
        error = (const char* (*)(void*, int*))dlsym(lib, "BZ2_bzerror");
        if (!open || !read || !close) lib = nullptr;
    }
};
Bz2& bz2()
{
    static Bz2 b;
    return b;
}

}  // namespace

struct FastxReader::Impl {
    std::string filename;
    gzFile gz = nullptr;
    void* bz = nullptr;
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    bool read_error = false;
    size_t num_reads = 0;
    bool have_qualities = false;
    std::mutex mu;
    std::string pending_error;  // InvalidRead message raised by a batch after its good reads were returned
    int pending_kind = 0;       // 1 InvalidRead, 2 StreamReadError

    bool fill()
    {
        if (eof) return false;
        if (pos < end) {
            memmove(buf.data(), buf.data() + pos, end - pos);
        }
        end -= pos;
        pos = 0;
        if (buf.size() - end < (1u << 16)) buf.resize(buf.size() * 2);
        int want = (int)std::min<size_t>(buf.size() - end, 1u << 30);
        int got;
        if (bz) {
            got = bz2().read(bz, buf.data() + end, want);
            if (got < 0) {
                read_error = true;
                got = 0;
            }
        } else {
            got = gzread(gz, buf.data() + end, (unsigned)want);
            if (got < 0) {
                read_error = true;
                got = 0;
            } else if (got < want) {
                int errnum = 0;
                gzerror(gz, &errnum);
                if (errnum != Z_OK && errnum != Z_STREAM_END) read_error = true;  // truncated / corrupt stream
            }
        }
        if (got == 0) eof = true;
        end += (size_t)got;
        return got > 0;
    }

    // next line without its terminator; false at end of data
    bool get_line(const char*& s, size_t& n)
    {
        while (true) {
            char* nl = (char*)memchr(buf.data() + pos, '\n', end - pos);
            if (nl) {
                s = buf.data() + pos;
                n = (size_t)(nl - s);
                pos += n + 1;
                if (n && s[n - 1] == '\r') n--;
                return true;
            }
            if (!fill()) {
                if (pos < end) {  // last line without newline
                    s = buf.data() + pos;
                    n = end - pos;
                    pos = end;
                    if (n && s[n - 1] == '\r') n--;
                    return true;
                }
                return false;
            }
        }
    }

    int peek()
    {
        while (pos >= end)
            if (!fill()) return -1;
        return (unsigned char)buf[pos];
    }

    void skip_blank()
    {
        while (true) {
            int c = peek();
            if (c == '\n' || c == '\r') pos++;
            else return;
        }
    }

    bool at_end()
    {
        skip_blank();
        return peek() < 0;
    }

    // one record; returns 0 ok, 1 end of stream, 2 malformed
    int next_record(std::string& name, std::string& seq, std::string& qual)
    {
        skip_blank();
        int c = peek();
        if (c < 0) return 1;
        const char* s;
        size_t n;
        if (c == '>') {
            get_line(s, n);
            name.assign(s + 1, n - 1);
            seq.clear();
            qual.clear();
            while (true) {
                int d = peek();
                if (d < 0 || d == '>') break;
                get_line(s, n);
                seq.append(s, n);
            }
            return 0;
        }
        if (c == '@') {
            get_line(s, n);
            name.assign(s + 1, n - 1);
            seq.clear();
            qual.clear();
            while (true) {
                int d = peek();
                if (d < 0) return 2;
                if (d == '+') break;
                get_line(s, n);
                seq.append(s, n);
            }
            get_line(s, n);  // '+' line
            while (qual.size() < seq.size()) {
                if (!get_line(s, n)) break;
                qual.append(s, n);
            }
            return 0;
        }
        return 2;
    }
};

FastxReader::FastxReader(const std::string& filename) : _impl(new Impl())
{
    Impl& m = *_impl;
    m.filename = filename;
    m.buf.resize(1u << 20);
    const std::string bad = "File " + filename + " contains badly formatted sequence or does not exist.";
    FILE* probe = fopen(filename.c_str(), "rb");
    if (!probe) throw InvalidStream(bad);
    unsigned char magic[3] = {0, 0, 0};
    size_t got = fread(magic, 1, 3, probe);
    fclose(probe);
    if (got == 3 && magic[0] == 'B' && magic[1] == 'Z' && magic[2] == 'h') {
        if (!bz2().lib) throw InvalidStream("File " + filename + " is bzip2-compressed and libbz2 is not available.");
        m.bz = bz2().open(filename.c_str(), "rb");
        if (!m.bz) throw InvalidStream(bad);
    } else {
        m.gz = gzopen(filename.c_str(), "rb");
        if (!m.gz) throw InvalidStream(bad);
        gzbuffer(m.gz, 1u << 20);
    }
    // same two checks as FastxReader::_init (read_parsers.cc:259-274)
    if (m.at_end()) {
        if (m.read_error) throw InvalidStream(bad);
        throw InvalidStream("File " + filename + " does not contain any sequences!");
    }
    int c = m.peek();
    if (c != '>' && c != '@') throw InvalidStream(bad);
}

FastxReader::~FastxReader() { close(); }

void FastxReader::close()
{
    Impl& m = *_impl;
    if (m.gz) gzclose(m.gz);
    if (m.bz) bz2().close(m.bz);
    m.gz = nullptr;
    m.bz = nullptr;
    m.eof = true;
    m.pos = m.end = 0;
}

bool FastxReader::is_complete()
{
    std::lock_guard<std::mutex> g(_impl->mu);
    if (_impl->pending_kind) return false;
    return _impl->at_end() && !_impl->read_error;
}

size_t FastxReader::get_num_reads()
{
    return _impl->num_reads;
}

Read FastxReader::get_next_read()
{
    Impl& m = *_impl;
    std::lock_guard<std::mutex> g(m.mu);
    if (m.pending_kind) {
        int k = m.pending_kind;
        std::string msg = m.pending_error;
        m.pending_kind = 0;
        if (k == 1) throw InvalidRead(msg);
        throw StreamReadError();
    }
    Read read;
    int rc = m.next_record(read.name, read.sequence, read.quality);
    if (rc == 1) {
        if (m.read_error) throw StreamReadError();
        throw NoMoreReadsAvailable();
    }
    if (rc == 2 || m.read_error) throw StreamReadError();
    if (m.num_reads == 0 && read.quality.length() != 0) m.have_qualities = true;
    if (read.sequence.length() == 0) throw InvalidRead("Sequence is empty");
    if (m.have_qualities && read.sequence.length() != read.quality.length()) throw InvalidRead("Sequence and quality lengths differ");
    m.num_reads++;
    return read;
}

size_t FastxReader::read_batch(uint64_t max_bases, std::string& seqs, std::vector<uint64_t>& offsets)
{
    Impl& m = *_impl;
    std::lock_guard<std::mutex> g(m.mu);
    if (m.pending_kind) {
        int k = m.pending_kind;
        std::string msg = m.pending_error;
        m.pending_kind = 0;
        if (k == 1) throw InvalidRead(msg);
        throw StreamReadError();
    }
    if (offsets.empty()) offsets.push_back(seqs.size());
    size_t n = 0;
    std::string name, seq, qual;
    const uint64_t start = seqs.size();
    while (seqs.size() - start < max_bases) {
        int rc = m.next_record(name, seq, qual);
        const char* err = nullptr;
        int kind = 0;
        if (rc == 1) {
            if (m.read_error) kind = 2;
            else break;
        } else if (rc == 2 || m.read_error) {
            kind = 2;
        } else {
            if (m.num_reads == 0 && qual.length() != 0) m.have_qualities = true;
            if (seq.length() == 0) { err = "Sequence is empty"; kind = 1; }
            else if (m.have_qualities && seq.length() != qual.length()) { err = "Sequence and quality lengths differ"; kind = 1; }
        }
        if (kind) {
            if (n == 0) {
                if (kind == 1) throw InvalidRead(err);
                throw StreamReadError();
            }
            m.pending_kind = kind;  // the reads before it are consumed first, as the reference would
            m.pending_error = err ? err : "";
            break;
        }
        m.num_reads++;
        seqs += seq;
        offsets.push_back(seqs.size());
        n++;
    }
    return n;
}

template <>
ReadPair ReadParser<FastxReader>::get_next_read_pair()
{
    ReadPair p;
    p.first = get_next_read();
    p.second = get_next_read();
    return p;
}

}  // namespace read_parsers
}  // namespace oxli_b200
