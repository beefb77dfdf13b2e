// FASTA/FASTQ reader for the device feed.  Stands in for the reference's FastxReader over seqan
// (src/oxli/read_parsers.cc:259-372): same record semantics and error behaviour, but block-buffered and able
// to hand out whole batches of reads under one lock instead of one read per spin-lock round trip.
// Plain and gzip input go through zlib's gz* API (it reads uncompressed files transparently); bzip2 input
// through libbz2 loaded at run time.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <thread>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "oxli_b200.hh"
#include "../../include/kmgpu.h"

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

// Page-locking memory costs milliseconds per call, a batch buffer is needed per bulk call and thread: released buffers are kept
// (a handful, the largest ones) and handed out again.
namespace {
struct PinnedPool {
    std::mutex mu;
    std::vector<std::pair<char*, size_t>> free_list;
    ~PinnedPool()
    {
        // the CUDA context may be gone at exit: leave the pages to the process teardown
    }
    char* take(size_t want, size_t* got)
    {
        std::lock_guard<std::mutex> g(mu);
        for (size_t i = 0; i < free_list.size(); i++)
            if (free_list[i].second >= want) {
                char* p = free_list[i].first;
                *got = free_list[i].second;
                free_list.erase(free_list.begin() + i);
                return p;
            }
        return nullptr;
    }
    bool give(char* p, size_t cap)
    {
        std::lock_guard<std::mutex> g(mu);
        if (free_list.size() >= 8) return false;
        free_list.push_back(std::make_pair(p, cap));
        return true;
    }
};
PinnedPool& pinned_pool()
{
    static PinnedPool* p = new PinnedPool();   // never destroyed: see ~PinnedPool
    return *p;
}
}  // namespace

ReadBatch::~ReadBatch()
{
    if (seqs) {
        if (pinned) {
            if (!pinned_pool().give(seqs, cap)) kmgpu_free_pinned(seqs);
        } else {
            free(seqs);
        }
    }
}

void ReadBatch::reserve(size_t n_want_bases)
{
    const size_t n = pack ? (n_want_bases + 31) / 32 * 8 + 16 : n_want_bases;   // bytes
    if (n <= cap) return;
    size_t want = std::max(n, cap + cap / 2);
    void* np = nullptr;
    size_t got = 0;
    bool np_pinned = false;
    if (char* pooled = pinned_pool().take(want, &got)) {
        np = pooled;
        want = got;
        np_pinned = true;
    } else {
        np_pinned = kmgpu_alloc_pinned(want, &np) == 0 && np;   // no device: plain memory still parses
        if (!np_pinned) np = malloc(want);
    }
    if (!np) throw std::bad_alloc();
    if (n_bases) memcpy(np, seqs, pack ? (n_bases + 31) / 32 * 8 : n_bases);
    if (seqs) {
        if (pinned) {
            if (!pinned_pool().give(seqs, cap)) kmgpu_free_pinned(seqs);
        } else {
            free(seqs);
        }
    }
    seqs = (char*)np;
    cap = want;
    pinned = np_pinned;
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

namespace {

// cleaned 2-bit code of a sequence byte: Read::set_clean_seq (ACGT kept, acgt upper-cased, anything else 'A') followed by
// twobit_repr (kmer_hash.hh:70-72)
struct PackLut {
    uint8_t v[256];
    PackLut()
    {
        memset(v, 0, sizeof v);
        v[(unsigned char)'T'] = v[(unsigned char)'t'] = 1;
        v[(unsigned char)'C'] = v[(unsigned char)'c'] = 2;
        v[(unsigned char)'G'] = v[(unsigned char)'g'] = 3;
    }
};
const PackLut g_pack_lut;

#if defined(__x86_64__)
// 32 sequence bytes -> 64 bits, first base in the two most significant bits; bytes outside ACGTacgt give 0 ('A', as the cleaning
// does).  (c >> 1) & 3 separates the four letters in either case (A 0, C 1, T 2, G 3); a byte shuffle turns that into khmer's code.
__attribute__((target("avx2"))) inline uint64_t pack32_avx2(const char* p)
{
    const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
    const __m256i t = _mm256_and_si256(_mm256_srli_epi16(x, 1), _mm256_set1_epi8(3));
    const __m256i lut = _mm256_setr_epi8(0, 2, 1, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i up = _mm256_and_si256(x, _mm256_set1_epi8((char)0xDF));
    const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(up, _mm256_set1_epi8('C'))),
                                       _mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(up, _mm256_set1_epi8('T'))));
    const __m256i code = _mm256_and_si256(_mm256_shuffle_epi8(lut, t), ok);
    const __m256i nib = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0104));          // pairs of bases: first * 4 + second
    const __m256i byt = _mm256_madd_epi16(nib, _mm256_set1_epi32(0x00010010));         // pairs of those: first * 16 + second
    // one byte (four bases) per 32-bit lane; gather them, first lane last, so that each half reads as a big-endian group
    const __m256i sel = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                         -1);
    const __m256i g = _mm256_shuffle_epi8(byt, sel);
    const uint32_t hi = (uint32_t)_mm256_extract_epi32(g, 0), lo = (uint32_t)_mm256_extract_epi32(g, 4);
    return ((uint64_t)hi << 32) | lo;
}
inline bool have_avx2()
{
    static const bool v = __builtin_cpu_supports("avx2") && !(getenv("KMGPU_NO_SIMD") && *getenv("KMGPU_NO_SIMD") == '1');
    return v;
}
#endif

// Appends bases to a 2-bit stream at base position `at`.  Several packers may work on disjoint base ranges of one stream:
// words that a range shares with its neighbours (the first and the last one it touches) are OR-ed in atomically — they
// are zeroed, or hold an earlier range's bits, before the packers start — all others are plain stores.
struct Packer {
    uint64_t* words;
    size_t at, first_word;
    uint64_t acc = 0;
    Packer(uint64_t* w, size_t start) : words(w), at(start), first_word(start >> 5) {}
    inline void flush_word(size_t w, bool shared)
    {
        if (shared) __atomic_fetch_or(&words[w], acc, __ATOMIC_RELAXED);
        else words[w] = acc;
        acc = 0;
    }
    inline void add_scalar(const char* p, size_t n)
    {
        const uint8_t* lut = g_pack_lut.v;
        for (size_t i = 0; i < n; i++) {
            const unsigned sh = 62 - 2 * (unsigned)(at & 31);
            acc |= (uint64_t)lut[(unsigned char)p[i]] << sh;
            at++;
            if ((at & 31) == 0) flush_word((at >> 5) - 1, (at >> 5) - 1 == first_word);
        }
    }
#if defined(__x86_64__)
    __attribute__((target("avx2"))) void add_avx2(const char* p, size_t n)
    {
        size_t i = 0;
        for (; i + 32 <= n; i += 32) {   // 32 bases complete the current word, whatever the offset within it
            const uint64_t v = pack32_avx2(p + i);
            const unsigned off = (unsigned)(at & 31);
            acc |= v >> (2 * off);
            const size_t w = at >> 5;
            flush_word(w, w == first_word);
            acc = off ? v << (64 - 2 * off) : 0;
            at += 32;
        }
        add_scalar(p + i, n - i);
    }
#endif
    inline void add(const char* p, size_t n)
    {
#if defined(__x86_64__)
        if (n >= 32 && have_avx2()) {
            add_avx2(p, n);
            return;
        }
#endif
        add_scalar(p, n);
    }
    inline void finish()
    {
        if (at & 31) flush_word(at >> 5, true);
    }
};

// a few long-lived worker threads: a batch is parsed by all of them twice (locate, then copy / pack), and a thread start per
// slice and pass costs as much as the work on small batches
class Workers {
public:
    explicit Workers(unsigned n) : n_(n)
    {
        for (unsigned t = 1; t < n_; t++) th_.emplace_back([this, t] { loop(t); });
    }
    ~Workers()
    {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto& x : th_) x.join();
    }
    // runs fn(0) .. fn(n - 1), fn(0) on the calling thread
    void run(const std::function<void(unsigned)>& fn)
    {
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &fn;
            left_ = n_ - 1;
            gen_++;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> g(mu_);
        done_.wait(g, [this] { return left_ == 0; });
        fn_ = nullptr;
    }
    unsigned size() const { return n_; }

private:
    void loop(unsigned t)
    {
        uint64_t seen = 0;
        while (true) {
            const std::function<void(unsigned)>* fn;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(t);
            {
                std::lock_guard<std::mutex> g(mu_);
                left_--;
            }
            done_.notify_one();
        }
    }
    unsigned n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(unsigned)>* fn_ = nullptr;
    unsigned left_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

}  // namespace

namespace {

unsigned parse_threads()
{
    static unsigned n = [] {
        const char* e = getenv("KMGPU_PARSE_THREADS");
        unsigned v = e && *e ? (unsigned)atoi(e) : std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        return std::max(1u, std::min(v, 64u));
    }();
    return n;
}

// Decompressed bytes of a gzip / bzip2 file, produced ahead of the parser by a thread of its own: inflating costs several times
// more than parsing, so it runs while the previous batch is parsed, packed and counted.  BGZF files (bgzip, the block-compressed
// gzip of htslib: members of at most 64 KB whose header carries the member's compressed size) are inflated by all parser
// threads at once; ordinary gzip has no block index and stays one sequential inflate (the reference reads it through seqan's
// single zlib stream, src/oxli/read_parsers.cc:259-372).
class Inflater {
public:
    static constexpr size_t BLOCK = 8u << 20;   // decompressed bytes per hand-over
    static constexpr size_t AHEAD = 8;          // blocks produced ahead of the parser

    Inflater(gzFile gz, void* bz, const std::string& path, bool bgzf) : gz_(gz), bz_(bz)
    {
        if (bgzf) {
            fp_ = fopen(path.c_str(), "rb");
            if (fp_) setvbuf(fp_, nullptr, _IOFBF, 4u << 20);
        }
        th_ = std::thread([this] { fp_ ? run_bgzf() : run_stream(); });
    }
    ~Inflater()
    {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_room_.notify_all();
        th_.join();
        if (fp_) fclose(fp_);
    }
    // next block in stream order into `out`; false at the end of the stream, *err set if it ended in an error
    bool next(std::vector<char>& out, bool* err)
    {
        std::unique_lock<std::mutex> g(mu_);
        cv_data_.wait(g, [this] { return !ready_.empty() || done_; });
        if (ready_.empty()) {
            *err = error_;
            return false;
        }
        out.swap(ready_.front());
        ready_.pop_front();
        g.unlock();
        cv_room_.notify_one();
        return true;
    }
    void recycle(std::vector<char>& v)
    {
        std::lock_guard<std::mutex> g(mu_);
        if (spare_.size() < AHEAD) spare_.emplace_back(std::move(v));
    }
    // is this file BGZF?  (first member: gzip, FEXTRA set, a 'B','C' subfield of two bytes)
    static bool is_bgzf(const std::string& path)
    {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        unsigned char h[4096];
        const size_t n = fread(h, 1, sizeof h, f);
        fclose(f);
        uint32_t bsize = 0, xlen = 0;
        return parse_header(h, n, &bsize, &xlen);
    }

private:
    static bool parse_header(const unsigned char* h, size_t n, uint32_t* block_size, uint32_t* xlen_out)
    {
        if (n < 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4) || (h[3] & ~4u)) return false;
        const uint32_t xlen = h[10] | (h[11] << 8);
        if (12 + (size_t)xlen > n) return false;
        for (uint32_t o = 0; o + 4 <= xlen;) {
            const unsigned char* sf = h + 12 + o;
            const uint32_t slen = sf[2] | (sf[3] << 8);
            if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && o + 6 <= xlen) {
                *block_size = (uint32_t)(sf[4] | (sf[5] << 8)) + 1;
                *xlen_out = xlen;
                return *block_size >= 12 + xlen + 8;
            }
            o += 4 + slen;
        }
        return false;
    }
    std::vector<char> fresh()
    {
        std::lock_guard<std::mutex> g(mu_);
        if (!spare_.empty()) {
            std::vector<char> v = std::move(spare_.back());
            spare_.pop_back();
            return v;
        }
        return std::vector<char>();
    }
    // false when the reader has gone away
    bool push(std::vector<char>& v)
    {
        std::unique_lock<std::mutex> g(mu_);
        cv_room_.wait(g, [this] { return ready_.size() < AHEAD || stop_; });
        if (stop_) return false;
        ready_.emplace_back(std::move(v));
        g.unlock();
        cv_data_.notify_one();
        return true;
    }
    void finish(bool error)
    {
        {
            std::lock_guard<std::mutex> g(mu_);
            done_ = true;
            error_ = error;
        }
        cv_data_.notify_all();
    }
    void run_stream()
    {
        bool error = false;
        while (true) {
            std::vector<char> v = fresh();
            v.resize(BLOCK);
            int got;
            if (bz_) {
                got = bz2().read(bz_, v.data(), (int)BLOCK);
                if (got < 0) { error = true; got = 0; }
            } else {
                got = gzread(gz_, v.data(), (unsigned)BLOCK);
                if (got < 0) {
                    error = true;
                    got = 0;
                } else if ((size_t)got < BLOCK) {
                    int errnum = 0;
                    gzerror(gz_, &errnum);
                    if (errnum != Z_OK && errnum != Z_STREAM_END) error = true;   // truncated / corrupt stream
                }
            }
            if (got == 0) break;
            v.resize((size_t)got);
            if (!push(v)) return;
            if (error) break;
        }
        finish(error);
    }
    struct Member {
        size_t c_off, c_len, out_off;
        uint32_t isize, crc;
    };
    void run_bgzf()
    {
        const unsigned T = parse_threads();
        Workers pool(T);
        std::vector<z_stream> zs(T);
        for (auto& z : zs) {
            memset(&z, 0, sizeof z);
            inflateInit2(&z, -15);
        }
        std::vector<unsigned char> comp;
        std::vector<Member> mem;
        bool error = false, eof = false;
        while (!eof && !error) {
            comp.clear();
            mem.clear();
            size_t total = 0;
            while (total < BLOCK) {
                unsigned char h[12];
                const size_t n = fread(h, 1, 12, fp_);
                if (n == 0) { eof = true; break; }
                if (n < 12 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { error = true; break; }
                const uint32_t xlen = h[10] | (h[11] << 8);
                const size_t at = comp.size();
                comp.resize(at + 12 + xlen);
                memcpy(comp.data() + at, h, 12);
                if (fread(comp.data() + at + 12, 1, xlen, fp_) != xlen) { error = true; break; }
                uint32_t bsize = 0, xl = 0;
                if (!parse_header(comp.data() + at, std::max<size_t>(18, 12 + xlen), &bsize, &xl)) { error = true; break; }
                const size_t rest = bsize - 12 - xlen;
                comp.resize(at + bsize);
                if (fread(comp.data() + at + 12 + xlen, 1, rest, fp_) != rest) { error = true; break; }
                const unsigned char* tr = comp.data() + at + bsize - 8;
                Member m;
                m.c_off = at + 12 + xlen;
                m.c_len = rest - 8;
                m.crc = tr[0] | (tr[1] << 8) | (tr[2] << 16) | ((uint32_t)tr[3] << 24);
                m.isize = tr[4] | (tr[5] << 8) | (tr[6] << 16) | ((uint32_t)tr[7] << 24);
                if (m.isize > (1u << 20)) { error = true; break; }   // a BGZF member holds at most 64 KB: a damaged trailer
                m.out_off = total;
                total += m.isize;
                mem.push_back(m);
            }
            if (mem.empty() || total == 0) continue;
            std::vector<char> v = fresh();
            v.resize(total);
            std::atomic<size_t> nextm(0);
            std::atomic<bool> bad(false);
            pool.run([&](unsigned t) {
                z_stream& z = zs[t];
                for (size_t i = nextm.fetch_add(1); i < mem.size(); i = nextm.fetch_add(1)) {
                    const Member& m = mem[i];
                    if (m.isize == 0) continue;
                    inflateReset2(&z, -15);
                    z.next_in = comp.data() + m.c_off;
                    z.avail_in = (uInt)m.c_len;
                    z.next_out = (Bytef*)v.data() + m.out_off;
                    z.avail_out = m.isize;
                    const int rc = inflate(&z, Z_FINISH);
                    if (rc != Z_STREAM_END || z.avail_out != 0 ||
                        crc32(crc32(0L, Z_NULL, 0), (const Bytef*)v.data() + m.out_off, m.isize) != m.crc)
                        bad.store(true);
                }
            });
            if (bad.load()) {
                error = true;
                break;
            }
            if (!push(v)) break;
        }
        for (auto& z : zs) inflateEnd(&z);
        finish(error);
    }

    gzFile gz_;
    void* bz_;
    FILE* fp_ = nullptr;
    std::thread th_;
    std::mutex mu_;
    std::condition_variable cv_data_, cv_room_;
    std::deque<std::vector<char>> ready_;
    std::vector<std::vector<char>> spare_;
    bool done_ = false, error_ = false, stop_ = false;
};

}  // namespace

namespace {
struct Seg {            // a piece of sequence inside the mapping
    const char* p;
    uint32_t n;
    uint32_t last;       // 1: this piece ends a record
};

struct Slice {
    const char* a;
    const char* b;
    std::vector<Seg> segs;
    size_t bytes = 0, reads = 0;
    bool ok = true;
};

}  // namespace

struct FastxReader::Impl {
    std::vector<Slice> slices;   // per-thread scratch of the parallel batch parser, reused
    std::string filename;
    gzFile gz = nullptr;
    void* bz = nullptr;
    std::unique_ptr<Inflater> inflater;   // compressed input: blocks inflated ahead by its thread(s)
    std::vector<char> spare;
    std::vector<char> buf;
    char* base = nullptr;      // parse window: buf.data() for compressed input, the mapping for plain files
    size_t pos = 0, end = 0;
    bool eof = false;
    // plain files are memory-mapped: no copy, and whole batches can be parsed by several threads
    void* map = nullptr;
    size_t map_len = 0;
    char kind = 0;             // '>' or '@' (first record marker), for the parallel path
    bool read_error = false;
    size_t num_reads = 0;
    bool have_qualities = false;
    std::mutex mu;
    std::unique_ptr<Workers> workers;   // parser threads of the parallel batch path (created on first use)
    std::string pending_error;  // InvalidRead message raised by a batch after its good reads were returned
    int pending_kind = 0;       // 1 InvalidRead, 2 StreamReadError

    bool fill()
    {
        if (eof) return false;
        if (pos && pos < end) memmove(buf.data(), buf.data() + pos, end - pos);
        end -= pos;
        pos = 0;
        if (inflater) {
            bool err = false;
            spare.clear();
            const bool have = inflater->next(spare, &err);
            if (err) read_error = true;
            const size_t got = have ? spare.size() : 0;
            if (buf.size() - end < got) buf.resize(std::max(buf.size() * 2, end + got));
            base = buf.data();
            if (got) memcpy(buf.data() + end, spare.data(), got);
            if (have) inflater->recycle(spare);
            if (got == 0) eof = true;
            end += got;
            return got > 0;
        }
        if (buf.size() - end < (1u << 16)) buf.resize(buf.size() * 2);
        base = buf.data();
        int want = (int)std::min<size_t>(buf.size() - end, 1u << 30);
        int got = gzread(gz, buf.data() + end, (unsigned)want);
        if (got < 0) {
            read_error = true;
            got = 0;
        } else if (got < want) {
            int errnum = 0;
            gzerror(gz, &errnum);
            if (errnum != Z_OK && errnum != Z_STREAM_END) read_error = true;  // truncated / corrupt stream
        }
        if (got == 0) eof = true;
        end += (size_t)got;
        return got > 0;
    }

    // next line without its terminator; false at end of data
    bool get_line(const char*& s, size_t& n)
    {
        while (true) {
            char* nl = (char*)memchr(base + pos, '\n', end - pos);
            if (nl) {
                s = base + pos;
                n = (size_t)(nl - s);
                pos += n + 1;
                if (n && s[n - 1] == '\r') n--;
                return true;
            }
            if (!fill()) {
                if (pos < end) {  // last line without newline
                    s = base + pos;
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
        return (unsigned char)base[pos];
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
        const char* s = nullptr;
        size_t n = 0;
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
        m.inflater.reset(new Inflater(nullptr, m.bz, filename, false));
    } else if (got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        m.gz = gzopen(filename.c_str(), "rb");
        if (!m.gz) throw InvalidStream(bad);
        gzbuffer(m.gz, 1u << 20);
        m.inflater.reset(new Inflater(m.gz, nullptr, filename, Inflater::is_bgzf(filename)));
    } else {
        // plain file: map it; the serial routines parse straight out of the mapping (no refills), and
        // read_batch can hand slices of it to several threads
        int fd = open(filename.c_str(), O_RDONLY);
        if (fd < 0) throw InvalidStream(bad);
        struct stat sb;
        if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
            ::close(fd);
            m.gz = gzopen(filename.c_str(), "rb");   // pipes and the like: stream through zlib's transparent mode
            if (!m.gz) throw InvalidStream(bad);
        } else {
            m.map_len = (size_t)sb.st_size;
            if (m.map_len) {
                m.map = mmap(nullptr, m.map_len, PROT_READ, MAP_PRIVATE, fd, 0);
                ::close(fd);
                if (m.map == MAP_FAILED) {
                    m.map = nullptr;
                    throw InvalidStream(bad);
                }
                madvise(m.map, m.map_len, MADV_SEQUENTIAL);
            } else {
                ::close(fd);
            }
            m.base = (char*)m.map;
            m.pos = 0;
            m.end = m.map_len;
            m.eof = true;   // nothing to refill
        }
    }
    // same two checks as FastxReader::_init (read_parsers.cc:259-274)
    if (m.at_end()) {
        if (m.read_error) throw InvalidStream(bad);
        throw InvalidStream("File " + filename + " does not contain any sequences!");
    }
    int c = m.peek();
    if (c != '>' && c != '@') throw InvalidStream(bad);
    m.kind = (char)c;
}

FastxReader::~FastxReader() { close(); }

void FastxReader::close()
{
    Impl& m = *_impl;
    m.inflater.reset();   // joins the inflating thread before its stream goes away
    if (m.gz) gzclose(m.gz);
    if (m.bz) bz2().close(m.bz);
    m.gz = nullptr;
    m.bz = nullptr;
    if (m.map) munmap(m.map, m.map_len);
    m.map = nullptr;
    m.base = m.buf.data();
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

namespace {

// first record start at or after p in [p, e): a line beginning with `kind`; for FASTQ the line two below must begin
// with '+' (a quality line may begin with '@', but then the line two below it is a sequence line)
const char* find_record(const char* b, const char* p, const char* e, char kind)
{
    while (p < e) {
        if (p != b) {   // move to the next line start unless we are at the very beginning of the data
            const char* nl = (const char*)memchr(p - 1, '\n', (size_t)(e - p + 1));
            if (!nl) return e;
            p = nl + 1;
            if (p >= e) return e;
        }
        if (*p == kind) {
            if (kind == '>') return p;
            const char* l2 = (const char*)memchr(p, '\n', (size_t)(e - p));
            const char* l3 = l2 ? (const char*)memchr(l2 + 1, '\n', (size_t)(e - l2 - 1)) : nullptr;
            if (!l3 || l3 + 1 >= e) return p;   // too close to the end to check: the caller only cuts far from it
            if (l3[1] == '+') return p;
        }
        p++;
    }
    return e;
}

// all records of [a, b) (a is a record start, b the next slice's record start or the end of the data); sequences are
// only located here, the copy into the batch happens once every slice's size is known
void parse_slice(Slice& s, char kind)
{
    const char* p = s.a;
    const char* e = s.b;
    s.segs.clear();
    s.bytes = s.reads = 0;
    s.ok = true;
    while (p < e) {
        while (p < e && (*p == '\n' || *p == '\r')) p++;
        if (p >= e) break;
        if (*p != kind) { s.ok = false; return; }
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        if (!nl) { s.ok = false; return; }   // header without a sequence
        p = nl + 1;
        size_t len = 0;
        if (kind == '>') {
            size_t first = s.segs.size();
            while (p < e && *p != '>') {
                nl = (const char*)memchr(p, '\n', (size_t)(e - p));
                const char* le = nl ? nl : e;
                size_t n = (size_t)(le - p);
                if (n && p[n - 1] == '\r') n--;
                if (n) {
                    if (n > 0xFFFFFFFFull) { s.ok = false; return; }
                    s.segs.push_back(Seg{p, (uint32_t)n, 0});
                    len += n;
                }
                p = nl ? nl + 1 : e;
            }
            if (len == 0 || s.segs.size() == first) { s.ok = false; return; }
            s.segs.back().last = 1;
        } else {
            // four-line FASTQ only; anything else is left to the serial parser
            nl = (const char*)memchr(p, '\n', (size_t)(e - p));
            if (!nl) { s.ok = false; return; }
            size_t n = (size_t)(nl - p);
            if (n && p[n - 1] == '\r') n--;
            const char* sp = p;
            p = nl + 1;
            if (p >= e || *p != '+') { s.ok = false; return; }
            nl = (const char*)memchr(p, '\n', (size_t)(e - p));
            if (!nl) { s.ok = false; return; }
            p = nl + 1;
            nl = (const char*)memchr(p, '\n', (size_t)(e - p));
            const char* le = nl ? nl : e;
            size_t q = (size_t)(le - p);
            if (q && p[q - 1] == '\r') q--;
            if (q != n || n == 0 || n > 0xFFFFFFFFull) { s.ok = false; return; }
            p = nl ? nl + 1 : e;
            s.segs.push_back(Seg{sp, (uint32_t)n, 1});
            len = n;
        }
        s.bytes += len;
        s.reads++;
    }
}

}  // namespace

// plain (memory-mapped) files and the inflated window of compressed ones: cut the next window into slices at record
// boundaries and parse them concurrently.  Any irregularity (empty sequence, quality length mismatch, multi-line FASTQ, stray bytes) makes the
// whole window fall back to the serial parser, which then reports it exactly as the reference would.
size_t FastxReader::read_batch(uint64_t max_bases, ReadBatch& out)
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
    if (out.offsets.empty()) out.offsets.push_back(out.n_bases);
    static const size_t par_min = [] {
        const char* e = getenv("KMGPU_PARSE_MIN_BYTES");
        return e && *e ? (size_t)strtoull(e, nullptr, 10) : (size_t)(4u << 20);
    }();
    // window: enough bytes for ~max_bases bases (FASTQ spends about half its bytes on qualities)
    const uint64_t want = m.kind == '@' ? max_bases * 2 + max_bases / 4 : max_bases + max_bases / 8;
    if (m.inflater && parse_threads() > 1)   // compressed input: gather the window (and the margin the cut below looks into) first
        while (!m.eof && m.end - m.pos < want + (2u << 20))
            if (!m.fill()) break;
    if ((m.map || m.inflater) && parse_threads() > 1 && m.end - m.pos > par_min) {
        const unsigned T = parse_threads();
        const char* b = m.base;
        const char* w0 = b + m.pos;
        const char* fe = b + m.end;
        const char* w1 = (uint64_t)(fe - w0) <= want + (1u << 20) ? fe : find_record(b, w0 + want, fe, m.kind);
        std::vector<Slice>& sl = m.slices;
        sl.resize(T);
        std::vector<const char*> cut(T + 1);
        cut[0] = w0;
        cut[T] = w1;
        for (unsigned t = 1; t < T; t++) cut[t] = find_record(b, w0 + (uint64_t)(w1 - w0) * t / T, w1, m.kind);
        for (unsigned t = 1; t <= T; t++)
            if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
        if (!m.workers || m.workers->size() != T) m.workers.reset(new Workers(T));
        for (unsigned t = 0; t < T; t++) {
            sl[t].a = cut[t];
            sl[t].b = cut[t + 1];
        }
        const char kind = m.kind;
        m.workers->run([&sl, kind](unsigned t) { parse_slice(sl[t], kind); });
        bool ok = true;
        size_t total = 0, nreads = 0;
        for (auto& x : sl) {
            ok = ok && x.ok;
            total += x.bytes;
            nreads += x.reads;
        }
        if (ok && nreads) {
            if (m.num_reads == 0 && m.kind == '@') m.have_qualities = true;
            const size_t base0 = out.n_bases;
            out.reserve(base0 + total);
            const size_t o0 = out.offsets.size();
            out.offsets.resize(o0 + nreads);
            std::vector<size_t> sb(T), rb(T);
            size_t accb = base0, accr = o0;
            for (unsigned t = 0; t < T; t++) {
                sb[t] = accb;
                rb[t] = accr;
                accb += sl[t].bytes;
                accr += sl[t].reads;
            }
            char* dst = out.seqs;
            uint64_t* offp = out.offsets.data();
            if (out.pack) {
                // words shared by two slices (and the tail word) start from zero — except the one the batch already ends in
                uint64_t* w = out.words();
                const size_t keep = base0 & 31 ? base0 >> 5 : (size_t)-1;
                for (unsigned t = 0; t <= T; t++) {
                    const size_t wi = (t < T ? sb[t] : base0 + total) >> 5;
                    if (wi != keep) w[wi] = 0;
                }
                m.workers->run([&](unsigned t) {
                    size_t r = rb[t];
                    Packer pk(w, sb[t]);
                    for (const Seg& sg : sl[t].segs) {
                        pk.add(sg.p, sg.n);
                        if (sg.last) offp[r++] = pk.at;
                    }
                    pk.finish();
                });
            } else {
                m.workers->run([&](unsigned t) {
                    size_t o = sb[t], r = rb[t];
                    for (const Seg& sg : sl[t].segs) {
                        memcpy(dst + o, sg.p, sg.n);
                        o += sg.n;
                        if (sg.last) offp[r++] = o;
                    }
                });
            }
            out.n_bases = base0 + total;
            m.pos = (size_t)(w1 - b);
            m.num_reads += nreads;
            return nreads;
        }
        // fall through: the serial parser handles this window
    }
    size_t n = 0;
    std::string name, seq, qual;
    const uint64_t start = out.n_bases;
    while (out.n_bases - start < max_bases) {
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
        out.reserve(out.n_bases + seq.size() + 32);
        if (out.pack) {
            // serial: the word the batch ends in keeps its bits, the words this read reaches for the first time start from zero
            uint64_t* w = out.words();
            const size_t w_first = (out.n_bases + 31) >> 5, w_last = (out.n_bases + seq.size()) >> 5;
            for (size_t wi = w_first; wi <= w_last; wi++) w[wi] = 0;
            const uint8_t* lut = g_pack_lut.v;
            for (size_t i = 0; i < seq.size(); i++) {
                const size_t at = out.n_bases + i;
                w[at >> 5] |= (uint64_t)lut[(unsigned char)seq[i]] << (62 - 2 * (unsigned)(at & 31));
            }
        } else {
            memcpy(out.seqs + out.n_bases, seq.data(), seq.size());
        }
        out.n_bases += seq.size();
        out.offsets.push_back(out.n_bases);
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
