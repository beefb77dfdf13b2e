// TEST INFRASTRUCTURE — not part of the product.
//
// Flat C wrapper around the UNMODIFIED reference liboxli (compiled from the
// sources where they lie under /root/reference by oracle/Makefile, outputs in
// oracle/_ref/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference arms may load the resulting
// oracle/_ref/libkhmer_ref.so.  Nothing in khmer_b200/ links or loads it.
//
// The wrapper follows the reference call sites it stands in for:
//   scripts/load-into-counting.py:145-158  (T threads share one parser, each in consume_seqfile)
//   khmer/_oxli/graphs.pyx:230-239,282-296 (consume_seqfile, abundance_distribution)
//   khmer/_oxli/graphs.pyx:172-185         (get_median_count, median_at_least)
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <memory>

#include "oxli/oxli.hh"
#include "oxli/hashtable.hh"
#include "oxli/hashgraph.hh"
#include "oxli/hllcounter.hh"
#include "oxli/kmer_hash.hh"
#include "oxli/read_parsers.hh"
#include "oxli/storage.hh"

using namespace oxli;
using namespace oxli::read_parsers;

static thread_local std::string g_err;

#define REF_TRY try {
#define REF_CATCH(rv)                                                      \
    }                                                                      \
    catch (oxli_file_exception & e) { g_err = std::string("file: ") + e.what(); return (rv); } \
    catch (oxli_exception & e) { g_err = std::string("oxli: ") + e.what(); return (rv); }      \
    catch (std::exception & e) { g_err = std::string("std: ") + e.what(); return (rv); }

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// kind: 0 Countgraph 1 SmallCountgraph 2 Nodegraph 3 Counttable 4 SmallCounttable 5 Nodetable
void* ref_new(int kind, int k, const uint64_t* sizes, int n)
{
    std::vector<uint64_t> v(sizes, sizes + n);
    REF_TRY
    switch (kind) {
    case 0: return new Countgraph(k, v);
    case 1: return new SmallCountgraph(k, v);
    case 2: return new Nodegraph(k, v);
    case 3: return new Counttable(k, v);
    case 4: return new SmallCounttable(k, v);
    case 5: return new Nodetable(k, v);
    }
    g_err = "bad kind";
    return nullptr;
    REF_CATCH(nullptr)
}

void ref_free(void* h) { delete (Hashtable*)h; }

int ref_primes(uint32_t n, uint64_t x, uint64_t* out)
{
    std::vector<uint64_t> p = get_n_primes_near_x(n, x);
    for (size_t i = 0; i < p.size(); i++) out[i] = p[i];
    return (int)p.size();
}

int ref_set_use_bigcount(void* h, int on)
{
    REF_TRY
    ((Hashtable*)h)->set_use_bigcount(on != 0);
    return 0;
    REF_CATCH(-1)
}

int ref_consume_seqfile(void* h, const char* fn, int threads, uint64_t* reads, uint64_t* kmers)
{
    REF_TRY
    Hashtable* ht = (Hashtable*)h;
    ReadParserPtr<FastxReader> parser = get_parser<FastxReader>(fn);
    unsigned int total_reads = 0;
    unsigned long long n_consumed = 0;
    if (threads <= 1) {
        ht->consume_seqfile<FastxReader>(parser, total_reads, n_consumed);
    } else {
        std::vector<std::thread> ts;
        std::vector<unsigned int> tr(threads, 0);
        std::vector<unsigned long long> tc(threads, 0);
        for (int t = 0; t < threads; t++) {
            ts.emplace_back([&, t]() {
                ht->consume_seqfile<FastxReader>(parser, tr[t], tc[t]);
            });
        }
        for (auto& t : ts) t.join();
        for (int t = 0; t < threads; t++) { total_reads += tr[t]; n_consumed += tc[t]; }
    }
    *reads = total_reads;
    *kmers = n_consumed;
    return 0;
    REF_CATCH(-1)
}

int ref_consume_seqfile_banding(void* h, const char* fn, unsigned num_bands, unsigned band,
                                uint64_t* reads, uint64_t* kmers)
{
    REF_TRY
    Hashtable* ht = (Hashtable*)h;
    unsigned int total_reads = 0;
    unsigned long long n_consumed = 0;
    ht->consume_seqfile_banding<FastxReader>(std::string(fn), num_bands, band, total_reads, n_consumed);
    *reads = total_reads;
    *kmers = n_consumed;
    return 0;
    REF_CATCH(-1)
}

int ref_consume_seqfile_with_mask(void* h, const char* fn, void* mask, unsigned threshold, int consume_masked,
                                  uint64_t* reads, uint64_t* kmers)
{
    REF_TRY
    Hashtable* ht = (Hashtable*)h;
    unsigned int total_reads = 0;
    unsigned long long n_consumed = 0;
    ht->consume_seqfile_with_mask<FastxReader>(std::string(fn), (Hashtable*)mask, threshold, total_reads,
                                               n_consumed, consume_masked != 0);
    *reads = total_reads;
    *kmers = n_consumed;
    return 0;
    REF_CATCH(-1)
}

// consume_string does NOT clean (src/oxli/hashtable.cc:280-294)
int64_t ref_consume_string(void* h, const char* s)
{
    REF_TRY
    return ((Hashtable*)h)->consume_string(std::string(s));
    REF_CATCH(-1)
}

int ref_hash_dna(void* h, const char* kmer, uint64_t* out)
{
    REF_TRY
    *out = ((Hashtable*)h)->hash_dna(kmer);
    return 0;
    REF_CATCH(-1)
}

int ref_add_hash(void* h, uint64_t hash) { REF_TRY return ((Hashtable*)h)->add(hash) ? 1 : 0; REF_CATCH(-1) }
int ref_get_count_hash(void* h, uint64_t hash) { REF_TRY return ((Hashtable*)h)->get_count(hash); REF_CATCH(-1) }
int ref_get_count_kmer(void* h, const char* kmer) { REF_TRY return ((Hashtable*)h)->get_count(kmer); REF_CATCH(-1) }

int64_t ref_get_kmer_hashes(void* h, const char* s, uint64_t* out, int64_t cap)
{
    REF_TRY
    std::vector<HashIntoType> v;
    ((Hashtable*)h)->get_kmer_hashes(std::string(s), v);
    for (size_t i = 0; i < v.size() && (int64_t)i < cap; i++) out[i] = v[i];
    return (int64_t)v.size();
    REF_CATCH(-1)
}

int64_t ref_get_kmer_counts(void* h, const char* s, uint16_t* out, int64_t cap)
{
    REF_TRY
    std::vector<BoundedCounterType> v;
    ((Hashtable*)h)->get_kmer_counts(std::string(s), v);
    for (size_t i = 0; i < v.size() && (int64_t)i < cap; i++) out[i] = v[i];
    return (int64_t)v.size();
    REF_CATCH(-1)
}

int ref_get_median_count(void* h, const char* s, uint16_t* med, float* avg, float* sd)
{
    REF_TRY
    BoundedCounterType m = 0;
    float a = 0, d = 0;
    ((Hashtable*)h)->get_median_count(std::string(s), m, a, d);
    *med = m; *avg = a; *sd = d;
    return 0;
    REF_CATCH(-1)
}

int ref_median_at_least(void* h, const char* s, unsigned cutoff)
{
    REF_TRY
    return ((Hashtable*)h)->median_at_least(std::string(s), cutoff) ? 1 : 0;
    REF_CATCH(-1)
}

int ref_abundance_distribution(void* h, const char* fn, void* tracking, uint64_t* dist /*65536*/)
{
    REF_TRY
    uint64_t* d = ((Hashtable*)h)->abundance_distribution<FastxReader>(std::string(fn), (Hashtable*)tracking);
    memcpy(dist, d, sizeof(uint64_t) * (MAX_BIGCOUNT + 1));
    delete[] d;
    return 0;
    REF_CATCH(-1)
}

int ref_save(void* h, const char* fn) { REF_TRY ((Hashtable*)h)->save(fn); return 0; REF_CATCH(-1) }
int ref_load(void* h, const char* fn) { REF_TRY ((Hashtable*)h)->load(fn); return 0; REF_CATCH(-1) }

uint64_t ref_n_unique_kmers(void* h) { return ((Hashtable*)h)->n_unique_kmers(); }
uint64_t ref_n_occupied(void* h) { return ((Hashtable*)h)->n_occupied(); }
int ref_n_tables(void* h) { return (int)((Hashtable*)h)->n_tables(); }
int ref_ksize(void* h) { return (int)((Hashtable*)h)->ksize(); }
void ref_tablesizes(void* h, uint64_t* out)
{
    std::vector<uint64_t> v = ((Hashtable*)h)->get_tablesizes();
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
}
const uint8_t* ref_raw_table(void* h, int i) { return ((Hashtable*)h)->get_raw_tables()[i]; }

int ref_nodegraph_update(void* dst, void* src)
{
    REF_TRY
    ((Nodegraph*)dst)->update_from(*(Nodegraph*)src);
    return 0;
    REF_CATCH(-1)
}

// free hash functions (src/khmer/_cpy_khmer.cc:60-311 exposes these to Python)
int ref_hash_twobit(const char* kmer, int k, uint64_t* canon, uint64_t* f, uint64_t* r)
{
    REF_TRY
    HashIntoType hf = 0, hr = 0;
    *canon = _hash(kmer, (WordLength)k, hf, hr);
    *f = hf; *r = hr;
    return 0;
    REF_CATCH(-1)
}
int ref_hash_murmur(const char* kmer, int k, uint64_t* canon, uint64_t* f, uint64_t* r)
{
    REF_TRY
    HashIntoType hf = 0, hr = 0;
    *canon = _hash_murmur(std::string(kmer, k), (WordLength)k, hf, hr);
    *f = hf; *r = hr;
    return 0;
    REF_CATCH(-1)
}
int ref_revhash(uint64_t h, int k, char* out)
{
    REF_TRY
    std::string s = _revhash(h, (WordLength)k);
    memcpy(out, s.c_str(), s.size() + 1);
    return 0;
    REF_CATCH(-1)
}

// parse a file the way the bulk loaders see it: cleaned sequences, concatenated.
// Returns number of reads; seq_out may be NULL to size.
int64_t ref_parse_clean(const char* fn, char* seq_out, uint64_t seq_cap, uint64_t* offsets, uint64_t off_cap,
                        uint64_t* total_bases)
{
    REF_TRY
    ReadParserPtr<FastxReader> parser = get_parser<FastxReader>(fn);
    uint64_t n = 0, pos = 0;
    Read read;
    while (!parser->is_complete()) {
        try {
            read = parser->get_next_read();
        } catch (NoMoreReadsAvailable&) {
            break;
        }
        read.set_clean_seq();
        if (offsets && n < off_cap) offsets[n] = pos;
        if (seq_out && pos + read.cleaned_seq.size() <= seq_cap)
            memcpy(seq_out + pos, read.cleaned_seq.data(), read.cleaned_seq.size());
        pos += read.cleaned_seq.size();
        n++;
    }
    if (offsets && n < off_cap) offsets[n] = pos;
    *total_bases = pos;
    return (int64_t)n;
    REF_CATCH(-1)
}

// tagging (Hashgraph): khmer/_oxli/graphs.pyx:588-600, :630-637 — only for kinds 0..2 (the *graph classes)
int ref_consume_seqfile_and_tag(void* h, const char* fn, uint64_t* reads, uint64_t* n_consumed_out)
{
    REF_TRY
    Hashgraph* hg = dynamic_cast<Hashgraph*>((Hashtable*)h);
    if (!hg) { g_err = "not a Hashgraph"; return -1; }
    unsigned int total_reads = 0;
    unsigned long long n_consumed = 0;
    hg->consume_seqfile_and_tag<FastxReader>(std::string(fn), total_reads, n_consumed);
    *reads = total_reads;
    *n_consumed_out = n_consumed;
    return 0;
    REF_CATCH(-1)
}

int64_t ref_n_tags(void* h)
{
    REF_TRY
    Hashgraph* hg = dynamic_cast<Hashgraph*>((Hashtable*)h);
    if (!hg) { g_err = "not a Hashgraph"; return -1; }
    return (int64_t)hg->n_tags();
    REF_CATCH(-1)
}

int ref_save_tagset(void* h, const char* fn)
{
    REF_TRY
    Hashgraph* hg = dynamic_cast<Hashgraph*>((Hashtable*)h);
    if (!hg) { g_err = "not a Hashgraph"; return -1; }
    hg->save_tagset(fn);
    return 0;
    REF_CATCH(-1)
}


// HLLCounter (khmer/_oxli/hllcounter.pyx; scripts/unique-kmers.py): n_counters = 2^p registers
void* ref_hll_new(int n_counters, int ksize)
{
    REF_TRY
    return new HLLCounter(n_counters, (WordLength)ksize);
    REF_CATCH(nullptr)
}

void ref_hll_free(void* c) { delete (HLLCounter*)c; }

int64_t ref_hll_consume_string(void* c, const char* s)
{
    REF_TRY
    return (int64_t)((HLLCounter*)c)->consume_string(std::string(s));
    REF_CATCH(-1)
}

int ref_hll_consume_seqfile(void* c, const char* fn, uint64_t* reads, uint64_t* n_consumed_out)
{
    REF_TRY
    unsigned int total_reads = 0;
    unsigned long long n_consumed = 0;
    ((HLLCounter*)c)->consume_seqfile<FastxReader>(std::string(fn), false, total_reads, n_consumed);
    *reads = total_reads;
    *n_consumed_out = n_consumed;
    return 0;
    REF_CATCH(-1)
}

int ref_hll_counters(void* c, uint8_t* out)
{
    REF_TRY
    std::vector<uint8_t> v = ((HLLCounter*)c)->get_counters();
    memcpy(out, v.data(), v.size());
    return (int)v.size();
    REF_CATCH(-1)
}

int ref_hll_set_counters(void* c, const uint8_t* in, int n)
{
    REF_TRY
    ((HLLCounter*)c)->set_counters(std::vector<uint8_t>(in, in + n));
    return 0;
    REF_CATCH(-1)
}

int64_t ref_hll_estimate(void* c)
{
    REF_TRY
    return (int64_t)((HLLCounter*)c)->estimate_cardinality();
    REF_CATCH(-1)
}

}  // extern "C"
