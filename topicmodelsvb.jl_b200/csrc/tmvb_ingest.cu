// tmvb_ingest.cu -- host-side corpus ingest: the docfile half of readcorp (Corpus.jl:277-296) and the flattening of
// update_buffer! (modelutils.jl:371-380, 443-472) in one native, multi-threaded pass: text -> packed 0-based Int32 CSR,
// i.e. exactly what tmvb_*_set_corpus32 uploads.  SURVEY.md 8(f) row 2: once an outer iteration costs ~2 ms, parsing
// 79 MB of text with split/parse per line and vcat-splatting 128 804 vectors dominates the wall time of a training run.
//
// Format (Corpus.jl:288-295): the lines of the file are partitioned into blocks of 1 + counts + readers + ratings lines,
// one block per document -- terms, [counts], [readers], [ratings] -- each a `delim`-separated list of integers
// (parse(Int, .) tolerates surrounding blanks and a sign); a shorter last block yields a document whose missing fields take
// their defaults (counts = 1, no readers, ratings = 1), as Iterators.partition + zip do.  Every document must pass
// check_doc (Corpus.jl:41-49); the first one that does not, or does not parse, raises the reference's
//     CorpusError("document d beginning on line l failed to load.")
// No device code in this file; it lives in libtmvb.so so that the binding stays a single library.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "tmvb_common.cuh"

namespace {

struct Line {
    const char *b, *e;  // [b, e) without the line terminator
};

// number of fields of a line = delimiters + 1 (an empty line has one, empty, field: it fails to parse, as in Julia)
inline int64_t count_fields(const Line &l, char delim)
{
    int64_t n = 1;
    for (const char *p = l.b; p < l.e; p++) n += (*p == delim);
    return n;
}

// parse(Int, field) for every field of the line into out[0..n); false on anything Julia would throw on (or > Int32)
inline bool parse_line(const Line &l, char delim, int32_t *out, int64_t n, int64_t sub, bool positive)
{
    const char *p = l.b;
    auto blank = [delim](char ch) { return (ch == ' ' || ch == '\t') && ch != delim; };
    for (int64_t k = 0; k < n; k++) {
        while (p < l.e && blank(*p)) p++;
        bool neg = false;
        if (p < l.e && (*p == '+' || *p == '-')) neg = (*p++ == '-');
        if (p >= l.e || *p < '0' || *p > '9') return false;
        int64_t v = 0;
        while (p < l.e && *p >= '0' && *p <= '9') {
            v = v * 10 + (*p++ - '0');
            if (v > 2147483647ll) return false;
        }
        while (p < l.e && blank(*p)) p++;
        if (k + 1 < n) {
            if (p >= l.e || *p != delim) return false;
            p++;
        } else if (p != l.e) {
            return false;
        }
        if (neg) v = -v;
        if (positive && v <= 0) return false;  // check_doc: all terms / counts / readers / ratings must be positive
        out[k] = (int32_t)(v - sub);
    }
    return true;
}

}  // namespace

extern "C" {

int tmvb_free_csr(tmvb_csr *c)
{
    if (!c) return 0;
    free(c->N_cumsum);
    free(c->terms);
    free(c->counts);
    free(c->R_cumsum);
    free(c->readers);
    free(c->ratings);
    memset(c, 0, sizeof(*c));
    return 0;
}

int tmvb_read_docfile(const char *path, char delim, int counts, int readers, int ratings, int nthreads, tmvb_csr *out)
{
    using tmvb::fail;
    TMVB_CHECK_ARG(path != nullptr && out != nullptr, "NULL argument");
    memset(out, 0, sizeof(*out));
    if (ratings && !readers) ratings = 0;  // "ratings require readers, ratings switch set to false." (Corpus.jl:278)

    FILE *f = fopen(path, "rb");
    if (!f) return fail(-1, "invalid argument: cannot open docfile %s", path);
    fseek(f, 0, SEEK_END);
    const long fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)std::max<long>(fsz, 0) + 1);
    const size_t got = fsz > 0 ? fread(buf.data(), 1, (size_t)fsz, f) : 0;
    fclose(f);
    if ((long)got != fsz) return fail(-1, "invalid argument: short read on docfile %s", path);

    // readlines(): split at '\n', drop one trailing '\r'; no empty last line after a final newline
    std::vector<Line> lines;
    {
        const char *p = buf.data(), *end = buf.data() + got;
        while (p < end) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *e = nl ? nl : end;
            Line l{p, (e > p && e[-1] == '\r') ? e - 1 : e};
            lines.push_back(l);
            p = nl ? nl + 1 : end;
        }
    }
    const int L = 1 + (counts != 0) + (readers != 0) + (ratings != 0);
    const int64_t nl = (int64_t)lines.size(), M = (nl + L - 1) / L;
    const int li_counts = counts ? 1 : -1, li_readers = readers ? 1 + (counts != 0) : -1, li_ratings = ratings ? li_readers + 1 : -1;
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = (int)std::min<int64_t>(nthreads, std::max<int64_t>(1, M / 256));

    out->M = M;
    out->N_cumsum = (int64_t *)calloc((size_t)M + 1, 8);
    out->R_cumsum = (int64_t *)calloc((size_t)M + 1, 8);
    if (!out->N_cumsum || !out->R_cumsum) {
        tmvb_free_csr(out);
        return fail(-7, "out of host memory");
    }
    auto line_of = [&](int64_t d, int k) -> const Line * { return (k >= 0 && d * L + k < nl) ? &lines[(size_t)(d * L + k)] : nullptr; };
    auto run = [&](auto &&body) {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back([&, t] { body(M * t / nthreads, M * (t + 1) / nthreads); });
        for (auto &x : th) x.join();
    };

    // pass 1: document lengths
    run([&](int64_t d0, int64_t d1) {
        for (int64_t d = d0; d < d1; d++) {
            out->N_cumsum[d + 1] = count_fields(*line_of(d, 0), delim);
            const Line *r = line_of(d, li_readers);
            out->R_cumsum[d + 1] = r ? count_fields(*r, delim) : 0;
        }
    });
    for (int64_t d = 0; d < M; d++) {
        out->N_cumsum[d + 1] += out->N_cumsum[d];
        out->R_cumsum[d + 1] += out->R_cumsum[d];
    }
    const int64_t nnz = out->N_cumsum[M], nr = out->R_cumsum[M];
    out->nnz = nnz;
    out->nr = nr;
    out->terms = (int32_t *)malloc((size_t)std::max<int64_t>(nnz, 1) * 4);
    out->counts = (int32_t *)malloc((size_t)std::max<int64_t>(nnz, 1) * 4);
    out->readers = (int32_t *)malloc((size_t)std::max<int64_t>(nr, 1) * 4);
    out->ratings = (int32_t *)malloc((size_t)std::max<int64_t>(nr, 1) * 4);
    if (!out->terms || !out->counts || !out->readers || !out->ratings) {
        tmvb_free_csr(out);
        return fail(-7, "out of host memory");
    }

    // pass 2: parse + check_doc; the smallest failing document wins (the reference stops at the first one)
    std::atomic<int64_t> bad(M);
    std::vector<int64_t> tmax((size_t)nthreads, 0), rmax((size_t)nthreads, 0);
    std::atomic<int> tid_gen(0);
    run([&](int64_t d0, int64_t d1) {
        const int me = tid_gen++;
        int64_t tm = 0, rm = 0;
        for (int64_t d = d0; d < d1; d++) {
            const int64_t o = out->N_cumsum[d], n = out->N_cumsum[d + 1] - o, ro = out->R_cumsum[d], rn = out->R_cumsum[d + 1] - ro;
            bool ok = parse_line(*line_of(d, 0), delim, out->terms + o, n, 1, true);
            if (const Line *c = line_of(d, li_counts))
                ok = ok && count_fields(*c, delim) == n && parse_line(*c, delim, out->counts + o, n, 0, true);
            else
                std::fill(out->counts + o, out->counts + o + n, 1);
            if (const Line *r = line_of(d, li_readers)) ok = ok && parse_line(*r, delim, out->readers + ro, rn, 1, true);
            if (const Line *g = line_of(d, li_ratings))
                ok = ok && count_fields(*g, delim) == rn && parse_line(*g, delim, out->ratings + ro, rn, 0, true);
            else
                std::fill(out->ratings + ro, out->ratings + ro + rn, 1);
            if (!ok) {
                int64_t cur = bad.load();
                while (d < cur && !bad.compare_exchange_weak(cur, d)) {
                }
                continue;
            }
            for (int64_t k = 0; k < n; k++) tm = std::max<int64_t>(tm, out->terms[o + k] + 1);
            for (int64_t k = 0; k < rn; k++) rm = std::max<int64_t>(rm, out->readers[ro + k] + 1);
        }
        tmax[(size_t)me] = tm;
        rmax[(size_t)me] = rm;
    });
    if (bad.load() < M) {
        const int64_t d = bad.load();
        tmvb_free_csr(out);
        // Corpus.jl:293: "document $d beginning on line $((d - 1) * (counts + readers + ratings) + d) failed to load."
        return fail(-6, "document %lld beginning on line %lld failed to load.", (long long)(d + 1), (long long)(d * L + 1));
    }
    out->max_term = *std::max_element(tmax.begin(), tmax.end());
    out->max_reader = *std::max_element(rmax.begin(), rmax.end());
    return 0;
}

}  // extern "C"
