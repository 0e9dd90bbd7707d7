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
#include <fcntl.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "tmvb_common.cuh"

namespace {

// read-only mapping of a whole file (an empty file maps to an empty range)
struct Mapping {
    const char *data = nullptr;
    size_t size = 0;
    bool open(const char *path)
    {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
            ::close(fd);
            return false;
        }
        size = (size_t)st.st_size;
        if (size > 0) {
            void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) {
                ::close(fd);
                size = 0;
                return false;
            }
            madvise(m, size, MADV_WILLNEED);
            data = (const char *)m;
        }
        ::close(fd);
        return true;
    }
    ~Mapping()
    {
        if (data) munmap((void *)data, size);
    }
};

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

// parse(Int, field) for every field of the line into out[0..n); false on anything Julia would throw on (or > Int32).
// Fast path per field: a bare run of 1-10 digits followed by the delimiter (or the end of the line for the last field); anything
// else -- surrounding blanks, a sign, garbage -- takes the general scan from the start of the field.
inline bool parse_line(const Line &l, char delim, int32_t *out, int64_t n, int64_t sub, bool positive)
{
    const char *p = l.b;
    const char *const e = l.e;
    auto blank = [delim](char ch) { return (ch == ' ' || ch == '\t') && ch != delim; };
    for (int64_t k = 0; k < n; k++) {
        const bool last = (k + 1 == n);
        {
            const char *q = p;
            uint64_t v = 0;
            while (q < e && (unsigned)(*q - '0') <= 9u) v = v * 10 + (unsigned)(*q++ - '0');
            const int64_t nd = q - p;
            if (nd >= 1 && nd <= 10 && (last ? q == e : (q < e && *q == delim))) {
                if (v > 2147483647ull) return false;
                if (positive && v == 0) return false;  // check_doc: all terms / counts / readers / ratings must be positive
                out[k] = (int32_t)((int64_t)v - sub);
                p = last ? q : q + 1;
                continue;
            }
        }
        while (p < e && blank(*p)) p++;
        bool neg = false;
        if (p < e && (*p == '+' || *p == '-')) neg = (*p++ == '-');
        if (p >= e || *p < '0' || *p > '9') return false;
        int64_t v = 0;
        while (p < e && *p >= '0' && *p <= '9') {
            v = v * 10 + (*p++ - '0');
            if (v > 2147483647ll) return false;
        }
        while (p < e && blank(*p)) p++;
        if (!last) {
            if (p >= e || *p != delim) return false;
            p++;
        } else if (p != e) {
            return false;
        }
        if (neg) v = -v;
        if (positive && v <= 0) return false;
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

    // the file is mapped, not copied (79 MB at NSF: a read() into a zero-initialised buffer cost more than the parse on 8 threads)
    Mapping map;
    if (!map.open(path)) return fail(-1, "invalid argument: cannot open docfile %s", path);
    const char *base = map.data;
    const size_t got = map.size;
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());

    // readlines(): split at '\n', drop one trailing '\r'; no empty last line after a final newline.  The newline positions are
    // collected by all threads (one slice of the file each), then turned into lines in order.
    std::vector<Line> lines;
    {
        const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)nthreads, got / (1u << 20)));
        std::vector<std::vector<const char *>> nls((size_t)T);
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                const char *p = base + got * (size_t)t / (size_t)T, *end = base + got * (size_t)(t + 1) / (size_t)T;
                std::vector<const char *> &v = nls[(size_t)t];
                v.reserve((size_t)(end - p) / 128 + 16);
                while (p < end) {
                    const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
                    if (!nl) break;
                    v.push_back(nl);
                    p = nl + 1;
                }
            });
        for (auto &x : th) x.join();
        size_t total = 0;
        for (auto &v : nls) total += v.size();
        lines.reserve(total + 1);
        const char *p = base, *end = base + got;
        for (auto &v : nls)
            for (const char *nl : v) {
                lines.push_back(Line{p, (nl > p && nl[-1] == '\r') ? nl - 1 : nl});
                p = nl + 1;
            }
        if (p < end) lines.push_back(Line{p, (end[-1] == '\r') ? end - 1 : end});
    }
    const int L = 1 + (counts != 0) + (readers != 0) + (ratings != 0);
    const int64_t nl = (int64_t)lines.size(), M = (nl + L - 1) / L;
    const int li_counts = counts ? 1 : -1, li_readers = readers ? 1 + (counts != 0) : -1, li_ratings = ratings ? li_readers + 1 : -1;
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
                ok = ok && parse_line(*c, delim, out->counts + o, n, 0, true);   // exactly n fields, or it fails
            else
                std::fill(out->counts + o, out->counts + o + n, 1);
            if (const Line *r = line_of(d, li_readers)) ok = ok && parse_line(*r, delim, out->readers + ro, rn, 1, true);
            if (const Line *g = line_of(d, li_ratings))
                ok = ok && parse_line(*g, delim, out->ratings + ro, rn, 0, true);
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
