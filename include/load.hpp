// load.hpp -- Matrix Market (coordinate) loader and COO -> CSR conversion, host side.
//
// Same public surface as reference/include/load.hpp: coo_t / csr_t with the same member names
// (load.hpp:131-161), LoadCoo<index_t, offset_t, value_t>(filename) (load.hpp:268-408),
// ToCsr(coo) (load.hpp:420-474), exception_t / throw_if_exception (load.hpp:116-128), the same
// messages and exits on unreadable files, bad banners, dense ("array") files, unsupported
// fields, zero-based indices and index overflow.  On every file the reference loads
// correctly, the arrays produced here are identical (tests/test_loader.py checks them against
// CSR arrays produced by the reference's own loader).
//
// What is different, by design:
//   * the file is read once into memory and tokenised by hand instead of one fscanf per
//     entry (load.hpp:323-324, :346-347) -- the loader is the wall-clock bottleneck of the
//     reference driver.  Files with one entry per line (all of them in practice) are cut at
//     line boundaries and parsed by all host threads, numbers with an exact fast path
//     (mantissa < 2^53, |decimal exponent| <= 22: one correctly rounded multiply or divide, the
//     same double strtod returns) and strtod for everything else; anything unusual -- several
//     entries on a line, a malformed line, a zero index -- sends the file through the
//     sequential tokenizer, which reports errors exactly where the reference does.  ToCsr
//     splits the rows among the threads (each scans the row indices and scatters only its own
//     rows, so the order within a row stays the input order).  SPMV_LOADER_THREADS overrides
//     the thread count;
//   * every counter that can reach nnz is 64-bit (size_t / offset_t); the reference's entry
//     loops, symmetric expansion and ToCsr cursors are index_t (load.hpp:321, :341, :364-371,
//     :448-452, :458-459, :467-471) and break at 2^31 entries;
//   * `skew-symmetric` files get their mirrored entries negated and `hermitian` (real) files
//     are mirrored like symmetric ones; the reference parses both banners (load.hpp:228-231)
//     but expands neither and silently loads half a matrix.
#pragma once

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <iostream>
#include <limits>
#include <string>
#include <thread>
#include <vector>

/****************************** Exception ***********************************/

struct exception_t : std::exception {
    std::string report;

    exception_t(std::string _message = "") { report = _message; }
    virtual const char *what() const noexcept { return report.c_str(); }
};

inline void throw_if_exception(bool is_exception, std::string message = "") {
    if (is_exception) throw exception_t(message);
}

/****************************** Containers **********************************/

template <typename index_t, typename offset_t, typename value_t>
struct coo_t {
    coo_t(index_t n_rows, index_t n_cols, offset_t nnz)
        : number_of_rows(n_rows), number_of_columns(n_cols), number_of_nonzeros(nnz),
          row_indices((size_t)nnz), column_indices((size_t)nnz), nonzero_values((size_t)nnz) {}

    index_t number_of_rows;
    index_t number_of_columns;
    offset_t number_of_nonzeros;
    std::vector<index_t> row_indices;
    std::vector<index_t> column_indices;
    std::vector<value_t> nonzero_values;
};

template <typename index_t, typename offset_t, typename value_t>
struct csr_t {
    using index_type = index_t;
    using offset_type = offset_t;
    using value_type = value_t;

    index_t number_of_rows;
    index_t number_of_columns;
    offset_t number_of_nonzeros;

    std::vector<offset_t> row_offsets;    // Ap
    std::vector<index_t> column_indices;  // Aj
    std::vector<value_t> nonzero_values;  // Ax
};

/****************************** Matrix Market header ************************/

enum matrix_market_format_t { coordinate, array };
enum matrix_market_data_t { real, complex, pattern, integer };
enum matrix_market_storage_scheme_t { general, hermitian, symmetric, skew };

struct mm_header_t {
    matrix_market_format_t format = coordinate;
    matrix_market_data_t data = real;
    matrix_market_storage_scheme_t scheme = general;
    std::size_t rows = 0, cols = 0, entries = 0;
};

// status codes of the non-exiting front end (the numbers the reference's C helpers return,
// load.hpp:60-66)
enum {
    MM_OK = 0,
    MM_COULD_NOT_READ_FILE = 11,
    MM_PREMATURE_EOF = 12,
    MM_NOT_MTX = 13,
    MM_NO_HEADER = 14,
    MM_UNSUPPORTED_TYPE = 15,
    MM_LINE_TOO_LONG = 16
};

namespace mm_detail {

inline std::string lower(std::string s) {
    for (auto &c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

// whole file -> memory
inline bool slurp(const std::string &filename, std::string &out) {
    FILE *f = std::fopen(filename.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? (size_t)sz : 0);
    size_t got = sz > 0 ? std::fread(&out[0], 1, (size_t)sz, f) : 0;
    std::fclose(f);
    out.resize(got);
    return true;
}

inline int loader_threads() {
    if (const char *e = std::getenv("SPMV_LOADER_THREADS")) {
        const int v = std::atoi(e);
        if (v >= 1) return v > 256 ? 256 : v;
    }
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 1;
    return (int)(hc > 32 ? 32 : hc);
}

template <typename F>
inline void parallel_for_threads(int n_threads, F &&fn) {  // fn(thread index)
    if (n_threads <= 1) {
        fn(0);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve((size_t)n_threads - 1);
    for (int t = 1; t < n_threads; ++t) pool.emplace_back([&fn, t] { fn(t); });
    fn(0);
    for (auto &th : pool) th.join();
}

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// Decimal number at p (no leading blanks) -> double, exact fast path only: at most 19 significant
// digits with a mantissa below 2^53 and a decimal exponent within [-22, 22], where
// double(mantissa) * or / 10^k is a single correctly rounded operation (Clinger) and therefore
// the value strtod returns.  The token must end at a blank, a newline or the buffer end.
// Returns false (p untouched) for anything else; the caller then uses strtod.
inline bool fast_double(const char *&p, const char *end, double &out) {
    static const double pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                     1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char *s = p;
    bool neg = false;
    if (s < end && (*s == '-' || *s == '+')) neg = *s++ == '-';
    std::uint64_t m = 0;
    int sig = 0, exp10 = 0;
    bool any = false;
    while (s < end && *s >= '0' && *s <= '9') {
        any = true;
        if (m != 0 || *s != '0') {
            if (++sig > 19) return false;
            m = m * 10 + (std::uint64_t)(*s - '0');
        }
        ++s;
    }
    if (s < end && *s == '.') {
        ++s;
        while (s < end && *s >= '0' && *s <= '9') {
            any = true;
            if (m != 0 || *s != '0') {
                if (++sig > 19) return false;
                m = m * 10 + (std::uint64_t)(*s - '0');
            }
            --exp10;
            ++s;
        }
    }
    if (!any) return false;
    if (s < end && (*s == 'e' || *s == 'E')) {
        const char *q = s + 1;
        bool eneg = false;
        if (q < end && (*q == '-' || *q == '+')) eneg = *q++ == '-';
        if (q >= end || *q < '0' || *q > '9') return false;
        int e = 0;
        while (q < end && *q >= '0' && *q <= '9') {
            if (e < 10000) e = e * 10 + (*q - '0');
            ++q;
        }
        exp10 += eneg ? -e : e;
        s = q;
    }
    if (s < end && !(is_blank(*s) || *s == '\n')) return false;
    double d;
    if (m == 0) {
        d = 0.0;
    } else {
        if (m > (1ull << 53) || exp10 < -22 || exp10 > 22) return false;
        d = (double)m;
        d = exp10 < 0 ? d / pow10[-exp10] : d * pow10[exp10];
    }
    out = neg ? -d : d;
    p = s;
    return true;
}

struct cursor {
    const char *p;
    const char *end;
    bool at_end() const { return p >= end; }
    void skip_ws() {
        while (p < end && std::isspace((unsigned char)*p)) ++p;
    }
    // next line (without the newline); false at end of buffer
    bool line(std::string &out) {
        if (p >= end) return false;
        const char *q = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        const char *stop = q ? q : end;
        out.assign(p, stop);
        p = q ? q + 1 : end;
        return true;
    }
    bool read_size(std::size_t &v) {
        skip_ws();
        if (p >= end || !std::isdigit((unsigned char)*p)) return false;
        std::size_t acc = 0;
        while (p < end && std::isdigit((unsigned char)*p)) acc = acc * 10 + (std::size_t)(*p++ - '0');
        v = acc;
        return true;
    }
    bool read_double(double &v) {
        skip_ws();
        if (p >= end) return false;
        if (fast_double(p, end, v)) return true;
        char *stop = nullptr;
        errno = 0;
        v = std::strtod(p, &stop);  // the buffer is NUL-terminated (std::string)
        if (stop == p) return false;
        p = stop;
        return true;
    }
};

// One entry per line, parsed by several threads.  `body` is everything after the size line.
// Returns false -- and the caller falls back to the sequential tokenizer -- if the region does
// not look like that: fewer non-blank lines than entries, a line that is not exactly
// "row col [value]", a zero index or an index beyond the header's dimensions (so that errors
// are reported by one code path only).
template <typename index_t, typename value_t>
inline bool parse_entries_by_line(const char *begin, const char *end, std::size_t entries, bool has_value,
                                  std::size_t n_rows, std::size_t n_cols,
                                  std::vector<index_t> &I, std::vector<index_t> &J, std::vector<value_t> &V,
                                  int n_threads) {
    const std::size_t bytes = (std::size_t)(end - begin);
    if (n_threads > 1 && bytes < ((std::size_t)1 << 18)) n_threads = 1;  // not worth the threads
    // chunk boundaries at line starts
    std::vector<const char *> cut((size_t)n_threads + 1);
    cut[0] = begin;
    cut[(size_t)n_threads] = end;
    for (int t = 1; t < n_threads; ++t) {
        const char *q = begin + bytes / (std::size_t)n_threads * (std::size_t)t;
        if (q < cut[(size_t)t - 1]) q = cut[(size_t)t - 1];
        const char *nl = (const char *)std::memchr(q, '\n', (size_t)(end - q));
        cut[(size_t)t] = nl ? nl + 1 : end;
    }
    // pass 1: non-blank lines per chunk
    std::vector<std::size_t> first((size_t)n_threads + 1, 0);
    parallel_for_threads(n_threads, [&](int t) {
        std::size_t lines = 0;
        const char *p = cut[(size_t)t], *e = cut[(size_t)t + 1];
        while (p < e) {
            const char *nl = (const char *)std::memchr(p, '\n', (size_t)(e - p));
            const char *stop = nl ? nl : e;
            const char *q = p;
            while (q < stop && is_blank(*q)) ++q;
            if (q < stop) ++lines;
            p = nl ? nl + 1 : e;
        }
        first[(size_t)t + 1] = lines;
    });
    for (int t = 0; t < n_threads; ++t) first[(size_t)t + 1] += first[(size_t)t];
    if (first[(size_t)n_threads] < entries) return false;
    I.resize(entries);
    J.resize(entries);
    V.resize(entries);
    // pass 2: every chunk fills its own slice; lines beyond `entries` are ignored like the
    // reference's loop ignores them
    std::vector<char> bad((size_t)n_threads, 0);
    parallel_for_threads(n_threads, [&](int t) {
        std::size_t n = first[(size_t)t];
        const char *p = cut[(size_t)t], *e = cut[(size_t)t + 1];
        while (p < e && n < entries) {
            const char *nl = (const char *)std::memchr(p, '\n', (size_t)(e - p));
            const char *stop = nl ? nl : e;
            const char *q = p;
            p = nl ? nl + 1 : e;
            while (q < stop && is_blank(*q)) ++q;
            if (q >= stop) continue;  // blank line
            std::size_t r = 0, c = 0;
            const char *d0 = q;
            while (q < stop && *q >= '0' && *q <= '9') r = r * 10 + (std::size_t)(*q++ - '0');
            if (q == d0 || q - d0 > 18 || q >= stop || !is_blank(*q)) { bad[(size_t)t] = 1; return; }
            while (q < stop && is_blank(*q)) ++q;
            d0 = q;
            while (q < stop && *q >= '0' && *q <= '9') c = c * 10 + (std::size_t)(*q++ - '0');
            if (q == d0 || q - d0 > 18 || r == 0 || c == 0) { bad[(size_t)t] = 1; return; }
            // an index beyond the header's dimensions: the sequential path reports it
            if (r > n_rows || c > n_cols) { bad[(size_t)t] = 1; return; }
            double w = 1.0;
            if (has_value) {
                if (q >= stop || !is_blank(*q)) { bad[(size_t)t] = 1; return; }
                while (q < stop && is_blank(*q)) ++q;
                if (!fast_double(q, stop, w)) {
                    // strtod needs a terminated token: copy it out (rare: > 19 digits, big exponents)
                    const char *tok = q;
                    while (q < stop && !is_blank(*q)) ++q;
                    std::string tmp(tok, q);
                    char *fin = nullptr;
                    w = std::strtod(tmp.c_str(), &fin);
                    if (fin == tmp.c_str() || *fin != 0) { bad[(size_t)t] = 1; return; }
                }
            }
            while (q < stop && is_blank(*q)) ++q;
            if (q < stop) { bad[(size_t)t] = 1; return; }  // something else on the line
            I[n] = (index_t)(r - 1);
            J[n] = (index_t)(c - 1);
            V[n] = has_value ? (value_t)w : (value_t)1.0;
            ++n;
        }
    });
    for (char b : bad)
        if (b) return false;
    return true;
}

// banner: "%%MatrixMarket matrix <format> <field> <symmetry>", case-insensitive after the tag
inline int parse_banner(const std::string &line, mm_header_t &h) {
    char banner[65], mtx[65], crd[65], data_type[65], storage[65];
    if (std::sscanf(line.c_str(), "%64s %64s %64s %64s %64s", banner, mtx, crd, data_type, storage) != 5)
        return MM_PREMATURE_EOF;
    if (std::strncmp(banner, "%%MatrixMarket", 14) != 0) return MM_NO_HEADER;
    if (lower(mtx) != "matrix") return MM_UNSUPPORTED_TYPE;
    const std::string f = lower(crd), d = lower(data_type), s = lower(storage);
    if (f == "coordinate") h.format = coordinate;
    else if (f == "array") h.format = array;
    else return MM_UNSUPPORTED_TYPE;
    if (d == "real") h.data = real;
    else if (d == "complex") h.data = complex;
    else if (d == "pattern") h.data = pattern;
    else if (d == "integer") h.data = integer;
    else return MM_UNSUPPORTED_TYPE;
    if (s == "general") h.scheme = general;
    else if (s == "symmetric") h.scheme = symmetric;
    else if (s == "hermitian") h.scheme = hermitian;
    else if (s == "skew-symmetric") h.scheme = skew;
    else return MM_UNSUPPORTED_TYPE;
    return MM_OK;
}

}  // namespace mm_detail

/****************************** Loader **************************************/

// Non-exiting front end: fills `coo` (which is resized) and returns MM_OK, or returns an MM_*
// code / throws exception_t exactly where the reference throws.
template <typename index_t, typename offset_t, typename value_t>
int TryLoadCoo(const std::string &filename, coo_t<index_t, offset_t, value_t> &coo,
               mm_header_t *header_out = nullptr) {
    std::string text;
    if (!mm_detail::slurp(filename, text)) return MM_COULD_NOT_READ_FILE;
    mm_detail::cursor cur{text.data(), text.data() + text.size()};

    std::string line;
    mm_header_t h;
    if (!cur.line(line)) return MM_PREMATURE_EOF;
    if (int rc = mm_detail::parse_banner(line, h)) return rc;
    if (h.format == array) return MM_NOT_MTX;  // "File is not a sparse matrix"

    // comments, then the size line (blank lines before it are tolerated, as the reference's
    // fscanf fallback does, load.hpp:257-263)
    for (;;) {
        const char *mark = cur.p;
        if (!cur.line(line)) return MM_PREMATURE_EOF;
        if (!line.empty() && line[0] == '%') continue;
        cur.p = mark;
        break;
    }
    if (!cur.read_size(h.rows) || !cur.read_size(h.cols) || !cur.read_size(h.entries))
        return MM_PREMATURE_EOF;
    if (header_out) *header_out = h;

    throw_if_exception(h.rows >= (std::size_t)std::numeric_limits<index_t>::max() ||
                           h.cols >= (std::size_t)std::numeric_limits<index_t>::max(),
                       "vertex_t overflow");
    throw_if_exception(h.entries >= (std::size_t)std::numeric_limits<offset_t>::max(),
                       "edge_t overflow");
    if (h.data == complex) return MM_UNSUPPORTED_TYPE;  // "Unrecognized matrix market format type"

    const bool mirror = h.scheme != general;
    const value_t mirror_sign = h.scheme == skew ? (value_t)-1 : (value_t)1;

    std::vector<index_t> I, J;
    std::vector<value_t> V;
    const int n_threads = mm_detail::loader_threads();

    // fast path: one entry per line, all threads; the sequential tokenizer below takes over for
    // anything else and is the one place errors are raised from
    std::vector<index_t> I0, J0;
    std::vector<value_t> V0;
    const char *body = cur.p;
    {   // the rest of the size line belongs to the header
        const char *nl = (const char *)std::memchr(body, '\n', (size_t)(cur.end - body));
        body = nl ? nl + 1 : cur.end;
    }
    const bool by_line = h.entries > 0 && mm_detail::parse_entries_by_line<index_t, value_t>(
                                              body, cur.end, h.entries, h.data != pattern, h.rows, h.cols, I0, J0, V0, n_threads);
    if (by_line && !mirror) {
        I.swap(I0);
        J.swap(J0);
        V.swap(V0);
    } else if (by_line) {
        // off-diagonal entries are followed by their mirror image, in the order the reference
        // emits them (load.hpp:379-387): positions from a prefix sum over equal slices
        const int T = h.entries < ((std::size_t)1 << 16) ? 1 : n_threads;
        std::vector<std::size_t> start((size_t)T + 1, 0);
        const std::size_t per = (h.entries + (std::size_t)T - 1) / (std::size_t)T;
        mm_detail::parallel_for_threads(T, [&](int t) {
            const std::size_t a = std::min(h.entries, per * (std::size_t)t), b = std::min(h.entries, a + per);
            std::size_t out = b - a;
            for (std::size_t n = a; n < b; ++n) out += I0[n] != J0[n];
            start[(size_t)t + 1] = out;
        });
        for (int t = 0; t < T; ++t) start[(size_t)t + 1] += start[(size_t)t];
        I.resize(start[(size_t)T]);
        J.resize(start[(size_t)T]);
        V.resize(start[(size_t)T]);
        mm_detail::parallel_for_threads(T, [&](int t) {
            const std::size_t a = std::min(h.entries, per * (std::size_t)t), b = std::min(h.entries, a + per);
            std::size_t o = start[(size_t)t];
            for (std::size_t n = a; n < b; ++n) {
                I[o] = I0[n];
                J[o] = J0[n];
                V[o] = V0[n];
                ++o;
                if (I0[n] != J0[n]) {
                    I[o] = J0[n];
                    J[o] = I0[n];
                    V[o] = mirror_sign * V0[n];
                    ++o;
                }
            }
        });
    } else {
    const std::size_t reserve = mirror ? 2 * h.entries : h.entries;
    I.reserve(reserve);
    J.reserve(reserve);
    V.reserve(reserve);

    for (std::size_t n = 0; n < h.entries; ++n) {
        std::size_t r = 0, c = 0;
        double w = 1.0;
        const bool ok = cur.read_size(r) && cur.read_size(c) && (h.data == pattern || cur.read_double(w));
        throw_if_exception(!ok, h.data == pattern ? "Could not read edge from market file"
                                                  : "Could not read weighted edge from market file");
        throw_if_exception(r == 0, "Market file is zero-indexed");
        throw_if_exception(c == 0, "Market file is zero-indexed");
        // The reference stores such an entry unchecked (load.hpp:330-331) and its ToCsr then writes
        // past row_offsets (load.hpp:443-445); here it is an error before any array is indexed.
        throw_if_exception(r > h.rows, "Market file row index exceeds the number of rows");
        throw_if_exception(c > h.cols, "Market file column index exceeds the number of columns");
        const index_t ri = (index_t)(r - 1), ci = (index_t)(c - 1);
        const value_t v = h.data == pattern ? (value_t)1.0 : (value_t)w;
        I.push_back(ri);
        J.push_back(ci);
        V.push_back(v);
        // off-diagonal entries of a symmetric file are followed by their mirror image, in
        // the order the reference emits them (load.hpp:379-387)
        if (mirror && ri != ci) {
            I.push_back(ci);
            J.push_back(ri);
            V.push_back(mirror_sign * v);
        }
    }
    }
    throw_if_exception(I.size() >= (std::size_t)std::numeric_limits<offset_t>::max(), "edge_t overflow");

    coo.number_of_rows = (index_t)h.rows;
    coo.number_of_columns = (index_t)h.cols;
    coo.number_of_nonzeros = (offset_t)I.size();
    coo.row_indices.swap(I);
    coo.column_indices.swap(J);
    coo.nonzero_values.swap(V);
    return MM_OK;
}

// The reference's entry point and failure behaviour (load.hpp:268-408): messages on stderr and
// exit(1) for unreadable files / bad banners / dense files / unsupported fields; exception_t
// for overflow, short files and zero-based indices.
template <typename index_t, typename offset_t, typename value_t>
coo_t<index_t, offset_t, value_t> LoadCoo(std::string filename) {
    coo_t<index_t, offset_t, value_t> coo(0, 0, 0);
    const int rc = TryLoadCoo(filename, coo);
    switch (rc) {
        case MM_OK: break;
        case MM_COULD_NOT_READ_FILE:
            std::cerr << "File could not be opened: " << filename << std::endl;
            exit(1);
        case MM_NOT_MTX:
            std::cerr << "File is not a sparse matrix" << std::endl;
            exit(1);
        case MM_UNSUPPORTED_TYPE:
            std::cerr << "Unrecognized matrix market format type" << std::endl;
            exit(1);
        case MM_PREMATURE_EOF:
        case MM_NO_HEADER:
        default:
            std::cerr << "Could not process Matrix Market banner" << std::endl;
            exit(1);
    }
    return coo;
}

/**
 * COO -> CSR by a stable counting sort on the row index: duplicates kept, input order kept
 * within a row (so columns stay unsorted if the file had them unsorted), like
 * reference/include/load.hpp:420-474 -- with 64-bit cursors.
 */
template <typename index_t, typename offset_t, typename value_t>
csr_t<index_t, offset_t, value_t> ToCsr(const coo_t<index_t, offset_t, value_t> &coo) {
    csr_t<index_t, offset_t, value_t> csr;
    csr.number_of_rows = coo.number_of_rows;
    csr.number_of_columns = coo.number_of_columns;
    csr.number_of_nonzeros = coo.number_of_nonzeros;

    const std::size_t n_rows = (std::size_t)coo.number_of_rows;
    const std::size_t nnz = (std::size_t)coo.number_of_nonzeros;
    csr.row_offsets.assign(n_rows + 1, (offset_t)0);
    csr.column_indices.resize(nnz);
    csr.nonzero_values.resize(nnz);

    // histogram of row lengths, shifted by one so the prefix sum lands in place
    std::vector<std::size_t> cursor(n_rows + 1, 0);
    const int T = nnz < ((std::size_t)1 << 17) || n_rows < 1024 ? 1 : mm_detail::loader_threads();
    if (T == 1) {
        for (std::size_t n = 0; n < nnz; ++n) ++cursor[(std::size_t)coo.row_indices[n] + 1];
        for (std::size_t r = 0; r < n_rows; ++r) cursor[r + 1] += cursor[r];
        for (std::size_t r = 0; r <= n_rows; ++r) csr.row_offsets[r] = (offset_t)cursor[r];
        // scatter in input order; cursor[r] is the next free slot of row r
        for (std::size_t n = 0; n < nnz; ++n) {
            const std::size_t dest = cursor[(std::size_t)coo.row_indices[n]]++;
            csr.column_indices[dest] = coo.column_indices[n];
            csr.nonzero_values[dest] = coo.nonzero_values[n];
        }
        return csr;
    }
    // Several threads, no atomics, same result: every thread scans all row indices but counts /
    // scatters only the rows of its own range, so the order within a row is the input order.
    const index_t *rows = coo.row_indices.data();
    {
        const std::size_t per = (n_rows + (std::size_t)T - 1) / (std::size_t)T;
        mm_detail::parallel_for_threads(T, [&](int t) {
            const std::size_t lo = std::min(n_rows, per * (std::size_t)t), hi = std::min(n_rows, lo + per);
            for (std::size_t n = 0; n < nnz; ++n) {
                const std::size_t r = (std::size_t)rows[n];
                if (r >= lo && r < hi) ++cursor[r + 1];
            }
        });
    }
    for (std::size_t r = 0; r < n_rows; ++r) cursor[r + 1] += cursor[r];
    for (std::size_t r = 0; r <= n_rows; ++r) csr.row_offsets[r] = (offset_t)cursor[r];
    // row ranges holding about nnz / T entries each
    std::vector<std::size_t> cut((size_t)T + 1, n_rows);
    cut[0] = 0;
    for (int t = 1; t < T; ++t) {
        const std::size_t want = nnz / (std::size_t)T * (std::size_t)t;
        cut[(size_t)t] = (std::size_t)(std::lower_bound(cursor.begin(), cursor.end(), want) - cursor.begin());
        if (cut[(size_t)t] > n_rows) cut[(size_t)t] = n_rows;
        if (cut[(size_t)t] < cut[(size_t)t - 1]) cut[(size_t)t] = cut[(size_t)t - 1];
    }
    mm_detail::parallel_for_threads(T, [&](int t) {
        const std::size_t lo = cut[(size_t)t], hi = cut[(size_t)t + 1];
        if (lo >= hi) return;
        for (std::size_t n = 0; n < nnz; ++n) {
            const std::size_t r = (std::size_t)rows[n];
            if (r >= lo && r < hi) {
                const std::size_t dest = cursor[r]++;
                csr.column_indices[dest] = coo.column_indices[n];
                csr.nonzero_values[dest] = coo.nonzero_values[n];
            }
        }
    });
    return csr;  // CSR representation (with possible duplicates)
}
