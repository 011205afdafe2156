// cnf_io.cpp -- see cnf_io.h.  Own implementation; nothing here derives from the vendored cnf_io sources.
//
// The file is mapped once, the problem line is found sequentially, and the clause body is cut at line starts into
// one piece per worker thread.  Every worker tokenises its piece into signed literals and clause widths; because a
// clause may continue over a line break, a piece also reports how many literals precede its first terminator
// ("head": they close the clause left open by the pieces before it) and how many follow its last one ("tail").
// A short sequential pass stitches the widths, then the workers copy (and, for cnf_read_csr, re-encode) their
// literals into the final arrays.
#include "cnf_io.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Mapped {                         // read-only view of the whole file
    const char *data = nullptr;
    size_t size = 0;
    bool ok = false;
    explicit Mapped(const string &path)
    {
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) return;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); return; }
        size = (size_t)st.st_size;
        if (size == 0) { ok = true; close(fd); return; }
        void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        close(fd);
        if (p == MAP_FAILED) return;
        madvise(p, size, MADV_SEQUENTIAL);       // (advice values are enumerators, not flags: one call each)
        madvise(p, size, MADV_WILLNEED);
        data = static_cast<const char *>(p);
        ok = true;
    }
    ~Mapped() { if (data) munmap(const_cast<char *>(data), size); }
    Mapped(const Mapped &) = delete;
    Mapped &operator=(const Mapped &) = delete;
};

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
inline bool is_digit(char c) { return (unsigned)(c - '0') < 10u; }

struct Piece {                          // what one worker found in [begin, end)
    const char *begin = nullptr, *end = nullptr;
    vector<int> vals;                   // non-zero literals, in order
    vector<uint32_t> widths;            // widths of the clauses that START and END inside this piece
    uint64_t head = 0, tail = 0;        // literals before the first terminator / after the last one
    bool has_zero = false, bad = false, trailer = false;    // trailer: a '%' line ended the data here
};

// Tokenises whole lines.  Never reads at or beyond `end` (the file may lack a final '\n').
void scan_piece(Piece &pc)
{
    const char *p = pc.begin, *const end = pc.end;
    const size_t bytes = (size_t)(end - p);
    pc.vals.reserve(bytes / 3 + 16);                        // upper bound is bytes/2; untouched pages cost nothing
    pc.widths.reserve(bytes / 16 + 16);
    uint64_t width = 0;
    while (p < end) {
        const char first = *p;
        if (first == 'c' || first == 'C') {                 // comment line
            const void *nl = memchr(p, '\n', (size_t)(end - p));
            p = nl ? static_cast<const char *>(nl) + 1 : end;
            continue;
        }
        if (first == '%') { pc.trailer = true; break; }     // SATLIB trailer: nothing after it is data
        while (p < end && *p != '\n') {
            char c = *p;
            if (is_blank(c)) { p++; continue; }
            bool neg = false;
            if (c == '-' || c == '+') {
                neg = (c == '-');
                if (++p >= end) { pc.bad = true; return; }
                c = *p;
            }
            if (!is_digit(c)) { pc.bad = true; return; }    // not an integer token
            uint64_t x = 0;
            do {
                x = x * 10 + (uint64_t)(c - '0');
                if (x > 0x7fffffffull) { pc.bad = true; return; }
                if (++p >= end) break;
                c = *p;
            } while (is_digit(c));
            if (p < end && c != '\n' && !is_blank(c)) { pc.bad = true; return; }     // e.g. "12abc"
            if (x == 0) {
                if (!pc.has_zero) { pc.head = width; pc.has_zero = true; }
                else pc.widths.push_back((uint32_t)width);
                width = 0;
            } else {
                pc.vals.push_back(neg ? -(int)x : (int)x);
                width++;
            }
        }
        if (p < end) p++;                                   // the '\n'
    }
    if (pc.has_zero) pc.tail = width; else pc.head = width;
}

struct Parsed {
    bool ok = false;
    int v_num = 0, c_num = 0;           // from the problem line
    vector<Piece> pieces;
    vector<uint64_t> piece_clause0;     // index of the first clause that ENDS in piece i
    vector<uint64_t> piece_lit0;        // index of piece i's first literal in the concatenation
    uint64_t n_clauses = 0, n_lits = 0; // complete clauses / their literals (an unterminated tail is dropped)
};

// Problem line: first line that is neither a comment nor blank.  On success *body is the first byte after it.
bool parse_header(const char *p, const char *end, int *v_num, int *c_num, const char **body)
{
    while (p < end) {
        const char *line = p;
        const void *nl = memchr(p, '\n', (size_t)(end - p));
        const char *eol = nl ? static_cast<const char *>(nl) : end;
        p = nl ? eol + 1 : end;
        if (*line == 'c' || *line == 'C') continue;
        const char *q = line;
        while (q < eol && is_blank(*q)) q++;
        if (q == eol) continue;
        if (*line != 'p' && *line != 'P') return false;
        q = line + 1;
        if (q >= eol || !is_blank(*q)) return false;
        while (q < eol && is_blank(*q)) q++;
        if (eol - q < 3 || (q[0] | 0x20) != 'c' || (q[1] | 0x20) != 'n' || (q[2] | 0x20) != 'f') return false;
        q += 3;
        if (q >= eol || !is_blank(*q)) return false;
        long long num[2];
        for (int i = 0; i < 2; i++) {
            while (q < eol && is_blank(*q)) q++;
            if (q >= eol || !is_digit(*q)) return false;
            long long x = 0;
            while (q < eol && is_digit(*q)) { x = x * 10 + (*q - '0'); if (x > 0x7fffffffLL) return false; q++; }
            num[i] = x;
        }
        *v_num = (int)num[0];
        *c_num = (int)num[1];
        *body = p;
        return true;
    }
    return false;
}

int default_threads()
{
    if (const char *e = getenv("ALLL_CNF_THREADS")) { const int t = atoi(e); if (t > 0) return t; }
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::min(64u, std::max(1u, hc));
}

template <class F>
void for_each_piece(size_t n, F f)
{
    if (n <= 1) { for (size_t i = 0; i < n; i++) f(i); return; }
    vector<std::thread> pool;
    pool.reserve(n - 1);
    for (size_t i = 1; i < n; i++) pool.emplace_back([&f, i] { f(i); });
    f(0);
    for (auto &t : pool) t.join();
}

bool parse(const Mapped &file, int n_threads, Parsed &out)
{
    out = Parsed();
    const char *p = file.data, *end = file.data + file.size;
    const char *body = nullptr;
    if (!p || !parse_header(p, end, &out.v_num, &out.c_num, &body)) return false;

    // ---- cut the body at line starts; small files are not worth a thread each
    const size_t body_bytes = (size_t)(end - body);
    size_t n = (size_t)std::max(1, n_threads);
    size_t min_piece = 1u << 20;
    if (const char *e = getenv("ALLL_CNF_PIECE_BYTES")) { const long b = atol(e); if (b > 0) min_piece = (size_t)b; }   // tests
    n = std::min(n, body_bytes / min_piece + 1);
    out.pieces.resize(n);
    const char *cut = body;
    for (size_t i = 0; i < n; i++) {
        out.pieces[i].begin = cut;
        const char *want = body + body_bytes * (i + 1) / n;
        if (i + 1 == n || want >= end) cut = end;
        else {
            if (want < cut) want = cut;
            const void *nl = memchr(want, '\n', (size_t)(end - want));
            cut = nl ? static_cast<const char *>(nl) + 1 : end;
        }
        out.pieces[i].end = cut;
    }
    for_each_piece(n, [&](size_t i) { scan_piece(out.pieces[i]); });

    // ---- stitch: clause boundaries across pieces, '%' trailer, unterminated tail
    out.piece_clause0.assign(n, 0);
    out.piece_lit0.assign(n, 0);
    uint64_t open = 0, clauses = 0, lits = 0;
    size_t used = n;
    for (size_t i = 0; i < n; i++) {
        Piece &pc = out.pieces[i];
        if (pc.bad) return false;
        out.piece_clause0[i] = clauses;
        out.piece_lit0[i] = lits;
        lits += pc.vals.size();
        if (pc.has_zero) {
            clauses += 1 + pc.widths.size();
            open = pc.tail;
        } else open += pc.head;
        if (pc.trailer) { used = i + 1; break; }
    }
    out.pieces.resize(used);
    out.n_clauses = clauses;
    out.n_lits = lits - open;           // literals after the last 0 do not form a clause
    out.ok = true;
    return true;
}

// Writes widths (as counts) or offsets for all clauses, and the literals through `put(dst_index, signed_value)`.
template <class PutWidth, class PutLit>
void emit(const Parsed &ps, PutWidth put_width, PutLit put_lit)
{
    const size_t n = ps.pieces.size();
    // width of the clause that ends with piece i's first terminator = piece i's head + everything left open before it
    vector<uint64_t> first_width(n, 0);
    uint64_t open = 0;
    for (size_t i = 0; i < n; i++) {
        const Piece &pc = ps.pieces[i];
        if (pc.has_zero) { first_width[i] = open + pc.head; open = pc.tail; }
        else open += pc.head;
    }
    for_each_piece(n, [&](size_t i) {
        const Piece &pc = ps.pieces[i];
        if (pc.has_zero) {
            uint64_t c = ps.piece_clause0[i];
            put_width(c++, first_width[i]);
            for (uint32_t w : pc.widths) put_width(c++, (uint64_t)w);
        }
        const uint64_t l0 = ps.piece_lit0[i];
        const uint64_t cnt = std::min<uint64_t>(pc.vals.size(), ps.n_lits > l0 ? ps.n_lits - l0 : 0);
        for (uint64_t j = 0; j < cnt; j++) put_lit(l0 + j, pc.vals[j]);
    });
}

// cnf_header_read parses; the cnf_data_read that follows finds the result here and does not touch the file again.
struct Cache {
    string path;
    Parsed parsed;
} g_cache;

const Parsed *get(const string &path)
{
    if (!(g_cache.parsed.ok && g_cache.path == path)) {
        Mapped file(path);
        g_cache.path = path;
        if (!file.ok || !parse(file, default_threads(), g_cache.parsed)) { g_cache.parsed = Parsed(); return nullptr; }
    }
    return &g_cache.parsed;
}

} // namespace

bool cnf_header_read(const string &cnf_file_name, int *v_num, int *c_num, int *l_num)
{
    g_cache.parsed = Parsed();                                      // always re-read on a header call
    const Parsed *p = get(cnf_file_name);
    if (!p || p->n_lits > 0x7fffffffull) return true;
    *v_num = p->v_num;
    *c_num = p->c_num;
    *l_num = (int)p->n_lits;
    return false;
}

bool cnf_data_read(const string &cnf_file_name, int v_num, int c_num, int l_num, int l_c_num[], int l_val[])
{
    const Parsed *p = get(cnf_file_name);
    if (!p) return true;
    bool error = (p->v_num != v_num) || (p->n_clauses != (uint64_t)c_num) || (p->n_lits != (uint64_t)l_num);
    const uint64_t nc = (uint64_t)std::max(c_num, 0), nl = (uint64_t)std::max(l_num, 0);
    for (uint64_t c = p->n_clauses; c < nc; c++) l_c_num[c] = 0;   // never leave caller memory uninitialised
    emit(*p,
         [&](uint64_t c, uint64_t w) { if (c < nc) l_c_num[c] = (int)w; },
         [&](uint64_t l, int x) { if (l < nl) l_val[l] = x; });
    for (uint64_t l = 0; l < std::min(nl, p->n_lits); l++) {
        const long long a = l_val[l] < 0 ? -(long long)l_val[l] : l_val[l];
        if (a > v_num) { error = true; break; }                     // variable index beyond the header's V
    }
    g_cache.parsed = Parsed();                                      // release the text-sized cache
    return error;
}

bool cnf_read_csr(const string &cnf_file_name, int *v_num, int *c_num_header, vector<uint64_t> &off, vector<uint32_t> &lit,
                  int n_threads)
{
    const auto t0 = std::chrono::steady_clock::now();
    off.clear();
    lit.clear();
    Mapped file(cnf_file_name);
    const auto t1 = std::chrono::steady_clock::now();
    Parsed ps;
    if (!file.ok || !parse(file, n_threads > 0 ? n_threads : default_threads(), ps)) return true;
    const auto t2 = std::chrono::steady_clock::now();
    *v_num = ps.v_num;
    *c_num_header = ps.c_num;
    off.assign(ps.n_clauses + 1, 0);
    lit.resize(ps.n_lits);
    const uint64_t v_max = (uint64_t)ps.v_num;
    bool error = ps.n_clauses != (uint64_t)ps.c_num;
    std::atomic<bool> out_of_range{false};
    emit(ps,
         [&](uint64_t c, uint64_t w) { off[c + 1] = w; },
         [&](uint64_t l, int x) {
             const uint64_t a = x < 0 ? (uint64_t)(-(long long)x) : (uint64_t)x;
             if (a > v_max) out_of_range.store(true, std::memory_order_relaxed);
             lit[l] = (uint32_t)(x > 0 ? 2 * a - 2 : 2 * a - 1);    // main.cpp:168 of the reference
         });
    for (uint64_t c = 0; c < ps.n_clauses; c++) off[c + 1] += off[c];
    if (getenv("ALLL_CNF_TRACE")) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "cnf_read_csr: map %.1f ms, scan %.1f ms (%zu pieces), emit %.1f ms\n", ms(t0, t1), ms(t1, t2),
                ps.pieces.size(), ms(t2, t3));
    }
    return error || out_of_range.load();
}
