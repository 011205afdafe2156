// cnf_io.cpp -- see cnf_io.h.  Own implementation; nothing here derives from the vendored cnf_io sources.
#include "cnf_io.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {

struct Parsed {
    string path;
    bool ok = false;
    int v_num = 0, c_num = 0;          // from the problem line
    vector<int> l_c_num;               // per clause found
    vector<int> l_val;                 // all non-zero literals
};

Parsed g_cache;                         // last parse: header_read followed by data_read hits it

bool read_file(const string &path, vector<char> &buf)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); return false; }
    buf.resize((size_t)sz + 1);
    const size_t got = fread(buf.data(), 1, (size_t)sz, f);
    fclose(f);
    buf[got] = '\n';                    // sentinel: the last line always ends
    buf.resize(got + 1);
    return true;
}

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// Parses the whole file.  Returns false on a malformed header or token.
bool parse(const string &path, Parsed &out)
{
    out = Parsed();
    out.path = path;
    vector<char> buf;
    if (!read_file(path, buf)) return false;
    const char *p = buf.data(), *end = buf.data() + buf.size();

    // ---- problem line: first line that is neither a comment nor blank
    bool have_header = false;
    while (p < end && !have_header) {
        const char *line = p;
        while (*p != '\n') p++;
        const char *eol = p++;
        if (*line == 'c' || *line == 'C') continue;
        const char *q = line;
        while (q < eol && is_blank(*q)) q++;
        if (q == eol) continue;
        if (*line != 'p' && *line != 'P') return false;
        q = line + 1;
        if (q >= eol || !is_blank(*q)) return false;
        while (q < eol && is_blank(*q)) q++;
        if (eol - q < 3 || tolower(q[0]) != 'c' || tolower(q[1]) != 'n' || tolower(q[2]) != 'f') return false;
        q += 3;
        if (q >= eol || !is_blank(*q)) return false;
        char *after = nullptr;
        const long v = strtol(q, &after, 10);
        if (after == q) return false;
        q = after;
        const long c = strtol(q, &after, 10);
        if (after == q) return false;
        if (v < 0 || c < 0 || v > 0x7fffffffL || c > 0x7fffffffL) return false;
        out.v_num = (int)v;
        out.c_num = (int)c;
        have_header = true;
    }
    if (!have_header) return false;
    out.l_c_num.reserve((size_t)out.c_num);

    // ---- clause body
    int width = 0;
    while (p < end) {
        const char *line = p;
        if (*line == 'c' || *line == 'C') { while (*p != '\n') p++; p++; continue; }
        if (*line == '%') break;                                    // SATLIB trailer
        while (*p != '\n') {
            if (is_blank(*p)) { p++; continue; }
            bool neg = false;
            if (*p == '-' || *p == '+') { neg = (*p == '-'); p++; }
            if (!isdigit((unsigned char)*p)) return false;          // not an integer token
            long long x = 0;
            while (isdigit((unsigned char)*p)) { x = x * 10 + (*p - '0'); if (x > 0x7fffffffLL) return false; p++; }
            if (*p != '\n' && !is_blank(*p)) return false;          // e.g. "12abc"
            if (x == 0) { out.l_c_num.push_back(width); width = 0; }
            else { out.l_val.push_back((int)(neg ? -x : x)); width++; }
        }
        p++;
    }
    // literals after the last 0 do not form a clause (the reference ignores them too: no terminator, no count)
    if (width) out.l_val.resize(out.l_val.size() - (size_t)width);
    out.ok = true;
    return true;
}

const Parsed *get(const string &path)
{
    if (!(g_cache.ok && g_cache.path == path)) {
        if (!parse(path, g_cache)) { g_cache.ok = false; return nullptr; }
    }
    return &g_cache;
}

} // namespace

bool cnf_header_read(const string &cnf_file_name, int *v_num, int *c_num, int *l_num)
{
    g_cache.ok = false;                                             // always re-read on a header call
    const Parsed *p = get(cnf_file_name);
    if (!p) return true;
    *v_num = p->v_num;
    *c_num = p->c_num;
    *l_num = (int)p->l_val.size();
    return false;
}

bool cnf_data_read(const string &cnf_file_name, int v_num, int c_num, int l_num, int l_c_num[], int l_val[])
{
    const Parsed *p = get(cnf_file_name);
    if (!p) return true;
    bool error = (p->v_num != v_num) || ((int)p->l_c_num.size() != c_num) || ((int)p->l_val.size() != l_num);
    const size_t nc = p->l_c_num.size() < (size_t)c_num ? p->l_c_num.size() : (size_t)c_num;
    for (size_t c = 0; c < nc; c++) l_c_num[c] = p->l_c_num[c];
    for (size_t c = nc; c < (size_t)c_num; c++) l_c_num[c] = 0;     // never leave caller memory uninitialised
    const size_t nl = p->l_val.size() < (size_t)l_num ? p->l_val.size() : (size_t)l_num;
    for (size_t l = 0; l < nl; l++) {
        l_val[l] = p->l_val[l];
        const long long a = l_val[l] < 0 ? -(long long)l_val[l] : l_val[l];
        if (a > v_num) error = true;                                // variable index beyond the header's V
    }
    g_cache = Parsed();                                             // release the text-sized cache
    return error;
}

