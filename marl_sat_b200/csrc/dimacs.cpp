// Native DIMACS CNF reader of libmarlsat_b200.so (host only, no CUDA): the data-loader side of the hot path
// (SURVEY.md section 8f rank 3).  Semantics follow the reference parser src/utils/data_parser.py:8-42:
// lines starting with 'c' are comments, the 'p cnf <vars> <clauses>' line is the header, every other line is
// one clause whose last token (the terminating 0) is dropped.  Unless `strict`, blank lines are skipped and a
// line starting with '%' ends the file (SATLIB footer) -- the reference raises on both.
#include <stdlib.h>
#include <string.h>

#include "../../include/marl_sat_b200.h"

namespace {

struct Cursor {
    const char* p;
    const char* end;
};

// One line [b, e) without the trailing newline; returns false at end of input.
bool next_line(Cursor& c, const char*& b, const char*& e) {
    if (c.p >= c.end) return false;
    b = c.p;
    const char* nl = static_cast<const char*>(memchr(c.p, '\n', (size_t)(c.end - c.p)));
    e = nl ? nl : c.end;
    c.p = nl ? nl + 1 : c.end;
    while (b < e && (*b == ' ' || *b == '\t' || *b == '\r')) ++b;           // str.strip()
    while (e > b && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r')) --e;
    return true;
}

// Parses the integers of a clause line; returns the token count or -1 on a malformed token.
int parse_ints(const char* b, const char* e, int32_t* out, int cap) {
    int n = 0;
    while (b < e) {
        while (b < e && (*b == ' ' || *b == '\t')) ++b;
        if (b >= e) break;
        bool neg = false;
        if (*b == '-' || *b == '+') { neg = *b == '-'; ++b; }
        if (b >= e || *b < '0' || *b > '9') return -1;
        long long v = 0;
        while (b < e && *b >= '0' && *b <= '9') { v = v * 10 + (*b - '0'); if (v > 2147483647LL) return -1; ++b; }
        if (b < e && *b != ' ' && *b != '\t') return -1;
        if (out && n < cap) out[n] = (int32_t)(neg ? -v : v);
        ++n;
    }
    return n;
}

}  // namespace

extern "C" int msat_dimacs_parse(const char* text, size_t len, int32_t strict, int32_t* num_vars, int32_t* num_clauses,
                                 int32_t* clause_count, int32_t* max_width, int32_t* clauses, int32_t clause_stride) {
    if (!text || !clause_count || !max_width) return MSAT_EINVAL;
    Cursor c{text, text + len};
    const char *b, *e;
    int nv = 0, nc = 0, rows = 0, width = 0;
    int32_t tmp[64];
    while (next_line(c, b, e)) {
        if (b < e && *b == 'c') continue;
        if (b < e && *b == 'p') {
            // "p cnf <vars> <clauses>": tokens 2 and 3 (data_parser.py:33-35)
            const char* q = b;
            int tok = 0;
            while (q < e && tok < 4) {
                while (q < e && (*q == ' ' || *q == '\t')) ++q;
                const char* t0 = q;
                while (q < e && *q != ' ' && *q != '\t') ++q;
                if (tok == 2) nv = atoi(t0);
                if (tok == 3) nc = atoi(t0);
                ++tok;
            }
            if (tok < 4) return MSAT_EINVAL;
            continue;
        }
        if (!strict) {
            if (b == e) continue;
            if (*b == '%') break;
        }
        int n = parse_ints(b, e, tmp, 64);
        if (n < 0 || n > 64) return MSAT_EINVAL;          // int('%') raises in the reference as well
        const int w = n > 0 ? n - 1 : 0;                   // literals[:-1]
        if (clauses) {
            if (w > clause_stride) return MSAT_EINVAL;
            int32_t* row = clauses + (size_t)rows * clause_stride;
            for (int j = 0; j < clause_stride; ++j) row[j] = j < w ? tmp[j] : 0;
        }
        width = w > width ? w : width;
        ++rows;
    }
    if (num_vars) *num_vars = nv;
    if (num_clauses) *num_clauses = nc;
    *clause_count = rows;
    *max_width = width;
    return MSAT_OK;
}
