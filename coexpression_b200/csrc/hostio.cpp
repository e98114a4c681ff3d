// hostio.cpp -- host-side writer of the two large result objects of a permutation.
//
// The reference dumps eight objects per permutation with json.dump (1-make-network.py:523-540).  Six of them are small;
// `individualscores` (P x (P - 1) doubles keyed by region labels twice over) and `indjoined` (G x (G - 1)) are ~19 MB of
// text per permutation at P = 650, and Python needs ~0.7 s to build the dict-of-dicts and serialise it -- three orders of
// magnitude longer than the GPU needs for the permutation.  This file writes the same TEXT, byte for byte, straight from the
// score matrix: the keys arrive already JSON-encoded (json.dumps of the labels, done once per permutation in Python), the
// numbers are formatted exactly as Python 3's float.__repr__ does (shortest digits that round-trip -- std::to_chars --
// laid out by the rule of CPython's format_float_short for 'r': exponent form iff decpt <= -4 or decpt > 16, ".0" appended
// to integers, at least two exponent digits), separators ", " and ": " as json.dump's defaults.
//
// Pure host code, no CUDA: built by g++ into coex_hostio.so next to coexpression_v2.so; declared in
// include/coexpression_b200_io.h; bound by coexpression_b200/hostio.py; pinned against json.dumps / repr in
// tests/test_hostio.py.
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/coexpression_b200_io.h"

namespace {

thread_local std::string g_error;

// Python 3 repr(float) / json.dumps(float) of v into out (>= 32 bytes); returns the length.
int format_double(double v, char *out) {
    char *p = out;
    if (v != v) { std::memcpy(p, "NaN", 3); return 3; }
    if (std::signbit(v)) { *p++ = '-'; v = -v; }
    if (std::isinf(v)) { std::memcpy(p, "Infinity", 8); return (int)(p + 8 - out); }
    if (v == 0.0) { std::memcpy(p, "0.0", 3); return (int)(p + 3 - out); }
    char tmp[48];
    const auto res = std::to_chars(tmp, tmp + sizeof tmp, v, std::chars_format::scientific);   // d[.ddd]e[+-]XX
    char digits[24];
    int nd = 0;
    const char *q = tmp;
    digits[nd++] = *q++;
    if (*q == '.') { ++q; while (*q != 'e') digits[nd++] = *q++; }
    ++q;                                              // 'e'
    const bool eneg = (*q == '-');
    ++q;
    int e = 0;
    while (q < res.ptr) e = e * 10 + (*q++ - '0');
    if (eneg) e = -e;
    const int decpt = e + 1;                          // value = 0.d1d2... x 10^decpt
    if (decpt <= -4 || decpt > 16) {
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; std::memcpy(p, digits + 1, nd - 1); p += nd - 1; }
        *p++ = 'e';
        *p++ = eneg ? '-' : '+';
        const int ae = eneg ? -e : e;
        if (ae >= 100) { *p++ = (char)('0' + ae / 100); *p++ = (char)('0' + (ae / 10) % 10); *p++ = (char)('0' + ae % 10); }
        else { *p++ = (char)('0' + ae / 10); *p++ = (char)('0' + ae % 10); }
    } else if (decpt <= 0) {
        *p++ = '0'; *p++ = '.';
        for (int z = 0; z < -decpt; ++z) *p++ = '0';
        std::memcpy(p, digits, nd); p += nd;
    } else if (decpt >= nd) {
        std::memcpy(p, digits, nd); p += nd;
        for (int z = 0; z < decpt - nd; ++z) *p++ = '0';
        *p++ = '.'; *p++ = '0';
    } else {
        std::memcpy(p, digits, decpt); p += decpt;
        *p++ = '.';
        std::memcpy(p, digits + decpt, nd - decpt); p += nd - decpt;
    }
    return (int)(p - out);
}

}  // namespace

extern "C" {

const char *coex_io_last_error(void) { return g_error.c_str(); }

long long coex_io_format_doubles(const double *v, long long n, char *out, long long cap) {
    if (!v || !out || n < 0) { g_error = "coex_io_format_doubles: bad argument"; return -1; }
    long long w = 0;
    for (long long i = 0; i < n; ++i) {
        if (cap - w < 34) { g_error = "coex_io_format_doubles: output buffer too small (34 bytes per value)"; return -1; }
        w += format_double(v[i], out + w);
        out[w++] = '\n';
    }
    return w;
}

long long coex_io_write_json_matrix(const char *path, const double *S, long long ld, int R, const char *keys,
                                    const long long *key_off, int skip_diag) {
    if (!path || R < 0 || (R > 0 && (!S || !keys || !key_off || ld < R))) { g_error = "coex_io_write_json_matrix: bad argument"; return -1; }
    std::FILE *f = std::fopen(path, "wb");
    if (!f) { g_error = std::string("coex_io_write_json_matrix: cannot open ") + path + ": " + std::strerror(errno); return -1; }
    const long long keys_len = R > 0 ? key_off[R] - key_off[0] : 0;
    // one outer entry: ", " + key + ": {" + R x (", " + key + ": " + number) + "}"
    long long max_key = 0;
    for (int i = 0; i < R; ++i) max_key = std::max(max_key, key_off[i + 1] - key_off[i]);
    std::vector<char> buf((size_t)(max_key + keys_len + (long long)R * 40 + 64));   // a number is at most 26 characters
    long long total = 0;
    bool ok = std::fputc('{', f) != EOF;
    total += 1;
    for (int i = 0; ok && i < R; ++i) {
        char *p = buf.data();
        if (i) { *p++ = ','; *p++ = ' '; }
        const long long kl = key_off[i + 1] - key_off[i];
        std::memcpy(p, keys + key_off[i], (size_t)kl); p += kl;
        *p++ = ':'; *p++ = ' '; *p++ = '{';
        const double *row = S + (long long)i * ld;
        bool first = true;
        for (int j = 0; j < R; ++j) {
            if (skip_diag && j == i) continue;
            if (!first) { *p++ = ','; *p++ = ' '; }
            first = false;
            const long long jl = key_off[j + 1] - key_off[j];
            std::memcpy(p, keys + key_off[j], (size_t)jl); p += jl;
            *p++ = ':'; *p++ = ' ';
            p += format_double(row[j], p);
        }
        *p++ = '}';
        const size_t len = (size_t)(p - buf.data());
        ok = std::fwrite(buf.data(), 1, len, f) == len;
        total += (long long)len;
    }
    ok = ok && std::fputc('}', f) != EOF;
    total += 1;
    if (std::fclose(f) != 0) ok = false;
    if (!ok) { g_error = std::string("coex_io_write_json_matrix: write to ") + path + " failed: " + std::strerror(errno); return -1; }
    return total;
}

}  // extern "C"
