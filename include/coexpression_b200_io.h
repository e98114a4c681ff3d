/* coexpression_b200_io.h -- C ABI of coex_hostio.so: host-side writer of a permutation's two large result objects
 * (SURVEY.md section 8 row A13).  Pure host code, no CUDA, no device needed.
 *
 * Replaces, text for text:  json.dump(individualscores, o) and json.dump(indjoined, o) of reference
 * 1-make-network.py:523-540 (Python 3 text: separators ", " and ": ", floats as float.__repr__), where
 *   individualscores[a][b] = score of the ordered hit pair (a, b), a != b        1-make-network.py:316-320
 *   indjoined[g][h]        = individualscores[winner_g][winner_h], g != h        1-make-network.py:437-452
 * Reference-side binding: see INTEGRATION.md ("Result files").
 */
#ifndef COEXPRESSION_B200_IO_H
#define COEXPRESSION_B200_IO_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define COEX_IO_API __attribute__((visibility("default")))
#else
#define COEX_IO_API
#endif

/* Message of the last failed call of this thread ("" if none). */
COEX_IO_API const char *coex_io_last_error(void);

/* Writes  {key_0: {key_1: S[0][1], key_2: S[0][2], ...}, key_1: {key_0: S[1][0], ...}, ...}  to `path` (created or
 * truncated).  S: R x R doubles, row stride ld (in doubles, >= R); keys: the R JSON-encoded labels (quotes and escapes
 * included, as json.dumps(label) returns them) concatenated, key r = keys[key_off[r] .. key_off[r + 1]); skip_diag != 0
 * leaves out the entries with row == column (both reference objects do).  Numbers are written as Python 3's
 * repr(float) (NaN / Infinity / -Infinity as json.dump writes them).  R = 0 writes "{}".
 * Returns the number of bytes written, or -1 (coex_io_last_error() says why). */
COEX_IO_API long long coex_io_write_json_matrix(const char *path, const double *S, long long ld, int R, const char *keys,
                                                const long long *key_off, int skip_diag);

/* Formats n doubles as repr(float), each followed by '\n', into out (cap bytes; 34 bytes per value always suffice).
 * Returns the number of bytes written, or -1.  (The number format of the writer above, exposed for its parity test.) */
COEX_IO_API long long coex_io_format_doubles(const double *v, long long n, char *out, long long cap);

#ifdef __cplusplus
}
#endif
#endif
