/* C restatement of frame2note -- TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * Follows MIR_ST500/utils.py:82-149 of the reference.  The only non-obvious part is
 *     max(set(pitch_counter), key=pitch_counter.count)            (utils.py:123,132,146)
 * whose tie-break is the CPython set iteration order.  That order is reproduced by
 * emulating CPython 3.12 Objects/setobject.c for small non-negative ints (hash(n) == n):
 * open addressing, LINEAR_PROBES = 9, PERTURB_SHIFT = 5, table 8 -> resize when
 * fill*5 >= mask*3 to the first power of two > used*4, re-inserting in old slot order.
 * Checked against the Python restatement and the imported reference in tests/test_decode.py.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libframe2note_oracle.so oracle/frame2note_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LINEAR_PROBES 9
#define PERTURB_SHIFT 5
#define MAXTAB 1024

typedef struct { int key[MAXTAB]; unsigned char used[MAXTAB]; int mask, fill; } pyset;

static void insert_clean(int *key, unsigned char *used, int mask, int k) {
    size_t perturb = (size_t)k, i = (size_t)k & (size_t)mask;
    for (;;) {
        size_t e = i;
        int probes = (i + LINEAR_PROBES <= (size_t)mask) ? LINEAR_PROBES : 0;
        do {
            if (!used[e]) { used[e] = 1; key[e] = k; return; }
            e++;
        } while (probes--);
        perturb >>= PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & (size_t)mask;
    }
}

static void resize(pyset *s, int minused) {
    int newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    static int okey[MAXTAB]; static unsigned char oused[MAXTAB];
    int oldn = s->mask + 1;
    memcpy(okey, s->key, sizeof(int) * oldn);
    memcpy(oused, s->used, oldn);
    memset(s->used, 0, newsize);
    s->mask = newsize - 1;
    for (int j = 0; j < oldn; j++) if (oused[j]) insert_clean(s->key, s->used, s->mask, okey[j]);
}

static void set_add(pyset *s, int k) {
    size_t perturb = (size_t)k, i = (size_t)k & (size_t)s->mask;
    for (;;) {
        size_t e = i;
        int probes = (i + LINEAR_PROBES <= (size_t)s->mask) ? LINEAR_PROBES : 0;
        do {
            if (!s->used[e]) {
                s->used[e] = 1; s->key[e] = k; s->fill++;
                if ((size_t)s->fill * 5 >= (size_t)s->mask * 3) resize(s, s->fill * 4);
                return;
            }
            if (s->key[e] == k) return;
            e++;
        } while (probes--);
        perturb >>= PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & (size_t)s->mask;
    }
}

/* mode with CPython tie-break; values must be in [0, 255] */
static int py_mode(const int *c, int n) {
    static pyset s; int count[256];
    memset(s.used, 0, 8); s.mask = 7; s.fill = 0;
    memset(count, 0, sizeof(count));
    for (int i = 0; i < n; i++) { set_add(&s, c[i]); count[c[i]]++; }
    int best = -1, bestc = -1;
    for (int j = 0; j <= s.mask; j++)
        if (s.used[j] && count[s.key[j]] > bestc) { bestc = count[s.key[j]]; best = s.key[j]; }
    return best;
}

/* out: n_notes x 3 doubles [onset, offset, midi]; returns n_notes, or -1 if n==1 hits the
 * reference's np.amax-of-empty-slice ValueError (utils.py:115 with n == 1 and p_on >= thr). */
int frame2note_oracle(const float *p_on, const float *p_off, const int *oct, const int *pc, int n,
                      double on_thr, double off_thr, double frame_size, double *out, int max_notes) {
    int n_out = 0, have = 0, ncnt = 0;
    double cur = 0.0, t = 0.0;
    int *cnt = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    /* python: info[0] >= onset_thres with info[0] a 0-d fp32 tensor -> compared in fp32 */
    const float on_thr_f = (float)on_thr, off_thr_f = (float)off_thr;
    for (int i = 0; i < n; i++) {
        t = frame_size * (double)i;
        int lo = i - 3 < 0 ? 0 : i - 3;
        int hi = i + 4 > n - 1 ? n - 1 : i + 4;
        int is_on = 0;
        if (p_on[i] >= on_thr_f) {
            if (hi <= lo) { free(cnt); return -1; }
            float m = p_on[lo];
            for (int j = lo + 1; j < hi; j++) if (p_on[j] > m) m = p_on[j];
            is_on = (p_on[i] == m);
        }
        if (is_on) {
            if (have && ncnt > 0 && n_out < max_notes) {
                out[3 * n_out] = cur; out[3 * n_out + 1] = t; out[3 * n_out + 2] = py_mode(cnt, ncnt) + 36; n_out++;
            }
            cur = t; have = 1; ncnt = 0;
        } else if (p_off[i] >= off_thr_f) {
            if (have) {
                if (ncnt > 0 && n_out < max_notes) {
                    out[3 * n_out] = cur; out[3 * n_out + 1] = t; out[3 * n_out + 2] = py_mode(cnt, ncnt) + 36; n_out++;
                }
                have = 0; ncnt = 0;
            }
        }
        if (have && oct[i] != 4 && pc[i] != 12) cnt[ncnt++] = oct[i] * 12 + pc[i];
    }
    if (have && ncnt > 0 && n_out < max_notes) {
        out[3 * n_out] = cur; out[3 * n_out + 1] = t; out[3 * n_out + 2] = py_mode(cnt, ncnt) + 36; n_out++;
    }
    free(cnt);
    return n_out;
}
