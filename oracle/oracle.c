/*
 * TEST INFRASTRUCTURE — plain-C restatement of the reference's mass-explanation path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * the library built from this file (oracle/_build/liboracle.so), and only as the checker or as the
 * timed CPU baseline.  The product (spectrseqtools_b200) never links or loads it.
 *
 * Parity status: PINNED against the reference's own functions (oracle/gen_golden.py ->
 * tests/golden/, checked by tests/test_oracle.py).
 *
 * Citations are relative to /root/reference/spectrseqtools.
 *
 *   oracle_build_bit_table   mass_table.py:207-248   literal word-by-word in-place loop
 *   oracle_is_valid          mass_explanation.py:45-89
 *   oracle_explain           mass_explanation.py:92-201 (backtrack :118-188), memo keyed (mass,row)
 *   oracle_length_bound      mass_table.py:343-487
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OR_OK 0
#define OR_OUT_OF_TABLE 1
#define OR_BAD_COMPRESSION 2
#define OR_NOMEM 3

static uint64_t width_mask(int compression) {
    return compression == 32 ? ~0ULL : ((1ULL << (2 * compression)) - 1ULL);
}

static uint64_t load_cell(const void *t, int compression, int64_t idx) {
    switch (compression) {
        case 4: return ((const uint8_t *)t)[idx];
        case 8: return ((const uint16_t *)t)[idx];
        case 16: return ((const uint32_t *)t)[idx];
        default: return ((const uint64_t *)t)[idx];
    }
}

static void store_cell(void *t, int compression, int64_t idx, uint64_t v) {
    switch (compression) {
        case 4: ((uint8_t *)t)[idx] = (uint8_t)v; break;
        case 8: ((uint16_t *)t)[idx] = (uint16_t)v; break;
        case 16: ((uint32_t *)t)[idx] = (uint32_t)v; break;
        default: ((uint64_t *)t)[idx] = v; break;
    }
}

/* mass_table.py:207-248.  `out` holds R * max_col cells of 2*compression bits, zero-initialised here.
 * `last_col_mask` is the value numpy produces for `full << 2*(max_col - (max_mass+1) % max_col)`
 * (computed by the Python caller so that numpy's shift semantics are kept). */
int oracle_build_bit_table(const int64_t *weights, int R, int64_t max_mass, int compression,
                           uint64_t last_col_mask, void *out) {
    if (compression != 4 && compression != 8 && compression != 16 && compression != 32) return OR_BAD_COMPRESSION;
    const uint64_t wm = width_mask(compression);
    const uint64_t alt_first = 0xAAAAAAAAAAAAAAAAULL & wm; /* bit1 of every cell */
    const uint64_t alt_sec = 0x5555555555555555ULL & wm;   /* bit0 of every cell */
    const int64_t max_col = (max_mass + 1 + compression - 1) / compression;
    const int cell_bytes = compression / 4;
    memset(out, 0, (size_t)R * (size_t)max_col * (size_t)cell_bytes);
    store_cell(out, compression, 0, 3ULL << (2 * (compression - 1))); /* :216 init = both bits of mass 0 */

    for (int i = 1; i < R; i++) {
        const int64_t base = (int64_t)i * max_col, prev = (int64_t)(i - 1) * max_col;
        for (int64_t j = 0; j < max_col; j++) { /* :221-223 */
            uint64_t v = load_cell(out, compression, prev + j);
            store_cell(out, compression, base + j, (v | (v >> 1)) & alt_sec);
        }
        const int64_t step = weights[i] / compression; /* :226-227 */
        const int shift = (int)(weights[i] % compression);
        for (int64_t j = 0; j < max_col; j++) { /* :230-243, ascending and in place */
            if (step + j < max_col) {
                uint64_t x = load_cell(out, compression, base + j);
                uint64_t lo = x >> (2 * shift);
                uint64_t add = alt_first & (((lo << 1) & wm) | lo);
                store_cell(out, compression, base + j + step, load_cell(out, compression, base + j + step) | add);
            }
            if (shift != 0 && j + step + 1 < max_col) {
                uint64_t x = load_cell(out, compression, base + j); /* re-read: step==0 may have changed it */
                uint64_t hi = (x << (2 * (compression - shift))) & wm;
                uint64_t add = alt_first & (((hi << 1) & wm) | hi);
                store_cell(out, compression, base + j + step + 1,
                           load_cell(out, compression, base + j + step + 1) | add);
            }
        }
    }
    for (int i = 0; i < R; i++) { /* :246 */
        int64_t idx = (int64_t)i * max_col + max_col - 1;
        store_cell(out, compression, idx, load_cell(out, compression, idx) & last_col_mask);
    }
    return OR_OK;
}

/* shifted cell value as in mass_explanation.py:140-145 */
static uint64_t shifted_cell(const void *t, int compression, int64_t max_col, int row, int64_t m) {
    uint64_t w = load_cell(t, compression, (int64_t)row * max_col + m / compression);
    return w >> (2 * (compression - 1 - (int)(m % compression)));
}

/* mass_explanation.py:45-89 with target/threshold already integerised by the caller.
 * returns OR_OK and *valid, or OR_OUT_OF_TABLE when the ascending scan meets an out-of-table value
 * before any hit. */
int oracle_is_valid(const void *table, int R, int64_t max_col, int compression, int64_t target, int64_t thr,
                    int *valid) {
    *valid = 0;
    for (int64_t v = target - thr; v <= target + thr; v++) {
        if (v <= 0) continue;
        if (v >= max_col * compression) return OR_OUT_OF_TABLE;
        uint64_t c = shifted_cell(table, compression, max_col, R - 1, v);
        if (c % (uint64_t)compression == 0) continue;
        if (c & 3) { *valid = 1; return OR_OK; }
    }
    return OR_OK;
}

/* ---------- explain: literal memoised recursion ---------- */

typedef struct { int32_t row; int32_t tail; } Cons; /* list = list(tail) + [row]; tail -1 = [] */
typedef struct { int32_t *items; int64_t n; } SolList; /* handles into the cons arena; -1 = empty solution */

typedef struct {
    const void *table; int R; int64_t max_col; int compression;
    const int64_t *weights; const uint8_t *is_mod; const int64_t *ind; int with_memo;
    Cons *cons; int64_t n_cons, cap_cons;
    /* memo: open addressing on key = m * 256 + row */
    int64_t *keys; SolList *vals; int64_t cap_memo, n_memo;
    int err;
    int64_t nodes;
} Ctx;

static int32_t cons_new(Ctx *c, int32_t row, int32_t tail) {
    if (c->n_cons == c->cap_cons) {
        c->cap_cons = c->cap_cons ? c->cap_cons * 2 : 1 << 16;
        c->cons = (Cons *)realloc(c->cons, (size_t)c->cap_cons * sizeof(Cons));
        if (!c->cons) { c->err = OR_NOMEM; return -1; }
    }
    c->cons[c->n_cons].row = row; c->cons[c->n_cons].tail = tail;
    return (int32_t)c->n_cons++;
}

static uint64_t mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; return k; }

static void memo_grow(Ctx *c) {
    int64_t ncap = c->cap_memo ? c->cap_memo * 2 : 1 << 14;
    int64_t *nk = (int64_t *)malloc((size_t)ncap * sizeof(int64_t));
    SolList *nv = (SolList *)malloc((size_t)ncap * sizeof(SolList));
    for (int64_t i = 0; i < ncap; i++) nk[i] = -1;
    for (int64_t i = 0; i < c->cap_memo; i++) if (c->keys[i] >= 0) {
        uint64_t h = mix((uint64_t)c->keys[i]) & (uint64_t)(ncap - 1);
        while (nk[h] >= 0) h = (h + 1) & (uint64_t)(ncap - 1);
        nk[h] = c->keys[i]; nv[h] = c->vals[i];
    }
    free(c->keys); free(c->vals);
    c->keys = nk; c->vals = nv; c->cap_memo = ncap;
}

static SolList *memo_find(Ctx *c, int64_t key) {
    if (!c->cap_memo) return NULL;
    uint64_t h = mix((uint64_t)key) & (uint64_t)(c->cap_memo - 1);
    while (c->keys[h] >= 0) { if (c->keys[h] == key) return &c->vals[h]; h = (h + 1) & (uint64_t)(c->cap_memo - 1); }
    return NULL;
}

static void memo_put(Ctx *c, int64_t key, SolList v) {
    if ((c->n_memo + 1) * 2 > c->cap_memo) memo_grow(c);
    uint64_t h = mix((uint64_t)key) & (uint64_t)(c->cap_memo - 1);
    while (c->keys[h] >= 0) h = (h + 1) & (uint64_t)(c->cap_memo - 1);
    c->keys[h] = key; c->vals[h] = v; c->n_memo++;
}

static void list_push(SolList *l, int64_t *cap, int32_t h) {
    if (l->n == *cap) { *cap = *cap ? *cap * 2 : 8; l->items = (int32_t *)realloc(l->items, (size_t)*cap * sizeof(int32_t)); }
    l->items[l->n++] = h;
}

/* returns a list; `*owned` tells the caller whether it must free items (no-memo mode / base cases) */
static SolList visit(Ctx *c, int64_t m, int r, int64_t all_left, int64_t ind_left, int *owned) {
    SolList out = {NULL, 0};
    *owned = 1;
    if (c->err) return out;
    if (c->with_memo) { /* :122-123, before any other check */
        SolList *hit = memo_find(c, m * 256 + r);
        if (hit) { *owned = 0; return *hit; }
    }
    if (m < 0) return out; /* :126-127 */
    int64_t cap = 0;
    if (m == 0) { list_push(&out, &cap, -1); return out; } /* :130-131 */
    if (m >= c->max_col * c->compression) { c->err = OR_OUT_OF_TABLE; return out; } /* :134-138 */
    c->nodes++;
    uint64_t cell = shifted_cell(c->table, c->compression, c->max_col, r, m);
    if (cell % (uint64_t)c->compression == 0) return out; /* :148-149 (not memoised) */
    if (cell & 1) { /* :153-162 UP */
        int own; SolList up = visit(c, m, r - 1, all_left, c->ind[r - 1], &own);
        for (int64_t k = 0; k < up.n; k++) list_push(&out, &cap, up.items[k]);
        if (own) free(up.items);
    }
    if (cell & 2) { /* :165-182 LEFT */
        if (!c->is_mod[r] || (all_left > 0 && ind_left > 0)) {
            if (c->is_mod[r]) { all_left -= 1; ind_left -= 1; }
            int own; SolList left = visit(c, m - c->weights[r], r, all_left, ind_left, &own);
            for (int64_t k = 0; k < left.n; k++) list_push(&out, &cap, cons_new(c, r, left.items[k]));
            if (own) free(left.items);
        }
    }
    if (c->with_memo) { memo_put(c, m * 256 + r, out); *owned = 0; } /* :185-186 */
    return out;
}

typedef struct { uint8_t *rows; int64_t *off; int64_t n_sol, n_rows, nodes; } Result;

/* Runs the whole window loop (:192-201).  max_mods < 0 means "unbounded" (np.inf default).
 * On success *res_out is a heap Result to be read with oracle_result_* and freed with oracle_result_free.
 * Solutions keep the reference's order; each is its row indices in ASCENDING row order (the reference
 * appends the current weight after the deeper ones, :174-176). */
int oracle_explain(const void *table, int R, int64_t max_col, int compression, const int64_t *weights,
                   const uint8_t *is_mod, const int64_t *ind, int64_t target, int64_t thr, int64_t max_mods,
                   int with_memo, void **res_out) {
    Ctx c; memset(&c, 0, sizeof c);
    if (R > 255) return OR_BAD_COMPRESSION;
    c.table = table; c.R = R; c.max_col = max_col; c.compression = compression;
    c.weights = weights; c.is_mod = is_mod; c.ind = ind; c.with_memo = with_memo;
    const int64_t unbounded = (int64_t)1 << 60;
    SolList all = {NULL, 0}; int64_t cap = 0;
    for (int64_t v = target - thr; v <= target + thr && !c.err; v++) {
        int own; SolList s = visit(&c, v, R - 1, max_mods < 0 ? unbounded : max_mods, ind[R - 1], &own);
        for (int64_t k = 0; k < s.n; k++) list_push(&all, &cap, s.items[k]);
        if (own) free(s.items);
    }
    Result *res = NULL;
    if (!c.err) {
        res = (Result *)calloc(1, sizeof(Result));
        res->n_sol = all.n; res->nodes = c.nodes;
        res->off = (int64_t *)malloc((size_t)(all.n + 1) * sizeof(int64_t));
        int64_t total = 0;
        for (int64_t k = 0; k < all.n; k++) { res->off[k] = total; for (int32_t h = all.items[k]; h >= 0; h = c.cons[h].tail) total++; }
        res->off[all.n] = total; res->n_rows = total;
        res->rows = (uint8_t *)malloc((size_t)(total ? total : 1));
        for (int64_t k = 0; k < all.n; k++) { /* walking tails yields descending rows; store ascending */
            int64_t end = res->off[k + 1], p = end;
            for (int32_t h = all.items[k]; h >= 0; h = c.cons[h].tail) res->rows[--p] = (uint8_t)c.cons[h].row;
        }
    }
    free(all.items);
    if (c.with_memo) for (int64_t i = 0; i < c.cap_memo; i++) if (c.keys[i] >= 0) free(c.vals[i].items);
    free(c.keys); free(c.vals); free(c.cons);
    *res_out = res;
    return c.err;
}

int64_t oracle_result_count(const void *r) { return ((const Result *)r)->n_sol; }
int64_t oracle_result_rows(const void *r) { return ((const Result *)r)->n_rows; }
int64_t oracle_result_nodes(const void *r) { return ((const Result *)r)->nodes; }
void oracle_result_fetch(const void *r, uint8_t *rows, int64_t *off) {
    const Result *res = (const Result *)r;
    memcpy(rows, res->rows, (size_t)res->n_rows);
    memcpy(off, res->off, (size_t)(res->n_sol + 1) * sizeof(int64_t));
}
void oracle_result_free(void *r) { if (r) { Result *res = (Result *)r; free(res->rows); free(res->off); free(res); } }

/* ---------- batch front ends (the same per-call functions in a loop; what a full-batch parity gate calls) ---------- */

/* codes[i] = 0 not valid, 1 valid, 2 out of table (NotImplementedError of mass_explanation.py:70-74) */
int oracle_is_valid_batch(const void *table, int R, int64_t max_col, int compression, const int64_t *target,
                          const int64_t *thr, int64_t n, uint8_t *codes) {
    for (int64_t i = 0; i < n; i++) {
        int ok = 0;
        int rc = oracle_is_valid(table, R, max_col, compression, target[i], thr[i], &ok);
        codes[i] = rc == OR_OUT_OF_TABLE ? 2 : (uint8_t)ok;
    }
    return OR_OK;
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* Every call of a batch through oracle_explain.  Each composition becomes one 8-byte key: its row indices in ascending
 * order in bytes 0.., zero padded (compositions longer than 8 nucleotides: rc OR_BAD_COMPRESSION); the keys of a call
 * are sorted ascending as little-endian uint64 — the canonical sort the parity tests apply to the device records.
 * counts[i] = compositions of call i, or -1 when the call left the table.  keys holds up to cap keys; *n_keys is the
 * number needed (call again with a larger buffer when it exceeds cap). */
int oracle_explain_batch_keys(const void *table, int R, int64_t max_col, int compression, const int64_t *weights,
                              const uint8_t *is_mod, const int64_t *ind, const int64_t *target, const int64_t *thr,
                              int64_t n, int64_t max_mods, int with_memo, int64_t *counts, uint64_t *keys, int64_t cap,
                              int64_t *n_keys) {
    int64_t used = 0;
    for (int64_t i = 0; i < n; i++) {
        void *res = NULL;
        int rc = oracle_explain(table, R, max_col, compression, weights, is_mod, ind, target[i], thr[i], max_mods, with_memo, &res);
        if (rc == OR_OUT_OF_TABLE) { counts[i] = -1; continue; }
        if (rc) return rc;
        const Result *r = (const Result *)res;
        counts[i] = r->n_sol;
        int64_t first = used;
        for (int64_t k = 0; k < r->n_sol; k++) {
            int64_t len = r->off[k + 1] - r->off[k];
            if (len > 8) { oracle_result_free(res); return OR_BAD_COMPRESSION; }
            uint64_t key = 0;
            for (int64_t q = 0; q < len; q++) key |= (uint64_t)r->rows[r->off[k] + q] << (8 * q);
            if (used < cap) keys[used] = key;
            used++;
        }
        if (used <= cap) qsort(keys + first, (size_t)(used - first), sizeof(uint64_t), cmp_u64);
        oracle_result_free(res);
    }
    *n_keys = used;
    return OR_OK;
}

/* ---------- sequence length bound, mass_table.py:343-487 ---------- */

typedef struct {
    const void *table; int R; int64_t max_col; int compression;
    const int64_t *weights; const uint8_t *is_mod; const int64_t *ind;
    int64_t *keys; int64_t *vals; int64_t cap, n; int lower; int64_t dflt; int err;
} BCtx;

static void b_grow(BCtx *c) {
    int64_t ncap = c->cap ? c->cap * 2 : 1 << 14;
    int64_t *nk = (int64_t *)malloc((size_t)ncap * 8), *nv = (int64_t *)malloc((size_t)ncap * 8);
    for (int64_t i = 0; i < ncap; i++) nk[i] = -1;
    for (int64_t i = 0; i < c->cap; i++) if (c->keys[i] >= 0) {
        uint64_t h = mix((uint64_t)c->keys[i]) & (uint64_t)(ncap - 1);
        while (nk[h] >= 0) h = (h + 1) & (uint64_t)(ncap - 1);
        nk[h] = c->keys[i]; nv[h] = c->vals[i];
    }
    free(c->keys); free(c->vals); c->keys = nk; c->vals = nv; c->cap = ncap;
}

static int64_t b_visit(BCtx *c, int64_t m, int r, int64_t all_left, int64_t ind_left) {
    if (c->err) return c->dflt;
    int64_t key = m * 256 + r;
    if (c->cap) { /* :377-378 */
        uint64_t h = mix((uint64_t)key) & (uint64_t)(c->cap - 1);
        while (c->keys[h] >= 0) { if (c->keys[h] == key) return c->vals[h]; h = (h + 1) & (uint64_t)(c->cap - 1); }
    }
    if (m < 0) return c->dflt;
    if (m == 0) return 0;
    if (m >= c->max_col * c->compression) { c->err = OR_OUT_OF_TABLE; return c->dflt; }
    uint64_t cell = shifted_cell(c->table, c->compression, c->max_col, r, m);
    if (cell % (uint64_t)c->compression == 0) return c->dflt;
    int64_t best = c->dflt;
    if (cell & 1) {
        int64_t b = b_visit(c, m, r - 1, all_left, c->ind[r - 1]);
        best = c->lower ? (b < best ? b : best) : (b > best ? b : best);
    }
    if (cell & 2) {
        if (!c->is_mod[r] || (all_left > 0 && ind_left > 0)) {
            if (c->is_mod[r]) { all_left -= 1; ind_left -= 1; }
            int64_t b = b_visit(c, m - c->weights[r], r, all_left, ind_left) + 1;
            best = c->lower ? (b < best ? b : best) : (b > best ? b : best);
        }
    }
    if ((c->n + 1) * 2 > c->cap) b_grow(c);
    uint64_t h = mix((uint64_t)key) & (uint64_t)(c->cap - 1);
    while (c->keys[h] >= 0) h = (h + 1) & (uint64_t)(c->cap - 1);
    c->keys[h] = key; c->vals[h] = best; c->n++;
    return best;
}

int oracle_length_bound(const void *table, int R, int64_t max_col, int compression, const int64_t *weights,
                        const uint8_t *is_mod, const int64_t *ind, int64_t target, int64_t thr, int64_t max_mods,
                        int64_t max_len, int lower, int64_t *bound) {
    BCtx c; memset(&c, 0, sizeof c);
    c.table = table; c.R = R; c.max_col = max_col; c.compression = compression;
    c.weights = weights; c.is_mod = is_mod; c.ind = ind; c.lower = lower;
    c.dflt = lower ? max_len + 1 : -1;
    int64_t best = 0; int first = 1;
    for (int64_t v = target - thr; v <= target + thr && !c.err; v++) {
        int64_t b = b_visit(&c, v, R - 1, max_mods, ind[R - 1]);
        if (first) { best = b; first = 0; }
        else best = lower ? (b < best ? b : best) : (b > best ? b : best);
    }
    if (best == c.dflt) best = lower ? 1 : max_len;
    free(c.keys); free(c.vals);
    *bound = best;
    return c.err;
}
