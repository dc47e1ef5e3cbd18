/* CPU ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the one-pass table definitions of oracle/sia_onepass.py (which tests prove equal to the
 * line-for-line restatement of the reference's per-label loops, oracle/sia_loops.py; reference file:
 * src/vplants/tissue_analysis/spatial_image_analysis.py: nd.sum :1231, nd.find_objects :517, nd.center_of_mass :466,
 * one-sided dilations :695-716 / :947-956, 18-connected dilations :796-799 / :835-863).  It exists so that the CUDA
 * tables can be compared bit for bit at full benchmark sizes (1024^3), where the numpy version needs too much
 * memory.  Parity pinning is inherited from sia_loops.py (docstring known-answers only; otherwise unpinned).
 *
 * Memory axes: fast, mid, slow (the volume is a C array [slow][mid][fast]).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t key; uint32_t v[7]; } slot_t;   /* faces[6], wall18 */
typedef struct { slot_t* s; uint64_t cap, n; } table_t;

static uint64_t mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; return k ^ (k >> 33); }

static int table_grow(table_t* t);
static slot_t* table_get(table_t* t, uint64_t key) {
    for (;;) {
        uint64_t i = mix(key) & (t->cap - 1);
        for (;;) {
            slot_t* s = &t->s[i];
            if (s->key == key) return s;
            if (s->key == UINT64_MAX) {
                if ((t->n + 1) * 2 > t->cap) break;
                s->key = key; t->n++; return s;
            }
            i = (i + 1) & (t->cap - 1);
        }
        if (table_grow(t)) return NULL;
    }
}
static int table_init(table_t* t, uint64_t cap) {
    t->cap = cap; t->n = 0; t->s = (slot_t*)malloc(cap * sizeof(slot_t));
    if (!t->s) return -1;
    for (uint64_t i = 0; i < cap; ++i) { t->s[i].key = UINT64_MAX; memset(t->s[i].v, 0, sizeof t->s[i].v); }
    return 0;
}
static int table_grow(table_t* t) {
    table_t b;
    if (table_init(&b, t->cap * 2)) return -1;
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->s[i].key != UINT64_MAX) { slot_t* d = table_get(&b, t->s[i].key); memcpy(d->v, t->s[i].v, sizeof d->v); }
    free(t->s); *t = b; return 0;
}

static inline uint32_t vox(const void* vol, int elem, int64_t i) {
    return elem == 2 ? ((const uint16_t*)vol)[i] : ((const uint32_t*)vol)[i];
}
static inline uint64_t pkey(uint32_t a, uint32_t b) { return a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a; }
static int cmp_slot(const void* a, const void* b) {
    uint64_t x = ((const slot_t*)a)->key, y = ((const slot_t*)b)->key;
    return x < y ? -1 : x > y;
}

/* Label table: count[nrows], s1[nrows*3], s2[nrows*6] (ff fm fs mm ms ss), bbox[nrows*6] (min f,m,s, max f,m,s;
 * min > max when absent).  Pair table returned through *out (malloc'ed, sorted by key): rows of 9 uint32
 * lo, hi, faces[6], wall18.  Returns the number of pairs, or -1 on allocation failure / label >= nrows. */
int64_t ta_oracle_onepass(const void* vol, int elem, int64_t nf, int64_t nm, int64_t ns, uint32_t nrows,
                          uint64_t* count, uint64_t* s1, uint64_t* s2, int32_t* bbox, uint32_t** out) {
    table_t t;
    if (table_init(&t, 1u << 16)) return -1;
    memset(count, 0, (size_t)nrows * 8); memset(s1, 0, (size_t)nrows * 24); memset(s2, 0, (size_t)nrows * 48);
    for (uint32_t l = 0; l < nrows; ++l) for (int a = 0; a < 3; ++a) { bbox[l * 6 + a] = INT32_MAX; bbox[l * 6 + 3 + a] = -1; }
    int64_t off18[18]; int no = 0;
    for (int ds = -1; ds <= 1; ++ds) for (int dm = -1; dm <= 1; ++dm) for (int df = -1; df <= 1; ++df) {
        int l1 = abs(ds) + abs(dm) + abs(df);
        if (l1 >= 1 && l1 <= 2) off18[no++] = (ds * nm + dm) * nf + df;
    }
    for (int64_t s = 0; s < ns; ++s) for (int64_t m = 0; m < nm; ++m) {
        const int64_t row = (s * nm + m) * nf;
        for (int64_t f = 0; f < nf; ++f) {
            const uint32_t a = vox(vol, elem, row + f);
            if (a >= nrows) { free(t.s); return -1; }
            count[a]++;
            s1[a * 3] += f; s1[a * 3 + 1] += m; s1[a * 3 + 2] += s;
            uint64_t* q = &s2[(size_t)a * 6];
            q[0] += f * f; q[1] += f * m; q[2] += f * s; q[3] += m * m; q[4] += m * s; q[5] += s * s;
            int32_t* b = &bbox[(size_t)a * 6];
            if (f < b[0]) b[0] = (int32_t)f; if (m < b[1]) b[1] = (int32_t)m; if (s < b[2]) b[2] = (int32_t)s;
            if (f > b[3]) b[3] = (int32_t)f; if (m > b[4]) b[4] = (int32_t)m; if (s > b[5]) b[5] = (int32_t)s;
            /* faces: a face between p and p+e_axis goes to slot 2*axis if label(p) < label(p+e) else 2*axis+1 */
            if (f + 1 < nf) { uint32_t w = vox(vol, elem, row + f + 1); if (w != a) { slot_t* p = table_get(&t, pkey(a, w)); if (!p) return -1; p->v[a < w ? 0 : 1]++; } }
            if (m + 1 < nm) { uint32_t w = vox(vol, elem, row + nf + f); if (w != a) { slot_t* p = table_get(&t, pkey(a, w)); if (!p) return -1; p->v[a < w ? 2 : 3]++; } }
            if (s + 1 < ns) { uint32_t w = vox(vol, elem, row + nm * nf + f); if (w != a) { slot_t* p = table_get(&t, pkey(a, w)); if (!p) return -1; p->v[a < w ? 4 : 5]++; } }
            /* wall18: one count per distinct other label among the 18 neighbours */
            const int inner = f > 0 && f + 1 < nf && m > 0 && m + 1 < nm && s > 0 && s + 1 < ns;
            if (inner) {               /* fast reject: all 18 neighbours carry the same label */
                int any = 0;
                for (int k = 0; k < 18; ++k) any |= (vox(vol, elem, row + f + off18[k]) != a);
                if (!any) continue;
            }
            uint32_t seen[18]; int nseen = 0;
            for (int ds = -1; ds <= 1; ++ds) for (int dm = -1; dm <= 1; ++dm) for (int df = -1; df <= 1; ++df) {
                int l1 = abs(ds) + abs(dm) + abs(df);
                if (l1 < 1 || l1 > 2) continue;
                int64_t ff = f + df, mm = m + dm, ss = s + ds;
                if (ff < 0 || ff >= nf || mm < 0 || mm >= nm || ss < 0 || ss >= ns) continue;
                uint32_t b2 = vox(vol, elem, (ss * nm + mm) * nf + ff);
                if (b2 == a) continue;
                int dup = 0;
                for (int k = 0; k < nseen; ++k) dup |= (seen[k] == b2);
                if (dup) continue;
                seen[nseen++] = b2;
                slot_t* p = table_get(&t, pkey(a, b2)); if (!p) return -1;
                p->v[6]++;
            }
        }
    }
    slot_t* flat = (slot_t*)malloc((t.n ? t.n : 1) * sizeof(slot_t));
    uint64_t k = 0;
    for (uint64_t i = 0; i < t.cap; ++i) if (t.s[i].key != UINT64_MAX) flat[k++] = t.s[i];
    qsort(flat, k, sizeof(slot_t), cmp_slot);
    uint32_t* o = (uint32_t*)malloc((k ? k : 1) * 9 * sizeof(uint32_t));
    for (uint64_t i = 0; i < k; ++i) {
        o[i * 9] = (uint32_t)(flat[i].key >> 32); o[i * 9 + 1] = (uint32_t)flat[i].key;
        for (int j = 0; j < 7; ++j) o[i * 9 + 2 + j] = flat[i].v[j];
    }
    free(flat); free(t.s);
    *out = o;
    return (int64_t)k;
}

void ta_oracle_free(void* p) { free(p); }
