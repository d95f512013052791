/* oracle/sz3_oracle.c -- plain-C CPU restatement of SZ3's predict -> quantize -> encode path (and its inverse).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sz3_b200/ includes, links, loads or executes this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may (as the checker, never as the thing measured or
 * shipped).  It restates the reference algorithm function by function (paths under /root/reference/include/SZ3):
 *
 *   typed half (sz3_oracle_t.inc, instantiated for float and double): LinearQuantizer, Interpolators,
 *       InterpolationDecomposition, BlockwiseDecomposition with Lorenzo(1,2) / Regression / Composed predictors
 *   this file:  HuffmanEncoder  encoder/HuffmanEncoder.hpp:96-125 (preprocess_encode, save), :140-218 (encode),
 *                               :225-279 (decode, load), :414-470 (nodes, priority queue), :478-508 (build_code),
 *                               :516-561 (init), :563-628 (pad_tree, tree bytes)
 *               Config blob     utils/Config.hpp:312-354;  bit-packed dims utils/ByteUtil.hpp:195-238
 *               eb resolution   utils/Statistic.hpp:24-56
 *               framing         api/sz.hpp:43-82,117-157; api/impl/SZDispatcher.hpp:13-107;
 *                               compressor/SZGenericCompressor.hpp:38-84; lossless/Lossless_zstd.hpp:29-45
 *
 *               auto-tuner      api/impl/SZAlgoInterp.hpp:43-119 (trial compressions), :122-286 (decisions);
 *                               utils/Sample.hpp:9-127 (profiling_block), :202-289 (sampleBlocks)  [typed half]
 *
 *               OpenMP container api/impl/SZImplOMP.hpp:16-186 (slabs compressed one after the other; conf.openmp
 *                               = number of slabs, where the reference takes omp_get_num_threads())
 *
 * Pinning: tests/test_oracle_port.py diffs every entry point below against oracle/_ref (the reference itself compiled
 * in this container) on seeded inputs -- indices, blobs and whole streams byte for byte.  The reference's own tests
 * hold no golden vectors for this path.
 */
#include "sz3_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "zstd.h"

/* ------------------------------------------------------------------ byte helpers */
static void wr(uint8_t **c, const void *v, size_t n) {
    memcpy(*c, v, n);
    *c += n;
}
static void rd(const uint8_t **c, void *v, size_t n) {
    memcpy(v, *c, n);
    *c += n;
}
static void be32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24);
    p[1] = (uint8_t)(v >> 16);
    p[2] = (uint8_t)(v >> 8);
    p[3] = (uint8_t)v;
}
static uint32_t rbe32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

/* std::next_permutation on a short int array */
static void next_perm(int *a, int n) {
    int i = n - 2, j, t, l, r;
    while (i >= 0 && a[i] >= a[i + 1]) i--;
    if (i >= 0) {
        j = n - 1;
        while (a[j] <= a[i]) j--;
        t = a[i]; a[i] = a[j]; a[j] = t;
    }
    for (l = i + 1, r = n - 1; l < r; l++, r--) { t = a[l]; a[l] = a[r]; a[r] = t; }
}

/* ------------------------------------------------------------------ Huffman */
typedef struct hnode {
    struct hnode *left, *right;
    size_t freq;
    int c;
    unsigned char t; /* 1 = leaf */
} hnode;

typedef struct {
    int offset, state_num;
    unsigned node_count;
    hnode *pool, **qq, *root;
    size_t n_nodes;
    int qend;
    uint64_t *code;        /* MSB-aligned code word per state (codes up to 64 bits) */
    unsigned char *cout;   /* code length per state */
    unsigned n_inode;
} htree;

static hnode *h_new_leaf(htree *h, size_t freq, int c) {
    hnode *n = h->pool + h->n_nodes++;
    n->left = n->right = NULL;
    n->c = c;
    n->freq = freq;
    n->t = 1;
    return n;
}
static hnode *h_new_inner(htree *h, hnode *a, hnode *b) {
    hnode *n = h->pool + h->n_nodes++;
    n->left = a;
    n->right = b;
    n->freq = a->freq + b->freq;
    n->c = 0;
    n->t = 0;
    return n;
}
/* :440-470 -- 1-based binary heap; the comparison operators (<= when sifting up, < and > when sifting down) are what
 * fixes the tree among equal frequencies */
static void h_qinsert(htree *h, hnode *n) {
    int j, i = h->qend++;
    while ((j = (i >> 1))) {
        if (h->qq[j]->freq <= n->freq) break;
        h->qq[i] = h->qq[j];
        i = j;
    }
    h->qq[i] = n;
}
static hnode *h_qremove(htree *h) {
    int i = 1, l;
    hnode *n = h->qq[1], *p;
    if (h->qend < 2) return NULL;
    h->qend--;
    h->qq[1] = h->qq[h->qend];
    while ((l = (i << 1)) < h->qend) {
        if (l + 1 < h->qend && h->qq[l + 1]->freq < h->qq[l]->freq) l++;
        if (h->qq[i]->freq > h->qq[l]->freq) {
            p = h->qq[i];
            h->qq[i] = h->qq[l];
            h->qq[l] = p;
            i = l;
        } else {
            break;
        }
    }
    return n;
}
/* :478-508, codes longer than 64 bits are not produced by any test input (would need > 1e13 symbols) */
static int h_build_code(htree *h, const hnode *n, int len, uint64_t out) {
    if (n->t) {
        if (len > 64) return -1;
        h->code[n->c] = len ? out << (64 - len) : 0;
        h->cout[n->c] = (unsigned char)len;
        return 0;
    }
    if (h_build_code(h, n->left, len + 1, out << 1)) return -1;
    return h_build_code(h, n->right, len + 1, (out << 1) | 1);
}
static void h_free(htree *h) {
    free(h->pool);
    free(h->qq);
    free(h->code);
    free(h->cout);
    memset(h, 0, sizeof(*h));
}
/* preprocess_encode + init: :96-105, :516-561 */
static int h_init(htree *h, const int *s, size_t n) {
    size_t i;
    int mx, k;
    size_t *freq;
    unsigned distinct = 0;
    memset(h, 0, sizeof(*h));
    if (n == 0) return -1;
    mx = h->offset = s[0];
    for (i = 0; i < n; i++) {
        if (s[i] > mx) mx = s[i];
        if (s[i] < h->offset) h->offset = s[i];
    }
    h->state_num = mx - h->offset + 2;
    freq = (size_t *)calloc((size_t)h->state_num, sizeof(size_t));
    for (i = 0; i < n; i++) freq[s[i] - h->offset]++;
    h->pool = (hnode *)calloc((size_t)h->state_num * 2, sizeof(hnode));
    h->qq = (hnode **)calloc((size_t)h->state_num * 2 + 2, sizeof(hnode *));
    h->code = (uint64_t *)calloc((size_t)h->state_num, sizeof(uint64_t));
    h->cout = (unsigned char *)calloc((size_t)h->state_num, 1);
    h->qend = 1;
    for (k = 0; k < h->state_num; k++) /* leaves in ascending symbol order */
        if (freq[k]) {
            h_qinsert(h, h_new_leaf(h, freq[k], k));
            distinct++;
        }
    free(freq);
    while (h->qend > 2) {
        hnode *l = h_qremove(h), *r = h_qremove(h);
        h_qinsert(h, h_new_inner(h, l, r));
    }
    h->root = h->qq[1];
    h->node_count = distinct * 2 - 1;
    return h_build_code(h, h->root, 0, 0);
}
/* pad_tree :563-579 -- pre-order numbering */
static void h_pad(htree *h, uint32_t *L, uint32_t *R, int *C, unsigned char *t, unsigned i, const hnode *root) {
    C[i] = root->c;
    t[i] = root->t;
    if (root->left) {
        h->n_inode++;
        L[i] = h->n_inode;
        h_pad(h, L, R, C, t, h->n_inode, root->left);
    }
    if (root->right) {
        h->n_inode++;
        R[i] = h->n_inode;
        h_pad(h, L, R, C, t, h->n_inode, root->right);
    }
}
/* save :108-125 + convert_HuffTree_to_bytes_anyStates :601-628 */
static void h_save(htree *h, uint8_t **c) {
    unsigned nc = h->node_count, i;
    size_t lw = nc <= 256 ? 1 : (nc <= 65536 ? 2 : 4);
    uint32_t *L = (uint32_t *)calloc(nc, 4), *R = (uint32_t *)calloc(nc, 4);
    int *C = (int *)calloc(nc, sizeof(int));
    unsigned char *t = (unsigned char *)calloc(nc, 1);
    uint8_t *p = *c;
    h->n_inode = 0;
    h_pad(h, L, R, C, t, 0, h->root);
    memcpy(p, &h->offset, 4);
    be32(p + 4, nc);
    be32(p + 8, (uint32_t)(h->state_num / 2));
    p += 12;
    *p++ = 0; /* sysEndianType: little-endian host */
    for (i = 0; i < nc; i++) memcpy(p + i * lw, &L[i], lw);
    p += nc * lw;
    for (i = 0; i < nc; i++) memcpy(p + i * lw, &R[i], lw);
    p += nc * lw;
    memcpy(p, C, nc * sizeof(int));
    p += nc * sizeof(int);
    memcpy(p, t, nc);
    p += nc;
    *c = p;
    free(L); free(R); free(C); free(t);
}
/* encode :140-218 restated as plain MSB-first bit concatenation (the reference's byte juggling produces exactly
 * that); out: size_t outSize | bits, outSize = ceil(total bits / 8) */
static void h_encode(const htree *h, const int *s, size_t n, uint8_t **c) {
    uint8_t *p = *c + 8;
    uint64_t acc = 0;
    int nacc = 0;
    uint64_t nbytes = 0;
    size_t i;
    for (i = 0; i < n; i++) {
        int st = s[i] - h->offset;
        int len = h->cout[st];
        uint64_t code = len ? h->code[st] >> (64 - len) : 0;
        int rest = len;
        while (rest > 0) {
            int take = 64 - nacc < rest ? 64 - nacc : rest;
            uint64_t bits = take == 64 ? code : (code >> (rest - take)) & ((((uint64_t)1) << take) - 1);
            acc = take == 64 ? bits : (acc << take) | bits;
            nacc += take;
            rest -= take;
            while (nacc >= 8) {
                p[nbytes++] = (uint8_t)(acc >> (nacc - 8));
                nacc -= 8;
            }
        }
    }
    if (nacc) p[nbytes++] = (uint8_t)(acc << (8 - nacc));
    memcpy(*c, &nbytes, 8);
    *c = p + nbytes;
}
static void huff_write(const int *idx, size_t n, uint8_t **c) {
    htree h;
    if (h_init(&h, idx, n)) return;
    h_save(&h, c);
    h_encode(&h, idx, n, c);
    h_free(&h);
}
/* load :261-279 + decode :225-255: tree blob | size_t outSize | bits */
typedef struct {
    int offset;
    unsigned nc;
    uint32_t *L, *R;
    int *C;
    unsigned char *t;
} hdec;
static int hdec_load(hdec *d, const uint8_t **c) {
    const uint8_t *p = *c;
    unsigned i;
    size_t lw;
    memcpy(&d->offset, p, 4);
    d->nc = rbe32(p + 4);
    p += 13;
    if (d->nc == 0) return -1;
    lw = d->nc <= 256 ? 1 : (d->nc <= 65536 ? 2 : 4);
    d->L = (uint32_t *)calloc(d->nc, 4);
    d->R = (uint32_t *)calloc(d->nc, 4);
    d->C = (int *)calloc(d->nc, sizeof(int));
    d->t = (unsigned char *)calloc(d->nc, 1);
    for (i = 0; i < d->nc; i++) memcpy(&d->L[i], p + i * lw, lw);
    p += d->nc * lw;
    for (i = 0; i < d->nc; i++) memcpy(&d->R[i], p + i * lw, lw);
    p += d->nc * lw;
    memcpy(d->C, p, d->nc * sizeof(int));
    p += d->nc * sizeof(int);
    memcpy(d->t, p, d->nc);
    p += d->nc;
    *c = p;
    return 0;
}
static void hdec_free(hdec *d) { free(d->L); free(d->R); free(d->C); free(d->t); }
static int hdec_decode(const hdec *d, const uint8_t **c, size_t n, int *out) {
    uint64_t enc;
    const uint8_t *p;
    size_t cnt = 0;
    uint64_t b = 0;
    unsigned node = 0;
    memcpy(&enc, *c, 8);
    p = *c + 8;
    if (d->t[0]) {
        for (cnt = 0; cnt < n; cnt++) out[cnt] = d->C[0] + d->offset;
    } else {
        while (cnt < n) {
            unsigned bit;
            if (b >= enc * 8) return -1;
            bit = (p[b >> 3] >> (7 - (b & 7))) & 1u;
            b++;
            node = bit ? d->R[node] : d->L[node];
            if (node == 0 || node >= d->nc) return -1;
            if (d->t[node]) {
                out[cnt++] = d->C[node] + d->offset;
                node = 0;
            }
        }
    }
    *c = p + enc;
    return 0;
}
static int huff_read(const uint8_t **c, size_t n, int *out) {
    hdec d;
    int rc;
    if (hdec_load(&d, c)) return -1;
    rc = hdec_decode(&d, c, n, out);
    hdec_free(&d);
    return rc;
}

/* ------------------------------------------------------------------ typed halves */
static long long zwrap(const uint8_t *src, size_t len, uint8_t *dst, size_t cap);

#define T float
#define SUF f
#include "sz3_oracle_t.inc"
#undef T
#undef SUF
#define T double
#define SUF d
#include "sz3_oracle_t.inc"
#undef T
#undef SUF

/* ------------------------------------------------------------------ Config */
static uint64_t conf_num(const orc_config *c) {
    uint64_t n = 1;
    int i;
    for (i = 0; i < c->N; i++) n *= c->dims[i];
    return n;
}
/* Config(dims...) + setDims: utils/Config.hpp:146-177 and the field defaults :441-478 */
void orc_config_default(orc_config *c, int N, const uint64_t *dims) {
    int i, k = 0;
    memset(c, 0, sizeof(*c));
    for (i = 0; i < N && k < 4; i++)
        if (dims[i] > 1) c->dims[k++] = dims[i];
    if (k == 0) c->dims[k++] = 1;
    c->N = k;
    c->cmprAlgo = ORC_ALGO_INTERP_LORENZO;
    c->errorBoundMode = ORC_EB_ABS;
    c->absErrorBound = 1e-3;
    c->quantbinCnt = 65536;
    c->blockSize = k == 1 ? 128 : (k == 2 ? 16 : 6);
    c->lorenzo = 1;
    c->regression = 1;
    c->interpAlgo = 1;
    c->interpAnchorStride = -1;
    c->interpAlpha = 1.25;
    c->interpBeta = 2.0;
}
/* Config::save: utils/Config.hpp:312-354 */
size_t orc_config_save(const orc_config *c, unsigned char *out) {
    uint8_t *p = out + 1;
    uint64_t mx = 0, num = conf_num(c);
    uint8_t bw = 0, b;
    size_t bit = 0, nbytes;
    int i, j;
    int32_t i32;
    *p++ = (uint8_t)c->N;
    for (i = 0; i < c->N; i++)
        if (c->dims[i] > mx) mx = c->dims[i];
    while (mx) { mx >>= 1; bw++; }
    *p++ = bw;
    nbytes = ((size_t)bw * (size_t)c->N + 7) / 8;
    memset(p, 0, nbytes);
    for (i = 0; i < c->N; i++)
        for (j = 0; j < bw; j++, bit++)
            if ((c->dims[i] >> j) & 1) p[bit >> 3] |= (uint8_t)(1u << (bit & 7));
    p += nbytes;
    wr(&p, &num, 8);
    b = (uint8_t)c->cmprAlgo; wr(&p, &b, 1);
    b = (uint8_t)c->errorBoundMode; wr(&p, &b, 1);
    switch (c->errorBoundMode) {
        case ORC_EB_ABS: wr(&p, &c->absErrorBound, 8); break;
        case ORC_EB_REL: wr(&p, &c->relErrorBound, 8); break;
        case ORC_EB_PSNR: wr(&p, &c->psnrErrorBound, 8); break;
        case ORC_EB_L2NORM: wr(&p, &c->l2normErrorBound, 8); break;
        default: wr(&p, &c->absErrorBound, 8); wr(&p, &c->relErrorBound, 8); break;
    }
    b = (uint8_t)((c->lorenzo & 1) << 7 | (c->lorenzo2 & 1) << 6 | (c->regression & 1) << 5 | (c->regression2 & 1) << 4 |
                  ((c->openmp != 0) & 1) << 3);
    wr(&p, &b, 1);
    b = 0; wr(&p, &b, 1);                 /* dataType */
    i32 = c->quantbinCnt; wr(&p, &i32, 4);
    i32 = c->blockSize; wr(&p, &i32, 4);
    b = (uint8_t)c->N; wr(&p, &b, 1);     /* predDim */
    out[0] = (uint8_t)(p - out);
    return (size_t)(p - out);
}
static int conf_load(orc_config *c, const uint8_t *in) {
    const uint8_t *p = in, *end;
    uint8_t sz, bw, b;
    size_t bit = 0;
    int i, j;
    uint64_t num;
    int32_t i32;
    sz = *p++;
    end = in + sz;
    orc_config_default(c, 1, &(uint64_t){2});
    c->N = (int8_t)*p++;
    if (c->N < 1 || c->N > 4) return -1;
    bw = *p++;
    memset(c->dims, 0, sizeof(c->dims));
    for (i = 0; i < c->N; i++)
        for (j = 0; j < bw; j++, bit++) c->dims[i] |= (uint64_t)((p[bit >> 3] >> (bit & 7)) & 1u) << j;
    p += ((size_t)bw * (size_t)c->N + 7) / 8;
    rd(&p, &num, 8);
    rd(&p, &b, 1); c->cmprAlgo = b;
    rd(&p, &b, 1); c->errorBoundMode = b;
    switch (c->errorBoundMode) {
        case ORC_EB_ABS: rd(&p, &c->absErrorBound, 8); break;
        case ORC_EB_REL: rd(&p, &c->relErrorBound, 8); break;
        case ORC_EB_PSNR: rd(&p, &c->psnrErrorBound, 8); break;
        case ORC_EB_L2NORM: rd(&p, &c->l2normErrorBound, 8); break;
        default: rd(&p, &c->absErrorBound, 8); rd(&p, &c->relErrorBound, 8); break;
    }
    if (p < end) {
        rd(&p, &b, 1);
        c->lorenzo = (b >> 7) & 1; c->lorenzo2 = (b >> 6) & 1; c->regression = (b >> 5) & 1;
        c->regression2 = (b >> 4) & 1; c->openmp = (b >> 3) & 1;
    }
    if (p < end) rd(&p, &b, 1);
    if (p < end) { rd(&p, &i32, 4); c->quantbinCnt = i32; }
    if (p < end) { rd(&p, &i32, 4); c->blockSize = i32; }
    return 0;
}

/* calAbsErrorBound: utils/Statistic.hpp:24-56 */
double orc_abs_eb(int dtype, const orc_config *c, const void *data) {
    uint64_t n = conf_num(c);
    double range = 0;
    if (c->errorBoundMode != ORC_EB_ABS && c->errorBoundMode != ORC_EB_L2NORM)
        range = dtype == 0 ? (double)data_rangef((const float *)data, n) : (double)data_ranged((const double *)data, n);
    switch (c->errorBoundMode) {
        case ORC_EB_ABS: return c->absErrorBound;
        case ORC_EB_REL: return c->relErrorBound * range;
        case ORC_EB_PSNR: return range * pow(10, (c->psnrErrorBound + 10 * log10(1 - 2.0 / 3.0 * 0.99)) / (-20));
        case ORC_EB_L2NORM: return sqrt(3.0 / (double)n) * c->l2normErrorBound;
        case ORC_EB_ABS_AND_REL: return c->absErrorBound < c->relErrorBound * range ? c->absErrorBound : c->relErrorBound * range;
        case ORC_EB_ABS_OR_REL: return c->absErrorBound > c->relErrorBound * range ? c->absErrorBound : c->relErrorBound * range;
        default: return -1;
    }
}

/* ------------------------------------------------------------------ stage-level entry points */
static void fix_conf(orc_config *c) {
    static const int def[4] = {4096, 128, 32, 16};
    if (c->interpAnchorStride < 0) c->interpAnchorStride = def[c->N - 1];
    if (c->blockSize <= 0) c->blockSize = c->N == 1 ? 128 : (c->N == 2 ? 16 : 6);
}

long long orc_interp_decompose(int dtype, const orc_config *c0, double eb, void *data, int *quant, unsigned char *blob,
                               size_t *blob_len) {
    orc_config c = *c0;
    uint8_t *p = blob;
    long long n;
    fix_conf(&c);
    if (dtype == 0) {
        ipd_tf s;
        ipd_setupf(&s, &c, eb);
        s.quant = quant;
        ipd_runf(&s, (float *)data);
        n = (long long)s.qidx;
        if (blob) ipd_savef(&s, &p);
        qz_freef(&s.qz);
    } else {
        ipd_td s;
        ipd_setupd(&s, &c, eb);
        s.quant = quant;
        ipd_rund(&s, (double *)data);
        n = (long long)s.qidx;
        if (blob) ipd_saved(&s, &p);
        qz_freed(&s.qz);
    }
    if (blob_len) *blob_len = (size_t)(p - blob);
    return n;
}

long long orc_blockwise_decompose(int dtype, const orc_config *c0, double eb, void *data, int *quant, unsigned char *blob,
                                  size_t *blob_len) {
    orc_config c = *c0;
    uint8_t *p = blob;
    long long n;
    fix_conf(&c);
    if (c.lorenzo + c.lorenzo2 + c.regression == 0) return -1;
    if (dtype == 0) {
        bwd_tf s;
        bwd_setupf(&s, &c, eb);
        n = (long long)bwd_runf(&s, &c, (float *)data, quant, 0);
        if (blob) bwd_savef(&s, &p);
        bwd_freef(&s);
    } else {
        bwd_td s;
        bwd_setupd(&s, &c, eb);
        n = (long long)bwd_rund(&s, &c, (double *)data, quant, 0);
        if (blob) bwd_saved(&s, &p);
        bwd_freed(&s);
    }
    if (blob_len) *blob_len = (size_t)(p - blob);
    return n;
}

/* out = tree blob | size_t outSize | bits; *tree_len = bytes of the tree blob */
long long orc_huffman_encode(const int *q, size_t n, unsigned char *out, size_t *tree_len) {
    htree h;
    uint8_t *p = out;
    if (h_init(&h, q, n)) return -1;
    h_save(&h, &p);
    if (tree_len) *tree_len = (size_t)(p - out);
    h_encode(&h, q, n, &p);
    h_free(&h);
    return (long long)(p - out);
}
long long orc_huffman_decode(const unsigned char *in, size_t in_len, size_t n, int *out) {
    const uint8_t *p = in;
    (void)in_len;
    if (huff_read(&p, n, out)) return -1;
    return (long long)(p - in);
}

/* ------------------------------------------------------------------ whole streams */
#define ORC_MAGIC 0xF342F310u
#define ORC_DATAVER ((3u << 24) | (3u << 16) | (2u << 8))

size_t orc_size_bound(int dtype, const orc_config *c) {
    /* SZ_compress_size_bound: api/impl/SZImpl.hpp:34-44 (generous superset) */
    return 4096 + 2 * 256 + ZSTD_compressBound(conf_num(c) * (dtype == 0 ? 4 : 8));
}

/* Lossless_zstd::compress: size_t srcLen | frame, level 3 */
static long long zwrap(const uint8_t *src, size_t len, uint8_t *dst, size_t cap) {
    size_t r;
    uint64_t l = len;
    if (cap < 8 || cap - 8 < ZSTD_compressBound(len)) return -2; /* length_error in the reference */
    memcpy(dst, &l, 8);
    r = ZSTD_compress(dst + 8, cap - 8, src, len, 3);
    if (ZSTD_isError(r)) return -1;
    return (long long)r + 8;
}

/* SZGenericCompressor::compress (:38-63) for one decomposition */
static long long generic_compress(int dtype, orc_config *c, void *work, uint8_t *dst, size_t cap) {
    uint64_t n = conf_num(c);
    size_t esz = dtype == 0 ? 4 : 8;
    int *quant = (int *)malloc(n * sizeof(int));
    size_t bufsz = 2 * (n * esz + n * 4) + (1u << 20);
    uint8_t *buf = (uint8_t *)malloc(bufsz), *p = buf;
    size_t blen = 0;
    long long nq, r;
    if (c->cmprAlgo == ORC_ALGO_INTERP)
        nq = orc_interp_decompose(dtype, c, c->absErrorBound, work, quant, p, &blen);
    else
        nq = orc_blockwise_decompose(dtype, c, c->absErrorBound, work, quant, p, &blen);
    if (nq < 0) { free(quant); free(buf); return -1; }
    p += blen;
    {   /* encoder.save | size_t n | encode */
        htree h;
        uint64_t nn = (uint64_t)nq;
        if (h_init(&h, quant, (size_t)nq)) { free(quant); free(buf); return -1; }
        h_save(&h, &p);
        wr(&p, &nn, 8);
        h_encode(&h, quant, (size_t)nq, &p);
        h_free(&h);
    }
    r = zwrap(buf, (size_t)(p - buf), dst, cap);
    free(quant);
    free(buf);
    return r;
}

/* SZ_compress_dispatcher (api/impl/SZDispatcher.hpp:13-76): payload of one array (or one OpenMP slab) into dst;
 * *c leaves with what the reference's conf holds afterwards (resolved bound, tuned / fallback algorithm). */
static long long dispatch_compress(int dtype, orc_config *c, const void *data, uint8_t *p, size_t dcap) {
    uint64_t n = conf_num(c);
    size_t esz = dtype == 0 ? 4 : 8;
    long long payload = -1;
    fix_conf(c);
    c->absErrorBound = orc_abs_eb(dtype, c, data);
    c->errorBoundMode = ORC_EB_ABS;
    if (c->absErrorBound == 0) c->cmprAlgo = ORC_ALGO_LOSSLESS;
    if (c->cmprAlgo == ORC_ALGO_INTERP_LORENZO) { /* SZ_compress_Interp_lorenzo: tune, then one of the two below */
        if ((dtype == 0 ? tunef(c, (const float *)data) : tuned(c, (const double *)data)) != 0) return -1;
    }
    if (c->cmprAlgo == ORC_ALGO_INTERP || c->cmprAlgo == ORC_ALGO_LORENZO_REG) {
        void *work = malloc(n * esz);
        memcpy(work, data, n * esz);
        payload = generic_compress(dtype, c, work, p, dcap);
        free(work);
        if (payload == -1) return -1;
        if (payload >= 0 && (double)(n * esz) / (double)payload < 3) { /* :62-73 */
            size_t zcap = ZSTD_compressBound(n * esz) + 8;
            uint8_t *tmp = (uint8_t *)malloc(zcap);
            long long z = zwrap((const uint8_t *)data, n * esz, tmp, zcap);
            if (z >= 0 && z < payload && (size_t)z <= dcap) {
                c->cmprAlgo = ORC_ALGO_LOSSLESS;
                memcpy(p, tmp, (size_t)z);
                payload = z;
            }
            free(tmp);
        }
    } else if (c->cmprAlgo != ORC_ALGO_LOSSLESS) {
        return -1; /* NOPRED / BIOMD: not restated */
    }
    if (c->cmprAlgo == ORC_ALGO_LOSSLESS && payload < 0) {
        payload = zwrap((const uint8_t *)data, n * esz, p, dcap);
        if (payload < 0) return -1;
    } else if (payload == -2) {
        c->cmprAlgo = ORC_ALGO_LOSSLESS;
        payload = zwrap((const uint8_t *)data, n * esz, p, dcap);
        if (payload < 0) return -1;
    }
    return payload;
}

/* Config::setDims (utils/Config.hpp:154-177): unit dimensions dropped, rank-dependent defaults reset */
static void conf_set_dims(orc_config *c, int nd, const uint64_t *dims) {
    int i, k = 0;
    uint64_t d[4] = {0, 0, 0, 0};
    for (i = 0; i < nd && k < 4; i++)
        if (dims[i] > 1) d[k++] = dims[i];
    if (k == 0) d[k++] = 1;
    memcpy(c->dims, d, sizeof(d));
    c->N = k;
    c->blockSize = k == 1 ? 128 : (k == 2 ? 16 : 6);
}

/* SZ_compress_OMP (api/impl/SZImplOMP.hpp:16-117) with c->openmp slabs along the outermost dimension, executed one
 * after the other: int nThreads | Config of every slab (saved after its compression) | size_t sizes | payloads */
static long long omp_compress(int dtype, orc_config *c, const void *data, uint8_t *dst, size_t cap) {
    size_t esz = dtype == 0 ? 4 : 8;
    int nslabs = c->openmp, t;
    uint64_t row = conf_num(c) / c->dims[0];
    uint8_t **parts;
    uint64_t *sizes;
    orc_config *confs;
    uint8_t *p = dst;
    size_t need = 4;
    uint8_t blob[256];
    if ((uint64_t)nslabs > c->dims[0]) nslabs = (int)c->dims[0];
    if (c->errorBoundMode != ORC_EB_ABS) { /* per-thread min/max reduced to the global range (:57-68) */
        c->absErrorBound = orc_abs_eb(dtype, c, data);
        c->errorBoundMode = ORC_EB_ABS;
    }
    parts = (uint8_t **)calloc((size_t)nslabs, sizeof(*parts));
    sizes = (uint64_t *)calloc((size_t)nslabs, sizeof(*sizes));
    confs = (orc_config *)calloc((size_t)nslabs, sizeof(*confs));
    for (t = 0; t < nslabs; t++) {
        int lo = (int)((uint64_t)t * c->dims[0] / (uint64_t)nslabs), hi = (int)((uint64_t)(t + 1) * c->dims[0] / (uint64_t)nslabs);
        uint64_t d[4];
        size_t pcap;
        long long r;
        int i;
        for (i = 0; i < c->N; i++) d[i] = c->dims[i];
        d[0] = (uint64_t)(hi - lo);
        confs[t] = *c;
        conf_set_dims(&confs[t], c->N, d);
        pcap = ZSTD_compressBound(conf_num(&confs[t]) * esz);
        parts[t] = (uint8_t *)malloc(pcap);
        r = dispatch_compress(dtype, &confs[t], (const uint8_t *)data + (uint64_t)lo * row * esz, parts[t], pcap);
        if (r < 0) {
            for (i = 0; i <= t; i++) free(parts[i]);
            free(parts); free(sizes); free(confs);
            return -1;
        }
        sizes[t] = (uint64_t)r;
    }
    for (t = 0; t < nslabs; t++) need += orc_config_save(&confs[t], blob) + 8 + sizes[t];
    if (need <= cap) {
        int32_t ns = nslabs;
        wr(&p, &ns, 4);
        for (t = 0; t < nslabs; t++) p += orc_config_save(&confs[t], p);
        for (t = 0; t < nslabs; t++) wr(&p, &sizes[t], 8);
        for (t = 0; t < nslabs; t++) {
            memcpy(p, parts[t], sizes[t]);
            p += sizes[t];
        }
    }
    for (t = 0; t < nslabs; t++) free(parts[t]);
    free(parts); free(sizes); free(confs);
    return need <= cap ? (long long)(p - dst) : -1;
}

/* SZ_compress (api/sz.hpp:43-82): magic | data version | size_t payload size | payload | Config */
long long orc_compress(int dtype, const orc_config *c0, const void *data, char *out, size_t cap) {
    orc_config c = *c0;
    uint8_t *p = (uint8_t *)out, *size_pos;
    uint8_t blob[256];
    size_t conf_est, dcap;
    long long payload;
    uint32_t u;
    if (c.N < 1 || c.N > 4) return -1;
    if (!c.openmp) fix_conf(&c);
    if (cap < orc_size_bound(dtype, &c)) return -1;
    u = ORC_MAGIC; wr(&p, &u, 4);
    u = ORC_DATAVER; wr(&p, &u, 4);
    size_pos = p;
    p += 8;
    conf_est = orc_config_save(&c, blob);
    dcap = cap - 16 - 2 * conf_est;
    payload = c.openmp ? omp_compress(dtype, &c, data, p, dcap) : dispatch_compress(dtype, &c, data, p, dcap);
    if (payload < 0) return -1;
    {
        uint64_t ps = (uint64_t)payload;
        memcpy(size_pos, &ps, 8);
    }
    p += payload;
    p += orc_config_save(&c, p);
    return (long long)(p - (uint8_t *)out);
}

/* SZ_decompress_dispatcher (api/impl/SZDispatcher.hpp:79-107) + SZGenericCompressor::decompress (:65-84): one payload */
static int dispatch_decompress(int dtype, orc_config *cp, const uint8_t *p, uint64_t payload, void *out) {
    orc_config c = *cp;
    uint64_t raw_len, num;
    size_t esz = dtype == 0 ? 4 : 8;
    uint8_t *raw;
    const uint8_t *q;
    int rc = -1;
    if (payload < 8) return -1;
    num = conf_num(&c);
    memcpy(&raw_len, p, 8);
    if (c.cmprAlgo == ORC_ALGO_LOSSLESS) {
        size_t r = ZSTD_decompress(out, num * esz, p + 8, payload - 8);
        if (ZSTD_isError(r) || r != num * esz) return -1;
        return 0;
    }
    raw = (uint8_t *)malloc(raw_len + 16);
    {
        size_t r = ZSTD_decompress(raw, raw_len, p + 8, payload - 8);
        if (ZSTD_isError(r) || r != raw_len) { free(raw); return -1; }
    }
    q = raw;
    fix_conf(&c);
    if (c.cmprAlgo == ORC_ALGO_INTERP) {
        int *quant = (int *)malloc(num * sizeof(int));
        uint64_t nq;
        if (dtype == 0) {
            ipd_tf s;
            ipd_setupf(&s, &c, c.absErrorBound);
            if (!ipd_loadf(&s, &q)) {
                hdec d;
                if (!hdec_load(&d, &q)) {
                    rd(&q, &nq, 8);
                    if (nq == num && !hdec_decode(&d, &q, num, quant)) {
                        s.quant = quant;
                        s.decode = 1;
                        ipd_runf(&s, (float *)out);
                        rc = 0;
                    }
                    hdec_free(&d);
                }
            }
            qz_freef(&s.qz);
        } else {
            ipd_td s;
            ipd_setupd(&s, &c, c.absErrorBound);
            if (!ipd_loadd(&s, &q)) {
                hdec d;
                if (!hdec_load(&d, &q)) {
                    rd(&q, &nq, 8);
                    if (nq == num && !hdec_decode(&d, &q, num, quant)) {
                        s.quant = quant;
                        s.decode = 1;
                        ipd_rund(&s, (double *)out);
                        rc = 0;
                    }
                    hdec_free(&d);
                }
            }
            qz_freed(&s.qz);
        }
        free(quant);
    } else if (c.cmprAlgo == ORC_ALGO_LORENZO_REG) {
        int *quant = (int *)malloc(num * sizeof(int));
        uint64_t nq;
        if (dtype == 0) {
            bwd_tf s;
            bwd_setupf(&s, &c, c.absErrorBound);
            if (!bwd_loadf(&s, &q)) {
                hdec d;
                if (!hdec_load(&d, &q)) {
                    rd(&q, &nq, 8);
                    if (nq == num && !hdec_decode(&d, &q, num, quant)) {
                        bwd_runf(&s, &c, (float *)out, quant, 1);
                        rc = 0;
                    }
                    hdec_free(&d);
                }
            }
            bwd_freef(&s);
        } else {
            bwd_td s;
            bwd_setupd(&s, &c, c.absErrorBound);
            if (!bwd_loadd(&s, &q)) {
                hdec d;
                if (!hdec_load(&d, &q)) {
                    rd(&q, &nq, 8);
                    if (nq == num && !hdec_decode(&d, &q, num, quant)) {
                        bwd_rund(&s, &c, (double *)out, quant, 1);
                        rc = 0;
                    }
                    hdec_free(&d);
                }
            }
            bwd_freed(&s);
        }
        free(quant);
    }
    free(raw);
    return rc;
}

/* SZ_decompress (api/sz.hpp:117-157); the payload is one dispatcher stream or the container of SZ_decompress_OMP
 * (api/impl/SZImplOMP.hpp:120-186) */
int orc_decompress(int dtype, const char *cmp, size_t n, void *out, orc_config *conf_out) {
    const uint8_t *p = (const uint8_t *)cmp;
    uint32_t magic, ver;
    uint64_t payload;
    orc_config c;
    size_t esz = dtype == 0 ? 4 : 8;
    int rc;
    if (n < 17) return -1;
    rd(&p, &magic, 4);
    rd(&p, &ver, 4);
    rd(&p, &payload, 8);
    if (magic != ORC_MAGIC || ver != ORC_DATAVER || payload > n - 16) return -1;
    if (conf_load(&c, p + payload)) return -1;
    if (!c.openmp) {
        rc = dispatch_decompress(dtype, &c, p, payload, out);
    } else {
        const uint8_t *q = p, *end = p + payload;
        int32_t nslabs, t;
        uint64_t row = conf_num(&c) / c.dims[0], off = 0;
        orc_config *confs;
        uint64_t *sizes;
        if (payload < 4) return -1;
        rd(&q, &nslabs, 4);
        if (nslabs < 1 || (uint64_t)nslabs > c.dims[0]) return -1;
        confs = (orc_config *)calloc((size_t)nslabs, sizeof(*confs));
        sizes = (uint64_t *)calloc((size_t)nslabs, sizeof(*sizes));
        rc = 0;
        for (t = 0; t < nslabs && rc == 0; t++) {
            if (q >= end || q + q[0] > end || conf_load(&confs[t], q)) rc = -1;
            else q += q[0];
        }
        if (rc == 0 && (size_t)(end - q) < (size_t)nslabs * 8) rc = -1;
        for (t = 0; t < nslabs && rc == 0; t++) rd(&q, &sizes[t], 8);
        for (t = 0; t < nslabs && rc == 0; t++) {
            uint64_t lo = (uint64_t)t * c.dims[0] / (uint64_t)nslabs;
            if (off + sizes[t] > (uint64_t)(end - q)) { rc = -1; break; }
            rc = dispatch_decompress(dtype, &confs[t], q + off, sizes[t], (uint8_t *)out + lo * row * esz);
            off += sizes[t];
        }
        free(confs);
        free(sizes);
    }
    if (rc == 0 && conf_out) *conf_out = c;
    return rc;
}

/* the tuner inside SZ_compress_Interp_lorenzo (api/impl/SZAlgoInterp.hpp:122-286): resolves the bound like the
 * dispatcher does, then rewrites *c to the configuration the reference goes on to compress with */
int orc_tune(int dtype, orc_config *c, const void *data) {
    if (c->N < 1 || c->N > 4) return -1;
    fix_conf(c);
    c->absErrorBound = orc_abs_eb(dtype, c, data);
    c->errorBoundMode = ORC_EB_ABS;
    return dtype == 0 ? tunef(c, (const float *)data) : tuned(c, (const double *)data);
}
