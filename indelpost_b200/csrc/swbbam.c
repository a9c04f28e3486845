/* swbbam.c — libswbbam.so: native BAM / BAI / FASTA ingestion and the integer core of indelPost's dictize_read
 * (include/swbbam.h).  Written from the SAM/BAM specification (SAMv1 §4: BGZF, BAM records, §5.2: BAI
 * binning index) -- htslib / pysam are absent from the image.  Host code only (gcc, zlib); nothing here touches a GPU. */
#define _GNU_SOURCE
#include "../../include/swbbam.h"
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

static __thread char g_err[512];
static void set_err(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
const char* swb_bam_last_error(void) { return g_err; }
const char* swb_bam_version(void) { return "swbbam 0.2 (BGZF/BAM/BAI/FAI, SAMv1)"; }

static inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | p[1] << 8); }
static inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | (uint64_t)rd32(p + 4) << 32; }
static inline void wr16(uint8_t* p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; }
static inline void wr32(uint8_t* p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; p[2] = (v >> 16) & 255; p[3] = (v >> 24) & 255; }
static inline void wr64(uint8_t* p, uint64_t v) { wr32(p, (uint32_t)v); wr32(p + 4, (uint32_t)(v >> 32)); }

/* ================================================================== BGZF reader (SAMv1 §4.1) */
#define BGZF_MAX 65536
#define NCACHE 64
typedef struct {
    int64_t coff;   /* file offset of the block, -1 = empty slot */
    int32_t clen;   /* compressed size (whole block)            */
    int32_t ulen;
    uint64_t stamp;
    uint8_t* data;
} bgzf_blk;
/* a run of consecutive blocks inflated together (by several threads) for one region query */
typedef struct { int64_t coff; int32_t clen, ulen; int64_t uoff; } span_blk;
typedef struct { int n; span_blk* b; uint8_t* data; } bgzf_span;
typedef struct {
    int fd;
    int64_t fsize;
    bgzf_blk cache[NCACHE];
    uint64_t tick;
    uint8_t* cbuf;
    bgzf_span span;       /* consulted before the cache */
    bgzf_blk span_view;   /* the span block last handed out, dressed as a cache entry */
} bgzf_rd;

static int bgzf_rd_init(bgzf_rd* z, const char* path) {
    memset(z, 0, sizeof *z);
    z->fd = open(path, O_RDONLY);
    if (z->fd < 0) { set_err("cannot open %s: %s", path, strerror(errno)); return -1; }
    struct stat st;
    if (fstat(z->fd, &st) != 0) { set_err("fstat %s: %s", path, strerror(errno)); close(z->fd); return -1; }
    z->fsize = st.st_size;
    for (int i = 0; i < NCACHE; i++) z->cache[i].coff = -1;
    z->cbuf = (uint8_t*)malloc(BGZF_MAX + 64);
    return z->cbuf ? 0 : -1;
}
static void span_free(bgzf_rd* z) { free(z->span.b); free(z->span.data); z->span.b = NULL; z->span.data = NULL; z->span.n = 0; }
static void bgzf_rd_free(bgzf_rd* z) {
    span_free(z);
    for (int i = 0; i < NCACHE; i++) free(z->cache[i].data);
    free(z->cbuf);
    if (z->fd >= 0) close(z->fd);
}
/* the block at file offset coff, inflated (cached); NULL at end of file or on a damaged block */
static bgzf_blk* bgzf_load(bgzf_rd* z, int64_t coff) {
    if (z->span.n && coff >= z->span.b[0].coff && coff <= z->span.b[z->span.n - 1].coff) {
        int lo = 0, hi = z->span.n - 1;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (z->span.b[mid].coff < coff) lo = mid + 1; else hi = mid; }
        if (z->span.b[lo].coff == coff) {
            z->span_view.coff = coff; z->span_view.clen = z->span.b[lo].clen; z->span_view.ulen = z->span.b[lo].ulen;
            z->span_view.data = z->span.data + z->span.b[lo].uoff;
            return &z->span_view;
        }
    }
    int slot = 0;
    for (int i = 0; i < NCACHE; i++) {
        if (z->cache[i].coff == coff) { z->cache[i].stamp = ++z->tick; return &z->cache[i]; }
        if (z->cache[i].stamp < z->cache[slot].stamp) slot = i;
    }
    if (coff >= z->fsize) return NULL;
    uint8_t h[18];
    if (pread(z->fd, h, 18, coff) != 18) { set_err("truncated BGZF header at %lld", (long long)coff); return NULL; }
    if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { set_err("not a BGZF block at %lld", (long long)coff); return NULL; }
    int xlen = rd16(h + 10);
    /* the BC subfield is normally first; walk the extra field in case it is not */
    int bsize = -1;
    if (h[12] == 'B' && h[13] == 'C' && rd16(h + 14) == 2) bsize = rd16(h + 16);
    uint8_t* x = NULL;
    if (bsize < 0) {
        x = (uint8_t*)malloc((size_t)xlen);
        if (pread(z->fd, x, (size_t)xlen, coff + 12) != xlen) { free(x); set_err("truncated BGZF extra field"); return NULL; }
        for (int o = 0; o + 4 <= xlen;) {
            int sl = rd16(x + o + 2);
            if (x[o] == 'B' && x[o + 1] == 'C' && sl == 2 && o + 6 <= xlen) { bsize = rd16(x + o + 4); break; }
            o += 4 + sl;
        }
        free(x);
        if (bsize < 0) { set_err("BGZF block without BC subfield at %lld", (long long)coff); return NULL; }
    }
    int clen = bsize + 1;
    if (clen < 12 + xlen + 8 || coff + clen > z->fsize) { set_err("bad BGZF block size at %lld", (long long)coff); return NULL; }
    if (pread(z->fd, z->cbuf, (size_t)clen, coff) != clen) { set_err("truncated BGZF block at %lld", (long long)coff); return NULL; }
    uint32_t crc = rd32(z->cbuf + clen - 8), isize = rd32(z->cbuf + clen - 4);
    if (isize > BGZF_MAX) { set_err("BGZF ISIZE too large"); return NULL; }
    bgzf_blk* b = &z->cache[slot];
    if (!b->data) b->data = (uint8_t*)malloc(BGZF_MAX);
    if (!b->data) return NULL;
    b->coff = -1;
    z_stream s; memset(&s, 0, sizeof s);
    if (inflateInit2(&s, -15) != Z_OK) { set_err("inflateInit2 failed"); return NULL; }
    s.next_in = z->cbuf + 12 + xlen; s.avail_in = (uInt)(clen - 12 - xlen - 8);
    s.next_out = b->data; s.avail_out = BGZF_MAX;
    int rc = inflate(&s, Z_FINISH);
    inflateEnd(&s);
    if (rc != Z_STREAM_END || s.total_out != isize) { set_err("BGZF inflate failed at %lld", (long long)coff); return NULL; }
    if ((uint32_t)crc32(crc32(0L, NULL, 0), b->data, isize) != crc) { set_err("BGZF CRC mismatch at %lld", (long long)coff); return NULL; }
    b->coff = coff; b->clen = clen; b->ulen = (int32_t)isize; b->stamp = ++z->tick;
    return b;
}

/* ---- parallel inflate of a block range: BGZF blocks are independent deflate streams, and every block states its compressed
 * size in the header (BSIZE) and its inflated size in the trailer (ISIZE), so the output layout is known before inflating */
#include <pthread.h>
typedef struct { const uint8_t* craw; bgzf_span* sp; int64_t base; int t, nt; int fail; } span_job;
static int inflate_one(const uint8_t* blk, int clen, uint8_t* dst, int ulen) {
    int xlen = rd16(blk + 10);
    z_stream s; memset(&s, 0, sizeof s);
    if (inflateInit2(&s, -15) != Z_OK) return -1;
    s.next_in = (Bytef*)(blk + 12 + xlen); s.avail_in = (uInt)(clen - 12 - xlen - 8);
    s.next_out = dst; s.avail_out = (uInt)ulen;
    int rc = inflate(&s, Z_FINISH);
    inflateEnd(&s);
    if (rc != Z_STREAM_END || (int)s.total_out != ulen) return -1;
    if ((uint32_t)crc32(crc32(0L, NULL, 0), dst, (uInt)ulen) != rd32(blk + clen - 8)) return -1;
    return 0;
}
static void* span_worker(void* a) {
    span_job* j = (span_job*)a;
    for (int i = j->t; i < j->sp->n; i += j->nt) {
        span_blk* b = &j->sp->b[i];
        if (inflate_one(j->craw + (b->coff - j->base), b->clen, j->sp->data + b->uoff, b->ulen) != 0) { j->fail = 1; return NULL; }
    }
    return NULL;
}
static int bam_threads(void) {
    static int n = 0;
    if (!n) {
        const char* e = getenv("SWB_BAM_THREADS");
        n = e ? atoi(e) : 4;
        long c = sysconf(_SC_NPROCESSORS_ONLN);
        if (n > c) n = (int)c;
        if (n < 1) n = 1;
        if (n > 32) n = 32;
    }
    return n;
}
/* inflate every block that starts in [c0, c1] (file offsets of block starts) into z->span; 0 = done, 1 = not worth it / not
 * possible (the caller falls back to block-by-block reading), never an error by itself */
#define SPAN_MAX_BYTES (256LL << 20)
static int span_load(bgzf_rd* z, int64_t c0, int64_t c1) {
    span_free(z);
    if (c1 >= z->fsize) c1 = z->fsize - 1;
    if (c1 < c0) return 1;
    int64_t want = c1 - c0 + BGZF_MAX + 64;                   /* the last block starts at c1 and may be a full block long */
    if (c0 + want > z->fsize) want = z->fsize - c0;
    if (want > SPAN_MAX_BYTES || want < 4 * 1024) return 1;   /* tiny ranges: the per-block path and its cache do fine */
    uint8_t* craw = (uint8_t*)malloc((size_t)want);
    if (!craw) return 1;
    if (pread(z->fd, craw, (size_t)want, c0) != want) { free(craw); return 1; }
    int cap = 64, n = 0; span_blk* b = (span_blk*)malloc(sizeof(span_blk) * (size_t)cap);
    int64_t o = 0, utot = 0;
    while (c0 + o <= c1 && o + 18 <= want) {
        const uint8_t* h = craw + o;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4) || h[12] != 'B' || h[13] != 'C' || rd16(h + 14) != 2) { free(craw); free(b); return 1; }
        int clen = rd16(h + 16) + 1;
        if (o + clen > want || clen < 26) { free(craw); free(b); return 1; }
        uint32_t isize = rd32(h + clen - 4);
        if (isize > BGZF_MAX) { free(craw); free(b); return 1; }
        if (n == cap) { cap *= 2; span_blk* q = (span_blk*)realloc(b, sizeof(span_blk) * (size_t)cap); if (!q) { free(craw); free(b); return 1; } b = q; }
        b[n].coff = c0 + o; b[n].clen = clen; b[n].ulen = (int32_t)isize; b[n].uoff = utot; n++;
        utot += isize; o += clen;
    }
    if (n < 2) { free(craw); free(b); return 1; }
    z->span.b = b; z->span.n = n; z->span.data = (uint8_t*)malloc((size_t)utot + 1);
    if (!z->span.data) { free(craw); span_free(z); return 1; }
    int nt = bam_threads(); if (nt > n / 4) nt = n / 4 ? n / 4 : 1;
    span_job jobs[32]; pthread_t th[32]; int started[32] = { 0 };
    for (int t = 0; t < nt; t++) { jobs[t].craw = craw; jobs[t].sp = &z->span; jobs[t].base = c0; jobs[t].t = t; jobs[t].nt = nt; jobs[t].fail = 0; }
    for (int t = 1; t < nt; t++) started[t] = pthread_create(&th[t], NULL, span_worker, &jobs[t]) == 0;
    span_worker(&jobs[0]);
    int fail = jobs[0].fail;
    for (int t = 1; t < nt; t++) { if (started[t]) pthread_join(th[t], NULL); else span_worker(&jobs[t]); fail |= jobs[t].fail; }
    free(craw);
    if (fail) { span_free(z); return 1; }      /* the sequential path will meet the damaged block and report it */
    return 0;
}
typedef struct { bgzf_rd* z; int64_t coff; int32_t uoff; } bgzf_cur;
static inline uint64_t cur_voff(const bgzf_cur* c) { return ((uint64_t)c->coff << 16) | (uint32_t)c->uoff; }
/* reads n bytes; returns n, 0 at a clean end of file (nothing read), -1 on error / truncation */
static int64_t bgzf_read(bgzf_cur* c, void* dst, int64_t n) {
    uint8_t* d = (uint8_t*)dst;
    int64_t got = 0;
    while (got < n) {
        bgzf_blk* b = bgzf_load(c->z, c->coff);
        if (!b) { if (c->coff >= c->z->fsize && got == 0) return 0; if (c->coff >= c->z->fsize) set_err("unexpected end of BAM"); return -1; }
        if (c->uoff >= b->ulen) { c->coff += b->clen; c->uoff = 0; continue; }
        int64_t k = b->ulen - c->uoff; if (k > n - got) k = n - got;
        memcpy(d + got, b->data + c->uoff, (size_t)k);
        got += k; c->uoff += (int32_t)k;
    }
    return got;
}

/* ================================================================== BAM header + BAI */
typedef struct { uint64_t beg, end; } chunk_t;
typedef struct { uint32_t bin; int32_t n; chunk_t* c; } bin_t;
typedef struct { int32_t n_bin; bin_t* bins; int32_t n_intv; uint64_t* ioff; } refidx_t;
struct swb_bam {
    bgzf_rd z;
    char* text; int64_t l_text;
    int32_t n_ref; char** names; int64_t* lens;
    uint64_t first_rec;       /* virtual offset of the first record */
    int has_idx; int32_t idx_nref; refidx_t* idx;
    char* path;               /* for the per-thread reader clones of swb_bam_fetch_pack4 */
};

static void free_index(swb_bam* b) {
    if (!b->idx) return;
    for (int i = 0; i < b->idx_nref; i++) {
        for (int k = 0; k < b->idx[i].n_bin; k++) free(b->idx[i].bins[k].c);
        free(b->idx[i].bins); free(b->idx[i].ioff);
    }
    free(b->idx); b->idx = NULL; b->has_idx = 0;
}
static int load_bai(swb_bam* b, const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return 0;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* buf = (uint8_t*)malloc((size_t)sz + 1);
    if (!buf || fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); return 0; }
    fclose(f);
    int ok = 0; long o = 8;
    if (sz < 8 || memcmp(buf, "BAI\1", 4) != 0) goto done;
    b->idx_nref = (int32_t)rd32(buf + 4);
    if (b->idx_nref < 0 || b->idx_nref > (1 << 24)) goto done;
    b->idx = (refidx_t*)calloc((size_t)b->idx_nref + 1, sizeof(refidx_t));
    for (int i = 0; i < b->idx_nref; i++) {
        refidx_t* r = &b->idx[i];
        if (o + 4 > sz) goto done;
        r->n_bin = (int32_t)rd32(buf + o); o += 4;
        r->bins = (bin_t*)calloc((size_t)r->n_bin + 1, sizeof(bin_t));
        for (int k = 0; k < r->n_bin; k++) {
            if (o + 8 > sz) goto done;
            r->bins[k].bin = rd32(buf + o); r->bins[k].n = (int32_t)rd32(buf + o + 4); o += 8;
            if (r->bins[k].n < 0 || o + 16L * r->bins[k].n > sz) goto done;
            r->bins[k].c = (chunk_t*)malloc(sizeof(chunk_t) * (size_t)(r->bins[k].n + 1));
            for (int c = 0; c < r->bins[k].n; c++) { r->bins[k].c[c].beg = rd64(buf + o); r->bins[k].c[c].end = rd64(buf + o + 8); o += 16; }
        }
        if (o + 4 > sz) goto done;
        r->n_intv = (int32_t)rd32(buf + o); o += 4;
        if (r->n_intv < 0 || o + 8L * r->n_intv > sz) goto done;
        r->ioff = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(r->n_intv + 1));
        for (int k = 0; k < r->n_intv; k++) { r->ioff[k] = rd64(buf + o); o += 8; }
    }
    ok = 1;
done:
    free(buf);
    if (!ok) free_index(b); else b->has_idx = 1;
    return ok;
}

swb_bam* swb_bam_open(const char* path) {
    swb_bam* b = (swb_bam*)calloc(1, sizeof *b);
    if (!b) return NULL;
    if (bgzf_rd_init(&b->z, path) != 0) { free(b); return NULL; }
    b->path = strdup(path);
    bgzf_cur c = { &b->z, 0, 0 };
    uint8_t h[8];
    if (bgzf_read(&c, h, 8) != 8 || memcmp(h, "BAM\1", 4) != 0) { if (!g_err[0] || memcmp(h, "BAM\1", 4)) set_err("%s is not a BAM file", path); goto fail; }
    b->l_text = (int32_t)rd32(h + 4);
    if (b->l_text < 0) { set_err("bad l_text"); goto fail; }
    b->text = (char*)malloc((size_t)b->l_text + 1);
    if (b->l_text && bgzf_read(&c, b->text, b->l_text) != b->l_text) goto fail;
    b->text[b->l_text] = 0;
    if (bgzf_read(&c, h, 4) != 4) goto fail;
    b->n_ref = (int32_t)rd32(h);
    if (b->n_ref < 0) { set_err("bad n_ref"); goto fail; }
    b->names = (char**)calloc((size_t)b->n_ref + 1, sizeof(char*));
    b->lens = (int64_t*)calloc((size_t)b->n_ref + 1, sizeof(int64_t));
    for (int i = 0; i < b->n_ref; i++) {
        if (bgzf_read(&c, h, 4) != 4) goto fail;
        int32_t ln = (int32_t)rd32(h);
        if (ln <= 0 || ln > (1 << 20)) { set_err("bad reference name length"); goto fail; }
        b->names[i] = (char*)malloc((size_t)ln + 1);
        if (bgzf_read(&c, b->names[i], ln) != ln) goto fail;
        b->names[i][ln] = 0;
        if (bgzf_read(&c, h, 4) != 4) goto fail;
        b->lens[i] = (int32_t)rd32(h);
    }
    b->first_rec = cur_voff(&c);
    {   /* <path>.bai, then <stem>.bai */
        size_t L = strlen(path);
        char* p = (char*)malloc(L + 8);
        sprintf(p, "%s.bai", path);
        if (!load_bai(b, p) && L > 4 && strcmp(path + L - 4, ".bam") == 0) { memcpy(p, path, L - 4); strcpy(p + L - 4, ".bai"); load_bai(b, p); }
        free(p);
    }
    return b;
fail:
    swb_bam_close(b);
    return NULL;
}
void swb_bam_close(swb_bam* b) {
    if (!b) return;
    free_index(b);
    for (int i = 0; i < b->n_ref; i++) if (b->names) free(b->names[i]);
    free(b->names); free(b->lens); free(b->text); free(b->path);
    bgzf_rd_free(&b->z);
    free(b);
}
int32_t swb_bam_n_ref(const swb_bam* b) { return b->n_ref; }
const char* swb_bam_ref_name(const swb_bam* b, int32_t tid) { return (tid >= 0 && tid < b->n_ref) ? b->names[tid] : NULL; }
int64_t swb_bam_ref_len(const swb_bam* b, int32_t tid) { return (tid >= 0 && tid < b->n_ref) ? b->lens[tid] : -1; }
int32_t swb_bam_tid(const swb_bam* b, const char* name) {
    for (int i = 0; i < b->n_ref; i++) if (strcmp(b->names[i], name) == 0) return i;
    return -1;
}
const char* swb_bam_header_text(const swb_bam* b, int64_t* len) { if (len) *len = b->l_text; return b->text; }
int swb_bam_has_index(const swb_bam* b) { return b->has_idx; }

/* ================================================================== growable columnar batch */
typedef struct {
    swb_bam_batch pub;
    int64_t cap_n, cap_names, cap_seq, cap_seq4, cap_cigar;
} batch_t;
#define GROW(ptr, cap, need, type) do { if ((need) > (cap)) { int64_t nc_ = (cap) ? (cap) * 2 : 256; while (nc_ < (need)) nc_ *= 2; \
    void* np_ = realloc((ptr), (size_t)nc_ * sizeof(type)); if (!np_) return -1; (ptr) = (type*)np_; (cap) = nc_; } } while (0)
static int batch_reserve_rows(batch_t* t, int64_t need) {
    if (need <= t->cap_n) return 0;
    int64_t nc = t->cap_n ? t->cap_n * 2 : 256; while (nc < need) nc *= 2;
    swb_bam_batch* p = &t->pub;
#define RS(f, type) do { void* q = realloc(p->f, (size_t)nc * sizeof(type)); if (!q) return -1; p->f = (type*)q; } while (0)
    RS(tid, int32_t); RS(pos, int32_t); RS(end, int32_t); RS(flag, uint16_t); RS(mapq, uint8_t); RS(l_seq, int32_t); RS(n_cigar, int32_t);
    RS(next_tid, int32_t); RS(next_pos, int32_t); RS(tlen, int32_t); RS(name_off, int64_t); RS(seq_off, int64_t); RS(cigar_off, int64_t);
    RS(seq4_off, int64_t);
#undef RS
    t->cap_n = nc;
    return 0;
}
void swb_bam_batch_free(swb_bam_batch* q) {
    if (!q) return;
    free(q->tid); free(q->pos); free(q->end); free(q->flag); free(q->mapq); free(q->l_seq); free(q->n_cigar); free(q->next_tid);
    free(q->next_pos); free(q->tlen); free(q->name_off); free(q->seq_off); free(q->cigar_off); free(q->names); free(q->seq);
    free(q->qual); free(q->seq4); free(q->seq4_off); free(q->cigar);
    free(q);
}
static const char NT16[] = "=ACMGRSVTWYHKDBN";
static inline int cigar_reflen(const uint8_t* cig, int n) {
    int r = 0;
    for (int i = 0; i < n; i++) { uint32_t w = rd32(cig + 4 * i); int op = w & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) r += (int)(w >> 4); }
    return r;
}
/* append one raw record (the bytes after block_size) */
static int batch_add(batch_t* t, const uint8_t* r, int32_t bs) {
    swb_bam_batch* p = &t->pub;
    int l_name = r[8], n_cig = rd16(r + 12); int32_t l_seq = (int32_t)rd32(r + 16);
    int64_t need = 32 + (int64_t)l_name + 4LL * n_cig + (l_seq + 1) / 2 + l_seq;
    if (l_seq < 0 || need > bs) { set_err("corrupt BAM record"); return -1; }
    if (batch_reserve_rows(t, p->n + 1) != 0) return -1;
    int64_t i = p->n;
    const uint8_t* name = r + 32; const uint8_t* cig = name + l_name; const uint8_t* sq = cig + 4 * n_cig; const uint8_t* ql = sq + (l_seq + 1) / 2;
    p->tid[i] = (int32_t)rd32(r); p->pos[i] = (int32_t)rd32(r + 4); p->mapq[i] = r[9]; p->flag[i] = rd16(r + 14);
    p->l_seq[i] = l_seq; p->n_cigar[i] = n_cig;
    p->next_tid[i] = (int32_t)rd32(r + 20); p->next_pos[i] = (int32_t)rd32(r + 24); p->tlen[i] = (int32_t)rd32(r + 28);
    p->end[i] = ((p->flag[i] & SWB_BAM_FUNMAP) || n_cig == 0) ? -1 : p->pos[i] + cigar_reflen(cig, n_cig);
    GROW(p->names, t->cap_names, p->names_len + l_name + 1, char);
    p->name_off[i] = p->names_len; memcpy(p->names + p->names_len, name, (size_t)l_name);
    if (l_name == 0 || name[l_name - 1] != 0) { p->names[p->names_len + l_name] = 0; p->names_len += l_name + 1; } else p->names_len += l_name;
    GROW(p->cigar, t->cap_cigar, p->cigar_len + n_cig + 1, uint32_t);
    p->cigar_off[i] = p->cigar_len;
    for (int k = 0; k < n_cig; k++) p->cigar[p->cigar_len + k] = rd32(cig + 4 * k);
    p->cigar_len += n_cig;
    {   /* seq and qual share capacity */
        int64_t needs = p->seq_len + l_seq + 1;
        if (needs > t->cap_seq) {
            int64_t nc = t->cap_seq ? t->cap_seq * 2 : 4096; while (nc < needs) nc *= 2;
            void* a = realloc(p->seq, (size_t)nc); if (!a) return -1; p->seq = (uint8_t*)a;
            void* b = realloc(p->qual, (size_t)nc); if (!b) return -1; p->qual = (uint8_t*)b;
            t->cap_seq = nc;
        }
    }
    p->seq_off[i] = p->seq_len;
    uint8_t* ds = p->seq + p->seq_len;
    for (int32_t k = 0; k + 1 < l_seq; k += 2) { uint8_t v = sq[k >> 1]; ds[k] = (uint8_t)NT16[v >> 4]; ds[k + 1] = (uint8_t)NT16[v & 15]; }
    if (l_seq & 1) ds[l_seq - 1] = (uint8_t)NT16[sq[l_seq >> 1] >> 4];
    memcpy(p->qual + p->seq_len, ql, (size_t)l_seq);
    p->seq_len += l_seq;
    GROW(p->seq4, t->cap_seq4, p->seq4_len + (l_seq + 1) / 2 + 1, uint8_t);
    p->seq4_off[i] = p->seq4_len; memcpy(p->seq4 + p->seq4_len, sq, (size_t)((l_seq + 1) / 2)); p->seq4_len += (l_seq + 1) / 2;
    p->n++;
    return 0;
}

/* ================================================================== region query (SAMv1 §5.3) */
static int reg2bins(int64_t beg, int64_t end, uint16_t* list) {
    int i = 0, k; if (end > (1LL << 29)) end = 1LL << 29; --end;
    if (beg < 0) beg = 0;
    list[i++] = 0;
    for (k = 1 + (int)(beg >> 26); k <= 1 + (int)(end >> 26); ++k) list[i++] = (uint16_t)k;
    for (k = 9 + (int)(beg >> 23); k <= 9 + (int)(end >> 23); ++k) list[i++] = (uint16_t)k;
    for (k = 73 + (int)(beg >> 20); k <= 73 + (int)(end >> 20); ++k) list[i++] = (uint16_t)k;
    for (k = 585 + (int)(beg >> 17); k <= 585 + (int)(end >> 17); ++k) list[i++] = (uint16_t)k;
    for (k = 4681 + (int)(beg >> 14); k <= 4681 + (int)(end >> 14); ++k) list[i++] = (uint16_t)k;
    return i;
}
static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}
static int cmp_chunk(const void* a, const void* b) {
    const chunk_t* x = (const chunk_t*)a; const chunk_t* y = (const chunk_t*)b;
    return x->beg < y->beg ? -1 : x->beg > y->beg ? 1 : 0;
}
/* chunks that may hold records of tid overlapping [beg, end), sorted and merged; *n_out = 0: nothing there */
static chunk_t* region_chunks(const swb_bam* b, int32_t tid, int64_t beg, int64_t end, int* n_out) {
    *n_out = 0;
    if (tid >= b->idx_nref) return NULL;
    const refidx_t* r = &b->idx[tid];
    uint64_t min_off = 0;
    if (r->n_intv > 0) { int64_t w = beg >> 14; min_off = r->ioff[w >= r->n_intv ? r->n_intv - 1 : w]; }
    static __thread uint16_t bins[40000];
    static __thread uint8_t want[37450];             /* membership map of the region's bins (cleared again below) */
    int nb = reg2bins(beg, end, bins), cap = 16, n = 0;
    for (int q = 0; q < nb; q++) if (bins[q] < 37450) want[bins[q]] = 1;
    chunk_t* out = (chunk_t*)malloc(sizeof(chunk_t) * (size_t)cap);
    for (int k = 0; k < r->n_bin; k++) {
        if (r->bins[k].bin >= 37450) continue;       /* pseudo-bin: metadata */
        if (!want[r->bins[k].bin]) continue;
        for (int c = 0; c < r->bins[k].n; c++) {
            if (r->bins[k].c[c].end <= min_off) continue;
            if (n == cap) { cap *= 2; out = (chunk_t*)realloc(out, sizeof(chunk_t) * (size_t)cap); }
            out[n++] = r->bins[k].c[c];
        }
    }
    for (int q = 0; q < nb; q++) if (bins[q] < 37450) want[bins[q]] = 0;
    if (!n) { free(out); return NULL; }
    qsort(out, (size_t)n, sizeof(chunk_t), cmp_chunk);
    int m = 0;
    /* chunks that overlap, touch, or start within one block's reach of the previous end become one run: a writer closes a
       chunk at every block boundary, and reading the few records in between costs less than a seek (every record is
       tested against the region anyway) */
    for (int i = 1; i < n; i++) {
        if (out[i].beg <= out[m].end || (out[i].beg >> 16) <= (out[m].end >> 16) + BGZF_MAX) { if (out[i].end > out[m].end) out[m].end = out[i].end; }
        else out[++m] = out[i];
    }
    *n_out = m + 1;
    return out;
}

typedef int (*rec_cb)(void* ctx, const uint8_t* rec, int32_t bs);
/* walks the records of a region (or of the whole file when tid < 0) and hands the selected ones to cb */
static int64_t scan_region(swb_bam* b, int32_t tid, int64_t beg, int64_t end, uint32_t require, uint32_t exclude, rec_cb cb, void* ctx) {
    chunk_t whole = { b->first_rec, ~0ULL };
    chunk_t* chunks = &whole; int nch = 1; int owned = 0;
    if (tid >= b->n_ref) { set_err("reference index %d out of range", tid); return -1; }
    if (tid >= 0 && b->has_idx) {
        chunks = region_chunks(b, tid, beg, end, &nch);
        owned = 1;
        if (!nch) return 0;
    }
    int64_t count = 0; int rc = 0, stop = 0;
    uint8_t* rec = NULL; int32_t cap = 0;
    for (int ci = 0; ci < nch && !stop; ci++) {
        bgzf_cur c = { &b->z, (int64_t)(chunks[ci].beg >> 16), (int32_t)(chunks[ci].beg & 0xffff) };
        /* the chunk's blocks, inflated together by several threads (whole-file scans: everything up to the last block) */
        span_load(&b->z, c.coff, chunks[ci].end == ~0ULL ? b->z.fsize - 1 : (int64_t)(chunks[ci].end >> 16));
        for (;;) {
            if (cur_voff(&c) >= chunks[ci].end) break;
            uint8_t h[4];
            int64_t g = bgzf_read(&c, h, 4);
            if (g == 0) { stop = 1; break; }
            if (g != 4) { rc = -1; stop = 1; break; }
            int32_t bs = (int32_t)rd32(h);
            if (bs < 32) { set_err("corrupt BAM record (block_size %d)", bs); rc = -1; stop = 1; break; }
            if (bs > cap) { cap = bs * 2; uint8_t* nr = (uint8_t*)realloc(rec, (size_t)cap); if (!nr) { rc = -1; stop = 1; break; } rec = nr; }
            if (bgzf_read(&c, rec, bs) != bs) { rc = -1; stop = 1; break; }
            int32_t rtid = (int32_t)rd32(rec), rpos = (int32_t)rd32(rec + 4);
            if (tid >= 0) {
                if (rtid != tid) { if (rtid > tid || rtid < 0) { stop = 1; break; } continue; }
                if (rpos >= end) { stop = 1; break; }
                int n_cig = rd16(rec + 12), l_name = rec[8];
                if (32 + l_name + 4 * n_cig > bs) { set_err("corrupt BAM record"); rc = -1; stop = 1; break; }
                int rl = cigar_reflen(rec + 32 + l_name, n_cig);
                if ((int64_t)rpos + (rl ? rl : 1) <= beg) continue;
            }
            uint32_t fl = rd16(rec + 14);
            if ((fl & exclude) || (fl & require) != require) continue;
            count++;
            if (cb && cb(ctx, rec, bs) != 0) { rc = -1; stop = 1; break; }
        }
    }
    free(rec);
    span_free(&b->z);
    if (owned) free(chunks);
    return rc ? -1 : count;
}
static int add_cb(void* ctx, const uint8_t* rec, int32_t bs) { return batch_add((batch_t*)ctx, rec, bs); }

swb_bam_batch* swb_bam_fetch(swb_bam* b, int32_t tid, int64_t beg, int64_t end, uint32_t require, uint32_t exclude) {
    batch_t* t = (batch_t*)calloc(1, sizeof *t);
    if (!t) return NULL;
    g_err[0] = 0;
    if (batch_reserve_rows(t, 1) != 0 || scan_region(b, tid, beg, end, require, exclude, add_cb, t) < 0) {
        if (!g_err[0]) set_err("out of memory");
        swb_bam_batch_free(&t->pub); return NULL;
    }
    /* arenas are never NULL so that callers can form views of empty batches */
    if (!t->pub.names) t->pub.names = (char*)calloc(1, 1);
    if (!t->pub.seq) { t->pub.seq = (uint8_t*)calloc(1, 1); t->pub.qual = (uint8_t*)calloc(1, 1); }
    if (!t->pub.seq4) t->pub.seq4 = (uint8_t*)calloc(1, 1);
    if (!t->pub.cigar) t->pub.cigar = (uint32_t*)calloc(1, 4);
    return &t->pub;
}
int64_t swb_bam_count(swb_bam* b, int32_t tid, int64_t beg, int64_t end, uint32_t require, uint32_t exclude) {
    g_err[0] = 0;
    return scan_region(b, tid, beg, end, require, exclude, NULL, NULL);
}

/* BAM nibble -> DNA_BASE_LUT code (sswpy.pyx:16-29): A(1)->0 C(2)->1 G(4)->2 T(8)->3, every other symbol -> 4;
   a BAM byte holds base k in its HIGH nibble, SWB_SEQ_PACKED4 in its LOW nibble */
static const uint8_t* pack4_lut(void) {
    static uint8_t lut[256]; static int ready = 0;
    if (!ready) {
        uint8_t code[16]; for (int i = 0; i < 16; i++) code[i] = 4;
        code[1] = 0; code[2] = 1; code[4] = 2; code[8] = 3;
        for (int v = 0; v < 256; v++) lut[v] = (uint8_t)(code[v >> 4] | code[v & 15] << 4);
        __sync_synchronize(); ready = 1;
    }
    return lut;
}
int64_t swb_bam_batch_pack4(const swb_bam_batch* q, uint8_t* dst, int64_t* dst_off) {
    const uint8_t* lut = pack4_lut();
    int64_t o = 0;
    for (int64_t i = 0; i < q->n; i++) {
        int32_t L = q->l_seq[i]; int64_t nb = (L + 1) / 2;
        const uint8_t* s = q->seq4 + q->seq4_off[i];
        dst_off[i] = o;
        for (int64_t k = 0; k < nb; k++) dst[o + k] = lut[s[k]];
        if (L & 1) dst[o + nb - 1] &= 0x0f;       /* the pad nibble of an odd-length entry is zero */
        o += nb;
    }
    return o;
}

/* ================================================================== many regions -> one packed read table, host threads */
typedef struct {
    uint8_t* tab; int64_t tab_len, tab_cap;
    int32_t* len; int32_t* pos; int32_t* end; uint16_t* flag; uint8_t* mapq; int64_t n, cap;
    int need_cigar, drop_pos0, oom;
} reg_out;
static int pack_cb(void* ctx, const uint8_t* r, int32_t bs) {
    reg_out* o = (reg_out*)ctx;
    int l_name = r[8], n_cig = rd16(r + 12); int32_t l_seq = (int32_t)rd32(r + 16), pos = (int32_t)rd32(r + 4);
    if (l_seq < 0 || 32 + (int64_t)l_name + 4LL * n_cig + (l_seq + 1) / 2 > bs) { set_err("corrupt BAM record"); return -1; }
    if ((o->need_cigar && n_cig == 0) || (o->drop_pos0 && pos == 0)) return 0;
    if (o->n == o->cap) {
        int64_t nc = o->cap ? o->cap * 2 : 512;
        void *a = realloc(o->len, (size_t)nc * 4), *b = realloc(o->pos, (size_t)nc * 4), *c = realloc(o->end, (size_t)nc * 4), *d = realloc(o->flag, (size_t)nc * 2), *e = realloc(o->mapq, (size_t)nc);
        if (a) o->len = (int32_t*)a;
        if (b) o->pos = (int32_t*)b;
        if (c) o->end = (int32_t*)c;
        if (d) o->flag = (uint16_t*)d;
        if (e) o->mapq = (uint8_t*)e;
        if (!a || !b || !c || !d || !e) { o->oom = 1; return -1; }
        o->cap = nc;
    }
    int64_t nb = (l_seq + 1) / 2;
    if (o->tab_len + nb > o->tab_cap) {
        int64_t nc = o->tab_cap ? o->tab_cap * 2 : 65536; while (nc < o->tab_len + nb) nc *= 2;
        void* t = realloc(o->tab, (size_t)nc); if (!t) { o->oom = 1; return -1; }
        o->tab = (uint8_t*)t; o->tab_cap = nc;
    }
    const uint8_t* lut = pack4_lut();
    const uint8_t* sq = r + 32 + l_name + 4 * n_cig; uint8_t* dst = o->tab + o->tab_len;
    for (int64_t k = 0; k < nb; k++) dst[k] = lut[sq[k]];
    if (l_seq & 1) dst[nb - 1] &= 0x0f;
    o->tab_len += nb;
    uint16_t fl = rd16(r + 14);
    o->len[o->n] = l_seq; o->pos[o->n] = pos; o->flag[o->n] = fl; o->mapq[o->n] = r[9];
    o->end[o->n] = ((fl & SWB_BAM_FUNMAP) || n_cig == 0) ? -1 : pos + cigar_reflen(r + 32 + l_name, n_cig);
    o->n++;
    return 0;
}
typedef struct {
    const swb_bam* src; int64_t n_regions; const int32_t* tid; const int64_t* beg; const int64_t* end; uint32_t require, exclude;
    reg_out* out; volatile int64_t* next; int fail; char err[256];
} pack_job;
static void* pack_worker(void* a) {
    pack_job* j = (pack_job*)a;
    swb_bam local = *j->src;                       /* header and index are shared read-only; the file handle, caches and span are its own */
    if (bgzf_rd_init(&local.z, j->src->path) != 0) { j->fail = 1; snprintf(j->err, sizeof j->err, "%.250s", g_err); return NULL; }
    for (;;) {
        int64_t r = __sync_fetch_and_add(j->next, 1);
        if (r >= j->n_regions) break;
        g_err[0] = 0;
        if (j->tid[r] < 0 || scan_region(&local, j->tid[r], j->beg[r], j->end[r], j->require, j->exclude, pack_cb, &j->out[r]) < 0) {
            j->fail = 1; snprintf(j->err, sizeof j->err, "region %lld: %.200s", (long long)r, j->tid[r] < 0 ? "unknown contig" : (j->out[r].oom ? "out of memory" : g_err));
            break;
        }
    }
    bgzf_rd_free(&local.z);
    return NULL;
}
void swb_bam_pack_free(swb_bam_pack* p) {
    if (!p) return;
    free(p->region_first); free(p->read_off); free(p->read_len); free(p->table); free(p->pos); free(p->end); free(p->flag); free(p->mapq); free(p);
}
swb_bam_pack* swb_bam_fetch_pack4(swb_bam* b, int64_t n_regions, const int32_t* tid, const int64_t* beg, const int64_t* end,
                                  uint32_t require, uint32_t exclude, int need_cigar, int drop_pos0, int threads) {
    if (n_regions < 0) { set_err("negative region count"); return NULL; }
    reg_out* out = (reg_out*)calloc((size_t)n_regions + 1, sizeof(reg_out));
    swb_bam_pack* p = (swb_bam_pack*)calloc(1, sizeof *p);
    if (!out || !p) { free(out); free(p); set_err("out of memory"); return NULL; }
    for (int64_t r = 0; r < n_regions; r++) { out[r].need_cigar = need_cigar; out[r].drop_pos0 = drop_pos0; }
    if (threads <= 0) { threads = (int)sysconf(_SC_NPROCESSORS_ONLN); if (threads > 16) threads = 16; }
    if (threads > n_regions) threads = n_regions ? (int)n_regions : 1;
    if (threads > 64) threads = 64;
    volatile int64_t next = 0;
    (void)pack4_lut(); (void)bam_threads();          /* one-time tables set up before the workers start */
    pack_job jobs[64]; pthread_t th[64]; int started[64] = { 0 };
    for (int t = 0; t < threads; t++) {
        jobs[t].src = b; jobs[t].n_regions = n_regions; jobs[t].tid = tid; jobs[t].beg = beg; jobs[t].end = end; jobs[t].require = require; jobs[t].exclude = exclude;
        jobs[t].out = out; jobs[t].next = &next; jobs[t].fail = 0; jobs[t].err[0] = 0;
    }
    for (int t = 1; t < threads; t++) started[t] = pthread_create(&th[t], NULL, pack_worker, &jobs[t]) == 0;
    pack_worker(&jobs[0]);
    int fail = jobs[0].fail; const char* msg = jobs[0].err;
    for (int t = 1; t < threads; t++) { if (started[t]) pthread_join(th[t], NULL); if (jobs[t].fail && !fail) { fail = 1; msg = jobs[t].err; } }
    if (!fail) {
        int64_t n = 0, tb = 0;
        for (int64_t r = 0; r < n_regions; r++) { n += out[r].n; tb += out[r].tab_len; }
        p->n_regions = n_regions; p->n_reads = n; p->table_len = tb;
        p->region_first = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_regions + 1));
        p->read_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1)); p->read_len = (int32_t*)malloc(4 * (size_t)(n + 1));
        p->pos = (int32_t*)malloc(4 * (size_t)(n + 1)); p->end = (int32_t*)malloc(4 * (size_t)(n + 1));
        p->flag = (uint16_t*)malloc(2 * (size_t)(n + 1)); p->mapq = (uint8_t*)malloc((size_t)n + 1); p->table = (uint8_t*)malloc((size_t)tb + 1);
        if (!p->region_first || !p->read_off || !p->read_len || !p->pos || !p->end || !p->flag || !p->mapq || !p->table) { fail = 1; msg = "out of memory"; }
        else {
            int64_t i = 0, o = 0;
            for (int64_t r = 0; r < n_regions; r++) {
                p->region_first[r] = i;
                if (out[r].tab_len) memcpy(p->table + o, out[r].tab, (size_t)out[r].tab_len);
                int64_t oo = o;
                for (int64_t k = 0; k < out[r].n; k++) { p->read_off[i + k] = oo; oo += (out[r].len[k] + 1) / 2; }
                if (out[r].n) {
                    memcpy(p->read_len + i, out[r].len, 4 * (size_t)out[r].n); memcpy(p->pos + i, out[r].pos, 4 * (size_t)out[r].n);
                    memcpy(p->end + i, out[r].end, 4 * (size_t)out[r].n); memcpy(p->flag + i, out[r].flag, 2 * (size_t)out[r].n); memcpy(p->mapq + i, out[r].mapq, (size_t)out[r].n);
                }
                i += out[r].n; o += out[r].tab_len;
            }
            p->region_first[n_regions] = i;
        }
    }
    for (int64_t r = 0; r < n_regions; r++) { free(out[r].tab); free(out[r].len); free(out[r].pos); free(out[r].end); free(out[r].flag); free(out[r].mapq); }
    free(out);
    if (fail) { char tmp[256]; snprintf(tmp, sizeof tmp, "%s", msg && msg[0] ? msg : "fetch failed"); swb_bam_pack_free(p); set_err("%s", tmp); return NULL; }
    return p;
}

/* ================================================================== pileup columns (dictize_read's integer core) */
typedef struct { int64_t n, cap; int32_t* v; } ivec;
static int ivec_push(ivec* a, int32_t x) {
    if (a->n == a->cap) { int64_t nc = a->cap ? a->cap * 2 : 64; int32_t* q = (int32_t*)realloc(a->v, (size_t)nc * 4); if (!q) return -1; a->v = q; a->cap = nc; }
    a->v[a->n++] = x; return 0;
}
/* Python's s[a:b] on a sequence of length len -> [*lo, *hi) */
static void py_slice(int64_t len, int64_t a, int64_t b, int64_t* lo, int64_t* hi) {
    if (a < 0) { a += len; if (a < 0) a = 0; } else if (a > len) a = len;
    if (b < 0) { b += len; if (b < 0) b = 0; } else if (b > len) b = len;
    if (b < a) b = a;
    *lo = a; *hi = b;
}
/* utilities.split (utilities.pyx:429-503), reverse = False: the index j + diff its slices are cut at */
static int32_t split_index(const uint32_t* cig, int n, int32_t target_pos, int32_t string_pos, int is_for_ref) {
    int64_t j = 0, sp = (int64_t)string_pos - 1;
    for (int k = 0; k < n; k++) {
        int op = cig[k] & 15; int64_t len = cig[k] >> 4, g, d;
        if (op == 3) { d = 0; g = len; }
        else if (op == 1) { g = 0; d = is_for_ref ? 0 : len; }
        else if (op == 2) { g = len; d = is_for_ref ? len : 0; }
        else if (op == 5 || op == 6) { d = 0; g = 0; }
        else { g = len; d = len; }
        if (sp < target_pos) sp += g; else break;
        j += d;
    }
    return (int32_t)(j + (target_pos - sp));
}

swb_pileup_cols* swb_pileup_columns(const swb_bam_batch* q, int32_t pos, int32_t rpos, int32_t thresh,
                                    const uint8_t* contig, int64_t contig_start, int64_t contig_len,
                                    int64_t local_start, int64_t local_len) {
    swb_pileup_cols* out = (swb_pileup_cols*)calloc(1, sizeof *out);
    if (!out) return NULL;
    out->n = q->n;
    out->reads = (swb_pileup_read*)calloc((size_t)q->n + 1, sizeof(swb_pileup_read));
    out->ref_seq_off = (int64_t*)calloc((size_t)q->n + 2, sizeof(int64_t));
    ivec sub = { 0, 0, NULL }, ind = { 0, 0, NULL };
    int64_t rcap = 0;
    if (!out->reads || !out->ref_seq_off) goto oom;
    for (int64_t i = 0; i < q->n; i++) {
        swb_pileup_read* R = &out->reads[i];
        const uint32_t* cig = q->cigar + q->cigar_off[i]; int nc = q->n_cigar[i]; int32_t L = q->l_seq[i];
        const uint8_t* seq = q->seq + q->seq_off[i]; const uint8_t* ql = q->qual + q->seq_off[i];
        out->ref_seq_off[i] = out->ref_seq_len;
        R->low_qual_base_num = -1;
        R->subread_off = sub.n / 2; R->indel_off = ind.n / 4;
        if (nc == 0) continue;
        int reflen = 0, n_N = 0;
        for (int k = 0; k < nc; k++) { int op = cig[k] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += (int)(cig[k] >> 4); if (op == 3) n_N++; }
        R->aln_start = q->pos[i] + 1;
        R->start_offset = ((cig[0] & 15) == 4) ? (int32_t)(cig[0] >> 4) : 0;
        R->read_start = R->aln_start - R->start_offset;
        /* reference_end is None for a record flagged unmapped: the reference then adds the reference length to the
           1-based start (pileup.pyx:181-183) */
        R->aln_end = (q->flag[i] & SWB_BAM_FUNMAP) ? R->aln_start + reflen : q->pos[i] + reflen;
        R->end_offset = ((cig[nc - 1] & 15) == 4) ? (int32_t)(cig[nc - 1] >> 4) : 0;
        R->read_end = R->aln_end + R->end_offset;
        R->n_count_gt1 = n_N > 1;
        /* is_dirty / is_end_dirty (pileup.pyx:214, 345-365) */
        { int c = 0; for (int32_t k = 0; k < L; k++) c += ql[k] <= thresh; R->is_dirty = L > 0 && (double)c / (double)L > 0.15; }
        {
            int dl = pos - R->read_start, dr = R->read_end - pos, lefty;
            if (dl < 0) lefty = 1; else if (dr < 0) lefty = 0; else lefty = dl <= dr;
            if (n_N > 1 || L == 0) R->is_end_dirty = 0;
            else {
                int m = 255, k3 = L < 3 ? L : 3;
                if (lefty) { for (int k = 0; k < k3; k++) if (ql[k] < m) m = ql[k]; }
                else { for (int k = L - k3; k < L; k++) if (ql[k] < m) m = ql[k]; }
                R->is_end_dirty = m < thresh;
            }
        }
        /* get_spliced_subreads (utilities.pyx:243-278) */
        {
            int64_t s0 = sub.n;
            if (n_N == 0) { if (ivec_push(&sub, R->read_start) || ivec_push(&sub, R->read_end)) goto oom; }
            else {
                int32_t cur = R->read_start; int prev = -1;
                if (ivec_push(&sub, cur)) goto oom;
                for (int k = 0; k < nc; k++) {
                    int op = cig[k] & 15; int32_t len = (int32_t)(cig[k] >> 4);
                    if (op == 3) { if (ivec_push(&sub, cur - 1)) goto oom; }
                    else if (prev == 3) { if (ivec_push(&sub, cur)) goto oom; }
                    if (!(op == 1 || op == 5 || op == 6)) cur += len;
                    prev = op;
                }
                if (prev != 3) { if (ivec_push(&sub, R->read_end)) goto oom; }
                if ((sub.n - s0) & 1) { if (ivec_push(&sub, sub.v[sub.n - 1])) goto oom; }  /* CIGAR ending in N: unpaired position */
            }
            R->n_subreads = (int32_t)((sub.n - s0) / 2);
            /* parse_spliced_read (pileup.pyx:391-436) */
            int32_t p = pos;
            for (int k = 0; k < R->n_subreads; k++) {
                int32_t a = sub.v[s0 + 2 * k], e = sub.v[s0 + 2 * k + 1];
                if (a <= p && p <= e) { R->is_covering = 1; R->covering_start = a; R->covering_end = e; }
                else if (a <= rpos && rpos <= e) { R->is_covering = 1; R->covering_start = a; R->covering_end = e; p = rpos; }
            }
            R->splice_pos = p;
            if (R->n_subreads > 1) {
                R->is_spliced = 1;
                for (int k = 0; k + 1 < R->n_subreads; k++) {
                    int32_t st = sub.v[s0 + 2 * k + 1] + 1, en = sub.v[s0 + 2 * k + 2] - 1;
                    if (st - 4 <= p && p <= en) { R->intron_start = st; R->intron_end = en; }
                }
            }
        }
        /* locate_indels (utilities.pyx:307-328) + the split indices of leftalign_indel_read (pileup.pyx:315-323) */
        for (int pass = 0; pass < 2; pass++) {
            int32_t cur = R->read_start - 1;
            for (int k = 0; k < nc; k++) {
                int op = cig[k] & 15; int32_t len = (int32_t)(cig[k] >> 4);
                if (op == 1) {
                    if (pass == 0) { if (ivec_push(&ind, cur) || ivec_push(&ind, len) || ivec_push(&ind, split_index(cig, nc, cur, R->read_start, 0)) || ivec_push(&ind, split_index(cig, nc, cur, R->aln_start, 1))) goto oom; R->n_ins++; }
                } else if (op == 2) {
                    if (pass == 1) { if (ivec_push(&ind, cur) || ivec_push(&ind, len) || ivec_push(&ind, split_index(cig, nc, cur, R->read_start, 0)) || ivec_push(&ind, split_index(cig, nc, cur, R->aln_start, 1))) goto oom; R->n_del++; }
                    cur += len;
                } else if (op == 5 || op == 6) { }
                else cur += len;
            }
        }
        /* get_ref_seq (pileup.pyx:269-299) */
        if (contig) {
            int64_t need = out->ref_seq_len + reflen + 8;
            if (need > rcap) { int64_t ncap = rcap ? rcap * 2 : 1 << 16; while (ncap < need) ncap *= 2; uint8_t* nr = (uint8_t*)realloc(out->ref_seq, (size_t)ncap); if (!nr) goto oom; out->ref_seq = nr; rcap = ncap; }
            uint8_t* dst = out->ref_seq + out->ref_seq_len; int64_t w = 0;
            if (n_N == 0) {
                /* UnsplicedLocalReference.get_ref_seq(aln_start - 1, aln_end): a Python slice of the local reference */
                int64_t a = (int64_t)(R->aln_start - 1) - local_start, lo, hi;
                py_slice(local_len, a, a + (R->aln_end - (R->aln_start - 1)), &lo, &hi);
                int64_t g0 = local_start + lo - contig_start;
                if (g0 >= 0 && g0 + (hi - lo) <= contig_len) { memcpy(dst, contig + g0, (size_t)(hi - lo)); w = hi - lo; }
                else for (int64_t x = lo; x < hi; x++) { int64_t g = local_start + x - contig_start; dst[w++] = (g >= 0 && g < contig_len) ? contig[g] : 'N'; }
            } else {
                int64_t cur = R->aln_start - 1;
                for (int k = 0; k < nc; k++) {
                    int op = cig[k] & 15; int64_t len = cig[k] >> 4;
                    if (op == 0 || op == 2) {
                        for (int64_t x = cur; x < cur + len; x++) { int64_t g = x - contig_start; if (g >= 0 && g < contig_len) dst[w++] = contig[g]; }
                        cur += len;
                    } else if (op == 1 || op == 4 || op == 5 || op == 6) { }
                    else cur += len;
                }
            }
            out->ref_seq_len += w;
            /* count_lowqual_non_ref_bases (utilities.pyx:187-218) */
            {
                int64_t a = 0, j = 0; int cnt = 0, ok = 1;
                for (int k = 0; k < nc && ok; k++) {
                    int op = cig[k] & 15; int64_t len = cig[k] >> 4;
                    if (op == 0 || op == 7 || op == 8) {
                        for (int64_t x = 0; x < len; x++) { if (a >= L || j >= w) { ok = 0; break; } if (seq[a] != dst[j] && ql[a] < thresh) cnt++; a++; j++; }
                    } else if (op == 1 || op == 4) {
                        for (int64_t x = 0; x < len; x++) { if (a >= L) { ok = 0; break; } if (ql[a] < thresh) cnt++; a++; }
                    } else if (op == 2) j += len;
                }
                R->low_qual_base_num = cnt;
            }
            R->is_reference_seq = (w == L) && memcmp(seq, dst, (size_t)L) == 0;
        }
    }
    out->ref_seq_off[q->n] = out->ref_seq_len;
    out->subreads = sub.v; out->n_subreads = sub.n / 2;
    out->indels = (swb_pileup_indel*)ind.v; out->n_indels = ind.n / 4;
    if (!out->subreads) out->subreads = (int32_t*)calloc(2, 4);
    if (!out->indels) out->indels = (swb_pileup_indel*)calloc(1, sizeof(swb_pileup_indel));
    if (!out->ref_seq) out->ref_seq = (uint8_t*)calloc(1, 1);
    return out;
oom:
    set_err("out of memory");
    free(sub.v); free(ind.v);
    swb_pileup_cols_free(out);
    return NULL;
}
void swb_pileup_cols_free(swb_pileup_cols* c) {
    if (!c) return;
    free(c->reads); free(c->subreads); free(c->indels); free(c->ref_seq); free(c->ref_seq_off); free(c);
}

/* ================================================================== BGZF / BAM / BAI writer */
#define WBLOCK 0xff00
typedef struct { uint64_t beg, end; } wchunk;
typedef struct { int n, cap; wchunk* c; } wbin;
typedef struct { wbin* bins; /* 37450 */ uint64_t* ioff; int n_intv, cap_intv; int used; } wref;
struct swb_bam_writer {
    FILE* f; char* path; int level;
    uint8_t ubuf[BGZF_MAX]; int ufill;
    uint8_t cbuf[BGZF_MAX + 1024];
    int64_t coff;
    int32_t n_ref; wref* refs;
    int32_t last_tid, last_pos;
    int failed;
};
static int w_flush(swb_bam_writer* w) {
    if (w->ufill == 0) return 0;
    z_stream s; memset(&s, 0, sizeof s);
    if (deflateInit2(&s, w->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    s.next_in = w->ubuf; s.avail_in = (uInt)w->ufill; s.next_out = w->cbuf + 18; s.avail_out = BGZF_MAX - 18 - 8;
    int rc = deflate(&s, Z_FINISH);
    deflateEnd(&s);
    if (rc != Z_STREAM_END) { set_err("deflate failed (incompressible block)"); return -1; }
    int clen = (int)s.total_out + 26;
    static const uint8_t hd[12] = { 31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0 };
    memcpy(w->cbuf, hd, 12); w->cbuf[12] = 'B'; w->cbuf[13] = 'C'; wr16(w->cbuf + 14, 2); wr16(w->cbuf + 16, (uint32_t)(clen - 1));
    wr32(w->cbuf + clen - 8, (uint32_t)crc32(crc32(0L, NULL, 0), w->ubuf, (uInt)w->ufill));
    wr32(w->cbuf + clen - 4, (uint32_t)w->ufill);
    if (fwrite(w->cbuf, 1, (size_t)clen, w->f) != (size_t)clen) { set_err("write failed: %s", strerror(errno)); return -1; }
    w->coff += clen; w->ufill = 0;
    return 0;
}
static int w_put(swb_bam_writer* w, const void* src, int64_t n) {
    const uint8_t* s = (const uint8_t*)src;
    while (n > 0) {
        int k = WBLOCK - w->ufill; if (k > n) k = (int)n;
        memcpy(w->ubuf + w->ufill, s, (size_t)k); w->ufill += k; s += k; n -= k;
        if (w->ufill == WBLOCK && w_flush(w) != 0) return -1;
    }
    return 0;
}
static inline uint64_t w_tell(const swb_bam_writer* w) { return ((uint64_t)w->coff << 16) | (uint32_t)w->ufill; }

swb_bam_writer* swb_bam_create(const char* path, const char* text, int32_t n_ref, const char* const* names, const int64_t* lens, int level) {
    swb_bam_writer* w = (swb_bam_writer*)calloc(1, sizeof *w);
    if (!w) return NULL;
    w->f = fopen(path, "wb");
    if (!w->f) { set_err("cannot create %s: %s", path, strerror(errno)); free(w); return NULL; }
    w->path = strdup(path); w->level = (level < 0 || level > 9) ? 6 : level; w->n_ref = n_ref; w->last_tid = 0; w->last_pos = -1;
    w->refs = (wref*)calloc((size_t)n_ref + 1, sizeof(wref));
    uint8_t h[8]; memcpy(h, "BAM\1", 4);
    size_t lt = text ? strlen(text) : 0;
    wr32(h + 4, (uint32_t)lt);
    int rc = w_put(w, h, 8);
    if (!rc && lt) rc = w_put(w, text, (int64_t)lt);
    wr32(h, (uint32_t)n_ref); if (!rc) rc = w_put(w, h, 4);
    for (int i = 0; i < n_ref && !rc; i++) {
        size_t ln = strlen(names[i]) + 1;
        wr32(h, (uint32_t)ln); rc = w_put(w, h, 4);
        if (!rc) rc = w_put(w, names[i], (int64_t)ln);
        wr32(h, (uint32_t)lens[i]); if (!rc) rc = w_put(w, h, 4);
    }
    if (!rc) rc = w_flush(w);       /* records start on a block boundary */
    if (rc) { w->failed = 1; swb_bam_writer_close(w, 0); return NULL; }
    return w;
}
static const uint8_t ASCII2NT16[256] = {
    ['='] = 0, ['A'] = 1, ['C'] = 2, ['M'] = 3, ['G'] = 4, ['R'] = 5, ['S'] = 6, ['V'] = 7, ['T'] = 8, ['W'] = 9, ['Y'] = 10, ['H'] = 11, ['K'] = 12, ['D'] = 13, ['B'] = 14, ['N'] = 15,
    ['a'] = 1, ['c'] = 2, ['m'] = 3, ['g'] = 4, ['r'] = 5, ['s'] = 6, ['v'] = 7, ['t'] = 8, ['w'] = 9, ['y'] = 10, ['h'] = 11, ['k'] = 12, ['d'] = 13, ['b'] = 14, ['n'] = 15 };
static int w_index_add(swb_bam_writer* w, int32_t tid, int64_t beg, int64_t end, int bin, uint64_t v0, uint64_t v1) {
    wref* r = &w->refs[tid];
    if (!r->bins) { r->bins = (wbin*)calloc(37450, sizeof(wbin)); if (!r->bins) return -1; }
    wbin* b = &r->bins[bin];
    if (b->n && (b->c[b->n - 1].end >> 16) == (v0 >> 16)) b->c[b->n - 1].end = v1;      /* same block: extend the last chunk */
    else {
        if (b->n == b->cap) { int nc = b->cap ? b->cap * 2 : 4; wchunk* q = (wchunk*)realloc(b->c, sizeof(wchunk) * (size_t)nc); if (!q) return -1; b->c = q; b->cap = nc; }
        b->c[b->n].beg = v0; b->c[b->n].end = v1; b->n++;
    }
    int64_t w0 = beg >> 14, w1 = (end - 1) >> 14;
    if (w1 + 1 > r->cap_intv) { int nc = r->cap_intv ? r->cap_intv : 64; while (nc < w1 + 1) nc *= 2; uint64_t* q = (uint64_t*)realloc(r->ioff, sizeof(uint64_t) * (size_t)nc); if (!q) return -1; memset(q + r->cap_intv, 0, sizeof(uint64_t) * (size_t)(nc - r->cap_intv)); r->ioff = q; r->cap_intv = nc; }
    for (int64_t k = w0; k <= w1; k++) if (r->ioff[k] == 0) r->ioff[k] = v0;
    if (w1 + 1 > r->n_intv) r->n_intv = (int)(w1 + 1);
    r->used = 1;
    return 0;
}
int swb_bam_write(swb_bam_writer* w, int64_t n, const int32_t* tid, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq,
                  const int32_t* l_seq, const int32_t* n_cigar, const int64_t* name_off, const int64_t* seq_off, const int64_t* cigar_off,
                  const char* names, const uint8_t* seq, const uint8_t* qual, const uint32_t* cigar,
                  const int32_t* next_tid, const int32_t* next_pos, const int32_t* tlen) {
    if (w->failed) { set_err("writer already failed"); return -1; }
    uint8_t* rec = NULL; int64_t cap = 0;
    for (int64_t i = 0; i < n; i++) {
        const char* nm = names + name_off[i]; size_t ln = strlen(nm) + 1;
        int32_t L = l_seq[i], nc = n_cigar[i];
        if (ln > 255 || L < 0 || nc < 0 || nc > 65535 || tid[i] >= w->n_ref) { set_err("record %lld: field out of range", (long long)i); free(rec); return -1; }
        /* coordinate order: mapped references ascending, unplaced (tid -1) last */
        uint32_t kt = (uint32_t)tid[i], lt = (uint32_t)w->last_tid;
        if (kt < lt || (kt == lt && pos[i] < w->last_pos)) { set_err("record %lld is not in coordinate order", (long long)i); free(rec); return -1; }
        w->last_tid = tid[i]; w->last_pos = pos[i];
        const uint32_t* cg = cigar + cigar_off[i];
        int reflen = 0;
        for (int k = 0; k < nc; k++) { int op = cg[k] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += (int)(cg[k] >> 4); }
        int64_t end = (int64_t)pos[i] + (((flag[i] & SWB_BAM_FUNMAP) || !reflen) ? 1 : reflen);
        int bin = tid[i] >= 0 && pos[i] >= 0 ? reg2bin(pos[i], end) : 4680;
        int64_t bs = 32 + (int64_t)ln + 4LL * nc + (L + 1) / 2 + L;
        if (bs + 4 > cap) { cap = (bs + 4) * 2; uint8_t* q = (uint8_t*)realloc(rec, (size_t)cap); if (!q) { free(rec); return -1; } rec = q; }
        uint8_t* p = rec;
        wr32(p, (uint32_t)bs); wr32(p + 4, (uint32_t)tid[i]); wr32(p + 8, (uint32_t)pos[i]);
        p[12] = (uint8_t)ln; p[13] = mapq[i]; wr16(p + 14, (uint32_t)bin); wr16(p + 16, (uint32_t)nc); wr16(p + 18, flag[i]);
        wr32(p + 20, (uint32_t)L); wr32(p + 24, (uint32_t)(next_tid ? next_tid[i] : -1)); wr32(p + 28, (uint32_t)(next_pos ? next_pos[i] : -1));
        wr32(p + 32, (uint32_t)(tlen ? tlen[i] : 0));
        memcpy(p + 36, nm, ln);
        uint8_t* q = p + 36 + ln;
        for (int k = 0; k < nc; k++) wr32(q + 4 * k, cg[k]);
        q += 4 * nc;
        const uint8_t* s = seq + seq_off[i];
        for (int32_t k = 0; k < L; k += 2) {
            uint8_t hi = ASCII2NT16[s[k]], lo = (k + 1 < L) ? ASCII2NT16[s[k + 1]] : 0;
            if (!hi && s[k] != '=') hi = 15;
            if (k + 1 < L && !lo && s[k + 1] != '=') lo = 15;
            q[k >> 1] = (uint8_t)(hi << 4 | lo);
        }
        q += (L + 1) / 2;
        if (qual) memcpy(q, qual + seq_off[i], (size_t)L); else memset(q, 0xff, (size_t)L);
        /* keep a record inside one block when it fits (the reader does not need it, the index is simpler to reason about) */
        if (bs + 4 <= WBLOCK && w->ufill + bs + 4 > WBLOCK && w_flush(w) != 0) { w->failed = 1; free(rec); return -1; }
        uint64_t v0 = w_tell(w);
        if (w_put(w, rec, bs + 4) != 0) { w->failed = 1; free(rec); return -1; }
        uint64_t v1 = w_tell(w);
        if (tid[i] >= 0 && pos[i] >= 0 && w_index_add(w, tid[i], pos[i], end, bin, v0, v1) != 0) { w->failed = 1; free(rec); return -1; }
    }
    free(rec);
    return 0;
}
int swb_bam_writer_close(swb_bam_writer* w, int index) {
    if (!w) return -1;
    int rc = w->failed ? -1 : 0;
    if (!rc) rc = w_flush(w);
    static const uint8_t eof[28] = { 31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    if (!rc && fwrite(eof, 1, 28, w->f) != 28) rc = -1;
    if (fclose(w->f) != 0) rc = -1;
    if (!rc && index) {
        size_t L = strlen(w->path); char* p = (char*)malloc(L + 8); sprintf(p, "%s.bai", w->path);
        FILE* f = fopen(p, "wb");
        if (!f) { set_err("cannot create %s", p); rc = -1; }
        else {
            uint8_t h[16]; memcpy(h, "BAI\1", 4); wr32(h + 4, (uint32_t)w->n_ref); fwrite(h, 1, 8, f);
            for (int t = 0; t < w->n_ref; t++) {
                wref* r = &w->refs[t]; int nb = 0;
                if (r->bins) for (int k = 0; k < 37450; k++) nb += r->bins[k].n > 0;
                wr32(h, (uint32_t)nb); fwrite(h, 1, 4, f);
                if (r->bins) for (int k = 0; k < 37450; k++) {
                    wbin* b = &r->bins[k]; if (!b->n) continue;
                    wr32(h, (uint32_t)k); wr32(h + 4, (uint32_t)b->n); fwrite(h, 1, 8, f);
                    for (int c = 0; c < b->n; c++) { wr64(h, b->c[c].beg); wr64(h + 8, b->c[c].end); fwrite(h, 1, 16, f); }
                }
                /* windows no record starts in inherit the offset of the window before (htslib does the same) */
                for (int k = 1; k < r->n_intv; k++) if (r->ioff[k] == 0) r->ioff[k] = r->ioff[k - 1];
                wr32(h, (uint32_t)r->n_intv); fwrite(h, 1, 4, f);
                for (int k = 0; k < r->n_intv; k++) { wr64(h, r->ioff[k]); fwrite(h, 1, 8, f); }
            }
            if (fclose(f) != 0) rc = -1;
        }
        free(p);
    }
    for (int t = 0; t < w->n_ref; t++) {
        if (w->refs[t].bins) { for (int k = 0; k < 37450; k++) free(w->refs[t].bins[k].c); free(w->refs[t].bins); }
        free(w->refs[t].ioff);
    }
    free(w->refs); free(w->path); free(w);
    return rc;
}

/* ================================================================== FASTA + .fai */
typedef struct { char* name; int64_t len, off; int32_t line_bases, line_width; } fai_ent;
struct swb_fai { int fd; int64_t fsize; int32_t n; fai_ent* e; };

int swb_fasta_write(const char* path, int32_t n, const char* const* names, const char* const* seqs, const int64_t* lens, int lw) {
    if (lw <= 0) lw = 60;
    FILE* f = fopen(path, "wb");
    if (!f) { set_err("cannot create %s: %s", path, strerror(errno)); return -1; }
    size_t L = strlen(path); char* p = (char*)malloc(L + 8); sprintf(p, "%s.fai", path);
    FILE* x = fopen(p, "wb"); free(p);
    if (!x) { fclose(f); set_err("cannot create %s.fai", path); return -1; }
    int64_t off = 0;
    for (int i = 0; i < n; i++) {
        off += fprintf(f, ">%s\n", names[i]);
        fprintf(x, "%s\t%lld\t%lld\t%d\t%d\n", names[i], (long long)lens[i], (long long)off, lw, lw + 1);
        for (int64_t k = 0; k < lens[i]; k += lw) {
            int64_t m = lens[i] - k < lw ? lens[i] - k : lw;
            fwrite(seqs[i] + k, 1, (size_t)m, f); fputc('\n', f); off += m + 1;
        }
    }
    int rc = (fclose(f) != 0) | (fclose(x) != 0);
    return rc ? -1 : 0;
}
static int fai_push(swb_fai* f, const char* name, int64_t len, int64_t off, int lb, int lw) {
    fai_ent* q = (fai_ent*)realloc(f->e, sizeof(fai_ent) * (size_t)(f->n + 1));
    if (!q) return -1;
    f->e = q; f->e[f->n].name = strdup(name); f->e[f->n].len = len; f->e[f->n].off = off; f->e[f->n].line_bases = lb; f->e[f->n].line_width = lw; f->n++;
    return 0;
}
static int fai_build(swb_fai* f) {     /* one pass over the FASTA (samtools faidx's rules: uniform line length but for the last line) */
    int64_t bufsz = 1 << 20, pos = 0; uint8_t* buf = (uint8_t*)malloc((size_t)bufsz);
    char name[1024]; int in_name = 0, nl = 0, have = 0; int64_t len = 0, off = 0, line_b = 0, line_w = 0, cur_b = 0, cur_w = 0; int first_line = 1, name_done = 0;
    for (;;) {
        ssize_t g = pread(f->fd, buf, (size_t)bufsz, pos);
        if (g <= 0) break;
        for (ssize_t k = 0; k < g; k++) {
            uint8_t c = buf[k];
            if (in_name) {
                if (c == '\n') { in_name = 0; name[nl] = 0; off = pos + k + 1; len = 0; first_line = 1; cur_b = cur_w = 0; line_b = line_w = 0; }
                else if (!name_done) { if (c == ' ' || c == '\t' || c == '\r') name_done = 1; else if (nl < 1023) name[nl++] = (char)c; }
                continue;
            }
            if (c == '>' && cur_w == 0) {
                if (have && fai_push(f, name, len, off, (int)line_b, (int)line_w) != 0) { free(buf); return -1; }
                have = 1; in_name = 1; nl = 0; name_done = 0; continue;
            }
            cur_w++;
            if (c == '\n') { if (first_line && cur_b) { line_b = cur_b; line_w = cur_w; first_line = 0; } cur_b = cur_w = 0; }
            else if (c != '\r') { cur_b++; len++; }
        }
        pos += g;
    }
    if (have) { if (first_line && cur_b) { line_b = cur_b; line_w = cur_b + 1; } if (fai_push(f, name, len, off, (int)line_b, (int)line_w) != 0) { free(buf); return -1; } }
    free(buf);
    return 0;
}
swb_fai* swb_fai_open(const char* path) {
    swb_fai* f = (swb_fai*)calloc(1, sizeof *f);
    if (!f) return NULL;
    f->fd = open(path, O_RDONLY);
    if (f->fd < 0) { set_err("cannot open %s: %s", path, strerror(errno)); free(f); return NULL; }
    struct stat st; fstat(f->fd, &st); f->fsize = st.st_size;
    size_t L = strlen(path); char* p = (char*)malloc(L + 8); sprintf(p, "%s.fai", path);
    FILE* x = fopen(p, "r"); free(p);
    if (x) {
        char line[4096];
        while (fgets(line, sizeof line, x)) {
            char nm[2048]; long long len, off; int lb, lw;
            if (sscanf(line, "%2047[^\t]\t%lld\t%lld\t%d\t%d", nm, &len, &off, &lb, &lw) == 5) fai_push(f, nm, len, off, lb, lw);
        }
        fclose(x);
    } else if (fai_build(f) != 0) { swb_fai_close(f); set_err("cannot index %s", path); return NULL; }
    return f;
}
void swb_fai_close(swb_fai* f) {
    if (!f) return;
    for (int i = 0; i < f->n; i++) free(f->e[i].name);
    free(f->e); if (f->fd >= 0) close(f->fd); free(f);
}
int32_t swb_fai_n(const swb_fai* f) { return f->n; }
const char* swb_fai_name(const swb_fai* f, int32_t i) { return (i >= 0 && i < f->n) ? f->e[i].name : NULL; }
static const fai_ent* fai_find(const swb_fai* f, const char* name) {
    for (int i = 0; i < f->n; i++) if (strcmp(f->e[i].name, name) == 0) return &f->e[i];
    return NULL;
}
int64_t swb_fai_len(const swb_fai* f, const char* name) { const fai_ent* e = fai_find(f, name); return e ? e->len : -1; }
int64_t swb_fai_fetch(const swb_fai* f, const char* name, int64_t beg, int64_t end, char* dst) {
    const fai_ent* e = fai_find(f, name);
    if (!e) { set_err("sequence %s not in the FASTA index", name); return -1; }
    if (beg < 0) beg = 0;
    if (end > e->len) end = e->len;
    if (end <= beg) return 0;
    int64_t lb = e->line_bases > 0 ? e->line_bases : e->len, lw = e->line_width > 0 ? e->line_width : lb + 1;
    int64_t fo0 = e->off + (beg / lb) * lw + beg % lb, fo1 = e->off + ((end - 1) / lb) * lw + (end - 1) % lb + 1;
    int64_t nraw = fo1 - fo0;
    uint8_t* raw = (uint8_t*)malloc((size_t)nraw + 1);
    if (!raw) return -1;
    if (pread(f->fd, raw, (size_t)nraw, fo0) != nraw) { free(raw); set_err("short read from FASTA"); return -1; }
    int64_t w = 0;
    for (int64_t k = 0; k < nraw; k++) if (raw[k] != '\n' && raw[k] != '\r') dst[w++] = (char)raw[k];
    free(raw);
    return w;
}

/* CIGAR strings of a batch ("70M1D80M"): entry i at dst + off[i] (NUL terminated), off has n + 1 entries.  Returns the bytes
 * needed; nothing is written when cap is smaller. */
int64_t swb_bam_batch_cigar_text(const swb_bam_batch* q, char* dst, int64_t cap, int64_t* off) {
    static const char OPS[] = "MIDNSHP=XB??????";
    int64_t need = 0;
    for (int64_t i = 0; i < q->n; i++) need += 11LL * q->n_cigar[i] + 1;
    if (cap < need || !dst) return need;
    int64_t o = 0;
    for (int64_t i = 0; i < q->n; i++) {
        off[i] = o;
        const uint32_t* c = q->cigar + q->cigar_off[i];
        for (int k = 0; k < q->n_cigar[i]; k++) { o += sprintf(dst + o, "%u", c[k] >> 4); dst[o++] = OPS[c[k] & 15]; }
        dst[o++] = 0;
    }
    off[q->n] = o;
    return o;
}
int32_t swb_pileup_read_size(void) { return (int32_t)sizeof(swb_pileup_read); }

/* n slices back to back into dst: slice i at dst + off[i] (off has n + 1 entries).  Returns the bytes written, or -(bytes needed) when
 * cap is too small, or INT64_MIN on a missing sequence. */
int64_t swb_fai_fetch_many(const swb_fai* f, int64_t n, const char* const* names, const int64_t* beg, const int64_t* end, char* dst, int64_t cap, int64_t* off) {
    int64_t need = 0;
    for (int64_t i = 0; i < n; i++) {
        const fai_ent* e = fai_find(f, names[i]);
        if (!e) { set_err("sequence %s not in the FASTA index", names[i]); return INT64_MIN; }
        int64_t b = beg[i] < 0 ? 0 : beg[i], x = end[i] > e->len ? e->len : end[i];
        need += x > b ? x - b : 0;
    }
    if (need > cap) return -need;
    int64_t o = 0;
    for (int64_t i = 0; i < n; i++) {
        off[i] = o;
        int64_t w = swb_fai_fetch(f, names[i], beg[i], end[i], dst + o);
        if (w < 0) return INT64_MIN;
        o += w;
    }
    off[n] = o;
    return o;
}
