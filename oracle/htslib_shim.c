/*
 * oracle/htslib_shim.c  --  TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference's hot-path sources (src/corelib/{bam_info,build_mod_bam,bam_mod_parser}.cpp,
 * src/app/hifimeth/eval_kmer_features.cpp) compile here from /root/reference/src but leave nine
 * htslib aux accessors unresolved, because libhts itself is not in this image.
 * This file supplies those nine functions, written from the BAM specification (SAMv1 s4.2.4,
 * "auxiliary data") and the contracts documented in the vendored header src/htslib/sam.h:1700-1930.
 * It is our own code, not a copy of htslib.  Struct layout comes from the vendored header.
 */
#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <htslib/sam.h>

static int aux_elem_size(int t)
{
    switch (t) {
    case 'A': case 'c': case 'C': return 1;
    case 's': case 'S': return 2;
    case 'i': case 'I': case 'f': return 4;
    case 'd': return 8;
    default: return 0;
    }
}

/* s points at the type byte; returns pointer one past the field, or NULL if corrupt. */
static uint8_t *aux_skip(uint8_t *s, uint8_t *end)
{
    if (s >= end) return NULL;
    int t = *s++;
    if (t == 'Z' || t == 'H') {
        while (s < end && *s) ++s;
        return s < end ? s + 1 : NULL;
    }
    if (t == 'B') {
        if (end - s < 5) return NULL;
        int es = aux_elem_size(*s++);
        uint32_t n;
        memcpy(&n, s, 4);
        s += 4;
        if (es == 0 || (uint64_t)(end - s) < (uint64_t)es * n) return NULL;
        return s + (size_t)es * n;
    }
    int es = aux_elem_size(t);
    if (es == 0 || end - s < es) return NULL;
    return s + es;
}

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2])
{
    uint8_t *s = bam_get_aux(b);
    uint8_t *end = b->data + b->l_data;
    while (s != NULL && end - s >= 3) {
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return s + 2;
        s = aux_skip(s + 2, end);
    }
    errno = (s == NULL) ? EINVAL : ENOENT;
    return NULL;
}

uint32_t bam_auxB_len(const uint8_t *s)
{
    if (s[0] != 'B') { errno = EINVAL; return 0; }
    uint32_t n;
    memcpy(&n, s + 2, 4);
    return n;
}

int64_t bam_auxB2i(const uint8_t *s, uint32_t idx)
{
    uint32_t n = bam_auxB_len(s);
    if (idx >= n) { errno = ERANGE; return 0; }
    const uint8_t *p = s + 6;
    switch (s[1]) {
    case 'c': return (int8_t)p[idx];
    case 'C': return p[idx];
    case 's': { int16_t v; memcpy(&v, p + 2 * (size_t)idx, 2); return v; }
    case 'S': { uint16_t v; memcpy(&v, p + 2 * (size_t)idx, 2); return v; }
    case 'i': { int32_t v; memcpy(&v, p + 4 * (size_t)idx, 4); return v; }
    case 'I': { uint32_t v; memcpy(&v, p + 4 * (size_t)idx, 4); return v; }
    default: errno = EINVAL; return 0;
    }
}

int64_t bam_aux2i(const uint8_t *s)
{
    switch (s[0]) {
    case 'c': return (int8_t)s[1];
    case 'C': return s[1];
    case 's': { int16_t v; memcpy(&v, s + 1, 2); return v; }
    case 'S': { uint16_t v; memcpy(&v, s + 1, 2); return v; }
    case 'i': { int32_t v; memcpy(&v, s + 1, 4); return v; }
    case 'I': { uint32_t v; memcpy(&v, s + 1, 4); return v; }
    default: errno = EINVAL; return 0;
    }
}

char *bam_aux2Z(const uint8_t *s)
{
    if (s[0] == 'Z' || s[0] == 'H') return (char *)(s + 1);
    errno = EINVAL;
    return NULL;
}

int bam_aux_del(bam1_t *b, uint8_t *s)
{
    uint8_t *end = b->data + b->l_data;
    uint8_t *next = aux_skip(s, end);
    if (!next) { errno = EINVAL; return -1; }
    uint8_t *from = s - 2;
    memmove(from, next, (size_t)(end - next));
    b->l_data -= (int)(next - from);
    return 0;
}

static int grow(bam1_t *b, size_t extra)
{
    size_t need = (size_t)b->l_data + extra;
    if (need > 0x7fffffffu) { errno = ENOMEM; return -1; }
    if (need <= b->m_data) return 0;
    size_t cap = need + need / 2 + 64;
    uint8_t *p = (uint8_t *)realloc(b->data, cap);
    if (!p) { errno = ENOMEM; return -1; }
    b->data = p;
    b->m_data = (uint32_t)cap;
    return 0;
}

/* Replace an existing field (located at type byte s) by new_len payload bytes starting at the
 * type byte; keeps the field's position (sam.h: "will not change the ordering of tags"). */
static int replace_field(bam1_t *b, uint8_t *s, const uint8_t *payload, size_t new_len)
{
    uint8_t *end = b->data + b->l_data;
    uint8_t *next = aux_skip(s, end);
    if (!next) { errno = EINVAL; return -1; }
    size_t old_len = (size_t)(next - s);
    size_t s_off = (size_t)(s - b->data);
    size_t tail = (size_t)(end - next);
    if (new_len > old_len) {
        if (grow(b, new_len - old_len)) return -1;
    }
    s = b->data + s_off;
    memmove(s + new_len, s + old_len, tail);
    memcpy(s, payload, new_len);
    b->l_data = (int)((size_t)b->l_data + new_len - old_len);
    return 0;
}

static int append_field(bam1_t *b, const char tag[2], const uint8_t *payload, size_t len)
{
    if (grow(b, len + 2)) return -1;
    uint8_t *p = b->data + b->l_data;
    p[0] = (uint8_t)tag[0];
    p[1] = (uint8_t)tag[1];
    memcpy(p + 2, payload, len);
    b->l_data += (int)(len + 2);
    return 0;
}

int bam_aux_update_str(bam1_t *b, const char tag[2], int len, const char *data)
{
    if (len < 0) len = (int)strlen(data);
    int need_nul = (len == 0 || data[len - 1] != '\0');
    size_t plen = 1 + (size_t)len + (size_t)need_nul;
    uint8_t *payload = (uint8_t *)malloc(plen);
    if (!payload) { errno = ENOMEM; return -1; }
    payload[0] = 'Z';
    memcpy(payload + 1, data, (size_t)len);
    if (need_nul) payload[plen - 1] = 0;
    int rc;
    uint8_t *s = bam_aux_get(b, tag);
    if (s) {
        if (s[0] != 'Z') { free(payload); errno = EINVAL; return -1; }
        rc = replace_field(b, s, payload, plen);
    } else if (errno == ENOENT) {
        rc = append_field(b, tag, payload, plen);
    } else {
        rc = -1;
    }
    free(payload);
    return rc;
}

int bam_aux_update_int(bam1_t *b, const char tag[2], int64_t val)
{
    /* htslib 1.19.1 sam.c (not vendored; contract in sam.h:1844-1866): the smallest type that holds val, compared with `<`
     * (255 -> 'S', 65535 -> 'I'); an existing integer field keeps its width when val fits in it. */
    uint8_t payload[5];
    size_t plen;
    int sz, neg = val < 0;
    if (val < INT32_MIN || val > UINT32_MAX) { errno = EOVERFLOW; return -1; }
    if (val < INT16_MIN) sz = 4;
    else if (val < INT8_MIN) sz = 2;
    else if (val < 0) sz = 1;
    else if (val < UINT8_MAX) sz = 1;
    else if (val < UINT16_MAX) sz = 2;
    else sz = 4;
    uint8_t *s = bam_aux_get(b, tag);
    if (s) {
        int old_sz;
        switch (s[0]) {
        case 'c': case 'C': old_sz = 1; break;
        case 's': case 'S': old_sz = 2; break;
        case 'i': case 'I': old_sz = 4; break;
        default: errno = EINVAL; return -1;
        }
        if (old_sz > sz) sz = old_sz;
    }
    payload[0] = (uint8_t)(neg ? (sz == 1 ? 'c' : sz == 2 ? 's' : 'i') : (sz == 1 ? 'C' : sz == 2 ? 'S' : 'I'));
    {
        uint32_t v = (uint32_t)(int32_t)val;
        if (!neg) v = (uint32_t)val;
        memcpy(payload + 1, &v, (size_t)sz);
    }
    plen = 1 + (size_t)sz;
    if (s) return replace_field(b, s, payload, plen);
    if (errno != ENOENT) return -1;
    return append_field(b, tag, payload, plen);
}

int bam_aux_update_array(bam1_t *b, const char tag[2], uint8_t type, uint32_t items, void *data)
{
    int es = aux_elem_size(type);
    if (es == 0 || type == 'A' || type == 'd') { errno = EINVAL; return -1; }
    size_t plen = 6 + (size_t)es * items;
    uint8_t *payload = (uint8_t *)malloc(plen);
    if (!payload) { errno = ENOMEM; return -1; }
    payload[0] = 'B';
    payload[1] = type;
    memcpy(payload + 2, &items, 4);
    memcpy(payload + 6, data, (size_t)es * items);
    int rc;
    uint8_t *s = bam_aux_get(b, tag);
    if (s) {
        if (s[0] != 'B') { free(payload); errno = EINVAL; return -1; }
        rc = replace_field(b, s, payload, plen);
    } else if (errno == ENOENT) {
        rc = append_field(b, tag, payload, plen);
    } else {
        rc = -1;
    }
    free(payload);
    return rc;
}
