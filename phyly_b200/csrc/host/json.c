#include "json.h"

#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const char *p;
    const char *start;
    char *err;
    size_t errlen;
    int depth;
    size_t hint;        /* length of the array parsed last: rows of a matrix have one length, so the next one is sized right at once */
} parser;

static void fail(parser *ps, const char *msg)
{
    if (ps->err && ps->errlen && !ps->err[0]) {
        int line = 1;
        for (const char *q = ps->start; q < ps->p; q++) if (*q == '\n') line++;
        snprintf(ps->err, ps->errlen, "error on json line %d: %s", line, msg);
    }
}

static void skip_ws(parser *ps)
{
    while (*ps->p == ' ' || *ps->p == '\t' || *ps->p == '\n' || *ps->p == '\r') ps->p++;
}

static void free_contents(jv *v)
{
    if (v->type == JV_STRING) free(v->u.s);
    else if (v->type == JV_ARRAY) {
        for (uint32_t i = 0; i < v->len; i++) if (v->u.items[i].type >= JV_STRING) free_contents(&v->u.items[i]);
        free(v->u.items);
    } else if (v->type == JV_OBJECT) {
        for (uint32_t i = 0; i < 2 * v->len; i++) free_contents(&v->u.items[i]);
        free(v->u.items);
    }
    v->type = JV_NULL;
}

static int parse_value(parser *ps, jv *out);

static int parse_string(parser *ps, jv *out)
{
    /* ps->p points at the opening quote */
    const char *q = ps->p + 1;
    size_t cap = 16, len = 0;
    char *s = malloc(cap);
    if (!s) { fail(ps, "out of memory"); return -1; }
    for (;;) {
        unsigned char ch = (unsigned char)*q;
        if (ch == 0) { free(s); ps->p = q; fail(ps, "premature end of input in string"); return -1; }
        if (ch == '"') { q++; break; }
        if (ch < 0x20) { free(s); ps->p = q; fail(ps, "control character in string"); return -1; }
        unsigned out_cp = ch;
        if (ch == '\\') {
            q++;
            switch (*q) {
            case '"': out_cp = '"'; break;
            case '\\': out_cp = '\\'; break;
            case '/': out_cp = '/'; break;
            case 'b': out_cp = '\b'; break;
            case 'f': out_cp = '\f'; break;
            case 'n': out_cp = '\n'; break;
            case 'r': out_cp = '\r'; break;
            case 't': out_cp = '\t'; break;
            case 'u': {
                unsigned cp = 0;
                for (int k = 1; k <= 4; k++) {
                    char h = q[k];
                    cp <<= 4;
                    if (h >= '0' && h <= '9') cp |= (unsigned)(h - '0');
                    else if (h >= 'a' && h <= 'f') cp |= (unsigned)(h - 'a' + 10);
                    else if (h >= 'A' && h <= 'F') cp |= (unsigned)(h - 'A' + 10);
                    else { free(s); ps->p = q; fail(ps, "invalid \\u escape"); return -1; }
                }
                q += 4;
                out_cp = cp;
                break;
            }
            default: free(s); ps->p = q; fail(ps, "invalid escape"); return -1;
            }
        }
        if (len + 5 > cap) { cap *= 2; char *t = realloc(s, cap); if (!t) { free(s); fail(ps, "out of memory"); return -1; } s = t; }
        if (out_cp < 0x80) s[len++] = (char)out_cp;
        else if (out_cp < 0x800) { s[len++] = (char)(0xC0 | (out_cp >> 6)); s[len++] = (char)(0x80 | (out_cp & 0x3F)); }
        else { s[len++] = (char)(0xE0 | (out_cp >> 12)); s[len++] = (char)(0x80 | ((out_cp >> 6) & 0x3F)); s[len++] = (char)(0x80 | (out_cp & 0x3F)); }
        q++;
    }
    s[len] = 0;
    out->type = JV_STRING; out->len = (uint32_t)len; out->u.s = s;
    ps->p = q;
    return 0;
}

static int parse_number(parser *ps, jv *out)
{
    const char *q = ps->p;
    int is_real = 0;
    if (*q == '-') q++;
    if (*q == '0') q++;
    else if (*q >= '1' && *q <= '9') { while (*q >= '0' && *q <= '9') q++; }
    else { fail(ps, "invalid token"); return -1; }
    if (*q == '.') {
        is_real = 1; q++;
        if (!(*q >= '0' && *q <= '9')) { ps->p = q; fail(ps, "invalid number"); return -1; }
        while (*q >= '0' && *q <= '9') q++;
    }
    if (*q == 'e' || *q == 'E') {
        is_real = 1; q++;
        if (*q == '+' || *q == '-') q++;
        if (!(*q >= '0' && *q <= '9')) { ps->p = q; fail(ps, "invalid number"); return -1; }
        while (*q >= '0' && *q <= '9') q++;
    }
    if (!is_real) {
        /* fast path for short integers (character codes, indices) */
        size_t nd = (size_t)(q - ps->p);
        if (nd <= 17) {
            const char *r = ps->p;
            int neg = 0;
            if (*r == '-') { neg = 1; r++; }
            int64_t v = 0;
            for (; r < q; r++) v = v * 10 + (*r - '0');
            out->type = JV_INT; out->u.i = neg ? -v : v; out->len = 0;
            ps->p = q;
            return 0;
        }
        errno = 0;
        long long v = strtoll(ps->p, NULL, 10);
        if (errno == ERANGE) { fail(ps, "too big integer"); return -1; }
        out->type = JV_INT; out->u.i = v; out->len = 0;
    } else {
        errno = 0;
        double d = strtod(ps->p, NULL);
        if (errno == ERANGE && (d == HUGE_VAL || d == -HUGE_VAL)) { fail(ps, "real number overflow"); return -1; }
        out->type = JV_REAL; out->u.d = d; out->len = 0;
    }
    ps->p = q;
    return 0;
}

static int parse_array(parser *ps, jv *out)
{
    size_t cap = (ps->hint >= 8 && ps->hint <= (1u << 20)) ? ps->hint : 8, len = 0;
    jv *items = malloc(cap * sizeof(jv));
    if (!items) { fail(ps, "out of memory"); return -1; }
    ps->p++;
    skip_ws(ps);
    if (*ps->p == ']') { ps->p++; goto done; }
    for (;;) {
        if (len == cap) {
            cap *= 2;
            jv *t = realloc(items, cap * sizeof(jv));
            if (!t) { fail(ps, "out of memory"); goto bad; }
            items = t;
        }
        skip_ws(ps);
        {
            /* fast path for what a large document consists of: short non-negative integers (character codes) directly
             * followed by ',' or ']'.  Anything else -- signs, fractions, exponents, leading zeros -- takes the general
             * route and its checks. */
            const char *q = ps->p;
            if (*q >= '1' && *q <= '9') {
                int64_t v = *q++ - '0';
                int nd = 1;
                while (*q >= '0' && *q <= '9' && nd < 9) { v = v * 10 + (*q++ - '0'); nd++; }
                if (*q == ',' || *q == ']') {
                    items[len].type = JV_INT; items[len].len = 0; items[len].u.i = v;
                    len++;
                    ps->p = q + 1;
                    if (*q == ',') continue;
                    break;
                }
            } else if (*q == '0' && (q[1] == ',' || q[1] == ']')) {
                items[len].type = JV_INT; items[len].len = 0; items[len].u.i = 0;
                len++;
                ps->p = q + 2;
                if (q[1] == ',') continue;
                break;
            }
        }
        if (parse_value(ps, &items[len])) goto bad;
        len++;
        skip_ws(ps);
        if (*ps->p == ',') { ps->p++; continue; }
        if (*ps->p == ']') { ps->p++; break; }
        fail(ps, "']' expected");
        goto bad;
    }
done:
    if (len < cap && len > 0) { jv *t = realloc(items, len * sizeof(jv)); if (t) items = t; }
    ps->hint = len;
    out->type = JV_ARRAY; out->len = (uint32_t)len; out->u.items = items;
    return 0;
bad:
    for (size_t i = 0; i < len; i++) free_contents(&items[i]);
    free(items);
    return -1;
}

static int parse_object(parser *ps, jv *out)
{
    size_t cap = 8, len = 0;   /* pairs */
    jv *items = malloc(2 * cap * sizeof(jv));
    if (!items) { fail(ps, "out of memory"); return -1; }
    ps->p++;
    skip_ws(ps);
    if (*ps->p == '}') { ps->p++; goto done; }
    for (;;) {
        if (len == cap) {
            cap *= 2;
            jv *t = realloc(items, 2 * cap * sizeof(jv));
            if (!t) { fail(ps, "out of memory"); goto bad; }
            items = t;
        }
        skip_ws(ps);
        if (*ps->p != '"') { fail(ps, "string or '}' expected"); goto bad; }
        if (parse_string(ps, &items[2 * len])) goto bad;
        skip_ws(ps);
        if (*ps->p != ':') { free_contents(&items[2 * len]); fail(ps, "':' expected"); goto bad; }
        ps->p++;
        skip_ws(ps);
        if (parse_value(ps, &items[2 * len + 1])) { free_contents(&items[2 * len]); goto bad; }
        /* jansson: a repeated key replaces the earlier value */
        {
            int dup = 0;
            for (size_t i = 0; i < len; i++) {
                if (!strcmp(items[2 * i].u.s, items[2 * len].u.s)) {
                    free_contents(&items[2 * i + 1]);
                    items[2 * i + 1] = items[2 * len + 1];
                    free_contents(&items[2 * len]);
                    dup = 1;
                    break;
                }
            }
            if (!dup) len++;
        }
        skip_ws(ps);
        if (*ps->p == ',') { ps->p++; continue; }
        if (*ps->p == '}') { ps->p++; break; }
        fail(ps, "'}' expected");
        goto bad;
    }
done:
    out->type = JV_OBJECT; out->len = (uint32_t)len; out->u.items = items;
    return 0;
bad:
    for (size_t i = 0; i < 2 * len; i++) free_contents(&items[i]);
    free(items);
    return -1;
}

static int parse_value(parser *ps, jv *out)
{
    out->type = JV_NULL; out->len = 0; out->u.i = 0;
    if (++ps->depth > 2048) { fail(ps, "maximum parsing depth reached"); return -1; }
    int rc;
    skip_ws(ps);
    char ch = *ps->p;
    if (ch == '{') rc = parse_object(ps, out);
    else if (ch == '[') rc = parse_array(ps, out);
    else if (ch == '"') rc = parse_string(ps, out);
    else if (ch == '-' || (ch >= '0' && ch <= '9')) rc = parse_number(ps, out);
    else if (!strncmp(ps->p, "true", 4)) { out->type = JV_TRUE; ps->p += 4; rc = 0; }
    else if (!strncmp(ps->p, "false", 5)) { out->type = JV_FALSE; ps->p += 5; rc = 0; }
    else if (!strncmp(ps->p, "null", 4)) { out->type = JV_NULL; ps->p += 4; rc = 0; }
    else if (ch == 0) { fail(ps, "unexpected end of input"); rc = -1; }
    else { fail(ps, "invalid token"); rc = -1; }
    ps->depth--;
    return rc;
}

jv *json_parse(const char *text, char *err, size_t errlen)
{
    parser ps = {text, text, err, errlen, 0, 0};
    if (err && errlen) err[0] = 0;
    jv *root = malloc(sizeof(jv));
    if (!root) { fail(&ps, "out of memory"); return NULL; }
    skip_ws(&ps);
    /* jansson's json_loads without JSON_DECODE_ANY accepts only arrays and objects */
    if (*ps.p != '{' && *ps.p != '[') { fail(&ps, "'[' or '{' expected"); free(root); return NULL; }
    if (parse_value(&ps, root)) { free(root); return NULL; }
    skip_ws(&ps);
    if (*ps.p != 0) { fail(&ps, "end of file expected"); free_contents(root); free(root); return NULL; }
    return root;
}

void json_free(jv *v)
{
    if (!v) return;
    free_contents(v);
    free(v);
}

const jv *jv_get(const jv *obj, const char *key)
{
    if (!obj || obj->type != JV_OBJECT) return NULL;
    for (uint32_t i = 0; i < obj->len; i++)
        if (!strcmp(obj->u.items[2 * i].u.s, key)) return &obj->u.items[2 * i + 1];
    return NULL;
}

int jv_unpack_strict(const jv *obj, const char *const *keys, const jv **out)
{
    if (!obj || obj->type != JV_OBJECT) {
        fprintf(stderr, "error: on line -1: Expected object, got %s\n", obj ? "another type" : "NULL");
        return -1;
    }
    size_t nk = 0;
    while (keys[nk]) nk++;
    size_t found = 0;
    for (size_t k = 0; k < nk; k++) {
        int optional = keys[k][0] == '?';
        const char *name = keys[k] + (optional ? 1 : 0);
        out[k] = jv_get(obj, name);
        if (out[k]) found++;
        else if (!optional) {
            fprintf(stderr, "error: on line -1: Object item not found: %s\n", name);
            return -1;
        }
    }
    if (found != obj->len) {
        /* JSON_STRICT: every key of the object must have been unpacked */
        for (uint32_t i = 0; i < obj->len; i++) {
            const char *have = obj->u.items[2 * i].u.s;
            int known = 0;
            for (size_t k = 0; k < nk; k++) if (!strcmp(have, keys[k] + (keys[k][0] == '?' ? 1 : 0))) known = 1;
            if (!known) {
                fprintf(stderr, "error: on line -1: %u object item(s) left unpacked: %s\n",
                        (unsigned)(obj->len - found), have);
                return -1;
            }
        }
    }
    return 0;
}

void jbuf_init(jbuf *b) { b->p = NULL; b->len = 0; b->cap = 0; }

static void jbuf_reserve(jbuf *b, size_t extra)
{
    if (b->len + extra + 1 > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 256;
        while (nc < b->len + extra + 1) nc *= 2;
        char *t = realloc(b->p, nc);
        if (!t) { fprintf(stderr, "out of memory building JSON output\n"); abort(); }
        b->p = t; b->cap = nc;
    }
}

void jbuf_puts(jbuf *b, const char *s)
{
    size_t n = strlen(s);
    jbuf_reserve(b, n);
    memcpy(b->p + b->len, s, n);
    b->len += n;
    b->p[b->len] = 0;
}

void jbuf_int(jbuf *b, long long v)
{
    char tmp[32];
    snprintf(tmp, sizeof tmp, "%lld", v);
    jbuf_puts(b, tmp);
}

void jbuf_real(jbuf *b, double d)
{
    char tmp[64];
    if (d == 0.0) d = 0.0;                 /* scrub -0.0, util.c:44-48 */
    if (d == 0.0 && signbit(d)) d = fabs(d);
    snprintf(tmp, sizeof tmp, "%.17g", d);
    /* jansson: make sure the token reads back as a real */
    if (!strchr(tmp, '.') && !strchr(tmp, 'e') && !strchr(tmp, 'n') && !strchr(tmp, 'i')) strcat(tmp, ".0");
    /* jansson strips the '+' and leading zeros of the exponent: 1e-05 -> 1e-5 */
    char *e = strchr(tmp, 'e');
    if (e) {
        char *r = e + 1, *w = e + 1;
        if (*r == '-') { w++; r++; }
        else if (*r == '+') r++;
        while (*r == '0' && r[1]) r++;
        memmove(w, r, strlen(r) + 1);
    }
    jbuf_puts(b, tmp);
}

char *jbuf_take(jbuf *b)
{
    if (!b->p) { jbuf_reserve(b, 0); b->p[0] = 0; }
    char *r = b->p;
    b->p = NULL; b->len = b->cap = 0;
    return r;
}
