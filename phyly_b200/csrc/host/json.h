/*
 * Minimal strict JSON reader / writer for the arbplf front end.
 *
 * The reference uses jansson (json_loads / json_unpack_ex with JSON_STRICT /
 * json_dumps, runjson.c:10-66); no jansson header exists in this image, so the
 * host layer carries its own.  What matters for parity:
 *   - integer and real tokens are distinguished (json_is_integer,
 *     parsemodel.c:592,771; parsereduction.c:47);
 *   - duplicate / unknown keys can be rejected by the caller;
 *   - reals are printed with 17 significant digits and always carry a '.0' or
 *     an exponent, like jansson's json_dumps.
 */
#ifndef PLF_JSON_H
#define PLF_JSON_H

#include <stddef.h>
#include <stdint.h>

enum { JV_NULL = 0, JV_FALSE, JV_TRUE, JV_INT, JV_REAL, JV_STRING, JV_ARRAY, JV_OBJECT };

/* storage flags of an array (json.c only): a large matrix of short integers -- 'character_data' -- is read by several
 * threads into a few big blocks instead of one allocation per row */
enum { JV_F_BORROWED = 1,         /* items live in a block owned by the enclosing array */
       JV_F_BLOCKS = 2,           /* items[len].len blocks follow the rows: items[len + 1 + t].u.items */
       JV_F_HASHED = 4,           /* with JV_F_BLOCKS: three more hidden items hold (h1, h2, length) of the matrix's text */
       JV_F_COMPACT = 8 };        /* no items at all: u.items is a jv_codes (see below) */

/*
 * A process that calls arbplf_* again and again on one alignment (an optimiser varying the edge rates) sends the same
 * 'character_data' text every time.  The reader keeps the codes of the last large 'character_data' matrix, keyed by a
 * 128-bit hash of its text: when the same bytes come again the matrix is not read at all, its node says JV_F_COMPACT
 * and carries the shared codes.  Only the consumer of 'character_data' (model.c) ever sees such a node; it also
 * publishes the codes after a normal read (json_codes_publish).  ARBPLF_NO_DATA_CACHE=1 turns all of it off.
 */
typedef struct jv_codes {
    void *codes;                  /* [rows][row_len], uint8 or int32 */
    int code_bytes;
    int64_t rows;
    int row_len, max_code;
    uint64_t h1, h2;              /* of the text the codes were read from */
    size_t text_len;
    int refs;                     /* holders: the cache, parsed documents, models, the driver's record of the upload */
} jv_codes;

typedef struct jv {
    uint8_t type;
    uint8_t flags;
    uint32_t len;                 /* array: items; object: pairs; string: bytes */
    union {
        int64_t i;
        double d;
        char *s;
        struct jv *items;         /* array: len values; object: 2*len (key, value) */
    } u;
} jv;

/* Parse a whole document.  On failure returns NULL and writes a message. */
jv *json_parse(const char *text, char *err, size_t errlen);
void json_free(jv *v);

static inline int jv_is_number(const jv *v) { return v && (v->type == JV_INT || v->type == JV_REAL); }
static inline int jv_is_int(const jv *v) { return v && v->type == JV_INT; }
static inline int jv_is_string(const jv *v) { return v && v->type == JV_STRING; }
static inline int jv_is_array(const jv *v) { return v && v->type == JV_ARRAY; }
static inline int jv_is_object(const jv *v) { return v && v->type == JV_OBJECT; }
static inline int jv_is_null(const jv *v) { return v && v->type == JV_NULL; }
static inline double jv_number(const jv *v) { return v->type == JV_INT ? (double)v->u.i : v->u.d; }
const jv *jv_get(const jv *obj, const char *key);

/* the shared codes of a JV_F_COMPACT node, or NULL */
static inline jv_codes *jv_compact_codes(const jv *v) { return (v && v->type == JV_ARRAY && (v->flags & JV_F_COMPACT)) ? (jv_codes *)v->u.items : NULL; }
/* 1 and the hash of its text if the array was read as a hashed matrix ('character_data', threaded reader) */
int jv_matrix_text_hash(const jv *v, uint64_t *h1, uint64_t *h2, size_t *text_len);
/* hands `codes` (malloc'd) to the cache under that hash; returns the shared record with one reference for the caller */
jv_codes *json_codes_publish(uint64_t h1, uint64_t h2, size_t text_len, void *codes, int code_bytes, int64_t rows, int row_len, int max_code);
void json_codes_retain(jv_codes *d);
void json_codes_release(jv_codes *d);

/*
 * json_unpack_ex(..., JSON_STRICT, "{s:o, s?o ...}") equivalent: keys is a
 * NULL-terminated list; a leading '?' marks an optional key.  out[i] receives
 * the value or NULL.  Returns 0 or -1 (message on stderr like the reference:
 * "error: on line %d: %s").
 */
int jv_unpack_strict(const jv *obj, const char *const *keys, const jv **out);

/* growing output buffer */
typedef struct { char *p; size_t len, cap; } jbuf;
void jbuf_init(jbuf *b);
void jbuf_puts(jbuf *b, const char *s);
void jbuf_append(jbuf *b, const char *s, size_t n);
void jbuf_int(jbuf *b, long long v);
void jbuf_real(jbuf *b, double d);   /* jansson formatting, -0.0 scrubbed (util.c:44-48) */
char *jbuf_take(jbuf *b);            /* malloc'd, NUL terminated; caller frees */

#endif
