/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See tamp_oracle.h for scope and the parity pin.
 *
 * A deliberately simple, whole-object restatement of the Tamp codec: exhaustive match search,
 * eager byte output, snapshot window copies.  It favours obviousness over speed so that it can
 * serve as the checker for the CUDA path.  Every function cites the reference lines it restates
 * (paths relative to /root/reference, commit 48880ad).
 */
#include "tamp_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---- static tables ------------------------------------------------------------------------- */

/* Static Huffman code for (match_len - min_pattern) 0..13 and FLUSH (14); bit counts include the
 * leading 0 "is-token" flag.  compressor.c:33-36, docs/source/specification.rst:159-186. */
static const uint8_t k_code[15] = {0x00, 0x03, 0x08, 0x0b, 0x14, 0x24, 0x26, 0x2b, 0x4b, 0x54, 0x94, 0x95, 0xaa, 0x27, 0xab};
static const uint8_t k_bits[15] = {2, 3, 5, 5, 6, 7, 7, 7, 8, 8, 9, 9, 9, 7, 9};

enum { SYM_RLE = 12, SYM_EXT = 13, SYM_FLUSH = 14, RLE_MAX = 241, RLE_WINDOW_MAX = 8, EXT_EXTRA_MAX = 120 };

/* ---- common.c ------------------------------------------------------------------------------ */

/* common.c:37-52 (tables :18-25, xorshift :28-35). */
void oracle_initialize_dictionary(uint8_t *buf, size_t size, int literal) {
    static const uint8_t tab8[16] = {' ', 0, '0', 'e', 'i', '>', 't', 'o', '<', 'a', 'n', 's', '\n', 'r', '/', '.'};
    static const char english[17] = " etaoinshrdlcumw";
    uint8_t tab[16];
    for (int i = 0; i < 16; i++) {
        if (literal <= 5)
            tab[i] = (uint8_t)(english[i] & 0x1F);
        else if (literal <= 6)
            tab[i] = (uint8_t)(english[i] & 0x3F);
        else
            tab[i] = tab8[i];
    }
    uint32_t s = 3758097560u;
    for (size_t base = 0; base < size; base += 8) {
        s ^= s << 13;
        s ^= s >> 17;
        s ^= s << 5;
        uint32_t r = s;
        for (size_t j = 0; j < 8 && base + j < size; j++, r >>= 4) buf[base + j] = tab[r & 15u];
    }
}

/* common.c:54-56 */
int oracle_min_pattern_size(int window, int literal) { return 2 + (window > 10 + 2 * (literal - 5)); }

/* ---- encoder ------------------------------------------------------------------------------- */

struct OracleEnc {
    OracleConf cf;
    int W, min_pat;
    uint8_t *win;
    int wpos;
    uint8_t q[16]; /* the 16-byte input ring, kept linear: q[0] is the oldest byte */
    int qn;
    uint64_t acc; /* pending (not yet byte-complete) output bits, MSb first */
    int nacc;
    uint8_t *out;
    size_t on, ocap;
    int rle, ext_n, ext_pos, last_flush;
    int cache_idx, cache_len; /* lazy matching */
};

static void out_byte(OracleEnc *e, uint8_t b) {
    if (e->on == e->ocap) {
        e->ocap = e->ocap ? e->ocap * 2 : 256;
        e->out = (uint8_t *)realloc(e->out, e->ocap);
    }
    e->out[e->on++] = b;
}

/* write_to_bit_buffer + partial_flush, compressor.c:49-75 (output is unbounded here). */
static void put(OracleEnc *e, uint32_t bits, int n) {
    e->acc = (e->acc << n) | (uint64_t)bits;
    e->nacc += n;
    while (e->nacc >= 8) {
        out_byte(e, (uint8_t)(e->acc >> (e->nacc - 8)));
        e->nacc -= 8;
    }
    e->acc &= (1ull << e->nacc) - 1ull;
}

/* write_extended_huffman, compressor.c:257-263 */
static void put_exthuff(OracleEnc *e, int v, int t) {
    int i = v >> t;
    put(e, ((uint32_t)k_code[i] << t) | (uint32_t)(v & ((1 << t) - 1)), k_bits[i] - 1 + t);
}

static void q_consume(OracleEnc *e, int n) {
    memmove(e->q, e->q + n, (size_t)(e->qn - n));
    e->qn -= n;
}

static uint8_t last_window_byte(const OracleEnc *e) { return e->win[(e->wpos - 1) & (e->W - 1)]; } /* :270-273 */

/* find_best_match: compressor_find_match_desktop.c:82-167 / compressor.c:113-172 /
 * fuzz/esp32_host/differential.cpp:50-67.  Whole physical buffer [0, W-2], never past W-1,
 * lowest index wins ties, needs >= 2 bytes. */
int oracle_find_best_match(const uint8_t *window, int W, const uint8_t *pat, int max_len, int *index) {
    int best = 0;
    if (max_len < 2) return 0;
    for (int i = 0; i + 1 < W; i++) {
        int lim = max_len < W - i ? max_len : W - i;
        int l = 0;
        while (l < lim && window[i + l] == pat[l]) l++;
        if (l >= 2 && l > best) {
            best = l;
            *index = i;
            if (best == max_len) break;
        }
    }
    return best;
}

static int enc_best_match(const OracleEnc *e, const uint8_t *pat, int avail, int *index) {
    /* MAX_PATTERN_SIZE, compressor.c:12-19; the 16-byte ring bounds `avail`. */
    int cap = e->cf.extended ? e->min_pat + 11 + EXT_EXTRA_MAX : e->min_pat + 13;
    if (avail < e->min_pat) return 0; /* compressor_find_match_desktop.c:85 */
    return oracle_find_best_match(e->win, e->W, pat, avail < cap ? avail : cap, index);
}

/* find_extended_match, compressor.c:297-333 */
static int enc_ext_search(const OracleEnc *e, int cur_pos, int cur_n, int *new_pos) {
    int best = 0;
    int maxp = cur_n + e->qn;
    if (maxp > e->min_pat + 11 + EXT_EXTRA_MAX) maxp = e->min_pat + 11 + EXT_EXTRA_MAX;
    for (int c = cur_pos; c + cur_n + 1 <= e->W; c++) {
        if (memcmp(e->win + c, e->win + cur_pos, (size_t)cur_n) != 0) continue;
        int lim = maxp < e->W - c ? maxp : e->W - c;
        int l = cur_n;
        while (l < lim && e->win[c + l] == e->q[l - cur_n]) l++;
        if (l > cur_n && l > best) {
            best = l;
            *new_pos = c;
            if (best == maxp) break;
        }
    }
    return best;
}

/* Window copy with the reference's semantics (common.c:58-86): forward/reverse choice there makes
 * the copy equal to "read all n source bytes first, then write them" — done literally here. */
static void win_copy(OracleEnc *e, int src, int n) {
    uint8_t tmp[256];
    memcpy(tmp, e->win + src, (size_t)n);
    for (int i = 0; i < n; i++) {
        e->win[e->wpos] = tmp[i];
        e->wpos = (e->wpos + 1) & (e->W - 1);
    }
}

/* write_rle_token, compressor.c:342-359 */
static void emit_rle(OracleEnc *e, int count) {
    uint8_t sym = last_window_byte(e);
    put(e, k_code[SYM_RLE], k_bits[SYM_RLE]);
    put_exthuff(e, count - 2, 4);
    int room = e->W - e->wpos;
    int nw = count < RLE_WINDOW_MAX ? count : RLE_WINDOW_MAX;
    if (nw > room) nw = room;
    for (int i = 0; i < nw; i++) {
        e->win[e->wpos] = sym;
        e->wpos = (e->wpos + 1) & (e->W - 1);
    }
}

/* write_extended_match_token, compressor.c:377-415 */
static void emit_ext(OracleEnc *e) {
    put(e, k_code[SYM_EXT], k_bits[SYM_EXT]);
    put_exthuff(e, e->ext_n - e->min_pat - 12, 3);
    put(e, (uint32_t)e->ext_pos, e->cf.window);
    int room = e->W - e->wpos;
    win_copy(e, e->ext_pos, e->ext_n < room ? e->ext_n : room);
    e->ext_n = 0;
}

static void emit_literal_from_window(OracleEnc *e) { /* compressor.c:512-523, :748-756 */
    uint8_t b = last_window_byte(e);
    put(e, (1u << e->cf.literal) | b, e->cf.literal + 1);
    e->win[e->wpos] = b;
    e->wpos = (e->wpos + 1) & (e->W - 1);
}

#define POLL_CONTINUE 127

/* poll_extended_handling, compressor.c:437-525 */
static int ext_handling(OracleEnc *e, int *m_idx, int *m_len) {
    if (e->ext_n) {
        int cap = e->min_pat + 11 + EXT_EXTRA_MAX;
        while (e->qn > 0) {
            if (e->ext_pos + e->ext_n >= e->W || e->ext_n >= cap) {
                emit_ext(e);
                return ORC_OK;
            }
            int npos = 0;
            int nlen = enc_ext_search(e, e->ext_pos, e->ext_n, &npos);
            if (nlen > e->ext_n) {
                q_consume(e, nlen - e->ext_n);
                e->ext_pos = npos;
                e->ext_n = nlen;
                continue;
            }
            emit_ext(e);
            return ORC_OK;
        }
        return ORC_OK;
    }
    uint8_t last = last_window_byte(e);
    int avail = 0;
    while (avail < e->qn && e->rle + avail < RLE_MAX && e->q[avail] == last) avail++;
    int total = e->rle + avail;
    int ended = (avail < e->qn) || (total >= RLE_MAX);
    if (!ended && total > 0) {
        e->rle = total;
        q_consume(e, avail);
        return ORC_OK;
    }
    if (total >= 2) {
        if (total == avail && total <= 6) {
            int idx = 0;
            int len = enc_best_match(e, e->q, e->qn, &idx);
            if (len > total) {
                e->rle = 0;
                *m_idx = idx;
                *m_len = len;
                return POLL_CONTINUE;
            }
        }
        q_consume(e, avail);
        emit_rle(e, total);
        e->rle = 0;
        return ORC_OK;
    }
    if (e->rle == 1) {
        emit_literal_from_window(e);
        e->rle = 0;
        return ORC_OK;
    }
    return POLL_CONTINUE;
}

/* tamp_compressor_poll, compressor.c:532-660 */
static int enc_poll(OracleEnc *e) {
    if (e->qn == 0) return ORC_OK;
    e->last_flush = 0;
    int idx = 0, len = 0;
    if (e->cf.extended) {
        int r = ext_handling(e, &idx, &len);
        if (r != POLL_CONTINUE) {
            e->cache_idx = -1;
            return r;
        }
    }
    if (e->cf.lazy_matching) { /* compressor.c:576-616 */
        if (e->cache_idx >= 0) {
            idx = e->cache_idx;
            len = e->cache_len;
            e->cache_idx = -1;
        } else if (len == 0) {
            len = enc_best_match(e, e->q, e->qn, &idx);
        }
        if (len >= e->min_pat && len <= 8 && e->qn > len + 2) {
            int nidx = 0;
            int nlen = enc_best_match(e, e->q + 1, e->qn - 1, &nidx);
            int no_overlap = e->wpos < nidx || e->wpos >= nidx + nlen; /* :185-188 */
            if (nlen > len && no_overlap) {
                e->cache_idx = nidx;
                e->cache_len = nlen;
                len = 0;
            } else {
                e->cache_idx = -1;
            }
        } else {
            e->cache_idx = -1;
        }
    } else if (len == 0) {
        len = enc_best_match(e, e->q, e->qn, &idx);
    }

    int n;
    if (len < e->min_pat) {
        uint8_t c = e->q[0];
        if (c >> e->cf.literal) return ORC_EXCESS_BITS;
        put(e, (1u << e->cf.literal) | c, e->cf.literal + 1);
        n = 1;
    } else if (e->cf.extended && len > e->min_pat + 11) {
        e->ext_n = len;
        e->ext_pos = idx;
        q_consume(e, len);
        return ORC_OK;
    } else {
        int h = len - e->min_pat;
        put(e, ((uint32_t)k_code[h] << e->cf.window) | (uint32_t)idx, k_bits[h] + e->cf.window);
        n = len;
    }
    for (int i = 0; i < n; i++) {
        e->win[e->wpos] = e->q[i];
        e->wpos = (e->wpos + 1) & (e->W - 1);
    }
    q_consume(e, n);
    return ORC_OK;
}

/* tamp_compressor_init, compressor.c:191-245 */
OracleEnc *oracle_enc_new(const OracleConf *conf, const uint8_t *dictionary, int *status) {
    OracleConf dflt = {10, 8, 0, 1, 0, 0};
    if (!conf) conf = &dflt;
    int st = ORC_OK;
    if (conf->window < 8 || conf->window > 15 || conf->literal < 5 || conf->literal > 8) st = ORC_INVALID_CONF;
    if (conf->use_custom_dictionary && !dictionary) st = ORC_INVALID_CONF;
    if (status) *status = st;
    if (st != ORC_OK) return NULL;
    OracleEnc *e = (OracleEnc *)calloc(1, sizeof *e);
    e->cf = *conf;
    e->W = 1 << conf->window;
    e->min_pat = oracle_min_pattern_size(conf->window, conf->literal);
    e->win = (uint8_t *)malloc((size_t)e->W);
    e->cache_idx = -1;
    if (conf->use_custom_dictionary)
        memcpy(e->win, dictionary, (size_t)e->W);
    else
        oracle_initialize_dictionary(e->win, (size_t)e->W, conf->extended ? conf->literal : 8);
    uint32_t header = ((uint32_t)(conf->window - 8) << 5) | ((uint32_t)(conf->literal - 5) << 3) |
                      ((uint32_t)!!conf->use_custom_dictionary << 2) | ((uint32_t)!!conf->extended << 1) |
                      (uint32_t)!!conf->dictionary_reset;
    put(e, header, 8);
    if (conf->dictionary_reset) put(e, 0, 8);
    return e;
}

/* tamp_compressor_compress_cb, compressor.c:681-722: sink, poll only when the ring is full. */
int oracle_enc_write(OracleEnc *e, const uint8_t *data, size_t n) {
    while (n > 0) {
        while (n > 0 && e->qn < 16) {
            e->q[e->qn++] = *data++;
            n--;
        }
        if (e->qn == 16) {
            int r = enc_poll(e);
            if (r != ORC_OK) return r;
        }
    }
    return ORC_OK;
}

/* tamp_compressor_flush, compressor.c:728-810 */
int oracle_enc_flush(OracleEnc *e, int write_token) {
    for (;;) {
        if (e->qn) {
            int r = enc_poll(e);
            if (r != ORC_OK) return r;
        } else if (e->cf.extended && e->rle >= 1) {
            if (e->rle == 1)
                emit_literal_from_window(e);
            else
                emit_rle(e, e->rle);
            e->rle = 0;
        } else if (e->cf.extended && e->ext_n) {
            emit_ext(e);
        } else {
            break;
        }
    }
    if (write_token && !e->last_flush && (e->nacc || e->cf.dictionary_reset)) {
        put(e, k_code[SYM_FLUSH], k_bits[SYM_FLUSH]);
        e->last_flush = 1;
    }
    if (e->nacc) {
        out_byte(e, (uint8_t)(e->acc << (8 - e->nacc)));
        e->nacc = 0;
        e->acc = 0;
    }
    return ORC_OK;
}

size_t oracle_enc_size(const OracleEnc *e) { return e->on; }
const uint8_t *oracle_enc_data(const OracleEnc *e) { return e->out; }
const uint8_t *oracle_enc_window(const OracleEnc *e) { return e->win; }
void oracle_enc_free(OracleEnc *e) {
    if (!e) return;
    free(e->win);
    free(e->out);
    free(e);
}

long oracle_compress(const OracleConf *conf, const uint8_t *dictionary, const uint8_t *in, size_t n, uint8_t *out,
                     size_t cap, int write_token) {
    int st;
    OracleEnc *e = oracle_enc_new(conf, dictionary, &st);
    if (!e) return st;
    st = oracle_enc_write(e, in, n);
    if (st == ORC_OK) st = oracle_enc_flush(e, write_token);
    long ret;
    if (st != ORC_OK)
        ret = st;
    else if (e->on > cap)
        ret = -100;
    else {
        memcpy(out, e->out, e->on);
        ret = (long)e->on;
    }
    oracle_enc_free(e);
    return ret;
}

/* ---- decoder ------------------------------------------------------------------------------- */

typedef struct {
    const uint8_t *p;
    size_t nbits; /* total bits in the frame */
    size_t pos;   /* next bit */
} BitIn;

static size_t bits_left(const BitIn *b) { return b->nbits - b->pos; }

/* peek n (<=24) bits, zero-filled past the end (decompressor.c:83-87 reads the LUT index that way) */
static uint32_t peek(const BitIn *b, int n) {
    uint32_t v = 0;
    for (int i = 0; i < n; i++) {
        size_t k = b->pos + (size_t)i;
        uint32_t bit = 0;
        if (k < b->nbits) bit = (b->p[k >> 3] >> (7 - (k & 7))) & 1u;
        v = (v << 1) | bit;
    }
    return v;
}

/* decode_huffman, decompressor.c:71-104.  Symbol decode for the code WITHOUT the is-literal flag.
 * Returns symbol (0..14) and advances, or -1 if the bits cannot complete the code + trailing. */
static int get_huffman(BitIn *b, int trailing, int *value) {
    if (bits_left(b) < (size_t)(1 + trailing)) return -1;
    int sym, used;
    if (peek(b, 1) == 0) {
        sym = 0;
        used = 1;
    } else {
        /* decode by walking the code table: codes are prefix-free, lengths (without flag) 1..8 */
        sym = -1;
        used = 0;
        for (int s = 1; s < 15 && sym < 0; s++) {
            int L = k_bits[s] - 1;
            uint32_t code = k_code[s] & ((1u << L) - 1u); /* k_code has the 0 flag as its top bit */
            if (peek(b, L) == code) {
                sym = s;
                used = L;
            }
        }
        if (sym < 0) return -1; /* unreachable: the code is complete */
        if (bits_left(b) < (size_t)(used + trailing)) return -1;
    }
    b->pos += (size_t)used;
    int tr = trailing ? (int)peek(b, trailing) : 0;
    b->pos += (size_t)trailing;
    *value = (sym << trailing) + tr;
    return sym;
}

/* tamp_decompressor_read_header :276-297, populate_from_conf :304-329, decompress_cb :371-578,
 * decode_rle :114-174, decode_extended_match :187-273. */
long oracle_decompress(const uint8_t *dictionary, int window_bits_max, const uint8_t *in, size_t n, uint8_t *out,
                       size_t cap, int *status) {
    int st_dummy;
    if (!status) status = &st_dummy;
    if (window_bits_max < 8 || window_bits_max > 15) {
        *status = ORC_INVALID_CONF;
        return 0;
    }
    if (n == 0) {
        *status = ORC_INPUT_EXHAUSTED;
        return 0;
    }
    size_t hdr = 1 + (in[0] & 1u);
    if (n < hdr) {
        *status = ORC_INPUT_EXHAUSTED;
        return 0;
    }
    if (hdr == 2 && in[1]) {
        *status = ORC_INVALID_CONF;
        return 0;
    }
    int window = ((in[0] >> 5) & 7) + 8, literal = ((in[0] >> 3) & 3) + 5;
    int custom = (in[0] >> 2) & 1, extended = (in[0] >> 1) & 1, dict_reset = in[0] & 1;
    if (window > window_bits_max || (custom && !dictionary)) {
        *status = ORC_INVALID_CONF;
        return 0;
    }
    int W = 1 << window, mask = W - 1, min_pat = oracle_min_pattern_size(window, literal);
    uint8_t *win = (uint8_t *)malloc((size_t)W);
    if (custom)
        memcpy(win, dictionary, (size_t)W);
    else
        oracle_initialize_dictionary(win, (size_t)W, extended ? literal : 8);
    int wpos = 0, last_flush = 0;
    BitIn b = {in + hdr, (n - hdr) * 8, 0};
    size_t on = 0;
    int st = ORC_INPUT_EXHAUSTED;

    for (;;) {
        if (bits_left(&b) == 0) break;           /* loop condition :433 + :463 */
        if (on == cap) {                          /* :434 */
            st = ORC_OUTPUT_FULL;
            break;
        }
        if (peek(&b, 1)) { /* literal :466-482 */
            if (bits_left(&b) < (size_t)(1 + literal)) break;
            b.pos += 1;
            uint8_t c = (uint8_t)peek(&b, literal);
            b.pos += (size_t)literal;
            out[on++] = c;
            win[wpos] = c;
            wpos = (wpos + 1) & mask;
            last_flush = 0;
            continue;
        }
        BitIn t = b; /* decode on a copy; commit only when the whole item is there (:486-489) */
        t.pos += 1;
        int v;
        int sym = get_huffman(&t, 0, &v);
        if (sym < 0) break;
        if (sym == SYM_FLUSH) { /* :501-514 */
            t.pos = (t.pos + 7) & ~(size_t)7;
            if (t.pos > t.nbits) t.pos = t.nbits;
            b = t;
            if (dict_reset && last_flush) {
                wpos = 0;
                oracle_initialize_dictionary(win, (size_t)W, extended ? literal : 8);
            }
            last_flush = 1;
            continue;
        }
        last_flush = 0;
        int len, off, nwin;
        int is_rle = 0;
        if (extended && sym == SYM_RLE) { /* decode_rle */
            b = t;                        /* symbol is consumed even if the count is not there (:522-525) */
            int raw;
            BitIn u = b;
            if (get_huffman(&u, 4, &raw) < 0) break;
            b = u;
            len = raw + 2;
            is_rle = 1;
            off = 0;
            nwin = len < RLE_WINDOW_MAX ? len : RLE_WINDOW_MAX;
            if (nwin > W - wpos) nwin = W - wpos;
        } else if (extended && sym == SYM_EXT) { /* decode_extended_match */
            b = t;
            int raw;
            BitIn u = b;
            if (get_huffman(&u, 3, &raw) < 0) break;
            b = u; /* size is committed before the offset is known to be there (:216-223) */
            len = raw + min_pat + 12;
            if (bits_left(&b) < (size_t)window) break;
            off = (int)peek(&b, window);
            b.pos += (size_t)window;
            if (off >= W || off + len > W) { /* :231-236 */
                st = ORC_OOB;
                break;
            }
            nwin = len < W - wpos ? len : W - wpos;
        } else { /* plain token :529-572 */
            if (bits_left(&t) < (size_t)window) break;
            len = sym + min_pat;
            off = (int)peek(&t, window);
            t.pos += (size_t)window;
            if (off >= W || off + len > W) { /* :540-544 */
                st = ORC_OOB;
                break;
            }
            b = t;
            nwin = len;
        }
        /* output: snapshot of the pre-update window */
        uint8_t tmp[256];
        if (is_rle)
            memset(tmp, win[(wpos - 1) & mask], (size_t)len);
        else
            memcpy(tmp, win + off, (size_t)len);
        size_t room = cap - on;
        size_t nout = (size_t)len < room ? (size_t)len : room;
        memcpy(out + on, tmp, nout);
        on += nout;
        if (nout < (size_t)len) { /* :555-557, :143-149, :243-250: partial token, window untouched */
            st = ORC_OUTPUT_FULL;
            break;
        }
        for (int i = 0; i < nwin; i++) {
            win[wpos] = tmp[i];
            wpos = (wpos + 1) & mask;
        }
    }
    free(win);
    *status = st;
    return (long)on;
}
