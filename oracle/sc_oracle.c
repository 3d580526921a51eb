/*
 * sc_oracle.c -- plain-int CPU restatement of the reference SC polar decoder.
 * TEST INFRASTRUCTURE ONLY (see sc_oracle.h).  Citations relative to /root/reference.
 *
 * Every value is kept as the raw width-w bit pattern the reference's sc_bigint<w> /
 * sc_biguint<w> would hold, so wrap-around on assignment (SystemC truncation) is modelled
 * too, not only the in-range behaviour.
 */
#include "sc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ bit helpers */
static inline uint32_t mask_w(int w) { return (w >= 32) ? 0xFFFFFFFFu : ((1u << w) - 1u); }
static inline int32_t sx(uint32_t p, int w) { /* two's complement value of a w-bit pattern */
    uint32_t m = 1u << (w - 1);
    p &= mask_w(w);
    return (int32_t)((p ^ m) - m);
}
static inline uint32_t tr(int64_t v, int w) { return (uint32_t)((uint64_t)v & mask_w(w)); }

int sco_sign(int w, uint32_t a) { return (int)((a >> (w - 1)) & 1u); } /* scalar.h:23-28,101-106 */

/* ------------------------------------------------------------------ CA2 (scalar.h:9-76) */
static inline uint32_t c2_abs(int w, uint32_t a) { /* qabs: (value[Q-1]==1) ? -value : +value, truncated */
    int32_t v = sx(a, w);
    return (v < 0) ? tr(-(int64_t)v, w) : tr(v, w);
}
static inline uint32_t c2_min(int w, uint32_t a, uint32_t b) { /* qmin: signed compare */
    return (sx(a, w) < sx(b, w)) ? a : b;
}
static inline uint32_t c2_sat(int w, int64_t v) { /* qsat<w>( (w+1)-bit value ) scalar.h:15-21 */
    int64_t maxv = ((int64_t)1 << (w - 1)) - 1;
    if (v > maxv) return tr(maxv, w);
    if (v < -maxv) return tr(-maxv, w);
    return tr(v, w);
}
static inline uint32_t c2_f(int w, uint32_t a, uint32_t b) { /* F_function_C2 functions.h:48-61 */
    uint32_t absa = c2_abs(w, a), absb = c2_abs(w, b);
    uint32_t mn = c2_min(w, absa, absb);
    int sig = sco_sign(w, a) ^ sco_sign(w, b);
    return sig ? tr(-(int64_t)sx(mn, w), w) : mn; /* qsign(a,b) scalar.h:30-36 */
}
static inline int64_t c2_addsub(int w, uint32_t a, uint32_t b, int s) {
    /* VECTOR_MUX(sub, add, sa): sub = lb - la where sa = 1, add = lb + la otherwise
     * (functions.h:66-70, vector.h:237-250) */
    int64_t va = sx(a, w), vb = sx(b, w);
    return s ? (vb - va) : (vb + va);
}

/* ------------------------------------------------------------------ SIGMAG (scalar.h:88-225) */
static inline uint32_t sm_mag(int w, uint32_t a) { return a & mask_w(w - 1); } /* qabs_sm */
static inline uint32_t sm_f(int w, uint32_t a, uint32_t b) { /* F_function_SM functions.h:124-145 */
    uint32_t ma = sm_mag(w, a), mb = sm_mag(w, b);
    uint32_t mn = (ma < mb) ? ma : mb; /* qmin_sm on (w-1)-bit magnitudes */
    uint32_t sig = (uint32_t)(sco_sign(w, a) ^ sco_sign(w, b));
    return (sig << (w - 1)) | mn; /* VECTOR_CONCAT_SM */
}
/* qfull_add_sub_sm<w> scalar.h:196-225: returns (w+1)-bit pattern (sign, w-bit sum). */
static inline uint32_t sm_addsub(int w, uint32_t a, uint32_t b, int s) {
    uint32_t mw = mask_w(w);
    uint32_t siga = (uint32_t)sco_sign(w, a) ^ (uint32_t)(s & 1);
    uint32_t sigb = (uint32_t)sco_sign(w, b);
    uint32_t xsig = siga ^ sigb;
    uint32_t absla = sm_mag(w, a), abslb = sm_mag(w, b); /* held in w-bit registers */
    uint32_t invla = (~absla) & mw, invlb = (~abslb) & mw;
    uint32_t is_min = (absla < abslb) ? 1u : 0u;
    uint32_t sel_a = xsig & is_min;
    uint32_t sel_b = xsig & (is_min ^ 1u);
    uint32_t absA = sel_a ? invla : absla;
    uint32_t absB = sel_b ? invlb : abslb;
    uint32_t somme = (absA + absB + xsig) & mw;
    uint32_t sig = is_min ? sigb : siga;
    return (sig << w) | somme;
}
static inline uint32_t sm_sat(int wout, uint32_t a) { /* qsat_sm<wout>( (wout+1)-bit ) scalar.h:94-99 */
    uint32_t maxv = mask_w(wout - 1); /* ((sc_uint<1>)0, (sc_uint<wout-1>)ones) */
    return (a > maxv) ? maxv : (a & mask_w(wout));
}
static inline uint32_t sm_g(int w, uint32_t a, uint32_t b, int s) { /* G_function_SM functions.h:186-194 */
    uint32_t somme = sm_addsub(w, a, b, s);           /* w+1 bits */
    uint32_t absv = somme & mask_w(w);                /* VECTOR_ABS_SM<P,w+1> : w bits */
    uint32_t sign = (somme >> w) & 1u;                /* VECTOR_SIGN_SM<P,w+1> */
    uint32_t sat = sm_sat(w - 1, absv);               /* VECTOR_SAT_SM<P,w-1> -> w-1 bits */
    return (sign << (w - 1)) | sat;                   /* VECTOR_CONCAT_SM<P,w> */
}
static inline uint32_t sm_conv(int w, uint32_t a) { /* qconv_format<w> scalar.h:229-239 */
    uint32_t mw = mask_w(w);
    uint32_t absv = a & mask_w(w - 1);
    uint32_t inv = (~absv) & mw;
    uint32_t add = (inv + 1u) & mw;
    return ((a >> (w - 1)) & 1u) ? add : (a & mw);
}

/* ------------------------------------------------------------------ dispatch (functions.h:287-347) */
uint32_t sco_input(int format, int q, int llr) {
    uint32_t p = tr(llr, q); /* sc_fifo<LLR> write of a short: truncation to Q bits */
    return (format == SCO_SIGMAG) ? sm_conv(q, p) : p; /* Adapt_format library.h:18-28 */
}
uint32_t sco_f(int format, int w, uint32_t a, uint32_t b) {
    return (format == SCO_SIGMAG) ? sm_f(w, a, b) : c2_f(w, a, b);
}
uint32_t sco_g(int format, int w, uint32_t a, uint32_t b, int s) {
    return (format == SCO_SIGMAG) ? sm_g(w, a, b, s) : c2_sat(w, c2_addsub(w, a, b, s));
}
uint32_t sco_g_ext(int format, int w, uint32_t a, uint32_t b, int s) {
    return (format == SCO_SIGMAG) ? sm_addsub(w, a, b, s) : tr(c2_addsub(w, a, b, s), w + 1);
}
int sco_f_simp(int format, int w, uint32_t a, uint32_t b, int fb) { /* functions.h:90-101,241-255 */
    (void)format;
    return (sco_sign(w, a) ^ sco_sign(w, b)) & fb;
}
int sco_g_simp(int format, int w, uint32_t a, uint32_t b, int s, int fb) { /* functions.h:103-118,257-281 */
    if (format == SCO_SIGMAG) {
        int sigla = sco_sign(w, a) ^ (s & 1);
        int siglb = sco_sign(w, b);
        int siga = sigla & fb, sigb = siglb & fb;
        int is_min = sm_mag(w, a) < sm_mag(w, b);
        return is_min ? sigb : siga;
    }
    uint32_t g = tr(c2_addsub(w, a, b, s), w + 1);
    return sco_sign(w + 1, g) & fb;
}
int sco_value(int format, int w, uint32_t a) {
    if (format == SCO_SIGMAG) {
        int m = (int)sm_mag(w, a);
        return sco_sign(w, a) ? -m : m;
    }
    return sx(a, w);
}

/* ------------------------------------------------------------------ leaf (functions.h:354-866) */
typedef struct {
    int format, extended, par, q, n;
    const uint8_t* flags;
    sco_stats* st;
    const int32_t* fsum; /* prefix sums of flags (stats only) */
} ctx_t;

static void leaf_rec(const ctx_t* c, int p, int w, const uint32_t* llr, const uint8_t* fb, uint8_t* bits,
                     uint32_t* scratch) {
    if (p == 1) { /* Spec_P1 :355-364 */
        bits[0] = (uint8_t)(sco_sign(w, llr[0]) & fb[0]);
        return;
    }
    if (p == 2) { /* Spec_P2 :367-384 */
        int sa1 = sco_f_simp(c->format, w, llr[0], llr[1], fb[0]);
        int sb1 = sco_g_simp(c->format, w, llr[0], llr[1], sa1, fb[1]);
        bits[0] = (uint8_t)(sa1 ^ sb1);
        bits[1] = (uint8_t)sb1;
        return;
    }
    int h = p / 2;
    uint32_t* child = scratch;      /* h entries */
    uint32_t* next = scratch + h;   /* scratch for deeper levels */
    for (int i = 0; i < h; i++) child[i] = sco_f(c->format, w, llr[i], llr[h + i]); /* Function_F<h,w> */
    leaf_rec(c, h, w, child, fb, bits, next);
    if (c->extended) { /* Spec_P*_ext: Function_G_ext, right child one bit wider */
        for (int i = 0; i < h; i++) child[i] = sco_g_ext(c->format, w, llr[i], llr[h + i], bits[i]);
        leaf_rec(c, h, w + 1, child, fb + h, bits + h, next);
    } else { /* Spec_P*: Function_G saturating at w */
        for (int i = 0; i < h; i++) child[i] = sco_g(c->format, w, llr[i], llr[h + i], bits[i]);
        leaf_rec(c, h, w, child, fb + h, bits + h, next);
    }
    for (int i = 0; i < h; i++) bits[i] ^= bits[h + i]; /* sa = sa1 ^ sb1 ; sb = sb1 */
}

void sco_leaf(int format, int extended, int p, int w, const uint32_t* llr, const uint8_t* info_flags,
              uint8_t* bits) {
    ctx_t c = {format, extended, p, w, p, info_flags, NULL, NULL};
    uint32_t scratch[1024];
    leaf_rec(&c, p, w, llr, info_flags, bits, scratch);
}

/* ------------------------------------------------------------------ tree above the leaf
 * Recursive statement of the F/R/G/H FSM (my_module.h:285-998): SURVEY.md App. A.          */
static void stats_node(const ctx_t* c, int n, int o, const uint32_t* alpha) {
    if (!c->st) return;
    int cnt = c->fsum[o + n] - c->fsum[o];
    if (cnt != n) return;
    /* maximal: parent (aligned block of 2n) is not all-information */
    if (n < c->n) {
        int po = o & ~(2 * n - 1);
        if (c->fsum[po + 2 * n] - c->fsum[po] == 2 * n) return;
    }
    int lg = 0;
    while ((1 << lg) < n) lg++;
    c->st->r1_visits[lg]++;
    for (int i = 0; i < n; i++) {
        uint32_t v = alpha[i];
        int zero = (c->format == SCO_SIGMAG) ? (sm_mag(c->q, v) == 0) : (sx(v, c->q) == 0);
        if (zero) {
            c->st->r1_with_zero[lg]++;
            break;
        }
    }
}

static void node_rec(const ctx_t* c, int n, int o, const uint32_t* alpha, uint8_t* beta, uint32_t* stack) {
    stats_node(c, n, o, alpha);
    if (n == c->par) { /* R_STATE :592-595 */
        leaf_rec(c, n, c->q, alpha, c->flags + o, beta, stack);
        return;
    }
    int h = n / 2;
    uint32_t* child = stack;
    uint32_t* next = stack + h;
    for (int i = 0; i < h; i++) child[i] = sco_f(c->format, c->q, alpha[i], alpha[h + i]); /* f_loop :373-445 */
    node_rec(c, h, o, child, beta, next);
    for (int i = 0; i < h; i++)
        child[i] = sco_g(c->format, c->q, alpha[i], alpha[h + i], beta[i]); /* g_loop :704-781 */
    node_rec(c, h, o + h, child, beta + h, next);
    for (int i = 0; i < h; i++) beta[i] ^= beta[h + i]; /* h_loop :903-932 */
}

/* ------------------------------------------------------------------ PRUNING_LEVEL 2 (REP / SPC shortcuts)
 * Restatement of the reference built with PRUNING_LEVEL 2, ELAG_R1 = ELAG_REP = ELAG_SPC = ELAG_H0 = 1,
 * ELAG_RARE = ELAG_REP2 = ELAG_SPC2 = 0 (src/module/config.h as checked in, PRUNING_LEVEL raised to 2).
 * It is NOT plain SC: a repetition child decides on a saturating running sum, a single-parity-check child
 * on signs + one flip.  Pinned on oracle/_ref/refdec_*_pl2.so (tests/test_oracle.py).                      */
enum { T_R0 = 0x0, T_R1 = 0xF, T_REP = 0x2, T_SPC = 0x4, T_RN = 0x8 }; /* my_module.h NODE_* codes */

/* do_prunning (my_module.h:61-166): the type of one PAR-wide word of the frozen table */
static int l2_word_type(const uint8_t* fb, int p) {
    int cnt = 0;
    for (int i = 0; i < p; i++) cnt += fb[i] ? 1 : 0;
    if (cnt == 0) return T_R0;
    if (cnt == p) return T_R1;
    if (cnt == 1 && fb[p - 1]) return T_REP;
    if (cnt == p - 1 && !fb[0]) return T_SPC;
    return T_RN;
}
/* F_STATE :337-545 / G_STATE: the child's type accumulated over its words */
static int l2_type(const ctx_t* c, int o, int n) {
    int p = c->par, m = n / p;
    int all_or = 0, all_and = 0xF, rep_or = 0, spc_and = 0xF, first = 0, last = 0;
    for (int k = 0; k < m; k++) {
        int t = l2_word_type(c->flags + o + k * p, p);
        all_or |= t;
        all_and &= t;
        if (k < m - 1) rep_or |= t;
        if (k > 0) spc_and &= t;
        if (k == 0) first = t;
        last = t;
    }
    if (all_or == 0) return T_R0;
    if (all_and == 0xF) return T_R1;
    if (rep_or == 0 && (last >> 1) == 1) return T_REP;
    if (spc_and == 0xF && (first >> 1) == 2) return T_SPC;
    return T_RN;
}
/* ADDER_TREE_<PAR> (functions.h:3141-3260): sum of one word plus the running sum, at width q + log2(par) + 1 */
static uint32_t l2_adder_tree(const ctx_t* c, const uint32_t* word, uint32_t old_sum) {
    int p = c->par, lg = 0;
    while ((1 << lg) < p) lg++;
    int wsum = c->q + lg + 1;
    if (c->format == SCO_CA2) { /* ADD_TREE_*_CA2 :2925-3030 is exact, then qadd (saturating) with old_sum */
        int64_t t = 0;
        for (int i = 0; i < p; i++) t += sx(word[i], c->q);
        return c2_sat(wsum, t + sx(old_sum, wsum));
    }
    /* ADD_TREE_*_SM :3032-3140: halves folded with qfull_adder_sm, one more bit per stage */
    uint32_t v[512];
    int w = c->q;
    for (int i = 0; i < p; i++) v[i] = word[i];
    for (int h = p / 2; h >= 1; h /= 2, w++)
        for (int i = 0; i < h; i++) v[i] = sm_addsub(w, v[i], v[i + h], 0);
    /* v[0]: w = q + lg bits (sign + magnitude); extended by one magnitude bit, then qfull_adder_sat_sm<wsum> */
    uint32_t ext = (((v[0] >> (w - 1)) & 1u) << (wsum - 1)) | (v[0] & mask_w(w - 1));
    uint32_t s = sm_addsub(wsum, ext, old_sum, 0);                        /* (sign, wsum-bit sum) */
    uint32_t sat = sm_sat(wsum - 1, s & mask_w(wsum));                    /* qsat_sm<wsum-1> */
    return (((s >> wsum) & 1u) << (wsum - 1)) | sat;
}
/* Min_Mask_TREE_<PAR> (functions.h:3450-3973): tournament over halves, the upper element wins only when
 * strictly smaller; returns the minimum magnitude and the one-hot position */
static void l2_min_mask(const ctx_t* c, const uint32_t* word, uint32_t* minv, int* pos) {
    int p = c->par;
    uint32_t v[512];
    int idx[512];
    for (int i = 0; i < p; i++) {
        v[i] = (c->format == SCO_CA2) ? c2_abs(c->q, word[i]) : sm_mag(c->q, word[i]);
        idx[i] = i;
    }
    for (int h = p / 2; h >= 1; h /= 2)
        for (int i = 0; i < h; i++) {
            int up = (c->format == SCO_CA2) ? (sx(v[i + h], c->q) < sx(v[i], c->q)) : (v[i + h] < v[i]);
            if (up) {
                v[i] = v[i + h];
                idx[i] = idx[i + h];
            }
        }
    *minv = v[0] & mask_w(c->q);
    *pos = idx[0];
}

static void l2_rec(const ctx_t* c, int n, int o, const uint32_t* alpha, uint8_t* beta, uint32_t* stack) {
    if (n == c->par) { /* R_STATE: leaves keep Spec_Polar_Decoder at PRUNING_LEVEL 2 (library.h:170-198) */
        leaf_rec(c, n, c->q, alpha, c->flags + o, beta, stack);
        return;
    }
    int h = n / 2, p = c->par;
    uint32_t* child = stack;
    uint32_t* next = stack + h;
    /* the root's own children are never shortcut (INIT :285-336 enters F_STATE / G_STATE directly) */
    int tl = (n == c->n) ? T_RN : l2_type(c, o, h);
    int tr_ = (n == c->n) ? T_RN : l2_type(c, o + h, h);
    if (tl == T_R0) { /* left child skipped, H0 instead of H (my_module.h H0_STATE) */
        memset(beta, 0, (size_t)h);
    } else {
        for (int i = 0; i < h; i++) child[i] = sco_f(c->format, c->q, alpha[i], alpha[h + i]);
        if (tl == T_REP) { /* F_REP_STATE :1292-1390 */
            uint32_t sum = 0;
            int lg = 0;
            while ((1 << lg) < p) lg++;
            for (int k = 0; k < h / p; k++) sum = l2_adder_tree(c, child + k * p, sum);
            memset(beta, sco_sign(c->q + lg + 1, sum), (size_t)h);
        } else {
            l2_rec(c, h, o, child, beta, next);
        }
    }
    for (int i = 0; i < h; i++)
        child[i] = sco_g(c->format, c->q, alpha[i], alpha[h + i], tl == T_R0 ? 0 : beta[i]);
    if (tr_ == T_R1) { /* G_R1_STATE :1571-1642 */
        for (int i = 0; i < h; i++) beta[h + i] = (uint8_t)sco_sign(c->q, child[i]);
    } else if (tr_ == T_SPC) { /* G_SPC_STATE :1737-1842 */
        int parity = 0, min_word = 0, min_pos = 0;
        uint32_t old_min = mask_w(c->q);
        for (int k = 0; k < h / p; k++) {
            for (int i = 0; i < p; i++) {
                beta[h + k * p + i] = (uint8_t)sco_sign(c->q, child[k * p + i]);
                parity ^= beta[h + k * p + i];
            }
            uint32_t mv;
            int pos;
            l2_min_mask(c, child + k * p, &mv, &pos);
            if (mv < old_min) {
                old_min = mv;
                min_word = k;
                min_pos = pos;
            }
        }
        if (parity) beta[h + min_word * p + min_pos] ^= 1;
    } else {
        l2_rec(c, h, o + h, child, beta + h, next);
    }
    for (int i = 0; i < h; i++) beta[i] ^= beta[h + i];
}

static int check_cfg(const sco_config* cfg) {
    if (!cfg) return -1;
    int n = cfg->n, p = cfg->par;
    if (n < 2 || (n & (n - 1))) return -2;
    if (p < 1 || (p & (p - 1)) || 2 * p > n) return -3; /* INIT needs N_DIVIDED>>1 >= 1 (my_module.h:294) */
    if (cfg->llr_bits < 3 || cfg->llr_bits > 12) return -4;
    if (cfg->format != SCO_CA2 && cfg->format != SCO_SIGMAG) return -5;
    if (p > 512) return -6;
    return 0;
}

static int decode_impl(const sco_config* cfg, const uint8_t* flags, const int8_t* llr, size_t nframes,
                       uint8_t* xhat, uint32_t* xhat32, sco_stats* st, int level2) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    int n = cfg->n;
    uint32_t* alpha = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n * 3 + 4096);
    uint8_t* beta = (uint8_t*)malloc((size_t)n);
    int32_t* fsum = NULL;
    if (!alpha || !beta) {
        free(alpha);
        free(beta);
        return -10;
    }
    ctx_t c = {cfg->format, cfg->extended, cfg->par, cfg->llr_bits, n, flags, st, NULL};
    if (st) {
        fsum = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
        fsum[0] = 0;
        for (int i = 0; i < n; i++) fsum[i + 1] = fsum[i] + (flags[i] ? 1 : 0);
        c.fsum = fsum;
    }
    for (size_t f = 0; f < nframes; f++) {
        const int8_t* in = llr + f * (size_t)n;
        for (int i = 0; i < n; i++) alpha[i] = sco_input(cfg->format, cfg->llr_bits, in[i]); /* wrapper_in.h:33-34 */
        if (level2) l2_rec(&c, n, 0, alpha, beta, alpha + n);
        else node_rec(&c, n, 0, alpha, beta, alpha + n);
        if (xhat) memcpy(xhat + f * (size_t)n, beta, (size_t)n);
        if (xhat32) {
            if (n >= 32) {
                uint32_t* out = xhat32 + f * (size_t)(n / 32);
                for (int wd = 0; wd < n / 32; wd++) {
                    uint32_t v = 0;
                    for (int b = 0; b < 32; b++) v |= (uint32_t)(beta[wd * 32 + b] & 1) << b;
                    out[wd] = v;
                }
            } else {
                uint32_t v = 0;
                for (int b = 0; b < n; b++) v |= (uint32_t)(beta[b] & 1) << b;
                xhat32[f] = v;
            }
        }
    }
    free(alpha);
    free(beta);
    free(fsum);
    return 0;
}

int sco_decode(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr, size_t nframes,
               uint8_t* xhat) {
    return decode_impl(cfg, info_flags, llr, nframes, xhat, NULL, NULL, 0);
}
int sco_decode_packed(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr, size_t nframes,
                      uint32_t* xhat32) {
    return decode_impl(cfg, info_flags, llr, nframes, NULL, xhat32, NULL, 0);
}
int sco_decode_stats(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr, size_t nframes,
                     uint8_t* xhat, sco_stats* st) {
    return decode_impl(cfg, info_flags, llr, nframes, xhat, NULL, st, 0);
}
int sco_decode_l2(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr, size_t nframes,
                  uint8_t* xhat) {
    return decode_impl(cfg, info_flags, llr, nframes, xhat, NULL, NULL, 1);
}
int sco_decode_l2_packed(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr, size_t nframes,
                         uint32_t* xhat32) {
    return decode_impl(cfg, info_flags, llr, nframes, NULL, xhat32, NULL, 1);
}

typedef struct {
    const sco_config* cfg;
    const uint8_t* flags;
    const int8_t* llr;
    size_t nframes;
    uint32_t* out;
    int rc;
} mt_job;
static void* mt_run(void* p) {
    mt_job* j = (mt_job*)p;
    j->rc = decode_impl(j->cfg, j->flags, j->llr, j->nframes, NULL, j->out, NULL, 0);
    return NULL;
}
int sco_decode_packed_mt(const sco_config* cfg, const uint8_t* info_flags, const int8_t* llr,
                         size_t nframes, uint32_t* xhat32, int nthreads) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > nframes) nthreads = nframes ? (int)nframes : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    mt_job* jobs = (mt_job*)malloc(sizeof(mt_job) * (size_t)nthreads);
    size_t wpf = cfg->n >= 32 ? (size_t)cfg->n / 32 : 1;
    size_t done = 0;
    for (int t = 0; t < nthreads; t++) {
        size_t cnt = nframes / (size_t)nthreads + ((size_t)t < nframes % (size_t)nthreads ? 1 : 0);
        jobs[t] = (mt_job){cfg, info_flags, llr + done * (size_t)cfg->n, cnt, xhat32 + done * wpf, 0};
        done += cnt;
        pthread_create(&th[t], NULL, mt_run, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(th);
    free(jobs);
    return rc;
}

/* ------------------------------------------------------------------ polar transform */
void sco_polar_transform(uint8_t* bits, int n) { /* x = u * F^{(x)log2 n}, natural order */
    for (int h = 1; h < n; h <<= 1)
        for (int b = 0; b < n; b += 2 * h)
            for (int i = 0; i < h; i++) bits[b + i] ^= bits[b + h + i];
}

/* ------------------------------------------------------------------ channel chain (App. C) */
void sco_xs128_seed(sco_xs128* a, sco_xs128* b, uint8_t seed) {
    uint32_t m = (uint32_t)seed * 0x01010101u; /* xMk = (mask,mask,mask,mask) sc_xorshift128.h:62 */
    a->x = 0x12311178u & m; /* :63-66 */
    a->y = 0x65498732u | m;
    a->z = 0xFEDCAA01u ^ m;
    a->w = 0xF489A179u + m;
    b->x = 0x98765432u & m; /* :101-104 */
    b->y = 0x12345678u | m;
    b->z = 0xFCBADEFFu ^ m;
    b->w = 0x12121212u + m;
}
uint32_t sco_xs128_next(sco_xs128* s) { /* :78-85 */
    uint32_t t = s->x;
    t ^= t << 11;
    t ^= t >> 8;
    s->x = s->y;
    s->y = s->z;
    s->z = s->w;
    s->w ^= s->w >> 19;
    s->w ^= t;
    return s->w;
}
float sco_xs128_uniform(uint32_t w) { /* :86  f = 1.0f - (float)w * (1.0f/4294967296.0f) */
    volatile float fw = (float)w;
    volatile float p = fw * (1.0f / 4294967296.0f);
    volatile float f = 1.0f - p;
    return f;
}
float sco_sigma(float ebn0_db, float rate) { /* main.cpp:91-98 */
    return 1.0f / sqrtf(2.f * rate * powf(10.f, ebn0_db / 10.f));
}
int sco_quantize(float y) { /* sc_quantizer.h:77-80, BETA=4, VSAT=+-31 (main.cpp:16-18) */
    volatile float scaled = y * 4.0f;
    short iv = (short)scaled;
    short mv = (iv > -31) ? iv : (short)-31;
    short rv = (mv < 31) ? mv : (short)31;
    return rv;
}
void sco_channel(int n, size_t frame0, size_t nframes, uint8_t seed, float sigma, const uint8_t* codeword,
                 int per_frame_cw, int8_t* llr) {
    sco_xs128 a, b;
    sco_xs128_seed(&a, &b, seed);
    size_t skip = frame0 * (size_t)(n / 2);
    for (size_t i = 0; i < skip; i++) {
        sco_xs128_next(&a);
        sco_xs128_next(&b);
    }
    const float _1PI = 3.14159265358979f; /* sc_awgn.h:61-62 */
    const float _2PI = 2.0f * _1PI;
    for (size_t f = 0; f < nframes; f++) {
        const uint8_t* cw = codeword ? (per_frame_cw ? codeword + f * (size_t)n : codeword) : NULL;
        for (int i = 0; i < n; i += 2) {
            float r1 = sco_xs128_uniform(sco_xs128_next(&a));
            float r2 = sco_xs128_uniform(sco_xs128_next(&b));
            /* G11: r1 == 0 makes logf(-inf) and the (short) cast undefined in the reference;
             * this restatement (and the device channel) clamp r1 to 2^-24. */
            if (r1 < 5.9604644775390625e-08f) r1 = 5.9604644775390625e-08f;
            volatile float y = _2PI * r2;         /* sc_awgn.h:67 */
            volatile float lg = -2.0f * logf(r1); /* :68 */
            volatile float x = sqrtf(lg);
            volatile float ph = x * sinf(y); /* :74-77 */
            volatile float qu = x * cosf(y);
            float noise[2] = {ph, qu};
            for (int k = 0; k < 2; k++) {
                int bit = cw ? cw[i + k] : 0;
                float obit = bit ? -1.0f : +1.0f;   /* sc_bpsk.h:53 */
                volatile float nn = noise[k] * sigma; /* sc_adder.h:139 */
                volatile float nbit = obit + nn;     /* :140 */
                llr[f * (size_t)n + (size_t)(i + k)] = (int8_t)sco_quantize(nbit);
            }
        }
    }
}

/* ------------------------------------------------------------------ error counter */
void sco_count_errors(int n, size_t nframes, const uint8_t* xhat, const uint8_t* ref, int per_frame_ref,
                      uint64_t counters[6]) {
    for (int k = 0; k < 6; k++) counters[k] = 0;
    for (size_t f = 0; f < nframes; f++) {
        const uint8_t* r = ref ? (per_frame_ref ? ref + f * (size_t)n : ref) : NULL;
        uint64_t err = 0;
        for (int i = 0; i < n; i++) err += (uint64_t)((xhat[f * (size_t)n + i] & 1) != (r ? (r[i] & 1) : 0));
        counters[0] += err;                 /* errBE += err      sc_error_counter.h:105 */
        counters[1] += (err != 0);          /* errFE += or_reduce :106 */
        counters[2] += (uint64_t)n;         /* pBits              :107 */
        counters[3] += 1;                   /* pFrames            :108 */
        counters[4] += (err & 1023u);       /* sc_uint<10> err    :70-71 (G9) */
        counters[5] += ((err & 1023u) != 0);
    }
}
