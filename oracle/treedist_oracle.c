/*
 * treedist_oracle.c -- CPU ORACLE for the treeCl inter-tree distance path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the algorithm behind
 *     treeCl.Collection.get_inter_tree_distances('rf'|'wrf'|'euc'|'geo')   (treeCl/collection.py:421-456)
 *     treeCl.treedist.{rf,wrf,euc,geo}dist / *_matrix                       (treeCl/treedist.py:30-118)
 * It is NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product path (treecl_b200 + libtreedist_cuda.so) never does.
 *
 * Where the arithmetic comes from.  treeCl delegates all four distances to the PyPI package
 * `tree_distance` (>=1.0.12, treeCl setup.py:117; a Cython wrapper round a C++ port of M. Owen's GTP code),
 * whose source is NOT under /root/reference and is not installed here.  So this oracle restates the
 * published algorithms (Robinson & Foulds 1981; Owen & Provan 2011 "A fast algorithm for computing geodesic
 * distances in tree space") and is PINNED against the reference's own golden vectors:
 *     tests/test.py:243-273 (8 scalar known answers), :280-294 (partial overlap), :320-326 / :437-467 (15-tree
 *     geodesic checksum), tests/data/cache/geo_dm.csv (105 geodesic values)          -> tests/test_oracle_golden.py
 * Parity is UNPINNED (no reference test exists) for: rooted input, polytomies (extended common-edge rule),
 * zero/missing branch lengths, rf/wrf/euc under partial overlap.  See DESIGN.md.
 *
 * Reference call structure followed here (file:line in /root/reference):
 *     rfdist_task & co         treeCl/tasks.py:41-75     parse BOTH Newick strings per pair -> orc_pair()
 *     _generic_distance_calc   treeCl/treedist.py:30-70  leaf-set check, min_overlap, prune to intersection
 *     _equalise_leaf_sets      treeCl/treedist.py:17-27
 *     Tree.prune_to_subset     treeCl/tree.py:989-1001   dendropy retain_taxa + encode_bipartitions
 *     Tree.rooted              treeCl/tree.py:858-862    rooted <=> root has exactly two children
 *     Tree.phylotree           treeCl/tree.py:830-845    PhyloTree(newick, rooted)
 *     scrape_args              treeCl/tasks.py:651-653   pair order = itertools.combinations = scipy condensed
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_RF = 0, ORC_WRF = 1, ORC_EUC = 2, ORC_GEO = 3 };
enum { ORC_OK = 0, ORC_EPARSE = 1, ORC_EMIXED = 2, ORC_EARG = 3, ORC_ENOMEM = 4 };

static __thread char g_err[256];
const char *orc_last_error(void) { return g_err; }
#define FAIL(code, ...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return (code); } while (0)

/* ------------------------------------------------------------------------------------------------ */
/* Newick -> node table.  Mirrors what reaches tree_distance after treeCl's dendropy round trip            */
/* (treeCl/tree.py:752-756, 808-828): internal node labels dropped, comments dropped, quoted labels         */
/* unquoted, underscores preserved, the root edge length ignored.                                           */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int parent, first_child, last_child, next_sib, nchild;
    double len;      /* edge to parent; 0 when absent (UNPINNED: reference behaviour for missing lengths) */
    char *label;     /* leaves only */
    int alive;
} node_t;

typedef struct {
    node_t *nd;
    int n, cap, root;
} tree_t;

static int tree_new_node(tree_t *t, int parent)
{
    if (t->n == t->cap) {
        t->cap = t->cap ? 2 * t->cap : 64;
        t->nd = (node_t *)realloc(t->nd, sizeof(node_t) * t->cap);
    }
    node_t *x = &t->nd[t->n];
    x->parent = parent; x->first_child = x->last_child = x->next_sib = -1; x->nchild = 0;
    x->len = 0.0; x->label = NULL; x->alive = 1;
    if (parent >= 0) {
        node_t *p = &t->nd[parent];
        if (p->last_child >= 0) t->nd[p->last_child].next_sib = t->n; else p->first_child = t->n;
        p->last_child = t->n; p->nchild++;
    }
    return t->n++;
}

static void tree_free(tree_t *t)
{
    for (int i = 0; i < t->n; i++) free(t->nd[i].label);
    free(t->nd); t->nd = NULL; t->n = t->cap = 0;
}

static const char *skip_ws(const char *s)
{
    for (;;) {
        while (*s && isspace((unsigned char)*s)) s++;
        if (*s == '[') { while (*s && *s != ']') s++; if (*s) s++; continue; }
        return s;
    }
}

static const char *read_label(const char *s, char **out)
{
    size_t cap = 32, n = 0; char *buf = (char *)malloc(cap);
    if (*s == '\'') {
        s++;
        for (;;) {
            if (!*s) { free(buf); return NULL; }
            if (*s == '\'') { if (s[1] == '\'') { s++; } else { s++; break; } }
            if (n + 2 > cap) { cap *= 2; buf = (char *)realloc(buf, cap); }
            buf[n++] = *s++;
        }
    } else {
        while (*s && !strchr("():,;[", *s) && !isspace((unsigned char)*s)) {
            if (n + 2 > cap) { cap *= 2; buf = (char *)realloc(buf, cap); }
            buf[n++] = *s++;
        }
    }
    buf[n] = 0; *out = buf; return s;
}

static int tree_parse(const char *s, tree_t *t)
{
    memset(t, 0, sizeof *t);
    int cur = tree_new_node(t, -1); t->root = cur;
    s = skip_ws(s);
    if (*s != '(') { /* single-leaf "A;" is not a tree for this path */
        tree_free(t); FAIL(ORC_EPARSE, "newick: expected '('");
    }
    int depth = 0;
    /* state machine: we are positioned at the start of a subtree description for node `cur` */
    int expecting_subtree = 1;
    while (*s) {
        s = skip_ws(s);
        if (expecting_subtree) {
            if (*s == '(') { depth++; cur = tree_new_node(t, cur); s++; continue; }   /* first child of cur's... */
            /* leaf label */
            char *lab = NULL; const char *e = read_label(s, &lab);
            if (!e) { tree_free(t); FAIL(ORC_EPARSE, "newick: unterminated quoted label"); }
            if (!lab[0]) { free(lab); tree_free(t); FAIL(ORC_EPARSE, "newick: empty leaf label at '%.20s'", s); }
            t->nd[cur].label = lab; s = e; expecting_subtree = 0; continue;
        }
        /* after a subtree: optional internal label, optional :length, then , ) or ; */
        if (*s == ':') {
            s = skip_ws(s + 1);
            char *e; double v = strtod(s, &e);
            if (e == s) { tree_free(t); FAIL(ORC_EPARSE, "newick: bad branch length at '%.20s'", s); }
            t->nd[cur].len = v; s = e; continue;
        }
        if (*s == ',') {
            if (depth == 0) { tree_free(t); FAIL(ORC_EPARSE, "newick: ',' outside parentheses"); }
            cur = tree_new_node(t, t->nd[cur].parent); s++; expecting_subtree = 1; continue;
        }
        if (*s == ')') {
            if (depth == 0) { tree_free(t); FAIL(ORC_EPARSE, "newick: unbalanced ')'"); }
            depth--; cur = t->nd[cur].parent; s++;
            continue;
        }
        if (*s == ';' || !*s) break;
        /* internal node label (support value): dropped, treeCl/tree.py:823-825 */
        { char *lab = NULL; const char *e = read_label(s, &lab);
          if (!e || e == s) { free(lab); tree_free(t); FAIL(ORC_EPARSE, "newick: unexpected '%c'", *s); }
          free(lab); s = e; }
    }
    if (depth != 0) { tree_free(t); FAIL(ORC_EPARSE, "newick: unbalanced '('"); }
    /* The parser above created the root node and then, on the first '(', a child: fix-up is not needed
     * because '(' always opens a child of `cur`; the root is node 0 and its own label/length are ignored. */
    if (cur != t->root) { tree_free(t); FAIL(ORC_EPARSE, "newick: truncated"); }
    for (int i = 0; i < t->n; i++)
        if (t->nd[i].nchild == 0 && !t->nd[i].label) { tree_free(t); FAIL(ORC_EPARSE, "newick: unlabeled leaf"); }
    return ORC_OK;
}

/* The state machine opens the root with the FIRST '(' as "child of node 0".  To keep node 0 == root we
 * strip that artificial level here: if the root has exactly one child and that child is internal and it was
 * produced by the outermost parentheses, promote it.  (Newick "(A,B,C);" parses as root0 -> X(A,B,C).) */
static void tree_strip_outer(tree_t *t)
{
    node_t *r = &t->nd[t->root];
    if (r->nchild == 1 && t->nd[r->first_child].nchild > 0) {
        int c = r->first_child;
        t->nd[c].parent = -1; t->nd[c].len = 0.0; r->alive = 0; t->root = c;
    }
}

/* number of leaves + label list */
static int tree_leaves(const tree_t *t, int *out /* may be NULL */)
{
    int k = 0;
    for (int i = 0; i < t->n; i++)
        if (t->nd[i].alive && t->nd[i].nchild == 0 && t->nd[i].label) { if (out) out[k] = i; k++; }
    return k;
}

/* ------------------------------------------------------------------------------------------------ */
/* Pruning at the TREE level: restates dendropy retain_taxa_with_labels + suppress_unifurcations +        */
/* encode_bipartitions(collapse_unrooted_basal_bifurcation) as used by treeCl/tree.py:989-1001.            */
/* Edge lengths merged by unifurcation suppression are SUMMED (pinned by reference tests/test.py:286).     */
/* ------------------------------------------------------------------------------------------------ */
static void unlink_child(tree_t *t, int c)
{
    int p = t->nd[c].parent; if (p < 0) return;
    int prev = -1, x = t->nd[p].first_child;
    while (x != c) { prev = x; x = t->nd[x].next_sib; }
    if (prev < 0) t->nd[p].first_child = t->nd[c].next_sib; else t->nd[prev].next_sib = t->nd[c].next_sib;
    if (t->nd[p].last_child == c) t->nd[p].last_child = prev;
    t->nd[p].nchild--; t->nd[c].parent = -1; t->nd[c].next_sib = -1;
}

static void link_child(tree_t *t, int p, int c)
{
    t->nd[c].parent = p; t->nd[c].next_sib = -1;
    if (t->nd[p].last_child >= 0) t->nd[t->nd[p].last_child].next_sib = c; else t->nd[p].first_child = c;
    t->nd[p].last_child = c; t->nd[p].nchild++;
}

static int keep_label(const char *lab, char *const *keep, int nkeep)
{
    int lo = 0, hi = nkeep - 1;                      /* keep[] is sorted */
    while (lo <= hi) { int m = (lo + hi) / 2; int c = strcmp(lab, keep[m]);
        if (c == 0) return 1; if (c < 0) hi = m - 1; else lo = m + 1; }
    return 0;
}

static void tree_prune(tree_t *t, char *const *keep, int nkeep, int was_rooted)
{
    /* 1. drop leaves whose label is not retained, then internal nodes left without children */
    int changed = 1;
    while (changed) {
        changed = 0;
        for (int i = 0; i < t->n; i++) {
            node_t *x = &t->nd[i];
            if (!x->alive || i == t->root || x->nchild != 0) continue;
            if (x->label && keep_label(x->label, keep, nkeep)) continue;
            unlink_child(t, i); x->alive = 0; changed = 1;
        }
    }
    /* 2. suppress unifurcations below the root: the child inherits the summed length */
    changed = 1;
    while (changed) {
        changed = 0;
        for (int i = 0; i < t->n; i++) {
            node_t *x = &t->nd[i];
            if (!x->alive || i == t->root || x->nchild != 1) continue;
            int c = x->first_child, p = x->parent;
            double l = x->len + t->nd[c].len;
            unlink_child(t, c); unlink_child(t, i); x->alive = 0;
            link_child(t, p, c); t->nd[c].len = l; changed = 1;
        }
    }
    /* 3. a root with one child: the child becomes the root, its stem edge is dropped */
    while (t->nd[t->root].nchild == 1 && t->nd[t->nd[t->root].first_child].nchild > 0) {
        int c = t->nd[t->root].first_child;
        unlink_child(t, c); t->nd[t->root].alive = 0; t->root = c; t->nd[c].len = 0.0;
    }
    /* 4. unrooted trees keep a trifurcating base: merge the two root edges (lengths summed) */
    if (!was_rooted && t->nd[t->root].nchild == 2) {
        int a = t->nd[t->root].first_child, b = t->nd[a].next_sib, del = -1, keepn = -1;
        if (t->nd[b].nchild >= 2) { keepn = a; del = b; } else if (t->nd[a].nchild >= 2) { keepn = b; del = a; }
        if (del >= 0) {
            t->nd[keepn].len += t->nd[del].len;
            int c = t->nd[del].first_child;
            unlink_child(t, del); t->nd[del].alive = 0;
            while (c >= 0) { int nx = t->nd[c].next_sib; t->nd[c].parent = -1; t->nd[c].next_sib = -1;
                             link_child(t, t->root, c); c = nx; }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Split encoding (what tree_distance.PhyloTree(newick, rooted) holds): one bit-set per internal edge     */
/* with its length, one length per leaf edge.  Unrooted: the side NOT containing the lowest present taxon. */
/* Rooted: the clade below the edge (both root edges are distinct edges).                                  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int W, n_leaves, n_splits, rooted, n_taxa;
    uint64_t *leafmask;   /* W */
    uint64_t *masks;      /* n_splits * W, sorted, unique */
    double *lens;         /* n_splits */
    double *leaflen;      /* n_taxa (0 where absent) */
} enc_t;

static void enc_free(enc_t *e) { free(e->leafmask); free(e->masks); free(e->lens); free(e->leaflen); memset(e, 0, sizeof *e); }

static int popc(const uint64_t *m, int W) { int c = 0; for (int w = 0; w < W; w++) c += __builtin_popcountll(m[w]); return c; }
static int lowbit(const uint64_t *m, int W) { for (int w = 0; w < W; w++) if (m[w]) return 64 * w + __builtin_ctzll(m[w]); return -1; }
static int mcmp(const uint64_t *a, const uint64_t *b, int W)
{ for (int w = W - 1; w >= 0; w--) { if (a[w] < b[w]) return -1; if (a[w] > b[w]) return 1; } return 0; }

typedef struct { const uint64_t *m; double l; } srec_t;
static int g_W_sort;  /* qsort context; guarded by being set per call in one thread (we use qsort_r) */
static int srec_cmp(const void *a, const void *b, void *ctx)
{ int W = *(int *)ctx; return mcmp(((const srec_t *)a)->m, ((const srec_t *)b)->m, W); }

/* taxa[]: sorted label table; returns index or -1 */
static int taxon_index(const char *lab, char *const *taxa, int ntaxa)
{
    int lo = 0, hi = ntaxa - 1;
    while (lo <= hi) { int m = (lo + hi) / 2; int c = strcmp(lab, taxa[m]);
        if (c == 0) return m; if (c < 0) hi = m - 1; else lo = m + 1; }
    return -1;
}

static int tree_encode(const tree_t *t, char *const *taxa, int ntaxa, enc_t *e)
{
    memset(e, 0, sizeof *e);
    int W = (ntaxa + 63) / 64; if (W < 1) W = 1;
    e->W = W; e->n_taxa = ntaxa;
    e->rooted = (t->nd[t->root].nchild == 2);                 /* treeCl/tree.py:858-862 */
    e->leafmask = (uint64_t *)calloc(W, 8);
    e->leaflen = (double *)calloc(ntaxa > 0 ? ntaxa : 1, sizeof(double));
    uint64_t *clade = (uint64_t *)calloc((size_t)t->n * W, 8);
    /* post-order accumulation: children were always created after their parent, so a reverse scan works */
    for (int i = t->n - 1; i >= 0; i--) {
        const node_t *x = &t->nd[i]; if (!x->alive) continue;
        if (x->nchild == 0) {
            int k = taxon_index(x->label, taxa, ntaxa);
            if (k < 0) { free(clade); enc_free(e); FAIL(ORC_EARG, "label '%s' not in taxon table", x->label); }
            if (e->leafmask[k >> 6] >> (k & 63) & 1) { free(clade); enc_free(e); FAIL(ORC_EPARSE, "duplicate leaf label '%s'", x->label); }
            clade[(size_t)i * W + (k >> 6)] |= 1ull << (k & 63);
            e->leafmask[k >> 6] |= 1ull << (k & 63);
        }
        if (x->parent >= 0) for (int w = 0; w < W; w++) clade[(size_t)x->parent * W + w] |= clade[(size_t)i * W + w];
    }
    e->n_leaves = popc(e->leafmask, W);
    int ref = lowbit(e->leafmask, W);
    srec_t *rec = (srec_t *)malloc(sizeof(srec_t) * (t->n + 1)); int nr = 0;
    for (int i = 0; i < t->n; i++) {
        const node_t *x = &t->nd[i]; if (!x->alive || i == t->root) continue;
        uint64_t *m = &clade[(size_t)i * W];
        int c = popc(m, W), cc = e->n_leaves - c;
        if (c == 0) continue;
        if (c == 1) { e->leaflen[lowbit(m, W)] += x->len; continue; }
        if (e->rooted) { if (cc == 0) continue; }
        else {
            if (cc == 0) continue;
            if (cc == 1) { for (int w = 0; w < W; w++) m[w] = e->leafmask[w] & ~m[w]; e->leaflen[lowbit(m, W)] += x->len; continue; }
            if (m[ref >> 6] >> (ref & 63) & 1) for (int w = 0; w < W; w++) m[w] = e->leafmask[w] & ~m[w];
        }
        rec[nr].m = m; rec[nr].l = x->len; nr++;
    }
    qsort_r(rec, nr, sizeof(srec_t), srec_cmp, &W);
    e->masks = (uint64_t *)malloc((size_t)(nr > 0 ? nr : 1) * W * 8);
    e->lens = (double *)malloc(sizeof(double) * (nr > 0 ? nr : 1));
    int ns = 0;
    for (int i = 0; i < nr; i++) {
        if (ns > 0 && mcmp(&e->masks[(size_t)(ns - 1) * W], rec[i].m, W) == 0) { e->lens[ns - 1] += rec[i].l; continue; }
        memcpy(&e->masks[(size_t)ns * W], rec[i].m, W * 8); e->lens[ns] = rec[i].l; ns++;
    }
    e->n_splits = ns;
    free(rec); free(clade);
    (void)g_W_sort;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* Distances on two encodings over the SAME leaf set.                                                   */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { long maxflows, ratios, pairs; } orc_stats_t;
static orc_stats_t g_stats;               /* not thread-exact; diagnostic only */
void orc_stats(long *maxflows, long *ratios, long *pairs) { *maxflows = g_stats.maxflows; *ratios = g_stats.ratios; *pairs = g_stats.pairs; }
void orc_stats_reset(void) { memset(&g_stats, 0, sizeof g_stats); }

static int splits_compatible(const uint64_t *a, const uint64_t *b, const uint64_t *leaf, int W, int rooted)
{
    /* two splits are compatible iff one of the four side-intersections is empty.  With the canonical side
     * (never containing the reference taxon; clades for rooted trees, where the virtual root plays that role)
     * the complements always intersect, so: a&b == 0, or a subset b, or b subset a. */
    (void)leaf; (void)rooted;
    int disjoint = 1, asub = 1, bsub = 1;
    for (int w = 0; w < W; w++) {
        uint64_t x = a[w] & b[w];
        if (x) disjoint = 0;
        if (x != a[w]) asub = 0;
        if (x != b[w]) bsub = 0;
    }
    return disjoint | asub | bsub;
}

/* Min-weight vertex cover of the bipartite incompatibility graph by max-flow (augmenting paths, BFS),
 * Owen & Provan 2011 section 4.  wa/wb: vertex weights; inc[i*nb+j] != 0 <=> a_i ~ b_j incompatible.
 * On return inA[i]/inB[j] flag the cover. */
static void min_vertex_cover(int na, int nb, const double *wa, const double *wb, const unsigned char *inc,
                             unsigned char *inA, unsigned char *inB)
{
    const double TOL = 1e-13;
    double *ra = (double *)malloc(sizeof(double) * na), *rb = (double *)malloc(sizeof(double) * nb);
    double *f = (double *)calloc((size_t)na * nb, sizeof(double));
    int *pa = (int *)malloc(sizeof(int) * na), *pb = (int *)malloc(sizeof(int) * nb);   /* BFS parents */
    int *q = (int *)malloc(sizeof(int) * (na + nb));
    memcpy(ra, wa, sizeof(double) * na); memcpy(rb, wb, sizeof(double) * nb);
    double scale = 0; for (int i = 0; i < na; i++) scale += wa[i];
    double tol = TOL * (scale > 1 ? scale : 1);
    for (;;) {
        /* BFS in the residual graph from the source. node ids: a_i -> i, b_j -> na + j */
        for (int i = 0; i < na; i++) pa[i] = -2; for (int j = 0; j < nb; j++) pb[j] = -2;
        int qh = 0, qt = 0, sink_b = -1;
        for (int i = 0; i < na; i++) if (ra[i] > tol) { pa[i] = -1; q[qt++] = i; }
        while (qh < qt && sink_b < 0) {
            int u = q[qh++];
            if (u < na) {
                for (int j = 0; j < nb; j++) if (inc[(size_t)u * nb + j] && pb[j] == -2) {
                    pb[j] = u; if (rb[j] > tol) { sink_b = j; break; } q[qt++] = na + j; }
            } else {
                int j = u - na;
                for (int i = 0; i < na; i++) if (pa[i] == -2 && f[(size_t)i * nb + j] > tol) { pa[i] = j; q[qt++] = i; }
            }
        }
        if (sink_b < 0) {
            /* no augmenting path: cover = (A not reachable) + (B reachable) */
            for (int i = 0; i < na; i++) inA[i] = (pa[i] == -2);
            for (int j = 0; j < nb; j++) inB[j] = (pb[j] != -2);
            break;
        }
        /* bottleneck */
        double d = rb[sink_b]; int j = sink_b;
        for (;;) { int i = pb[j]; if (pa[i] == -1) { if (ra[i] < d) d = ra[i]; break; }
                   int j2 = pa[i]; if (f[(size_t)i * nb + j2] < d) d = f[(size_t)i * nb + j2]; j = j2; }
        rb[sink_b] -= d; j = sink_b;
        for (;;) { int i = pb[j]; f[(size_t)i * nb + j] += d;
                   if (pa[i] == -1) { ra[i] -= d; break; }
                   int j2 = pa[i]; f[(size_t)i * nb + j2] -= d; j = j2; }
    }
    free(ra); free(rb); free(f); free(pa); free(pb); free(q);
}

/* GTP on the non-common edges: returns sum_k (||A_k|| + ||B_k||)^2 */
static double gtp_ratio_term(int na, int nb, const uint64_t *am, const double *al, const uint64_t *bm, const double *bl,
                             const uint64_t *leaf, int W, int rooted)
{
    if (na == 0 && nb == 0) return 0.0;
    unsigned char *inc = (unsigned char *)malloc((size_t)(na ? na : 1) * (nb ? nb : 1));
    for (int i = 0; i < na; i++) for (int j = 0; j < nb; j++)
        inc[(size_t)i * nb + j] = !splits_compatible(&am[(size_t)i * W], &bm[(size_t)j * W], leaf, W, rooted);
    /* explicit stack of ratios, each a pair of index lists; processed in sequence order */
    typedef struct { int *a, *b; int na, nb; } ratio_t;
    int cap = 2 * (na + nb) + 4, top = 0;
    ratio_t *st = (ratio_t *)malloc(sizeof(ratio_t) * cap);
    st[0].a = (int *)malloc(sizeof(int) * (na ? na : 1)); st[0].b = (int *)malloc(sizeof(int) * (nb ? nb : 1));
    for (int i = 0; i < na; i++) st[0].a[i] = i; for (int j = 0; j < nb; j++) st[0].b[j] = j;
    st[0].na = na; st[0].nb = nb; top = 1;
    double total = 0.0;
    double *wa = (double *)malloc(sizeof(double) * (na ? na : 1)), *wb = (double *)malloc(sizeof(double) * (nb ? nb : 1));
    unsigned char *sub = (unsigned char *)malloc((size_t)(na ? na : 1) * (nb ? nb : 1));
    unsigned char *inA = (unsigned char *)malloc(na ? na : 1), *inB = (unsigned char *)malloc(nb ? nb : 1);
    while (top > 0) {
        ratio_t r = st[--top];
        double sa = 0, sb = 0;
        for (int i = 0; i < r.na; i++) sa += al[r.a[i]] * al[r.a[i]];
        for (int j = 0; j < r.nb; j++) sb += bl[r.b[j]] * bl[r.b[j]];
        int final = 1;
        if (r.na > 0 && r.nb > 0 && sa > 0 && sb > 0 && (r.na > 1 || r.nb > 1)) {
            for (int i = 0; i < r.na; i++) wa[i] = al[r.a[i]] * al[r.a[i]] / sa;
            for (int j = 0; j < r.nb; j++) wb[j] = bl[r.b[j]] * bl[r.b[j]] / sb;
            for (int i = 0; i < r.na; i++) for (int j = 0; j < r.nb; j++) sub[(size_t)i * r.nb + j] = inc[(size_t)r.a[i] * nb + r.b[j]];
            min_vertex_cover(r.na, r.nb, wa, wb, sub, inA, inB);
            __sync_fetch_and_add(&g_stats.maxflows, 1);
            double cw = 0; int ca = 0, cb = 0;
            for (int i = 0; i < r.na; i++) if (inA[i]) { cw += wa[i]; ca++; }
            for (int j = 0; j < r.nb; j++) if (inB[j]) { cw += wb[j]; cb++; }
            if (cw < 1.0 - 1e-12 && ca > 0 && ca < r.na && cb > 0 && cb < r.nb) {
                /* split: (C&A, B\C) comes first, then (A\C, C&B): push in reverse so the first pops first */
                ratio_t r1, r2;
                r1.a = (int *)malloc(sizeof(int) * ca); r1.b = (int *)malloc(sizeof(int) * (r.nb - cb));
                r2.a = (int *)malloc(sizeof(int) * (r.na - ca)); r2.b = (int *)malloc(sizeof(int) * cb);
                r1.na = r1.nb = r2.na = r2.nb = 0;
                for (int i = 0; i < r.na; i++) if (inA[i]) r1.a[r1.na++] = r.a[i]; else r2.a[r2.na++] = r.a[i];
                for (int j = 0; j < r.nb; j++) if (inB[j]) r2.b[r2.nb++] = r.b[j]; else r1.b[r1.nb++] = r.b[j];
                if (top + 2 > cap) { cap *= 2; st = (ratio_t *)realloc(st, sizeof(ratio_t) * cap); }
                st[top++] = r2; st[top++] = r1; final = 0;
            }
        }
        if (final) { double s = sqrt(sa) + sqrt(sb); total += s * s; __sync_fetch_and_add(&g_stats.ratios, 1); }
        free(r.a); free(r.b);
    }
    free(st); free(wa); free(wb); free(sub); free(inA); free(inB); free(inc);
    return total;
}

static int enc_distance(const enc_t *x, const enc_t *y, int metric, int normalise, double *out)
{
    int W = x->W;
    if (x->rooted != y->rooted) FAIL(ORC_EMIXED, "cannot compare a rooted with an unrooted tree");
    /* classify internal splits by a merge over the two sorted lists */
    int nx = x->n_splits, ny = y->n_splits;
    int i = 0, j = 0, common = 0;
    double wrf = 0, sq = 0;               /* wrf: L1; sq: L2^2 over internal edges where both/one present */
    int *ax = (int *)malloc(sizeof(int) * (nx + 1)), *by = (int *)malloc(sizeof(int) * (ny + 1)); int na = 0, nb = 0;
    double common_sq = 0;
    while (i < nx || j < ny) {
        int c = (i >= nx) ? 1 : (j >= ny) ? -1 : mcmp(&x->masks[(size_t)i * W], &y->masks[(size_t)j * W], W);
        if (c == 0) { double d = x->lens[i] - y->lens[j]; wrf += fabs(d); sq += d * d; common_sq += d * d; common++; i++; j++; }
        else if (c < 0) { wrf += fabs(x->lens[i]); sq += x->lens[i] * x->lens[i]; ax[na++] = i; i++; }
        else { wrf += fabs(y->lens[j]); sq += y->lens[j] * y->lens[j]; by[nb++] = j; j++; }
    }
    double leaf1 = 0, leaf2 = 0;
    for (int k = 0; k < x->n_taxa; k++) { double d = x->leaflen[k] - y->leaflen[k]; leaf1 += fabs(d); leaf2 += d * d; }
    double d = 0;
    if (metric == ORC_RF) {
        d = (double)(na + nb);
        if (normalise) { int den = nx + ny; d = den ? d / den : 0.0; }
    } else if (metric == ORC_WRF) {
        d = wrf + leaf1;
        if (normalise) {
            double den = 0;
            for (int s = 0; s < nx; s++) den += x->lens[s]; for (int s = 0; s < ny; s++) den += y->lens[s];
            for (int k = 0; k < x->n_taxa; k++) den += x->leaflen[k] + y->leaflen[k];
            d = den != 0 ? d / den : 0.0;
        }
    } else {
        double val;
        if (metric == ORC_EUC) val = sq + leaf2;
        else {
            /* extended common-edge rule (GTP): an edge compatible with every split of the other tree behaves as
             * a common edge whose partner has length 0.  Cannot occur for two binary trees on one leaf set. */
            uint64_t *am = (uint64_t *)malloc((size_t)(na + 1) * W * 8), *bm = (uint64_t *)malloc((size_t)(nb + 1) * W * 8);
            double *al = (double *)malloc(sizeof(double) * (na + 1)), *bl = (double *)malloc(sizeof(double) * (nb + 1));
            int ka = 0, kb = 0; double ext = 0;
            for (int a = 0; a < na; a++) {
                int allc = 1;
                for (int s = 0; s < ny && allc; s++) allc = splits_compatible(&x->masks[(size_t)ax[a] * W], &y->masks[(size_t)s * W], x->leafmask, W, x->rooted);
                if (allc) ext += x->lens[ax[a]] * x->lens[ax[a]];
                else { memcpy(&am[(size_t)ka * W], &x->masks[(size_t)ax[a] * W], W * 8); al[ka++] = x->lens[ax[a]]; }
            }
            for (int b = 0; b < nb; b++) {
                int allc = 1;
                for (int s = 0; s < nx && allc; s++) allc = splits_compatible(&y->masks[(size_t)by[b] * W], &x->masks[(size_t)s * W], x->leafmask, W, x->rooted);
                if (allc) ext += y->lens[by[b]] * y->lens[by[b]];
                else { memcpy(&bm[(size_t)kb * W], &y->masks[(size_t)by[b] * W], W * 8); bl[kb++] = y->lens[by[b]]; }
            }
            double rs = gtp_ratio_term(ka, kb, am, al, bm, bl, x->leafmask, W, x->rooted);
            val = rs + common_sq + ext + leaf2;
            free(am); free(bm); free(al); free(bl);
            __sync_fetch_and_add(&g_stats.pairs, 1);
        }
        d = sqrt(val);
        if (normalise) {
            double n1 = 0, n2 = 0;
            for (int s = 0; s < nx; s++) n1 += x->lens[s] * x->lens[s]; for (int s = 0; s < ny; s++) n2 += y->lens[s] * y->lens[s];
            for (int k = 0; k < x->n_taxa; k++) { n1 += x->leaflen[k] * x->leaflen[k]; n2 += y->leaflen[k] * y->leaflen[k]; }
            double den = sqrt(n1) + sqrt(n2);
            d = den != 0 ? d / den : 0.0;
        }
    }
    free(ax); free(by);
    *out = d;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* Pair entry: _generic_distance_calc (treeCl/treedist.py:30-70) on two parsed trees.                   */
/* ------------------------------------------------------------------------------------------------ */
static int cmp_str(const void *a, const void *b) { return strcmp(*(char *const *)a, *(char *const *)b); }

static int sorted_labels(const tree_t *t, char ***out)
{
    int n = tree_leaves(t, NULL); int *idx = (int *)malloc(sizeof(int) * (n + 1)); tree_leaves(t, idx);
    char **l = (char **)malloc(sizeof(char *) * (n + 1));
    for (int i = 0; i < n; i++) l[i] = t->nd[idx[i]].label;
    qsort(l, n, sizeof(char *), cmp_str); free(idx); *out = l; return n;
}

static tree_t tree_copy(const tree_t *t)
{
    tree_t c = *t; c.nd = (node_t *)malloc(sizeof(node_t) * t->cap); memcpy(c.nd, t->nd, sizeof(node_t) * t->n);
    for (int i = 0; i < t->n; i++) if (t->nd[i].label) c.nd[i].label = strdup(t->nd[i].label);
    return c;
}

static int pair_trees(const tree_t *ta, const tree_t *tb, int metric, int normalise, int min_overlap, double fail, double *out)
{
    char **la, **lb; int na = sorted_labels(ta, &la), nb = sorted_labels(tb, &lb);
    int same = (na == nb);
    for (int i = 0; same && i < na; i++) same = !strcmp(la[i], lb[i]);
    int rc;
    if (same) {
        enc_t ea, eb;
        if ((rc = tree_encode(ta, la, na, &ea))) { free(la); free(lb); return rc; }
        if ((rc = tree_encode(tb, la, na, &eb))) { enc_free(&ea); free(la); free(lb); return rc; }
        rc = enc_distance(&ea, &eb, metric, normalise, out);
        enc_free(&ea); enc_free(&eb); free(la); free(lb); return rc;
    }
    /* t1 ^ t2 non-empty: intersection, min_overlap, prune (treedist.py:63-68) */
    char **common = (char **)malloc(sizeof(char *) * (na + 1)); int nc = 0;
    for (int i = 0, j = 0; i < na && j < nb;) { int c = strcmp(la[i], lb[j]); if (c == 0) { common[nc++] = la[i]; i++; j++; } else if (c < 0) i++; else j++; }
    if (nc < min_overlap) { *out = fail; free(common); free(la); free(lb); return ORC_OK; }
    tree_t pa = tree_copy(ta), pb = tree_copy(tb);
    /* `common` points into ta's labels; the copies own their own strings */
    tree_prune(&pa, common, nc, ta->nd[ta->root].nchild == 2);
    tree_prune(&pb, common, nc, tb->nd[tb->root].nchild == 2);
    enc_t ea, eb;
    rc = tree_encode(&pa, common, nc, &ea);
    if (!rc) { rc = tree_encode(&pb, common, nc, &eb); if (rc) enc_free(&ea); }
    if (!rc) { rc = enc_distance(&ea, &eb, metric, normalise, out); enc_free(&ea); enc_free(&eb); }
    tree_free(&pa); tree_free(&pb); free(common); free(la); free(lb);
    return rc;
}

static int parse_full(const char *nwk, tree_t *t)
{
    int rc = tree_parse(nwk, t); if (rc) return rc;
    tree_strip_outer(t);
    if (tree_leaves(t, NULL) < 2) { tree_free(t); FAIL(ORC_EPARSE, "newick: fewer than two leaves"); }
    return ORC_OK;
}

/* reference-shaped pair: two Newick strings in, one float out (treeCl/tasks.py:41-75) */
int orc_pair(const char *nwk_a, const char *nwk_b, int metric, int normalise, int min_overlap, double fail, double *out)
{
    if (metric < 0 || metric > 3) FAIL(ORC_EARG, "unknown metric %d", metric);
    tree_t ta, tb; int rc;
    if ((rc = parse_full(nwk_a, &ta))) return rc;
    if ((rc = parse_full(nwk_b, &tb))) { tree_free(&ta); return rc; }
    rc = pair_trees(&ta, &tb, metric, normalise, min_overlap, fail, out);
    tree_free(&ta); tree_free(&tb);
    return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* Kendall-Colijn metric.  Restates treeCl/utils/kendallcolijn.py:                          */
/*   _precompute (:38-51)   every internal node: number of edges and path length from the root (seed node)  */
/*   _get_vectors (:53-72)  for every pair of leaves, in sorted label order (itertools.combinations), the    */
/*                          two depths of their MRCA; then per leaf the constants 1 and its edge length      */
/*   get_vector (:74-79)    v = (1 - lambda) m + lambda M                                                    */
/*   get_distance (:81-95)  Euclidean distance of the two v; differing leaf sets: fewer than min_overlap      */
/*                          common leaves -> 0, else both trees pruned to the intersection first              */
/* PARITY UNPINNED: the reference has no test or golden value for this metric and its dendropy dependency   */
/* is not installed here; this follows the published definition (Kendall & Colijn 2016) and the code above.  */
/* ------------------------------------------------------------------------------------------------ */
static int kc_vector(const tree_t *t, char *const *taxa, int ntaxa, double lambda, double *out)
{
    int *leaf = (int *)malloc(sizeof(int) * (ntaxa + 1));
    int *stamp = (int *)calloc((size_t)t->n + 1, sizeof(int));
    int *de = (int *)calloc((size_t)t->n + 1, sizeof(int));
    double *dd = (double *)calloc((size_t)t->n + 1, sizeof(double));
    for (int k = 0; k < ntaxa; k++) leaf[k] = -1;
    for (int i = 0; i < t->n; i++)
        if (t->nd[i].alive && t->nd[i].nchild == 0 && t->nd[i].label) {
            int k = taxon_index(t->nd[i].label, taxa, ntaxa);
            if (k >= 0) leaf[k] = i;
        }
    int rc = ORC_OK;
    for (int k = 0; k < ntaxa && !rc; k++) if (leaf[k] < 0) { snprintf(g_err, sizeof g_err, "kc: leaf '%s' missing", taxa[k]); rc = ORC_EARG; }
    /* depths by climbing to the root */
    for (int i = 0; i < t->n && !rc; i++) {
        if (!t->nd[i].alive) continue;
        int e = 0; double d = 0.0;
        for (int v = i; v != t->root && v >= 0; v = t->nd[v].parent) { e++; d += t->nd[v].len; }
        de[i] = e; dd[i] = d;
    }
    long at = 0;
    for (int a = 0; a < ntaxa && !rc; a++) {
        for (int v = leaf[a]; v >= 0; v = t->nd[v].parent) stamp[v] = a + 1;       /* ancestors of a (incl. itself) */
        for (int b = a + 1; b < ntaxa; b++) {
            int v = leaf[b];
            while (v >= 0 && stamp[v] != a + 1) v = t->nd[v].parent;
            if (v < 0) { snprintf(g_err, sizeof g_err, "kc: no common ancestor"); rc = ORC_EARG; break; }
            out[at++] = (1.0 - lambda) * (double)de[v] + lambda * dd[v];
        }
    }
    for (int k = 0; k < ntaxa && !rc; k++) out[at++] = (1.0 - lambda) * 1.0 + lambda * t->nd[leaf[k]].len;
    free(leaf); free(stamp); free(de); free(dd);
    return rc;
}

static int kc_pair_trees(const tree_t *ta, const tree_t *tb, double lambda, int min_overlap, double *out)
{
    char **la, **lb; int na = sorted_labels(ta, &la), nb = sorted_labels(tb, &lb);
    int same = (na == nb);
    for (int i = 0; same && i < na; i++) same = !strcmp(la[i], lb[i]);
    char **taxa = la; int nt = na, rc = ORC_OK;
    tree_t pa, pb; const tree_t *xa = ta, *xb = tb; int pruned = 0;
    char **common = NULL;
    if (!same) {
        common = (char **)malloc(sizeof(char *) * (na + 1)); int nc = 0;
        for (int i = 0, j = 0; i < na && j < nb;) { int c = strcmp(la[i], lb[j]); if (c == 0) { common[nc++] = la[i]; i++; j++; } else if (c < 0) i++; else j++; }
        if (nc < min_overlap) { *out = 0.0; free(common); free(la); free(lb); return ORC_OK; }      /* kendallcolijn.py:84-85 */
        pa = tree_copy(ta); pb = tree_copy(tb); pruned = 1;
        tree_prune(&pa, common, nc, ta->nd[ta->root].nchild == 2);
        tree_prune(&pb, common, nc, tb->nd[tb->root].nchild == 2);
        xa = &pa; xb = &pb; taxa = common; nt = nc;
    }
    const long D = (long)nt * (nt - 1) / 2 + nt;
    double *va = (double *)malloc(sizeof(double) * (D + 1)), *vb = (double *)malloc(sizeof(double) * (D + 1));
    rc = kc_vector(xa, taxa, nt, lambda, va);
    if (!rc) rc = kc_vector(xb, taxa, nt, lambda, vb);
    if (!rc) { double s = 0.0; for (long k = 0; k < D; k++) { const double d = va[k] - vb[k]; s += d * d; } *out = sqrt(s); }
    free(va); free(vb);
    if (pruned) { tree_free(&pa); tree_free(&pb); }
    free(common); free(la); free(lb);
    return rc;
}

int orc_kc_pair(const char *nwk_a, const char *nwk_b, double lambda, int min_overlap, double *out)
{
    tree_t ta, tb; int rc;
    if ((rc = parse_full(nwk_a, &ta))) return rc;
    if ((rc = parse_full(nwk_b, &tb))) { tree_free(&ta); return rc; }
    rc = kc_pair_trees(&ta, &tb, lambda, min_overlap, out);
    tree_free(&ta); tree_free(&tb);
    return rc;
}

/* the vector of one tree (length n(n-1)/2 + n over its own sorted labels) */
int orc_kc_vector(const char *nwk, double lambda, double *out, long cap, long *need)
{
    tree_t t; int rc;
    if ((rc = parse_full(nwk, &t))) return rc;
    char **l; int n = sorted_labels(&t, &l);
    const long D = (long)n * (n - 1) / 2 + n;
    if (need) *need = D;
    if (out && cap >= D) rc = kc_vector(&t, l, n, lambda, out);
    free(l); tree_free(&t);
    return rc;
}

/* condensed all-pairs vector (scipy order), trees parsed once */
int orc_kc_matrix(const char *const *nwk, int n, double lambda, int min_overlap, double *out)
{
    tree_t *ts = (tree_t *)calloc((size_t)n, sizeof(tree_t)); int rc = ORC_OK, parsed = 0;
    for (; parsed < n && !rc; parsed++) rc = parse_full(nwk[parsed], &ts[parsed]);
    if (rc) parsed--;
    long at = 0;
    for (int i = 0; i < n && !rc; i++)
        for (int j = i + 1; j < n && !rc; j++) rc = kc_pair_trees(&ts[i], &ts[j], lambda, min_overlap, &out[at++]);
    for (int i = 0; i < parsed; i++) tree_free(&ts[i]);
    free(ts);
    return rc;
}

/* prune one tree and write it back as Newick-free encoding summary: used by tests to pin prune semantics */
int orc_encode_summary(const char *nwk, int *n_leaves, int *n_splits, int *rooted, double *sum_internal, double *sum_leaf)
{
    tree_t t; int rc; if ((rc = parse_full(nwk, &t))) return rc;
    char **l; int n = sorted_labels(&t, &l); enc_t e;
    rc = tree_encode(&t, l, n, &e);
    if (!rc) {
        *n_leaves = e.n_leaves; *n_splits = e.n_splits; *rooted = e.rooted; *sum_internal = 0; *sum_leaf = 0;
        for (int s = 0; s < e.n_splits; s++) *sum_internal += e.lens[s];
        for (int k = 0; k < n; k++) *sum_leaf += e.leaflen[k];
        enc_free(&e);
    }
    free(l); tree_free(&t); return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* All pairs, scipy-condensed order.  reparse_per_pair=1 is "reference-shaped" (every pair parses both    */
/* strings, as rfdist_task does); 0 parses/encodes each tree once ("best CPU").  Threads pull rows.        */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    const char *const *nwk; int n, metric, normalise, min_overlap, reparse; double fail; double *out;
    tree_t *trees; enc_t *encs; int all_same; long next_row; int rc; char err[256];
    long pair_begin, pair_end;   /* condensed index range to fill (bounded samples) */
} job_t;

static long row_start(long i, long n) { return i * n - i * (i + 1) / 2; }   /* condensed index of (i, i+1) */

/* row of condensed index p: the largest i with row_start(i) <= p */
static long row_of(long p, long n)
{
    double b = 2.0 * (double)n - 1.0;
    long i = (long)floor((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
    if (i < 0) i = 0;
    while (i > 0 && row_start(i, n) > p) i--;
    while (i + 1 < n - 1 && row_start(i + 1, n) <= p) i++;
    return i;
}

/* threads pull chunks of 256 consecutive pairs of [pair_begin, pair_end): a bounded sample that lies in one or two long rows still
 * keeps every thread busy */
static void *worker(void *arg)
{
    job_t *J = (job_t *)arg;
    const long chunk = 256;
    for (;;) {
        long c = __sync_fetch_and_add(&J->next_row, 1);
        long p = J->pair_begin + c * chunk, p_end = p + chunk < J->pair_end ? p + chunk : J->pair_end;
        if (p >= J->pair_end || J->rc) break;
        long i = row_of(p, J->n), j = i + 1 + (p - row_start(i, J->n));
        for (; p < p_end; p++) {
            int rc; double d = 0;
            if (J->reparse) rc = orc_pair(J->nwk[i], J->nwk[j], J->metric, J->normalise, J->min_overlap, J->fail, &d);
            else if (J->all_same) rc = enc_distance(&J->encs[i], &J->encs[j], J->metric, J->normalise, &d);
            else rc = pair_trees(&J->trees[i], &J->trees[j], J->metric, J->normalise, J->min_overlap, J->fail, &d);
            if (rc) { if (!__sync_lock_test_and_set(&J->rc, rc)) memcpy(J->err, g_err, sizeof J->err); break; }
            J->out[p - J->pair_begin] = d;
            if (++j >= J->n) { i++; j = i + 1; }
        }
    }
    return NULL;
}

int orc_matrix_range(const char *const *nwk, int n, int metric, int normalise, int min_overlap, double fail,
                     int nthreads, int reparse_per_pair, long pair_begin, long pair_end, double *out)
{
    if (metric < 0 || metric > 3) FAIL(ORC_EARG, "unknown metric %d", metric);
    job_t J; memset(&J, 0, sizeof J);
    J.nwk = nwk; J.n = n; J.metric = metric; J.normalise = normalise; J.min_overlap = min_overlap; J.fail = fail;
    J.out = out; J.reparse = reparse_per_pair; J.pair_begin = pair_begin; J.pair_end = pair_end;
    int rc = ORC_OK;
    if (!reparse_per_pair) {
        J.trees = (tree_t *)calloc(n, sizeof(tree_t));
        for (int i = 0; i < n; i++) if ((rc = parse_full(nwk[i], &J.trees[i]))) { for (int k = 0; k < i; k++) tree_free(&J.trees[k]); free(J.trees); return rc; }
        /* identical leaf sets everywhere? then encode once over the shared taxon table */
        char **l0; int n0 = sorted_labels(&J.trees[0], &l0); J.all_same = 1;
        for (int i = 1; i < n && J.all_same; i++) {
            char **li; int ni = sorted_labels(&J.trees[i], &li);
            if (ni != n0) J.all_same = 0;
            for (int k = 0; J.all_same && k < n0; k++) if (strcmp(l0[k], li[k])) J.all_same = 0;
            free(li);
        }
        if (J.all_same) {
            J.encs = (enc_t *)calloc(n, sizeof(enc_t));
            for (int i = 0; i < n && !rc; i++) rc = tree_encode(&J.trees[i], l0, n0, &J.encs[i]);
        }
        free(l0);
    }
    if (!rc) {
        if (nthreads < 1) nthreads = 1;
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
        for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &J);
        worker(&J);
        for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
        free(th);
        if (J.rc) { rc = J.rc; memcpy(g_err, J.err, sizeof g_err); }
    }
    if (J.encs) { for (int i = 0; i < n; i++) enc_free(&J.encs[i]); free(J.encs); }
    if (J.trees) { for (int i = 0; i < n; i++) tree_free(&J.trees[i]); free(J.trees); }
    return rc;
}

/* Selected rows of the SQUARE matrix: out[k * n + j] = d(tree rows[k], tree j) for every j (0 on the diagonal), trees parsed once
 * and in parallel.  For spot checks of matrices too large for orc_matrix (random rows across the whole triangle). */
typedef struct { job_t *J; const long *rows; int n_rows; long next; int phase; } rows_job_t;

static void *rows_worker(void *arg)
{
    rows_job_t *R = (rows_job_t *)arg; job_t *J = R->J;
    for (;;) {
        long w = __sync_fetch_and_add(&R->next, 1);
        if (J->rc) break;
        if (R->phase == 0) {                               /* parse tree w */
            if (w >= J->n) break;
            int rc = parse_full(J->nwk[w], &J->trees[w]);
            if (rc) { if (!__sync_lock_test_and_set(&J->rc, rc)) memcpy(J->err, g_err, sizeof J->err); break; }
        } else if (R->phase == 1) {                        /* encode tree w over the shared taxon table (kept in J->out as a pointer) */
            if (w >= J->n) break;
            char **l0; int n0 = sorted_labels(&J->trees[0], &l0);
            int rc = tree_encode(&J->trees[w], l0, n0, &J->encs[w]);
            free(l0);
            if (rc) { if (!__sync_lock_test_and_set(&J->rc, rc)) memcpy(J->err, g_err, sizeof J->err); break; }
        } else {                                           /* 64 columns of one requested row */
            const long chunks = ((long)J->n + 63) / 64;
            if (w >= chunks * R->n_rows) break;
            const long k = w / chunks, j0 = (w % chunks) * 64, i = R->rows[k];
            for (long j = j0; j < j0 + 64 && j < J->n; j++) {
                double d = 0; int rc = ORC_OK;
                if (j != i) {
                    const long a = i < j ? i : j, b = i < j ? j : i;
                    if (J->all_same) rc = enc_distance(&J->encs[a], &J->encs[b], J->metric, J->normalise, &d);
                    else rc = pair_trees(&J->trees[a], &J->trees[b], J->metric, J->normalise, J->min_overlap, J->fail, &d);
                }
                if (rc) { if (!__sync_lock_test_and_set(&J->rc, rc)) memcpy(J->err, g_err, sizeof J->err); break; }
                J->out[k * (long)J->n + j] = d;
            }
        }
    }
    return NULL;
}

static void rows_run(rows_job_t *R, int phase, int nthreads)
{
    R->phase = phase; R->next = 0;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, rows_worker, R);
    rows_worker(R);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
}

int orc_matrix_rows(const char *const *nwk, int n, int metric, int normalise, int min_overlap, double fail, int nthreads,
                    const long *rows, int n_rows, double *out)
{
    if (metric < 0 || metric > 3) FAIL(ORC_EARG, "unknown metric %d", metric);
    for (int k = 0; k < n_rows; k++) if (rows[k] < 0 || rows[k] >= n) FAIL(ORC_EARG, "row %ld outside [0,%d)", rows[k], n);
    if (nthreads < 1) nthreads = 1;
    job_t J; memset(&J, 0, sizeof J);
    J.nwk = nwk; J.n = n; J.metric = metric; J.normalise = normalise; J.min_overlap = min_overlap; J.fail = fail; J.out = out;
    J.trees = (tree_t *)calloc(n, sizeof(tree_t));
    rows_job_t R; memset(&R, 0, sizeof R); R.J = &J; R.rows = rows; R.n_rows = n_rows;
    rows_run(&R, 0, nthreads);
    if (!J.rc) {
        char **l0; int n0 = sorted_labels(&J.trees[0], &l0); J.all_same = 1;
        for (int i = 1; i < n && J.all_same; i++) {
            char **li; int ni = sorted_labels(&J.trees[i], &li);
            if (ni != n0) J.all_same = 0;
            for (int k = 0; J.all_same && k < n0; k++) if (strcmp(l0[k], li[k])) J.all_same = 0;
            free(li);
        }
        free(l0);
        if (J.all_same) { J.encs = (enc_t *)calloc(n, sizeof(enc_t)); rows_run(&R, 1, nthreads); }
    }
    if (!J.rc) rows_run(&R, 2, nthreads);
    int rc = J.rc;
    if (rc) memcpy(g_err, J.err, sizeof g_err);
    if (J.encs) { for (int i = 0; i < n; i++) enc_free(&J.encs[i]); free(J.encs); }
    for (int i = 0; i < n; i++) tree_free(&J.trees[i]);
    free(J.trees);
    return rc;
}

int orc_matrix(const char *const *nwk, int n, int metric, int normalise, int min_overlap, double fail,
               int nthreads, int reparse_per_pair, double *condensed)
{
    return orc_matrix_range(nwk, n, metric, normalise, min_overlap, fail, nthreads, reparse_per_pair,
                            0, (long)n * (n - 1) / 2, condensed);
}

/* prune a tree to a label subset and serialise it again (for the partial-overlap golden, which prunes the
 * fixture trees first: reference tests/test.py:281-284) */
static void write_newick(const tree_t *t, int i, char **buf, size_t *len, size_t *cap)
{
    #define PUT(...) do { if (*len + 64 + 1 > *cap) { *cap = 2 * *cap + 128; *buf = (char *)realloc(*buf, *cap); } \
                          *len += snprintf(*buf + *len, *cap - *len, __VA_ARGS__); } while (0)
    const node_t *x = &t->nd[i];
    if (x->nchild == 0) {
        size_t need = strlen(x->label) + 8; if (*len + need > *cap) { *cap = 2 * *cap + need; *buf = (char *)realloc(*buf, *cap); }
        *len += snprintf(*buf + *len, *cap - *len, "%s", x->label);
    } else {
        PUT("(");
        for (int c = x->first_child; c >= 0; c = t->nd[c].next_sib) { if (c != x->first_child) PUT(","); write_newick(t, c, buf, len, cap); }
        PUT(")");
    }
    if (i != t->root) PUT(":%.17g", x->len);
    #undef PUT
}

int orc_prune_newick(const char *nwk, const char *const *keep, int nkeep, char *out, size_t outcap)
{
    tree_t t; int rc; if ((rc = parse_full(nwk, &t))) return rc;
    char **k = (char **)malloc(sizeof(char *) * (nkeep + 1)); for (int i = 0; i < nkeep; i++) k[i] = (char *)keep[i];
    qsort(k, nkeep, sizeof(char *), cmp_str);
    tree_prune(&t, k, nkeep, t.nd[t.root].nchild == 2);
    char *buf = NULL; size_t len = 0, cap = 0;
    write_newick(&t, t.root, &buf, &len, &cap);
    if (len + 2 > outcap) { free(buf); free(k); tree_free(&t); FAIL(ORC_EARG, "output buffer too small"); }
    memcpy(out, buf, len); out[len] = ';'; out[len + 1] = 0;
    free(buf); free(k); tree_free(&t); return ORC_OK;
}
