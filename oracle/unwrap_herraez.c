/*
 * TEST INFRASTRUCTURE ONLY (part of oracle/): 2-D phase unwrapping by reliability
 * sorting along a non-continuous path.
 *
 * Restatement of the published algorithm of
 *   M. A. Herraez, D. R. Burton, M. J. Lalor, M. A. Gdeisat, "Fast two-dimensional
 *   phase-unwrapping algorithm based on sorting by reliability following a
 *   noncontinuous path", Applied Optics 41(35), 7437-7444 (2002),
 * which is what skimage.restoration.unwrap_phase (called at
 * /root/reference/pyfcd/fcd.py:119) implements.  scikit-image is neither vendored in the
 * reference nor installed here and the reference pins no version: parity unpinned for
 * this boundary; only the published algorithm is followed.
 *
 *  1. reliability of an interior pixel = sum of squares of the four wrapped second
 *     differences (horizontal, vertical, two diagonals); smaller = more reliable;
 *     border pixels get a very large value (least reliable).
 *  2. every horizontal and vertical neighbour pair is an edge whose cost is the sum of
 *     the two pixel reliabilities; edges are processed in ascending cost.
 *  3. pixels belong to groups; joining two groups shifts the smaller one by the integer
 *     number of 2*pi that removes the jump across the edge.
 *  4. output = input + 2*pi*increment.
 *
 * Build: gcc -O2 -shared -fPIC -o libunwrap_herraez.so unwrap_herraez.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static double wrap_pi(double d) {
    if (d > M_PI) return d - 2.0 * M_PI;
    if (d < -M_PI) return d + 2.0 * M_PI;
    return d;
}

/* number of 2*pi to add to b so that it is within pi of a */
static int jump_between(double a, double b) {
    double d = a - b;
    if (d > M_PI) return -1;
    if (d < -M_PI) return 1;
    return 0;
}

typedef struct {
    double cost;
    int32_t p, q; /* linear pixel indices */
    int32_t jump; /* jump_between(value[p], value[q]) */
    int32_t order;
} edge_t;

static int edge_cmp(const void *a, const void *b) {
    const edge_t *x = (const edge_t *)a, *y = (const edge_t *)b;
    if (x->cost < y->cost) return -1;
    if (x->cost > y->cost) return 1;
    return (x->order > y->order) - (x->order < y->order);
}

int unwrap_herraez_2d(const double *in, double *out, int rows, int cols) {
    const int64_t n = (int64_t)rows * cols;
    if (rows < 1 || cols < 1) return 1;
    double *rel = (double *)malloc(sizeof(double) * n);
    int32_t *inc = (int32_t *)calloc(n, sizeof(int32_t));
    int32_t *head = (int32_t *)malloc(sizeof(int32_t) * n);  /* group id = index of head pixel */
    int32_t *next = (int32_t *)malloc(sizeof(int32_t) * n);  /* linked list within group */
    int32_t *last = (int32_t *)malloc(sizeof(int32_t) * n);  /* valid at head */
    int32_t *count = (int32_t *)malloc(sizeof(int32_t) * n); /* valid at head */
    const int64_t n_edges = (int64_t)rows * (cols - 1) + (int64_t)(rows - 1) * cols;
    edge_t *edges = (edge_t *)malloc(sizeof(edge_t) * (n_edges > 0 ? n_edges : 1));
    if (!rel || !inc || !head || !next || !last || !count || !edges) return 2;

    /* deterministic filler for the border reliabilities */
    uint64_t lcg = 0x9E3779B97F4A7C15ull;
    for (int64_t i = 0; i < n; ++i) {
        lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
        rel[i] = 9999999.0 + (double)(lcg >> 11) / 9007199254740992.0;
        head[i] = (int32_t)i;
        next[i] = -1;
        last[i] = (int32_t)i;
        count[i] = 1;
    }
    for (int r = 1; r < rows - 1; ++r) {
        for (int c = 1; c < cols - 1; ++c) {
            const double *w = in + (int64_t)r * cols + c;
            double h = wrap_pi(w[-1] - w[0]) - wrap_pi(w[0] - w[1]);
            double v = wrap_pi(w[-cols] - w[0]) - wrap_pi(w[0] - w[cols]);
            double d1 = wrap_pi(w[-cols - 1] - w[0]) - wrap_pi(w[0] - w[cols + 1]);
            double d2 = wrap_pi(w[-cols + 1] - w[0]) - wrap_pi(w[0] - w[cols - 1]);
            rel[(int64_t)r * cols + c] = h * h + v * v + d1 * d1 + d2 * d2;
        }
    }
    int64_t e = 0;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c + 1 < cols; ++c) {
            int32_t p = (int32_t)((int64_t)r * cols + c), q = p + 1;
            edges[e].cost = rel[p] + rel[q];
            edges[e].p = p; edges[e].q = q;
            edges[e].jump = jump_between(in[p], in[q]);
            edges[e].order = (int32_t)e;
            ++e;
        }
    for (int r = 0; r + 1 < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            int32_t p = (int32_t)((int64_t)r * cols + c), q = p + cols;
            edges[e].cost = rel[p] + rel[q];
            edges[e].p = p; edges[e].q = q;
            edges[e].jump = jump_between(in[p], in[q]);
            edges[e].order = (int32_t)e;
            ++e;
        }
    qsort(edges, (size_t)n_edges, sizeof(edge_t), edge_cmp);

    for (int64_t i = 0; i < n_edges; ++i) {
        int32_t p = edges[i].p, q = edges[i].q;
        int32_t hp = head[p], hq = head[q];
        if (hp == hq) continue;
        /* value[p] + 2pi*inc[p] must be continuous with value[q] + 2pi*inc[q]:
           inc[q] should equal inc[p] - jump(p,q) where jump = jump_between(v[p], v[q])
           is the number of 2pi to add to v[q]... sign: d = v[p]-v[q] > pi  => jump=-1,
           i.e. q must be raised: inc[q] = inc[p] - jump. */
        int32_t want_q = inc[p] - edges[i].jump;
        if (count[hp] >= count[hq]) {
            int32_t delta = want_q - inc[q];
            for (int32_t k = hq; k != -1; k = next[k]) { head[k] = hp; inc[k] += delta; }
            next[last[hp]] = hq;
            last[hp] = last[hq];
            count[hp] += count[hq];
        } else {
            int32_t want_p = inc[q] + edges[i].jump;
            int32_t delta = want_p - inc[p];
            for (int32_t k = hp; k != -1; k = next[k]) { head[k] = hq; inc[k] += delta; }
            next[last[hq]] = hp;
            last[hq] = last[hp];
            count[hq] += count[hp];
        }
    }
    for (int64_t i = 0; i < n; ++i) out[i] = in[i] + 2.0 * M_PI * (double)inc[i];

    free(rel); free(inc); free(head); free(next); free(last); free(count); free(edges);
    return 0;
}
