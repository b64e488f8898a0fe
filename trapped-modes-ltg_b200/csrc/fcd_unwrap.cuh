// fcd_unwrap.cuh -- reliability-guided 2-D phase unwrapping on the device.
//
// Reference: skimage.restoration.unwrap_phase called at pyfcd/fcd.py:119, i.e. Herraez, Burton,
// Lalor, Gdeisat, Appl. Opt. 41, 7437 (2002): pixel reliability from wrapped second
// differences, edges (horizontal + vertical neighbour pairs) sorted by the sum of the two
// reliabilities, groups merged in that order, each merge shifting one group by the integer
// number of 2*pi that removes the jump across the edge.
//
// Merging in sorted order and skipping edges inside a group is Kruskal's algorithm: the
// result is "integrate the wrapped differences along the minimum spanning tree of the
// reliability-weighted grid graph", up to one global 2*pi*k (verified pixel for pixel against
// oracle/unwrap_herraez.c on the reference's own example pair, 337 and 250 residues).  The MST
// under a strict total order (weight, then edge index: horizontal edges before vertical ones,
// raster order -- the order the sequential algorithm creates them in) is unique, so Boruvka's
// parallel algorithm builds the same tree.  The union-find carries integer potentials:
// word = (offset to parent) << 32 | parent, unwrapped[p] = wrapped[p] + 2*pi*pot(p).
#pragma once
#include "fcd_mask.cuh"

namespace fcd {

constexpr double kPiD = 3.14159265358979323846;

FCD_HD double wrap_pi_d(double d) {
    if (d > kPiD) return d - 2.0 * kPiD;
    if (d < -kPiD) return d + 2.0 * kPiD;
    return d;
}
// number of 2*pi to add to b so that it is within pi of a (sign convention of the oracle)
FCD_HD int jump_between(double a, double b) {
    const double d = a - b;
    if (d > kPiD) return -1;
    if (d < -kPiD) return 1;
    return 0;
}

FCD_HD void atomic_min_u64(unsigned long long* a, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicMin(a, v);
#else
    if (v < *a) *a = v;
#endif
}
FCD_HD void atomic_min_u32(unsigned* a, unsigned v) {
#if defined(__CUDA_ARCH__)
    atomicMin(a, v);
#else
    if (v < *a) *a = v;
#endif
}

using po_t = unsigned long long;
FCD_HD po_t po_pack(int parent, int off) { return ((po_t)(unsigned)off << 32) | (po_t)(unsigned)parent; }
FCD_HD int po_parent(po_t v) { return (int)(unsigned)(v & 0xffffffffull); }
FCD_HD int po_off(po_t v) { return (int)(unsigned)(v >> 32); }

FCD_HD po_t po_load(const po_t* p) {
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const volatile po_t*>(p);
#else
    return *p;
#endif
}
// root of x and the potential of x relative to that root
FCD_HD void pot_find(const po_t* PO, int x, int& root, int& pot) {
    int acc = 0;
    for (;;) {
        const po_t v = po_load(PO + x);
        const int par = po_parent(v);
        if (par == x) { root = x; pot = acc; return; }
        acc += po_off(v);
        x = par;
    }
}
// The same walk with path halving: a non-root node is re-pointed at its grandparent with the summed offset.
// Links are never removed and pot(x) - pot(ancestor) never changes, so any (ancestor, offset) pair read at
// any time stays valid: concurrent 64-bit stores of such pairs are benign.  Keeps the chains that the
// unions of one round build under each other short.
FCD_HD void po_store(po_t* p, po_t v) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<volatile po_t*>(p) = v;
#else
    *p = v;
#endif
}
FCD_HD void pot_find_halving(po_t* PO, int x, int& root, int& pot) {
    int acc = 0;
    for (;;) {
        const po_t v = po_load(PO + x);
        const int par = po_parent(v);
        if (par == x) { root = x; pot = acc; return; }
        const po_t g = po_load(PO + par);
        const int gp = po_parent(g);
        if (gp != par) {
            po_store(PO + x, po_pack(gp, po_off(v) + po_off(g)));
            acc += po_off(v) + po_off(g);
            x = gp;
        } else {
            acc += po_off(v);
            x = par;
        }
    }
}
// edge e of a map with n = H*W pixels: e < n horizontal (p, p+1), e >= n vertical (p, p+W)
FCD_HD bool edge_ends(int e, int n, int H, int W, int& p, int& q) {
    if (e < n) { p = e; q = e + 1; return (e % W) != W - 1; }
    p = e - n; q = p + W; return p / W != H - 1;
}

// ---- reliability (double) ---------------------------------------------------------------------
struct MstRelParams {
    const float* w;            // [maps][n] wrapped phases
    const double* border;      // [2W + 2H] reliability of border pixels (large + deterministic filler):
                               // top row, bottom row, left column, right column
    double* rel;               // [maps][n]
    po_t* PO;                  // [maps][n] initialised to (self, 0)
    long long total;
    int H, W;
};
struct MstReliability : ElemBase {
    using Params = MstRelParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.total) return;
        const int n = p.H * p.W;
        const int px = (int)(i % n);
        const int r = px / p.W, c = px % p.W;
        p.PO[i] = po_pack(px, 0);
        if (r == 0 || c == 0 || r == p.H - 1 || c == p.W - 1) {
            const int b = r == 0 ? c : (r == p.H - 1 ? p.W + c : (c == 0 ? 2 * p.W + r : 2 * p.W + p.H + r));
            p.rel[i] = p.border[b];
            return;
        }
        const float* w = p.w + i;
        const int W = p.W;
        const double c0 = (double)w[0];
        const double h = wrap_pi_d((double)w[-1] - c0) - wrap_pi_d(c0 - (double)w[1]);
        const double v = wrap_pi_d((double)w[-W] - c0) - wrap_pi_d(c0 - (double)w[W]);
        const double d1 = wrap_pi_d((double)w[-W - 1] - c0) - wrap_pi_d(c0 - (double)w[W + 1]);
        const double d2 = wrap_pi_d((double)w[-W + 1] - c0) - wrap_pi_d(c0 - (double)w[W - 1]);
        p.rel[i] = h * h + v * v + d1 * d1 + d2 * d2;
    }
};

// ---- Boruvka rounds over a shrinking list of cross edges -------------------------------------------------
// Round 0 is local (MstRound0: every pixel is a component and picks its minimum incident edge).  Every later
// round works on the list of edges whose end points still lie in different components.  A list entry carries what
// the round needs, so that the selection passes stream through the list without chasing a single pointer:
//     key  bit pattern of rel[u] + rel[v] (a double >= 0: unsigned order == numeric order)
//     id   global edge id  = map * 2n + e
//     ru, rv   the ROOTS of the two end points, as global pixel indices (map * n + pixel)
// Between the compaction that wrote the entry and the unions of the next round the union-find does not change, so
// the cached roots are exact; the compaction after the unions re-finds them starting from the old roots (a chain as
// long as the unions stacked in that one round) and drops the edges that have become internal.
struct EdgeList {
    unsigned long long* key;
    unsigned* id;
    unsigned* ru;
    unsigned* rv;
};
struct MstRoundParams {
    const float* w;
    const double* rel;
    po_t* PO;                     // [maps][n]; parents are pixel indices WITHIN the map
    unsigned long long* best_w;   // [maps][n] per root: smallest outgoing weight
    unsigned* best_e;             // [maps][n] per root: smallest edge id among those of that weight
    EdgeList in, out;             // this round's list / the list being written (MstBuild, MstCompact)
    unsigned* counters;           // [1] length of `out`, [3] length of `in`
    unsigned char* isroot0;       // [maps][n] 1 for the roots of the round-0 forest (the only nodes later unions move)
    long long count;              // upper bound of the items of this launch (the exact list lengths are counters)
    int H, W;
    float* outp;                  // MstApply
};
FCD_HD unsigned atomic_add_u32(unsigned* a, unsigned v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(a, v);
#else
    const unsigned o = *a;
    *a += v;
    return o;
#endif
}
// block-aggregated append of up to two items per thread: one global atomic per thread block
struct AppendState { unsigned n; unsigned slot; };
FCD_HD void block_append_reserve(int ph, int tid, unsigned char* smem, AppendState& st, unsigned* counter) {
    unsigned* sc = reinterpret_cast<unsigned*>(smem);        // [0] block count, [1] block base
    if (ph == 0) { if (tid == 0) sc[0] = 0; }
    else if (ph == 1) { if (st.n) st.slot = atomic_add_u32(sc, st.n); }
    else if (ph == 2) { if (tid == 0) sc[1] = atomic_add_u32(counter, sc[0]); }
    else { st.slot += sc[1]; }
}
struct MstListBase : ElemBase {
    static constexpr int PHASES = 4, SMEM_BYTES = 16;
};
// Round 0: every pixel is its own component, so its minimum outgoing edge is the minimum over its (at most four)
// incident edges under the same (weight, edge index) order.  The picks form a forest once every mutual pick
// (the only possible cycle under a strict order) is rooted at its smaller pixel, so each pixel simply writes
// its own union-find word -- (picked neighbour, integer offset across the edge) -- with no atomics and no
// union at all; MstFlatten then points everything at the roots.
struct MstRound0 : ElemBase {
    using Params = MstRoundParams;     // count = maps * n pixels
    // best incident edge of pixel px: edge index and the pixel at its other end
    FCD_HD static void best_edge(const double* rel, int px, int H, int W, unsigned& be, int& other) {
        const int n = H * W;
        const int r = px / W, c = px - r * W;
        const double r0 = rel[px];
        unsigned long long bw = ~0ull;
        be = ~0u; other = -1;
        auto consider = [&](bool ok, int q, unsigned e) {
            if (!ok) return;
            const unsigned long long key = f64_bits(r0 + rel[q]);     // the sum is commutative: same key from both ends
            if (key < bw || (key == bw && e < be)) { bw = key; be = e; other = q; }
        };
        consider(c > 0, px - 1, (unsigned)(px - 1));                  // horizontal edge (px-1, px) has index px-1
        consider(c + 1 < W, px + 1, (unsigned)px);
        consider(r > 0, px - W, (unsigned)(n + px - W));              // vertical edge (px-W, px) has index n + px-W
        consider(r + 1 < H, px + W, (unsigned)(n + px));
    }
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.count) return;
        const int n = p.H * p.W;
        const long long o = (i / n) * n;
        const int px = (int)(i - o);
        const double* rel = p.rel + o;
        unsigned be, bq; int q, qq;
        best_edge(rel, px, p.H, p.W, be, q);
        if (q < 0) return;                                            // a 1 x 1 map: stays its own root
        best_edge(rel, q, p.H, p.W, bq, qq);
        if (bq == be && px < q) return;                               // mutual pick: the smaller pixel is the root
        // pot(v) = pot(u) - jump(u, v) for the edge (u, v), u < v   (same convention as MstHook)
        const float* w = p.w + o;
        const int off = q < px ? -jump_between((double)w[q], (double)w[px]) : jump_between((double)w[px], (double)w[q]);
        p.PO[o + px] = po_pack(q, off);
    }
};
struct MstFlatten : ElemBase {
    using Params = MstRoundParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.count) return;
        const int n = p.H * p.W;
        po_t* PO = p.PO + (i / n) * n;
        const int px = (int)(i % n);
        int root, pot;
        pot_find(PO, px, root, pot);
        if (root != px) PO[px] = po_pack(root, pot);     // one 64-bit store: concurrent walks stay consistent
    }
};
// After the round-0 forest has been flattened: the first edge list (every edge between two different trees, found by
// comparing each pixel's root with its right and lower neighbour's), the per-root minima cleared, the roots flagged.
struct MstBuild : MstListBase {
    using Params = MstRoundParams;     // count = maps * n pixels
    struct State { AppendState a; unsigned long long key[2]; unsigned id[2], ru[2], rv[2]; };
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem, State& st) {
        if constexpr (PH == 1) {
            st.a.n = 0;
            const long long i = (long long)bx * THREADS + tid;
            if (i < p.count) {
                const int n = p.H * p.W;
                const long long map = i / n;
                const long long o = map * n;
                const int px = (int)(i - o);
                const int r = px / p.W, c = px - r * p.W;
                const int rp = po_parent(p.PO[i]);
                const bool isroot = rp == px;
                p.isroot0[i] = isroot ? 1 : 0;
                if (isroot) { p.best_w[i] = ~0ull; p.best_e[i] = ~0u; }
                const double r0 = p.rel[i];
                auto consider = [&](bool ok, int q, unsigned e) {
                    if (!ok) return;
                    const int rq = po_parent(p.PO[o + q]);
                    if (rq == rp) return;
                    const unsigned k = st.a.n++;
                    st.key[k] = f64_bits(r0 + p.rel[o + q]);
                    st.id[k] = (unsigned)(map * 2LL * n + e);
                    st.ru[k] = (unsigned)(o + rp);
                    st.rv[k] = (unsigned)(o + rq);
                };
                consider(c + 1 < p.W, px + 1, (unsigned)px);
                consider(r + 1 < p.H, px + p.W, (unsigned)(n + px));
            }
        }
        block_append_reserve(PH, tid, smem, st.a, p.counters + 1);
        if constexpr (PH == 3) {
            for (unsigned k = 0; k < st.a.n; ++k) {
                const unsigned at = st.a.slot + k;
                p.out.key[at] = st.key[k]; p.out.id[at] = st.id[k]; p.out.ru[at] = st.ru[k]; p.out.rv[at] = st.rv[k];
            }
        }
    }
};
template <int PASS>   // 0: minimum weight per component, 1: minimum edge id among the minimum-weight edges
struct MstSelect : ElemBase {
    using Params = MstRoundParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= (long long)p.counters[3]) return;
        const unsigned long long key = p.in.key[i];
        const unsigned ru = p.in.ru[i], rv = p.in.rv[i];
        if (PASS == 0) {
            atomic_min_u64(p.best_w + ru, key);
            atomic_min_u64(p.best_w + rv, key);
        } else {
            const unsigned id = p.in.id[i];
            if (p.best_w[ru] == key) atomic_min_u32(p.best_e + ru, id);
            if (p.best_w[rv] == key) atomic_min_u32(p.best_e + rv, id);
        }
    }
};
// Hooking.  Every component selected exactly one edge, and that edge sits in exactly one list entry, so the entry
// that carries a component's pick links that component's root under the root at the other end by writing the root's
// OWN union-find word: one writer per word, no compare-and-swap, no retry.  The picks form a forest once every
// mutual pick (the only cycle a strict order allows) keeps its smaller root, so roots may hook under roots that
// hook elsewhere in the same launch.  Potentials of the end points are taken relative to the roots the entry
// caches -- the walk stops there, because the word above may already have been rewritten by its owner.
FCD_HD int pot_until(const po_t* PO, int x, int root) {
    int acc = 0;
    while (x != root) {
        const po_t v = po_load(PO + x);
        acc += po_off(v);
        x = po_parent(v);
    }
    return acc;
}
struct MstHook : ElemBase {
    using Params = MstRoundParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= (long long)p.counters[3]) return;
        const unsigned id = p.in.id[i], gu = p.in.ru[i], gv = p.in.rv[i];
        const bool cu = p.best_e[gu] == id, cv = p.best_e[gv] == id;
        if (!cu && !cv) return;
        const int n = p.H * p.W;
        const long long map = (long long)id / (2LL * n);
        int u, v;
        edge_ends((int)((long long)id - map * 2LL * n), n, p.H, p.W, u, v);
        const long long o = map * n;
        po_t* PO = p.PO + o;
        const int ru = (int)((long long)gu - o), rv = (int)((long long)gv - o);
        // value[u] + 2pi inc[u] continuous with value[v] + 2pi inc[v]:  inc[v] = inc[u] - jump(u, v)
        const int delta = -jump_between((double)p.w[o + u], (double)p.w[o + v]);
        const int pu = pot_until(PO, u, ru), pv = pot_until(PO, v, rv);
        const bool v_under_u = (cu && cv) ? (ru < rv) : cv;
        if (v_under_u) po_store(PO + rv, po_pack(ru, pu + delta - pv));
        else po_store(PO + ru, po_pack(rv, pv - delta - pu));
    }
};
// root of a node that was a root before this round's unions (roots only ever get linked under other roots)
FCD_HD unsigned root_from(po_t* PO_all, unsigned g, int n) {
    const unsigned o = (g / (unsigned)n) * (unsigned)n;
    int root, pot;
    pot_find_halving(PO_all + o, (int)(g - o), root, pot);
    return o + (unsigned)root;
}
// keep the edges that still cross components, with their roots brought up to date
struct MstCompact : MstListBase {
    using Params = MstRoundParams;
    struct State { AppendState a; unsigned long long key; unsigned id, ru, rv; };
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem, State& st) {
        if constexpr (PH == 1) {
            st.a.n = 0;
            const long long i = (long long)bx * THREADS + tid;
            if (i < (long long)p.counters[3]) {
                const int n = p.H * p.W;
                const unsigned ru = root_from(p.PO, p.in.ru[i], n), rv = root_from(p.PO, p.in.rv[i], n);
                if (ru != rv) {
                    st.a.n = 1;
                    st.key = p.in.key[i]; st.id = p.in.id[i]; st.ru = ru; st.rv = rv;
                    // the two components take part in the next round: clear their minima here (same value
                    // from every thread that touches them)
                    p.best_w[ru] = ~0ull; p.best_e[ru] = ~0u;
                    p.best_w[rv] = ~0ull; p.best_e[rv] = ~0u;
                }
            }
        }
        block_append_reserve(PH, tid, smem, st.a, p.counters + 1);
        if constexpr (PH == 3) {
            if (st.a.n) {
                const unsigned at = st.a.slot;
                p.out.key[at] = st.key; p.out.id[at] = st.id; p.out.ru[at] = st.ru; p.out.rv[at] = st.rv;
            }
        }
    }
};
// End of the rounds.  Only the roots of the round-0 forest were ever linked under other nodes; every other pixel
// still points at such a root (or, after path halving, at another one further up).  Point those roots straight at
// the root of the whole map ...
struct MstFlattenRoots : ElemBase {
    using Params = MstRoundParams;
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.count || !p.isroot0[i]) return;
        const int n = p.H * p.W;
        po_t* PO = p.PO + (i / n) * n;
        const int px = (int)(i % n);
        int root, pot;
        pot_find(PO, px, root, pot);
        if (root != px) PO[px] = po_pack(root, pot);
    }
};
// ... so that every pixel is at most two hops from it.
struct MstApply : ElemBase {
    using Params = MstRoundParams;
    FCD_HD static int pot_of(const po_t* PO, int px) {
        const po_t v = PO[px];
        const int q = po_parent(v);
        if (q == px) return 0;
        const po_t v2 = PO[q];
        return po_off(v) + (po_parent(v2) == q ? 0 : po_off(v2));
    }
    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char*, State&) {
        const long long i = (long long)bx * THREADS + tid;
        if (i >= p.count) return;
        const int n = p.H * p.W;
        const po_t* PO = p.PO + (i / n) * n;
        // normalisation: pixel (0, 0) keeps its wrapped value
        p.outp[i] = (float)((double)p.w[i] + 2.0 * kPiD * (double)(pot_of(PO, (int)(i % n)) - pot_of(PO, 0)));
    }
};

// ---- forward row transform of z = phi0 + i*phi1 read from a materialised phases array ---------------
struct RowPhaseFwdParams {
    const float* phases;   // [F][2][H][W]
    cf* w3;                // column-blocked, see w3_index
    const cf* tw;
    int H;
};
template <int L, int G>
struct RowPhaseFwd : AllPhases {
    using FF = Fft<L, -1, float, RowPlan<L>>;
    using GL = GroupLayout<L, G>;
    using Params = RowPhaseFwdParams;
    static constexpr bool BLOCKED_TILES = false;
    static constexpr bool PIPELINED = false;
    static constexpr int SYNC_THREADS = 0;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int TPF = GL::TPF, THREADS = GL::THREADS, PHASES = 4;
    using TW = SmemTwiddles<FF, THREADS>;
    static constexpr int SMEM_BYTES = TW::TW_BYTES + G * GL::STRIDE * (int)sizeof(cf);
    FCD_HD static void prologue(const Params& p, int tid, unsigned char* smem) { TW::load(p.tw, tid, smem); }
    struct State { cf v[16]; };

    template <int PH>
    FCD_HD static void phase(const Params& p, int bx, int by, int tid, unsigned char* smem_all, State& st) {
        const cf* tw = reinterpret_cast<const cf*>(smem_all);
        const int g = tid / TPF, t = tid % TPF;
        cf* s = reinterpret_cast<cf*>(smem_all + TW::TW_BYTES) + g * GL::STRIDE;
        const int W = L;
        const int y = bx * G + g, f = by;
        if constexpr (PH == 0) {
            const float* p0 = p.phases + (((long long)f * 2 + 0) * p.H + y) * W;
            const float* p1 = p.phases + (((long long)f * 2 + 1) * p.H + y) * W;
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) st.v[m] = mk<float>(p0[t + TPF * m], p1[t + TPF * m]);
            FF::stepA(st.v, t, s);
        } else if constexpr (PH == 1) {
            FF::stepB(st.v, t, s, tw);
        } else if constexpr (PH == 2) {
            FF::stepC(st.v, t, s);
        } else {
            FF::stepD(st.v, t, s, tw);
            FCD_UNROLL
            for (int m = 0; m < 16; ++m) p.w3[w3_index(f, t + TPF * m, y, p.H, W)] = st.v[m];
        }
    }
};

}  // namespace fcd
