// rtb_bvh.cpp — host-side acceleration-structure builders (new functionality; the reference has none).
//
//   buildBinary  multi-threaded binned-SAH binary tree over padded triangle boxes (temporary form)
//   buildBvh     -> 64-byte two-box nodes (BvhNode), breadth-first prefix then depth-first
//   buildCwbvh   -> 80-byte 8-wide compressed nodes (Node8) after Ylitie, Karras, Laine, "Efficient Incoherent Ray
//                   Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017): greedy surface-area collapse of the
//                   binary tree, octant-ordered child slots, 8-bit child boxes on a per-node power-of-two grid
#include "rtb_bvh.h"
#include "rtb_node8_encode.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <thread>

namespace rtb {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -std::numeric_limits<float>::infinity(); } }
    void grow(const Box& b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(const float* p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    float area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0.0f && dy >= 0.0f && dz >= 0.0f)) return 0.0f;
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};

struct Prim { Box b; float c[3]; };

// temporary binary node: links >= 0 index the owning vector, < 0 are leaf codes ~((first << 3) | (count - 1))
struct TNode { Box b0, b1; int32_t c0, c1; };

struct BinaryTree {
    std::vector<TNode> all;        // inner nodes
    int32_t root = 0;              // link of the root (a leaf code when the whole scene is one leaf)
    std::vector<uint32_t> idx;     // triangle order: every leaf owns a contiguous range
    Box rootBox;
    uint32_t maxDepth = 0;
};

constexpr int BINS = 16;
constexpr float C_TRAV = 1.0f;

inline int32_t leafLink(uint32_t first, uint32_t count) { return (int32_t)~((first << 3) | (count - 1u)); }
inline uint32_t leafFirst(int32_t link) { return (~(uint32_t)link) >> 3; }
inline uint32_t leafCount(int32_t link) { return ((~(uint32_t)link) & 7u) + 1u; }

struct Split { int axis; int bin; float cost; Box left, right; uint32_t nLeft; float cbLo, scale; };

struct Builder {
    const Prim* prims;
    uint32_t* idx;
    uint32_t maxLeaf;

    Box boundsOf(uint32_t b, uint32_t e) const { Box r; r.reset(); for (uint32_t i = b; i < e; ++i) r.grow(prims[idx[i]].b); return r; }

    // best binned split of [b, e); returns false when every centroid coincides
    bool findSplit(uint32_t b, uint32_t e, Split& best) const {
        Box cb; cb.reset();
        for (uint32_t i = b; i < e; ++i) cb.grow(prims[idx[i]].c);
        Box binBox[3][BINS]; uint32_t binCnt[3][BINS];
        float scale[3];
        bool any = false;
        for (int a = 0; a < 3; ++a) {
            const float ext = cb.hi[a] - cb.lo[a];
            scale[a] = ext > 0.0f ? (float)BINS * (1.0f - 1e-6f) / ext : 0.0f;
            any |= ext > 0.0f;
            for (int k = 0; k < BINS; ++k) { binBox[a][k].reset(); binCnt[a][k] = 0; }
        }
        if (!any) return false;
        for (uint32_t i = b; i < e; ++i) {
            const Prim& p = prims[idx[i]];
            for (int a = 0; a < 3; ++a) {
                if (scale[a] == 0.0f) continue;
                int k = (int)((p.c[a] - cb.lo[a]) * scale[a]);
                k = k < 0 ? 0 : (k >= BINS ? BINS - 1 : k);
                binBox[a][k].grow(p.b); binCnt[a][k]++;
            }
        }
        best.cost = std::numeric_limits<float>::infinity();
        for (int a = 0; a < 3; ++a) {
            if (scale[a] == 0.0f) continue;
            float rightArea[BINS]; uint32_t rightCnt[BINS]; Box rightBox[BINS];
            Box acc; acc.reset(); uint32_t cnt = 0;
            for (int k = BINS - 1; k > 0; --k) { acc.grow(binBox[a][k]); cnt += binCnt[a][k]; rightArea[k] = acc.area(); rightCnt[k] = cnt; rightBox[k] = acc; }
            acc.reset(); cnt = 0;
            for (int k = 1; k < BINS; ++k) {
                acc.grow(binBox[a][k - 1]); cnt += binCnt[a][k - 1];
                if (cnt == 0 || rightCnt[k] == 0) continue;
                const float cost = acc.area() * (float)cnt + rightArea[k] * (float)rightCnt[k];
                if (cost < best.cost) { best.cost = cost; best.axis = a; best.bin = k; best.left = acc; best.right = rightBox[k]; best.nLeft = cnt; best.cbLo = cb.lo[a]; best.scale = scale[a]; }
            }
        }
        return best.cost < std::numeric_limits<float>::infinity();
    }

    // partitions [b, e) and returns the middle; boxes of the halves in lb, rb
    uint32_t partition(uint32_t b, uint32_t e, uint32_t depth, const Box& box, Box& lb, Box& rb, bool& makeLeaf) const {
        const uint32_t n = e - b;
        makeLeaf = false;
        Split s;
        const bool forceMedian = depth >= 32;   // bounds the depth: 32 SAH levels + log2(n) median levels
        if (!forceMedian && findSplit(b, e, s)) {
            if (n <= maxLeaf) {
                const float leafCost = (float)n * box.area();
                if (leafCost <= C_TRAV * box.area() + s.cost) { makeLeaf = true; return b; }
            }
            const int axis = s.axis, bin = s.bin; const float lo = s.cbLo, sc = s.scale;
            uint32_t* mid = std::partition(idx + b, idx + e, [&](uint32_t id) {
                int k = (int)((prims[id].c[axis] - lo) * sc);
                k = k < 0 ? 0 : (k >= BINS ? BINS - 1 : k);
                return k < bin;
            });
            const uint32_t m = (uint32_t)(mid - idx);
            if (m > b && m < e) { lb = s.left; rb = s.right; return m; }
        }
        if (n <= maxLeaf) { makeLeaf = true; return b; }
        // object median along the widest axis of the box
        int axis = 0;
        for (int a = 1; a < 3; ++a) if (box.hi[a] - box.lo[a] > box.hi[axis] - box.lo[axis]) axis = a;
        const uint32_t m = b + n / 2;
        std::nth_element(idx + b, idx + m, idx + e, [&](uint32_t x, uint32_t y) { return prims[x].c[axis] < prims[y].c[axis]; });
        lb = boundsOf(b, m); rb = boundsOf(m, e);
        return m;
    }

    // recursive build of [b, e) into `out`; returns the link (leaf code or index into out)
    int32_t build(uint32_t b, uint32_t e, uint32_t depth, const Box& box, std::vector<TNode>& out) const {
        const uint32_t n = e - b;
        if (n == 1) return leafLink(b, 1);
        Box lb, rb; bool makeLeaf;
        const uint32_t m = partition(b, e, depth, box, lb, rb, makeLeaf);
        if (makeLeaf) return leafLink(b, n);
        const int32_t self = (int32_t)out.size();
        out.emplace_back();
        const int32_t l = build(b, m, depth + 1, lb, out);
        const int32_t r = build(m, e, depth + 1, rb, out);
        TNode& t = out[(size_t)self];
        t.b0 = lb; t.b1 = rb; t.c0 = l; t.c1 = r;
        return self;
    }
};

template <class F>
void parallelFor(int threads, uint32_t n, uint32_t chunk, F&& body) {
    if (threads <= 1 || n <= chunk) { body(0u, n); return; }
    std::atomic<uint32_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const uint32_t b = next.fetch_add(chunk);
            if (b >= n) break;
            body(b, std::min(n, b + chunk));
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}

void buildBinary(const TriangleRec* tris, uint32_t count, int threads, uint32_t maxLeaf, BinaryTree& out) {
    // ---- primitive boxes, padded ---------------------------------------------------------------------------
    float maxAbs = 0.0f;
    for (uint32_t i = 0; i < count; ++i) {
        const TriangleRec& t = tris[i];
        for (int a = 0; a < 3; ++a) maxAbs = std::max(maxAbs, std::max(std::fabs(t.p0[a]), std::max(std::fabs(t.p1[a]), std::fabs(t.p2[a]))));
    }
    if (!(maxAbs < std::numeric_limits<float>::infinity())) maxAbs = 1.0f;
    const float pad = std::max(maxAbs * 3.814697265625e-6f /* 2^-18 */, 1e-30f);
    std::vector<Prim> prims(count);
    out.idx.resize(count);
    uint32_t* idx = out.idx.data();
    parallelFor(threads, count, 1u << 16, [&](uint32_t b, uint32_t e) {
        for (uint32_t i = b; i < e; ++i) {
            const TriangleRec& t = tris[i];
            Prim& p = prims[i];
            p.b.reset(); p.b.grow(t.p0); p.b.grow(t.p1); p.b.grow(t.p2);
            for (int a = 0; a < 3; ++a) {
                p.c[a] = 0.5f * (p.b.lo[a] + p.b.hi[a]);
                p.b.lo[a] -= pad; p.b.hi[a] += pad;
                if (!(p.b.lo[a] <= p.b.hi[a])) { p.b.lo[a] = -maxAbs - pad; p.b.hi[a] = maxAbs + pad; p.c[a] = 0.0f; }   // NaN vertex: keep it reachable
            }
            idx[i] = i;
        }
    });
    Builder bl{prims.data(), idx, maxLeaf};

    // ---- top of the tree, sequential: split the largest open range until there is enough parallel work -------
    struct Open { uint32_t b, e, depth; Box box; int32_t parent; int side; };
    std::vector<TNode> top;
    std::vector<Open> open;
    out.rootBox = bl.boundsOf(0, count);
    open.push_back({0, count, 0, out.rootBox, -1, 0});
    const size_t wantRanges = threads > 1 ? (size_t)threads * 8 : 1;
    while (open.size() < wantRanges) {
        size_t big = 0;
        for (size_t i = 1; i < open.size(); ++i) if (open[i].e - open[i].b > open[big].e - open[big].b) big = i;
        Open o = open[big];
        if (o.e - o.b <= 4096) break;
        Box lb, rb; bool makeLeaf;
        const uint32_t m = bl.partition(o.b, o.e, o.depth, o.box, lb, rb, makeLeaf);
        if (makeLeaf) break;   // cannot happen for > maxLeaf primitives
        const int32_t self = (int32_t)top.size();
        top.emplace_back();
        top[(size_t)self].b0 = lb; top[(size_t)self].b1 = rb; top[(size_t)self].c0 = 0; top[(size_t)self].c1 = 0;
        if (o.parent >= 0) (o.side ? top[(size_t)o.parent].c1 : top[(size_t)o.parent].c0) = self;
        open[big] = {o.b, m, o.depth + 1, lb, self, 0};
        open.push_back({m, o.e, o.depth + 1, rb, self, 1});
    }

    // ---- subtrees in parallel -----------------------------------------------------------------------------
    std::vector<std::vector<TNode>> sub(open.size());
    std::vector<int32_t> subRoot(open.size());
    {
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= open.size()) break;
                sub[i].reserve((open[i].e - open[i].b) / 2 + 4);
                subRoot[i] = bl.build(open[i].b, open[i].e, open[i].depth, open[i].box, sub[i]);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads && (size_t)t < open.size(); ++t) pool.emplace_back(worker);
        worker();
        for (auto& t : pool) t.join();
    }

    // ---- merge into one temporary array ---------------------------------------------------------------------
    std::vector<TNode>& all = out.all;
    size_t total = top.size();
    for (auto& s : sub) total += s.size();
    all.reserve(total + 1);
    all = top;
    int32_t rootLink = 0;
    for (size_t i = 0; i < open.size(); ++i) {
        const int32_t off = (int32_t)all.size();
        for (TNode t : sub[i]) { if (t.c0 >= 0) t.c0 += off; if (t.c1 >= 0) t.c1 += off; all.push_back(t); }
        const int32_t link = subRoot[i] >= 0 ? subRoot[i] + off : subRoot[i];
        if (open[i].parent >= 0) (open[i].side ? all[(size_t)open[i].parent].c1 : all[(size_t)open[i].parent].c0) = link;
        else rootLink = link;
        std::vector<TNode>().swap(sub[i]);
    }
    if (!top.empty()) rootLink = 0;
    out.root = rootLink;

    // depth
    out.maxDepth = 0;
    if (rootLink >= 0) {
        std::vector<std::pair<int32_t, uint32_t>> st; st.push_back({rootLink, 1});
        while (!st.empty()) {
            auto [t, d] = st.back(); st.pop_back();
            out.maxDepth = std::max(out.maxDepth, d);
            if (all[(size_t)t].c0 >= 0) st.push_back({all[(size_t)t].c0, d + 1});
            if (all[(size_t)t].c1 >= 0) st.push_back({all[(size_t)t].c1, d + 1});
        }
    }
}

void fillTravTris(const TriangleRec* tris, const std::vector<uint32_t>& order, int threads, std::vector<TravTri>& travTris) {
    travTris.resize(order.size());
    parallelFor(threads, (uint32_t)order.size(), 1u << 16, [&](uint32_t b, uint32_t e) {
        for (uint32_t k = b; k < e; ++k) {
            const uint32_t id = order[k];
            const TriangleRec& t = tris[id];
            TravTri& o = travTris[k];
            for (int a = 0; a < 3; ++a) { o.p0[a] = t.p0[a]; o.e1[a] = t.p1[a] - t.p0[a]; o.e2[a] = t.p2[a] - t.p0[a]; }
            o.id = id; o.pad1 = 0; o.pad2 = 0;
        }
    });
}

int resolveThreads(int threads) {
    if (threads <= 0) { threads = (int)std::thread::hardware_concurrency(); if (threads <= 0) threads = 1; }
    return threads;
}

}  // namespace

void buildBvh(const TriangleRec* tris, uint32_t count, uint32_t topNodes, int threads, std::vector<BvhNode>& nodes,
              std::vector<TravTri>& travTris, BvhStats& stats) {
    const auto t0 = std::chrono::steady_clock::now();
    nodes.clear(); travTris.clear();
    stats = BvhStats();
    if (count == 0) return;
    threads = resolveThreads(threads);
    BinaryTree bt;
    buildBinary(tris, count, threads, BVH_MAX_LEAF, bt);
    std::vector<TNode>& all = bt.all;
    int32_t rootLink = bt.root;
    // a scene that is a single leaf still gets an inner root: both children are that leaf (a triangle tested twice is harmless)
    if (rootLink < 0) {
        TNode r;
        r.b0 = bt.rootBox; r.b1 = bt.rootBox;
        r.c0 = rootLink; r.c1 = rootLink;
        all.clear(); all.push_back(r);
        rootLink = 0;
    }

    // ---- final order: breadth-first prefix of `topNodes`, then depth-first ----------------------------------
    const size_t nNodes = all.size();
    std::vector<int32_t> order; order.reserve(nNodes);          // temp index by final position
    std::vector<int32_t> finalOf(nNodes, -1);
    {
        std::deque<int32_t> q; q.push_back(rootLink);
        while (!q.empty() && order.size() < topNodes) {
            const int32_t t = q.front(); q.pop_front();
            finalOf[(size_t)t] = (int32_t)order.size(); order.push_back(t);
            if (all[(size_t)t].c0 >= 0) q.push_back(all[(size_t)t].c0);
            if (all[(size_t)t].c1 >= 0) q.push_back(all[(size_t)t].c1);
        }
        std::vector<int32_t> stack;
        while (!q.empty()) {
            stack.push_back(q.front()); q.pop_front();
            while (!stack.empty()) {
                const int32_t t = stack.back(); stack.pop_back();
                finalOf[(size_t)t] = (int32_t)order.size(); order.push_back(t);
                if (all[(size_t)t].c1 >= 0) stack.push_back(all[(size_t)t].c1);
                if (all[(size_t)t].c0 >= 0) stack.push_back(all[(size_t)t].c0);
            }
        }
    }
    nodes.resize(order.size());
    parallelFor(threads, (uint32_t)order.size(), 1u << 15, [&](uint32_t b, uint32_t e) {
        for (uint32_t f = b; f < e; ++f) {
            const TNode& t = all[(size_t)order[f]];
            BvhNode& n = nodes[f];
            n.c0lox = t.b0.lo[0]; n.c0hix = t.b0.hi[0]; n.c0loy = t.b0.lo[1]; n.c0hiy = t.b0.hi[1];
            n.c1lox = t.b1.lo[0]; n.c1hix = t.b1.hi[0]; n.c1loy = t.b1.lo[1]; n.c1hiy = t.b1.hi[1];
            n.c0loz = t.b0.lo[2]; n.c0hiz = t.b0.hi[2]; n.c1loz = t.b1.lo[2]; n.c1hiz = t.b1.hi[2];
            n.child0 = t.c0 >= 0 ? finalOf[(size_t)t.c0] : t.c0;
            n.child1 = t.c1 >= 0 ? finalOf[(size_t)t.c1] : t.c1;
            n.pad0 = 0; n.pad1 = 0;
        }
    });
    fillTravTris(tris, bt.idx, threads, travTris);

    // ---- statistics -------------------------------------------------------------------------------------------
    {
        const float rootArea = std::max(bt.rootBox.area(), 1e-30f);
        double cost = C_TRAV;   // the root itself
        uint32_t leaves = 0;
        for (const TNode& n : all) {
            const Box* bx[2] = {&n.b0, &n.b1}; const int32_t ch[2] = {n.c0, n.c1};
            for (int s = 0; s < 2; ++s) {
                const float rel = bx[s]->area() / rootArea;
                if (ch[s] >= 0) cost += C_TRAV * rel;
                else { cost += rel * (float)leafCount(ch[s]); leaves++; }
            }
        }
        stats.nodeCount = (uint32_t)nodes.size(); stats.leafCount = leaves; stats.maxDepth = std::max(bt.maxDepth, 1u); stats.sahCost = (float)cost;
    }
    stats.buildMs = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// ------------------------------------------------------------------------------------------------------------
// 8-wide compressed BVH
// ------------------------------------------------------------------------------------------------------------
void buildCwbvh(const TriangleRec* tris, uint32_t count, int threads, std::vector<Node8>& nodes, std::vector<TravTri>& travTris, BvhStats& stats) {
    const auto t0 = std::chrono::steady_clock::now();
    nodes.clear(); travTris.clear();
    stats = BvhStats();
    if (count == 0) return;
    threads = resolveThreads(threads);
    BinaryTree bt;
    buildBinary(tris, count, threads, 1, bt);   // down to single triangles: the collapse below decides the leaves
    const std::vector<TNode>& all = bt.all;

    struct Child { int32_t link; Box box; };
    struct Work { int32_t bin; uint32_t node8; Box box; uint32_t depth; };   // bin < 0: the whole scene is one leaf
    std::vector<uint32_t> triOrder; triOrder.reserve(count);
    std::deque<Work> queue;
    nodes.emplace_back();
    queue.push_back({bt.root, 0u, bt.rootBox, 1u});
    const float rootArea = std::max(bt.rootBox.area(), 1e-30f);
    double cost = 0.0, leafExtentSum = 0.0;
    uint32_t leaves = 0, maxDepth = 0, leafNodes = 0;
    for (int a = 0; a < 3; ++a) { stats.lo[a] = bt.rootBox.lo[a]; stats.hi[a] = bt.rootBox.hi[a]; }

    // ---- which binary nodes become 8-wide nodes: the SAH-optimal collapse of Ylitie et al. (section 3.1) -----------
    //   c(n, i)  = cheapest way to represent subtree n with at most i roots (each root = one child slot of the parent)
    //   c(n, 1)  = min(leaf: A_n * P_n * C_PRIM if P_n <= 3,  inner: distribute(n, 8) + A_n * C_NODE)
    //   c(n, i)  = min(distribute(n, i), c(n, i - 1));   distribute(n, j) = min_k c(left, k) + c(right, j - k)
    // evaluated bottom-up (children always have larger indices than their parent), decisions replayed top-down.
    float C_NODE = 1.0f, C_PRIM = 1.0f;   // measured on B200: a triangle test costs the traversal about what a node costs
    if (const char* e = std::getenv("RTB_CPRIM")) C_PRIM = (float)std::atof(e);   // tuning hook
    const size_t nInner = all.size();
    struct Dp { float c[8]; uint8_t split[9]; uint8_t dec[8]; uint32_t triFirst, triCnt; };
    std::vector<Dp> dp(nInner);
    auto linkCost = [&](int32_t link, const Box& box, int i) { return link >= 0 ? dp[(size_t)link].c[i] : box.area() * (float)leafCount(link) * C_PRIM; };
    auto linkFirst = [&](int32_t link) { return link >= 0 ? dp[(size_t)link].triFirst : leafFirst(link); };
    auto linkCnt = [&](int32_t link) { return link >= 0 ? dp[(size_t)link].triCnt : leafCount(link); };
    for (size_t r = nInner; r-- > 0;) {
        const TNode& t = all[r];
        Dp& d = dp[r];
        Box nb; nb = t.b0; nb.grow(t.b1);
        const float A = nb.area();
        d.triFirst = std::min(linkFirst(t.c0), linkFirst(t.c1));
        d.triCnt = linkCnt(t.c0) + linkCnt(t.c1);
        float dist[9];
        for (int j = 2; j <= 8; ++j) {
            dist[j] = std::numeric_limits<float>::infinity(); d.split[j] = 1;
            for (int k = std::max(1, j - 7); k <= std::min(7, j - 1); ++k) {
                const float v = linkCost(t.c0, t.b0, k) + linkCost(t.c1, t.b1, j - k);
                if (v < dist[j]) { dist[j] = v; d.split[j] = (uint8_t)k; }
            }
        }
        const float cInner = dist[8] + A * C_NODE;
        const float cLeaf = d.triCnt <= CWBVH_MAX_LEAF ? A * (float)d.triCnt * C_PRIM : std::numeric_limits<float>::infinity();
        d.c[0] = 0.0f; d.dec[0] = 0;
        d.c[1] = std::min(cLeaf, cInner); d.dec[1] = cInner < cLeaf ? 1 : 0;
        for (int i = 2; i <= 7; ++i) {
            if (dist[i] < d.c[i - 1]) { d.c[i] = dist[i]; d.dec[i] = 1; } else { d.c[i] = d.c[i - 1]; d.dec[i] = 0; }
        }
    }

    while (!queue.empty()) {
        const Work w = queue.front(); queue.pop_front();
        if (w.depth > maxDepth) stats.levelFirst.push_back(w.node8);   // breadth-first: every level is one index range
        maxDepth = std::max(maxDepth, w.depth);
        Child ch[8]; int n = 0;
        if (w.bin < 0) { ch[0] = {w.bin, w.box}; n = 1; }
        else {
            // replay the decisions: the 8 slots of this node are distributed over the two binary children
            struct Item { int32_t link; Box box; int i; };
            Item st[16]; int sn = 0;
            const TNode& top = all[(size_t)w.bin];
            const int k0 = dp[(size_t)w.bin].split[8];
            st[sn++] = {top.c1, top.b1, 8 - k0};
            st[sn++] = {top.c0, top.b0, k0};
            while (sn > 0) {
                const Item it = st[--sn];
                if (it.link < 0) { ch[n++] = {it.link, it.box}; continue; }
                const Dp& d = dp[(size_t)it.link];
                if (it.i == 1) {
                    if (d.dec[1]) ch[n++] = {it.link, it.box};                                  // a new 8-wide node
                    else ch[n++] = {leafLink(d.triFirst, d.triCnt), it.box};                     // the whole subtree as one leaf slot
                    continue;
                }
                if (!d.dec[it.i]) { st[sn++] = {it.link, it.box, it.i - 1}; continue; }
                const TNode& t = all[(size_t)it.link];
                const int k = d.split[it.i];
                st[sn++] = {t.c1, t.b1, it.i - k};
                st[sn++] = {t.c0, t.b0, k};
            }
        }
        Box nb; nb.reset();
        for (int i = 0; i < n; ++i) nb.grow(ch[i].box);
        cost += C_TRAV * (nb.area() / rootArea);

        // ---- slot assignment: child i -> slot s maximising dot(centre_i - centre_node, d_s), d_s = (+-1,+-1,+-1) by the bits of s
        int slotOf[8]; bool slotUsed[8] = {false, false, false, false, false, false, false, false}; bool done[8] = {false, false, false, false, false, false, false, false};
        float costTab[8][8];
        for (int i = 0; i < n; ++i)
            for (int s = 0; s < 8; ++s) {
                float d = 0.0f;
                for (int a = 0; a < 3; ++a) {
                    const float cc = 0.5f * (ch[i].box.lo[a] + ch[i].box.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
                    d += ((s >> a) & 1) ? cc : -cc;
                }
                costTab[i][s] = d;
            }
        for (int k = 0; k < n; ++k) {
            int bi = -1, bs = -1; float bc = -std::numeric_limits<float>::infinity();
            for (int i = 0; i < n; ++i) if (!done[i]) for (int s = 0; s < 8; ++s) if (!slotUsed[s] && costTab[i][s] > bc) { bc = costTab[i][s]; bi = i; bs = s; }
            if (bi < 0) { for (int i = 0; i < n && bi < 0; ++i) if (!done[i]) bi = i; for (int s = 0; s < 8 && bs < 0; ++s) if (!slotUsed[s]) bs = s; }   // NaN costs
            done[bi] = true; slotUsed[bs] = true; slotOf[bi] = bs;
        }
        int childAt[8]; for (int s = 0; s < 8; ++s) childAt[s] = -1;
        for (int i = 0; i < n; ++i) childAt[slotOf[i]] = i;

        // ---- quantisation grid and child boxes: rtb_node8_encode.h (shared with the device refit) -------------------
        Node8 out;
        std::memset(&out, 0, sizeof out);
        double step[3];
        Box6 nb6; for (int a = 0; a < 3; ++a) { nb6.lo[a] = nb.lo[a]; nb6.hi[a] = nb.hi[a]; }
        node8Grid(nb6, out, step);
        out.childBase = (uint32_t)nodes.size();
        out.triBase = (uint32_t)triOrder.size();
        uint32_t innerCount = 0;
        for (int s = 0; s < 8; ++s) {
            const int i = childAt[s];
            if (i < 0) { node8Child(out, s, nullptr, step); continue; }
            Box6 cb; for (int a = 0; a < 3; ++a) { cb.lo[a] = ch[i].box.lo[a]; cb.hi[a] = ch[i].box.hi[a]; }
            node8Child(out, s, &cb, step);
            const float rel = ch[i].box.area() / rootArea;
            if (ch[i].link >= 0) {
                out.imask |= (uint8_t)(1u << s);
                const uint32_t id8 = (uint32_t)nodes.size() + innerCount;   // consecutive, in slot order
                ++innerCount;
                queue.push_back({ch[i].link, id8, ch[i].box, w.depth + 1});
            } else {
                const uint32_t first = leafFirst(ch[i].link), cnt = leafCount(ch[i].link);
                out.valid |= ((1u << cnt) - 1u) << (3 * s);   // unary count at the slot's three bits
                for (uint32_t k = 0; k < cnt; ++k) triOrder.push_back(bt.idx[first + k]);
                cost += rel * (float)cnt; ++leaves;
            }
        }
        out.valid |= (uint32_t)out.imask << 24;
        if (!innerCount) { leafExtentSum += ((double)(nb.hi[0] - nb.lo[0]) + (double)(nb.hi[1] - nb.lo[1]) + (double)(nb.hi[2] - nb.lo[2])) / 3.0; ++leafNodes; }
        nodes[w.node8] = out;
        for (uint32_t k = 0; k < innerCount; ++k) nodes.emplace_back();
    }
    fillTravTris(tris, triOrder, threads, travTris);
    stats.nodeCount = (uint32_t)nodes.size(); stats.leafCount = leaves; stats.maxDepth = maxDepth; stats.sahCost = (float)cost;
    stats.leafNodeExtent = leafNodes ? (float)(leafExtentSum / leafNodes) : 0.0f;
    stats.levelFirst.push_back((uint32_t)nodes.size());
    stats.buildMs = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace rtb
