// Host BVH builder: binned-SAH binary tree (multi-threaded) -> SAH-optimal collapse to 8-wide ->
// octant-ordered, quantised 64-byte nodes and 64-byte triangle units (48-byte TriAccel record + pad) in one array.
// Replaces Accel_QBVH::Build (/root/reference/src/liblightmetrica/accel/accel_qbvh.cpp:152-396).
// What is kept from the reference: triangles are world-space, every triangle's box is padded
// by Math::Eps() = 1e-4 (accel_qbvh.cpp:189-190) so node culling is never tighter than the
// reference's, and the triangle records come from the bit-exact TriAccel precompute.
// What is new: 3-axis binning (the reference bins the longest axis only), <=3-triangle leaves,
// 8-wide nodes with octant slot assignment, contiguous arrays instead of one heap node each.
#include "bvh.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <mutex>
#include <thread>

namespace lmb200 {

namespace {

struct Ref {            // 32 bytes, partitioned in place
    float lo[3];
    uint32_t id;
    float hi[3];
    uint32_t pad;
};

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; } }
    void grow(const float* l, const float* h) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]); } }
    void grow(const Box& b) { grow(b.lo, b.hi); }
    float half_area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

struct BinNode {
    Box box;
    uint32_t left;    // children are left, left+1 (count == 0)
    uint32_t first;   // the subtree covers refs[first, first+total)
    uint32_t count;   // > 0: binary leaf with `count` triangles
    uint32_t total;   // triangles in the subtree
};

constexpr int kBins = 16;
constexpr int kMaxLeaf = 3;           // triangles per leaf slot (2 bits per slot in Node64::counts)
constexpr float kCostNode = 1.0f;     // binary build SAH: one split step
constexpr float kCostTri = 1.0f;      // binary build SAH: one triangle test
// wide-node SAH used by the collapse (Ylitie et al. 2017, sec. 4.1): one 8-wide node visit vs one triangle test
#ifndef LMB_WIDE_COST_NODE
#define LMB_WIDE_COST_NODE 1.0f
#endif
#ifndef LMB_WIDE_COST_TRI
#define LMB_WIDE_COST_TRI 1.2f   // measured on B200 (profiles/r01_sweep.md): the TriAccel test (3 x LDG.128, one divide) costs about a node visit
#endif
constexpr uint32_t kParallelGrain = 1u << 14;
constexpr double kQSlack = 1.0 / 256.0;   // grid steps added around every quantised child box (see emit_node)

struct Builder {
    std::vector<Ref> refs;
    std::vector<BinNode> nodes;
    std::atomic<uint32_t> node_count{0};

    // work queue of subtrees for the thread pool
    struct Task { uint32_t node, begin, end; };
    std::vector<Task> tasks;
    std::mutex mu;
    std::atomic<int> pending{0};

    uint32_t alloc_pair() { return node_count.fetch_add(2); }

    static Box bounds_of(const Ref* r, uint32_t n) {
        Box b; b.reset();
        for (uint32_t i = 0; i < n; i++) b.grow(r[i].lo, r[i].hi);
        return b;
    }

    // Finds the best binned SAH split of refs[begin,end). Returns false if no split separates them.
    bool find_split(uint32_t begin, uint32_t end, const Box& box, int& best_axis, float& best_pos, float& best_cost) const {
        Box cb; cb.reset();
        for (uint32_t i = begin; i < end; i++) {
            const Ref& r = refs[i];
            float c[3] = {0.5f * (r.lo[0] + r.hi[0]), 0.5f * (r.lo[1] + r.hi[1]), 0.5f * (r.lo[2] + r.hi[2])};
            cb.grow(c, c);
        }
        best_cost = INFINITY; best_axis = -1;
        const float inv_area = 1.0f / std::max(box.half_area(), 1e-30f);
        for (int axis = 0; axis < 3; axis++) {
            const float cmin = cb.lo[axis], cmax = cb.hi[axis];
            if (!(cmax > cmin)) continue;
            const float scale = kBins / (cmax - cmin);
            Box bb[kBins]; uint32_t cnt[kBins];
            for (int b = 0; b < kBins; b++) { bb[b].reset(); cnt[b] = 0; }
            for (uint32_t i = begin; i < end; i++) {
                const Ref& r = refs[i];
                const float c = 0.5f * (r.lo[axis] + r.hi[axis]);
                int b = (int)((c - cmin) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                bb[b].grow(r.lo, r.hi); cnt[b]++;
            }
            float right_area[kBins]; uint32_t right_cnt[kBins];
            Box acc; acc.reset(); uint32_t n = 0;
            for (int b = kBins - 1; b > 0; b--) {
                if (cnt[b]) acc.grow(bb[b]);
                n += cnt[b];
                right_area[b] = n ? acc.half_area() : 0.f; right_cnt[b] = n;
            }
            acc.reset(); n = 0;
            for (int b = 0; b < kBins - 1; b++) {
                if (cnt[b]) acc.grow(bb[b]);
                n += cnt[b];
                if (n == 0 || right_cnt[b + 1] == 0) continue;
                const float cost = kCostNode + kCostTri * (acc.half_area() * n + right_area[b + 1] * right_cnt[b + 1]) * inv_area;
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_pos = cmin + (b + 1) / scale; }
            }
        }
        return best_axis >= 0;
    }

    // Splits node over refs[begin,end); returns mid, or 0 if it became a leaf.
    uint32_t split_node(uint32_t ni, uint32_t begin, uint32_t end) {
        BinNode& node = nodes[ni];
        const uint32_t n = end - begin;
        node.box = bounds_of(&refs[begin], n);
        node.count = 0; node.first = begin; node.left = 0; node.total = n;
        if (n == 1) { node.count = 1; return 0; }
        int axis; float pos, cost;
        uint32_t mid = 0;
        if (find_split(begin, end, node.box, axis, pos, cost)) {
            Ref* lo = &refs[begin]; Ref* hi = &refs[end];
            Ref* m = std::partition(lo, hi, [&](const Ref& r) { return 0.5f * (r.lo[axis] + r.hi[axis]) < pos; });
            mid = begin + (uint32_t)(m - lo);
        }
        if (mid == begin || mid == end || mid == 0) {
            // coincident centroids (or a degenerate bin edge): the reference recurses forever here
            // (accel_qbvh.cpp:375-377); we fall back to a leaf or an index-median split.
            if (n <= (uint32_t)kMaxLeaf) { node.count = n; return 0; }
            mid = begin + n / 2;
        }
        const uint32_t l = alloc_pair();
        nodes[ni].left = l;
        return mid;
    }

    void build_serial(uint32_t ni, uint32_t begin, uint32_t end) {
        // explicit stack: depth can be large for adversarial inputs
        std::vector<Task> st;
        st.push_back({ni, begin, end});
        while (!st.empty()) {
            const Task t = st.back(); st.pop_back();
            const uint32_t mid = split_node(t.node, t.begin, t.end);
            if (!mid) continue;
            const uint32_t l = nodes[t.node].left;
            st.push_back({l + 1, mid, t.end});
            st.push_back({l, t.begin, mid});
        }
    }

    void run(int num_threads) {
        const uint32_t n = (uint32_t)refs.size();
        nodes.resize(std::max<size_t>(1, 2 * (size_t)n));
        node_count = 1;
        if (n == 0) { nodes[0].box.reset(); nodes[0].count = 0; nodes[0].left = 0; nodes[0].first = 0; nodes[0].total = 0; return; }
        // Phase A: split the largest open ranges on this thread until there is enough parallel work.
        std::vector<Task> open;
        open.push_back({0, 0, n});
        const size_t want = (size_t)std::max(1, num_threads) * 8;
        while (open.size() < want) {
            size_t bi = 0;
            for (size_t i = 1; i < open.size(); i++) if (open[i].end - open[i].begin > open[bi].end - open[bi].begin) bi = i;
            const Task t = open[bi];
            if (t.end - t.begin <= kParallelGrain || num_threads <= 1) break;
            open.erase(open.begin() + bi);
            const uint32_t mid = split_node(t.node, t.begin, t.end);
            if (!mid) continue;
            const uint32_t l = nodes[t.node].left;
            open.push_back({l, t.begin, mid});
            open.push_back({l + 1, mid, t.end});
        }
        // Phase B: finish the subtrees in parallel.
        std::sort(open.begin(), open.end(), [](const Task& a, const Task& b) { return a.end - a.begin > b.end - b.begin; });
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= open.size()) break;
                build_serial(open[i].node, open[i].begin, open[i].end);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < num_threads; t++) th.emplace_back(worker);
        worker();
        for (auto& x : th) x.join();
    }
};

// ------------------------------------------------------------------------------------------------
// Collapse + emit

struct Emitter {
    const Builder& B;
    const std::vector<TriRecord>& recs;   // per input triangle
    HostBVH& out;
    double sah = 0;
    int max_depth = 0;

    struct Pending { uint32_t bin; uint32_t wide; int depth; };

    uint64_t num_nodes = 0;

    void emit_all() {
        out.units.clear();
        out.units.emplace_back();
        memset(&out.units[0], 0, sizeof(Unit64));
        num_nodes = 1;
        const uint32_t nrefs = (uint32_t)B.refs.size();
        if (nrefs == 0) {       // empty scene: a root whose eight slots are empty
            Node64& r = out.units[0].node;
            r.e[0] = r.e[1] = r.e[2] = 127;
            for (int s = 0; s < 8; s++) for (int a = 0; a < 3; a++) { r.qlo[a][s] = 255; r.qhi[a][s] = 0; }
            return;
        }
        out.units.reserve((size_t)nrefs + nrefs / 2);
        plan();
        std::vector<Pending> st;
        st.push_back({0, 0, 1});
        const float root_area = std::max(B.nodes[0].box.half_area(), 1e-30f);
        while (!st.empty()) {
            const Pending p = st.back(); st.pop_back();
            max_depth = std::max(max_depth, p.depth);
            emit_node(p, st, root_area);
        }
    }

    // ---- SAH-optimal collapse (dynamic programme over the binary tree) ----
    // cost[n][i-1], i = 1..7: cheapest way to represent subtree n as at most i wide-node slots.
    //   i = 1: either one leaf slot (<= 3 triangles) or one internal slot (a new wide node with up to 8 slots)
    //   i > 1: additionally the slots may be split between the two binary children.
    struct Choice { uint8_t take[7]; uint8_t dist_k[7]; uint8_t int_k; };   // take: 0 leaf, 1 internal, 2 distribute, 3 as i-1
    std::vector<float> cost;       // 7 per binary node
    std::vector<Choice> choice;

    void plan() {
        const uint32_t nb = B.node_count.load();
        cost.assign((size_t)nb * 7, 0.f);
        choice.assign(nb, Choice());
        const float inv_root = 1.0f / std::max(B.nodes[0].box.half_area(), 1e-30f);
        for (uint32_t n = nb; n-- > 0;) {          // children have larger indices than their parent
            const BinNode& N = B.nodes[n];
            float* c = &cost[(size_t)n * 7];
            Choice& ch = choice[n];
            const float area = N.box.half_area() * inv_root;
            const float leaf = N.total <= (uint32_t)kMaxLeaf ? area * N.total * LMB_WIDE_COST_TRI : INFINITY;
            if (N.count > 0) {                     // binary leaf: nothing to distribute
                for (int i = 0; i < 7; i++) { c[i] = leaf; ch.take[i] = 0; }
                continue;
            }
            const float* cl = &cost[(size_t)N.left * 7];
            const float* cr = &cost[(size_t)(N.left + 1) * 7];
            auto distribute = [&](int j, uint8_t& kbest) {   // best split of j slots (2..8) between the children
                float best = INFINITY; kbest = 1;
                for (int k = 1; k < j; k++) {
                    if (k > 7 || j - k > 7) continue;
                    const float v = cl[k - 1] + cr[j - k - 1];
                    if (v < best) { best = v; kbest = (uint8_t)k; }
                }
                return best;
            };
            const float internal = distribute(8, ch.int_k) + area * LMB_WIDE_COST_NODE;
            if (leaf <= internal) { c[0] = leaf; ch.take[0] = 0; } else { c[0] = internal; ch.take[0] = 1; }
            for (int i = 2; i <= 7; i++) {
                const float d = distribute(i, ch.dist_k[i - 1]);
                if (d < c[i - 2]) { c[i - 1] = d; ch.take[i - 1] = 2; } else { c[i - 1] = c[i - 2]; ch.take[i - 1] = 3; }
            }
        }
    }

    // slots of subtree n when it may use at most i slots
    void collect(uint32_t n, int i, uint32_t* out, bool* out_leaf, int& cnt) const {
        for (;;) {
            const uint8_t t = choice[n].take[i - 1];
            if (t == 3) { i--; continue; }
            if (t == 0) { out[cnt] = n; out_leaf[cnt++] = true; return; }
            if (t == 1) { out[cnt] = n; out_leaf[cnt++] = false; return; }
            const int k = choice[n].dist_k[i - 1];
            collect(B.nodes[n].left, k, out, out_leaf, cnt);
            n = B.nodes[n].left + 1; i = i - k;
        }
    }

    void emit_node(const Pending& p, std::vector<Pending>& st, float root_area) {
        // 1. the slots of this wide node as planned by the dynamic programme
        uint32_t ch[8]; bool is_leaf[8]; int n = 0;
        const BinNode& root = B.nodes[p.bin];
        if (root.count > 0 || root.total <= 1) { ch[0] = p.bin; is_leaf[0] = true; n = 1; }   // the whole tree is one leaf
        else {
            const int k = choice[p.bin].int_k;
            collect(root.left, k, ch, is_leaf, n);
            collect(root.left + 1, 8 - k, ch, is_leaf, n);
        }
        // 2. node box + quantisation grid. The origin is the node minimum snapped DOWN to the scene grid (16 bits per
        //    axis); child planes are quantised against the decoded origin, so the snap costs range, not tightness.
        Box nb; nb.reset();
        for (int i = 0; i < n; i++) nb.grow(B.nodes[ch[i]].box);
        Node64 node; memset(&node, 0, sizeof(node));
        double scale[3];
        float origin[3];
        for (int a = 0; a < 3; a++) {
            node.k[a] = grid_floor(out.grid, a, nb.lo[a]);
            origin[a] = grid_origin(out.grid, a, node.k[a]);
            const double ext = std::max(0.0, (double)nb.hi[a] - (double)origin[a]);
            // the traversal evaluates q*step + base as fma(1 + q*2^-15, step*2^15, base - step*2^15): child
            // boxes are widened by kQSlack grid steps to cover that rounding, and e+15 must stay a valid exponent
            int e = ext > 0 ? (int)std::ceil(std::log2(ext / 254.0)) : -126;
            e = std::max(-126, std::min(110, e));
            while (e < 110 && std::ceil(ext / std::ldexp(1.0, e) + 2 * kQSlack) > 255.0) e++;
            node.e[a] = (uint8_t)(e + 127);
            scale[a] = std::ldexp(1.0, e);
        }
        // 3. octant slot assignment (greedy on the centroid-offset score)
        float cen[8][3];
        for (int i = 0; i < n; i++) for (int a = 0; a < 3; a++) {
            const Box& b = B.nodes[ch[i]].box;
            cen[i][a] = 0.5f * (b.lo[a] + b.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
        }
        int slot_of[8]; bool slot_used[8] = {false}; bool child_done[8] = {false};
        for (int round = 0; round < n; round++) {
            int bc = -1, bs = -1; float bscore = -INFINITY;
            for (int i = 0; i < n; i++) {
                if (child_done[i]) continue;
                for (int s = 0; s < 8; s++) {
                    if (slot_used[s]) continue;
                    // slot s is visited first by rays whose direction is negative on the axes of its set bits
                    const float score = ((s & 1) ? cen[i][0] : -cen[i][0]) + ((s & 2) ? cen[i][1] : -cen[i][1]) + ((s & 4) ? cen[i][2] : -cen[i][2]);
                    if (score > bscore) { bscore = score; bc = i; bs = s; }
                }
            }
            slot_of[bc] = bs; slot_used[bs] = true; child_done[bc] = true;
        }
        int child_in_slot[8];
        for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
        for (int i = 0; i < n; i++) child_in_slot[slot_of[i]] = i;
        // 4. allocate the children in the unit array: internal children first (slot order), then the triangles of the
        //    leaf slots (slot order)
        uint32_t n_internal = 0, n_tris = 0;
        for (int i = 0; i < n; i++) { if (is_leaf[i]) n_tris += B.nodes[ch[i]].total; else n_internal++; }
        const uint32_t base = (uint32_t)out.units.size();
        out.units.resize(out.units.size() + n_internal + n_tris);
        node.base = base;
        uint32_t tri_off = 0;
        const float inv_root = 1.0f / root_area;
        for (int s = 0; s < 8; s++) {
            const int i = child_in_slot[s];
            if (i < 0) { for (int a = 0; a < 3; a++) { node.qlo[a][s] = 255; node.qhi[a][s] = 0; } continue; }
            const BinNode& c = B.nodes[ch[i]];
            for (int a = 0; a < 3; a++) {
                double ql = std::floor(((double)c.box.lo[a] - (double)origin[a]) / scale[a] - kQSlack);
                double qh = std::ceil(((double)c.box.hi[a] - (double)origin[a]) / scale[a] + kQSlack);
                ql = std::max(0.0, std::min(255.0, ql));
                qh = std::max(0.0, std::min(255.0, qh));
                node.qlo[a][s] = (uint8_t)ql; node.qhi[a][s] = (uint8_t)qh;
            }
            if (is_leaf[i]) {
                node.counts |= (uint16_t)(c.total << (2 * s));
                for (uint32_t k = 0; k < c.total; k++) {
                    const uint32_t id = B.refs[c.first + k].id;
                    TriUnit& tu = out.units[base + n_internal + tri_off + k].tri;
                    tu.rec = recs[id];
                    tu.pad[0] = tu.pad[1] = tu.pad[2] = tu.pad[3] = 0;
                }
                tri_off += c.total;
                sah += LMB_WIDE_COST_TRI * c.total * c.box.half_area() * inv_root;
            } else {
                node.imask |= (uint8_t)(1u << s);
            }
        }
        sah += LMB_WIDE_COST_NODE * nb.half_area() * inv_root;
        out.units[p.wide].node = node;
        num_nodes += n_internal;
        // 5. schedule the internal children (contiguous, slot order)
        uint32_t rel = 0;
        for (int s = 0; s < 8; s++) {
            const int i = child_in_slot[s];
            if (i < 0 || is_leaf[i]) continue;
            st.push_back({ch[i], base + rel, p.depth + 1});
            rel++;
        }
    }
};

}  // namespace

void make_scene_grid(const float lo[3], const float hi[3], SceneGrid& g)
{
    for (int a = 0; a < 3; a++) {
        const double l = std::isfinite(lo[a]) ? lo[a] : 0.0, h = std::isfinite(hi[a]) && hi[a] >= lo[a] ? hi[a] : l;
        int e = (int)std::ceil(std::log2(std::max(h - l, 1e-30) / 65535.0));
        e = std::max(-100, std::min(100, e));
        for (;; e++) {
            const double step = std::ldexp(1.0, e);
            const double gl = std::floor(l / step) * step;
            // 65535 steps must reach the top, and lo + k step must be exact in fp32 for every k (|value| / step < 2^23)
            if (gl + 65535.0 * step >= h && std::fabs(gl) / step < 4194304.0 && (std::fabs(gl) + 65535.0 * step) / step < 8388608.0) {
                g.lo[a] = (float)gl; g.step[a] = (float)step;
                break;
            }
        }
    }
}

uint16_t grid_floor(const SceneGrid& g, int axis, float x)
{
    double k = std::floor(((double)x - (double)g.lo[axis]) / (double)g.step[axis]);
    k = std::max(0.0, std::min(65535.0, k));
    uint32_t ki = (uint32_t)k;
    while (ki > 0 && grid_origin(g, axis, ki) > x) ki--;
    return (uint16_t)ki;
}

void build_bvh(const float* verts, uint64_t ntris, HostBVH& out, int num_threads)
{
    const auto t0 = std::chrono::steady_clock::now();
    if (num_threads <= 0) num_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    num_threads = std::min(num_threads, 64);

    // TriAccel records + scene bounds
    std::vector<TriRecord> recs(ntris);
    std::vector<uint8_t> valid(ntris, 0);
    {
        auto work = [&](uint64_t b, uint64_t e) {
            for (uint64_t i = b; i < e; i++) {
                const float* v = verts + 9 * i;
                const int degenerate = triaccel_load(recs[i], v, v + 3, v + 6, (uint32_t)i);
                bool finite = true;
                for (int k = 0; k < 9; k++) finite = finite && std::isfinite(v[k]);
                valid[i] = (!degenerate && finite) ? 1 : 0;
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < num_threads; t++) th.emplace_back(work, ntris * t / num_threads, ntris * (t + 1) / num_threads);
        work(0, ntris / num_threads);
        for (auto& x : th) x.join();
    }
    Box scene; scene.reset();
    uint64_t nvalid = 0;
    for (uint64_t i = 0; i < ntris; i++) {
        if (!valid[i]) continue;
        nvalid++;
        const float* v = verts + 9 * i;
        for (int k = 0; k < 3; k++) scene.grow(v + 3 * k, v + 3 * k);
    }
    if (nvalid == 0) { for (int a = 0; a < 3; a++) { scene.lo[a] = scene.hi[a] = 0.f; } }
    float extent = 0.f;
    for (int a = 0; a < 3; a++) extent = std::max(extent, std::max(std::fabs(scene.lo[a]), std::fabs(scene.hi[a])));
    // Reference pad (Math::Eps) plus a slack proportional to the scene size that covers the
    // rounding of the quantised slab test (about 4e-7 * distance), so culling stays conservative.
    const float pad = 1e-4f + 4e-6f * extent;

    Builder B;
    B.refs.reserve(nvalid);
    for (uint64_t i = 0; i < ntris; i++) {
        if (!valid[i]) continue;
        const float* v = verts + 9 * i;
        Ref r; r.id = (uint32_t)i; r.pad = 0;
        for (int a = 0; a < 3; a++) {
            r.lo[a] = std::min(v[a], std::min(v[3 + a], v[6 + a])) - pad;
            r.hi[a] = std::max(v[a], std::max(v[3 + a], v[6 + a])) + pad;
        }
        B.refs.push_back(r);
    }
    B.run(num_threads);

    {
        float glo[3], ghi[3];
        for (int a = 0; a < 3; a++) { glo[a] = nvalid ? scene.lo[a] - pad : 0.f; ghi[a] = nvalid ? scene.hi[a] + pad : 0.f; }
        make_scene_grid(glo, ghi, out.grid);
    }
    Emitter E{B, recs, out};
    E.emit_all();

    for (int a = 0; a < 3; a++) { out.scene_lo[a] = scene.lo[a]; out.scene_hi[a] = scene.hi[a]; }
    out.stats.num_triangles = ntris;
    out.stats.num_valid = nvalid;
    out.stats.num_nodes = E.num_nodes;
    out.stats.sah_cost = (float)E.sah;
    out.stats.max_depth = E.max_depth;
    out.stats.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace lmb200
