// pathfind_core.h -- the reference's path-finding assembly methods (src/djfind.rs:78-492) on raw arrays, compiled for the host AND for
// the device: one group per thread runs exactly this code on the GPU (rogtk_gpu.cu: k_pathfind), the host entries call it too
// (pathfind_host.h, when the node ids are the ascending k-mer order), so the two agree by construction and the CPU tests of the
// host entry pin the device's arithmetic.  See pathfind_host.h for what each step follows in the reference and in petgraph 0.7.1 /
// Rust's BinaryHeap; this file is the same procedure without containers:
//   * node ids = ascending k-mer order (no boomphf rank: callers that pass one stay on pathfind_host.h);
//   * petgraph's edge ids are the CSR positions (edges are added from nodes in id order, right bases ascending -- the CSR's order),
//     so a node's outgoing list "newest first" is its CSR row backwards and needs no storage; incoming lists are built once;
//   * edge weights come from a table of -ln(sum / 2) over all sums of two u16 counts, filled on the HOST with libm's log (the
//     device's log differs from glibc's in the last bit now and then; additions, comparisons and the divisions of the scoring
//     are IEEE on both sides, and nothing here may be contracted into an FMA: there is no multiply-add in the score arithmetic).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define PF_HD __host__ __device__ inline
#else
#define PF_HD inline
#endif

namespace rogtk_pf {

constexpr int PF_MAX_ITERATIONS = 1000;   // back-trace steps (src/djfind.rs:185)
constexpr int PF_MAX_PAIRS = 100;         // endpoint pairs the automatic method scores (src/djfind.rs:406)

struct HeapItem {
    double score;
    int32_t node;
    int32_t pad;
};

struct View {   // one group's pruned nodes (ascending) + successor CSR rows
    int K;
    int32_t n;
    const uint64_t *lo, *hi;     // hi == nullptr unless K > 32
    const uint16_t *counts;
    const int64_t *row_ptr;      // [n + 1] positions of the group's rows in col (any base: only differences and col[...] are used)
    const int32_t *col;          // successor index inside the group
    const double *wtab;          // [131071] -ln(sum / 2)
};

// what the search keeps per node / per edge, packed so that one 32-byte sector answers a relaxation (the device kernel is bound by
// scattered sector traffic: score, flags, count and list head of a node used to live in four arrays)
struct NodeRec {
    double score;
    int32_t in_head;
    uint16_t count;
    uint8_t has, visited;
};
struct EdgeRec {
    int32_t in_next, e_from;
};

struct Scratch {   // all sized by the group: n nodes, m = row_ptr[n] - row_ptr[0] edges
    NodeRec *node;       // [n]
    EdgeRec *edge;       // [m]
    HeapItem *heap;      // [m + 1]
    int32_t *path;       // [min(n, 1001) + 1]  back-trace, end first
    int32_t *best_path;  // [min(n, 1001) + 1]  best so far, start first
    int32_t *list_a;     // [n] start nodes / start candidates
    int32_t *list_b;     // [n] end nodes / end candidates
};

PF_HD unsigned kmer_base(const View &g, int32_t i, int p)   // base p (0 = first) of node i
{
    const int sh = 2 * (g.K - 1 - p);
    return sh >= 64 ? (unsigned)((g.hi[i] >> (sh - 64)) & 3u) : (unsigned)((g.lo[i] >> sh) & 3u);
}
PF_HD char base_char(unsigned c)
{
    return c == 0u ? 'A' : (c == 1u ? 'C' : (c == 2u ? 'G' : 'T'));
}
PF_HD bool kmer_starts_with(const View &g, int32_t i, const char *a, int len)
{
    if (len > g.K) return false;
    for (int p = 0; p < len; p++)
        if (a[p] != base_char(kmer_base(g, i, p))) return false;
    return true;
}
PF_HD bool kmer_ends_with(const View &g, int32_t i, const char *a, int len)
{
    if (len > g.K) return false;
    const int off = g.K - len;
    for (int p = 0; p < len; p++)
        if (a[p] != base_char(kmer_base(g, i, off + p))) return false;
    return true;
}

PF_HD int32_t edge_count(const View &g) { return (int32_t)(g.row_ptr[g.n] - g.row_ptr[0]); }
PF_HD double edge_weight(const View &g, const Scratch &s, int32_t from, int32_t to)
{
    return g.wtab[(uint32_t)s.node[from].count + (uint32_t)s.node[to].count];
}
PF_HD int32_t edge_to(const View &g, int32_t e) { return g.col[g.row_ptr[0] + e]; }
// node's outgoing edges newest first: e = out_first(i), then e - 1 while e >= out_last(i)
PF_HD int32_t out_first(const View &g, int32_t i) { return (int32_t)(g.row_ptr[i + 1] - g.row_ptr[0]) - 1; }
PF_HD int32_t out_last(const View &g, int32_t i) { return (int32_t)(g.row_ptr[i] - g.row_ptr[0]); }

// incoming lists (newest edge first), the source of every edge, the node records
PF_HD void build_in_lists(const View &g, const Scratch &s)
{
    for (int32_t i = 0; i < g.n; i++) {
        s.node[i].in_head = -1;
        s.node[i].count = g.counts[i];
        s.node[i].has = 0;
        s.node[i].visited = 0;
    }
    for (int32_t i = 0; i < g.n; i++)
        for (int32_t e = out_last(g, i); e <= out_first(g, i); e++) {
            const int32_t to = edge_to(g, e);
            s.edge[e].e_from = i;
            s.edge[e].in_next = s.node[to].in_head;
            s.node[to].in_head = e;
        }
}

// ---- Rust std::collections::BinaryHeap<MinScored<f64, NodeIndex>> (smallest score on top; equal scores compare Equal) ----------
PF_HD bool heap_le(const HeapItem &a, const HeapItem &b) { return !(a.score < b.score); }
PF_HD void heap_sift_up(HeapItem *d, int32_t start, int32_t pos)
{
    const HeapItem el = d[pos];
    while (pos > start) {
        const int32_t parent = (pos - 1) / 2;
        if (heap_le(el, d[parent])) break;
        d[pos] = d[parent];
        pos = parent;
    }
    d[pos] = el;
}
PF_HD void heap_push(HeapItem *d, int32_t &size, double score, int32_t node)
{
    d[size] = HeapItem{score, node, 0};
    size++;
    heap_sift_up(d, 0, size - 1);
}
PF_HD bool heap_pop(HeapItem *d, int32_t &size, HeapItem &out)
{
    if (size == 0) return false;
    HeapItem item = d[size - 1];
    size--;
    if (size > 0) {
        const HeapItem top = d[0];
        d[0] = item;
        item = top;
        // sift_down_to_bottom(0), then sift_up
        const int32_t end = size;
        int32_t pos = 0;
        const HeapItem el = d[0];
        int32_t child = 1;
        while (end >= 2 && child <= end - 2) {
            if (heap_le(d[child], d[child + 1])) child++;
            d[pos] = d[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            d[pos] = d[child];
            pos = child;
        }
        d[pos] = el;
        heap_sift_up(d, 0, pos);
    }
    out = item;
    return true;
}

// petgraph::algo::dijkstra(graph, start, None, |e| *e.weight()): every node is final on its first pop
PF_HD void dijkstra(const View &g, const Scratch &s, int32_t start)
{
    for (int32_t i = 0; i < g.n; i++) {
        s.node[i].has = 0;
        s.node[i].visited = 0;
    }
    int32_t size = 0;
    s.node[start].score = 0.0;
    s.node[start].has = 1;
    heap_push(s.heap, size, 0.0, start);
    HeapItem it;
    while (heap_pop(s.heap, size, it)) {
        const int32_t node = it.node;
        if (s.node[node].visited) continue;
        const uint32_t node_count = s.node[node].count;
        for (int32_t e = out_first(g, node); e >= out_last(g, node); e--) {   // graph.edges(node): newest edge first
            const int32_t next = edge_to(g, e);
            NodeRec &nx = s.node[next];
            if (nx.visited) continue;
            const double next_score = it.score + g.wtab[node_count + (uint32_t)nx.count];
            if (nx.has) {
                if (next_score < nx.score) {
                    nx.score = next_score;
                    heap_push(s.heap, size, next_score, next);
                }
            } else {
                nx.score = next_score;
                nx.has = 1;
                heap_push(s.heap, size, next_score, next);
            }
        }
        s.node[node].visited = 1;
    }
}

// the custom back-trace of djfind.rs:176-228 from `end` to `start` over the scores of one dijkstra run (s.path, end first);
// returns the number of nodes, 0 when it breaks off
PF_HD int32_t trace_back(const View &g, const Scratch &s, int32_t start, int32_t end)
{
    const double FLOAT_EPSILON = 1e-9;
    int32_t len = 0;
    s.path[len++] = end;
    int32_t current = end;
    bool path_valid = false;
    int32_t iterations = 0;
    while (current != start) {
        iterations++;
        if (iterations > PF_MAX_ITERATIONS) break;
        int32_t best_prev = -1;
        double best_dist = INFINITY;
        const double current_dist = s.node[current].score;
        for (int32_t e = s.node[current].in_head; e >= 0; e = s.edge[e].in_next) {   // neighbors_directed(Incoming)
            const int32_t neighbor = s.edge[e].e_from;
            if (!s.node[neighbor].has) continue;
            // graph.find_edge(neighbor, current): the first match in neighbor's outgoing list; a k-mer graph has no parallel edges, so
            // it is this edge, and its weight depends on the two end nodes only
            const double w = edge_weight(g, s, neighbor, current);
            const double neighbor_dist = s.node[neighbor].score;
            const double expected_dist = neighbor_dist + w;
            if (fabs(expected_dist - current_dist) < FLOAT_EPSILON) {
                if (neighbor_dist < best_dist) {
                    best_dist = neighbor_dist;
                    best_prev = neighbor;
                }
            }
        }
        if (best_prev < 0) break;
        s.path[len++] = best_prev;
        current = best_prev;
        if (current == start) path_valid = true;
    }
    return path_valid ? len : 0;
}

// djfind.rs:157-247: starts / ends in node-id order; best path into s.best_path (start first).  Returns its length (0: none).
PF_HD int32_t find_shortest_path(const View &g, const Scratch &s, const int32_t *starts, int32_t n_starts, const int32_t *ends, int32_t n_ends,
                                 double &best_weight)
{
    int32_t best_len = 0;
    double min_total_weight = INFINITY;
    for (int32_t a = 0; a < n_starts; a++) {
        const int32_t start = starts[a];
        dijkstra(g, s, start);
        for (int32_t b = 0; b < n_ends; b++) {
            const int32_t end = ends[b];
            if (!s.node[end].has) continue;
            const double total_weight = s.node[end].score;
            if (!(total_weight < min_total_weight)) continue;
            const int32_t len = trace_back(g, s, start, end);
            if (len > 0) {
                for (int32_t t = 0; t < len; t++) s.best_path[t] = s.path[len - 1 - t];
                best_len = len;
                best_weight = total_weight;
                min_total_weight = total_weight;
            }
        }
    }
    return best_len;
}

// djfind.rs:60-74: first K-mer, then the last base of every further node.  Returns the number of bytes written.
PF_HD int32_t path_sequence(const View &g, const int32_t *path, int32_t len, char *out)
{
    if (len <= 0) return 0;
    int32_t w = 0;
    for (int p = 0; p < g.K; p++) out[w++] = base_char(kmer_base(g, path[0], p));
    for (int32_t t = 1; t < len; t++) out[w++] = base_char(kmer_base(g, path[t], g.K - 1));
    return w;
}

// AssemblyMethod::ShortestPath (src/fracture.rs:385-416, src/djfind.rs:257-304): bytes written to out (0: anchors / path missing).
// out holds n + K - 1 bytes at least.
PF_HD int32_t anchored_contig(const View &g, const Scratch &s, const char *start_anchor, int sl, const char *end_anchor, int el, char *out)
{
    if (g.n <= 0) return 0;
    int32_t ns = 0, ne = 0;
    for (int32_t i = 0; i < g.n; i++) {
        if (kmer_starts_with(g, i, start_anchor, sl)) s.list_a[ns++] = i;
        if (kmer_ends_with(g, i, end_anchor, el)) s.list_b[ne++] = i;
    }
    if (ns == 0 || ne == 0) return 0;
    build_in_lists(g, s);
    double w = 0.0;
    const int32_t len = find_shortest_path(g, s, s.list_a, ns, s.list_b, ne, w);
    return path_sequence(g, s.best_path, len, out);
}

// AssemblyMethod::ShortestPathAuto (src/djfind.rs:309-492, src/fracture.rs:419-461)
PF_HD int32_t auto_contig(const View &g, const Scratch &s, char *out)
{
    if (g.n <= 0) return 0;
    build_in_lists(g, s);
    // find_endpoint_candidates
    double sum = 0.0;
    for (int32_t i = 0; i < g.n; i++) sum += (double)g.counts[i];
    const double avg = sum / (double)g.n;
    const double thr_f = fmax(avg * 0.1, 1.0);
    const uint16_t thr = thr_f >= 65535.0 ? (uint16_t)65535 : (uint16_t)thr_f;   // `as u16` truncates (and saturates)
    int32_t ns = 0, ne = 0;
    for (int32_t i = 0; i < g.n; i++) {
        if (g.counts[i] < thr) continue;
        const bool has_in = s.node[i].in_head >= 0;
        const int32_t out_degree = (int32_t)(g.row_ptr[i + 1] - g.row_ptr[i]);
        if (!has_in && out_degree > 0) s.list_a[ns++] = i;
        if (out_degree == 0 && has_in) s.list_b[ne++] = i;
    }
    if (ns == 0 || ne == 0) return 0;
    if (ns == 1 && ne == 1) {
        // the reference searches again with the two full K-mer strings as anchors: they match exactly these two nodes
        double w = 0.0;
        const int32_t len = find_shortest_path(g, s, s.list_a, 1, s.list_b, 1, w);
        return path_sequence(g, s.best_path, len, out);
    }
    // find_best_endpoint_pair: the first 100 pairs, best score wins (strictly greater replaces)
    int32_t evaluated = 0;
    bool have = false;
    double best_score = 0.0;
    int32_t best_len = 0;
    for (int32_t a = 0; a < ns; a++) {
        const int32_t st = s.list_a[a];
        bool have_dist = false;   // (the reference runs dijkstra(st) again for every end; the scores are the same)
        for (int32_t b = 0; b < ne; b++) {
            if (evaluated >= PF_MAX_PAIRS) break;
            evaluated++;
            const int32_t en = s.list_b[b];
            if (!have_dist) {
                dijkstra(g, s, st);
                have_dist = true;
            }
            if (!s.node[en].has) continue;
            const double w = s.node[en].score;
            if (!(w < INFINITY)) continue;
            const int32_t len = trace_back(g, s, st, en);
            if (len == 0) continue;
            // score_path
            const double path_length = (double)((int64_t)len * (int64_t)g.K);
            const double mean_coverage = 1.0 / (w / (double)len);
            const double normalized_length = fmin(path_length / 5000.0, 1.0);
            const double normalized_coverage = fmin(mean_coverage / 100.0, 1.0);
#ifdef __CUDA_ARCH__
            const double score = __dadd_rn(__dmul_rn(0.6, normalized_length), __dmul_rn(0.4, normalized_coverage));   // no FMA: the host rounds twice
#else
            const double score = 0.6 * normalized_length + 0.4 * normalized_coverage;
#endif
            if (!have || score > best_score) {
                have = true;
                best_score = score;
                for (int32_t t = 0; t < len; t++) s.best_path[t] = s.path[len - 1 - t];
                best_len = len;
            }
        }
    }
    return have ? path_sequence(g, s.best_path, best_len, out) : 0;
}

}  // namespace rogtk_pf
