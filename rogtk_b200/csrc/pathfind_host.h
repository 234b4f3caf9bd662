// pathfind_host.h -- host C++ restatement of the reference's path-finding assembly methods on top of the node arrays + successor
// CSR the GPU path emits.  Pure host code, no CUDA: included by rogtk_gpu.cu (product) only.
//
// What it follows, line by line:
//   * reference src/djfind.rs:78-121   convert_to_petgraph: one node per surviving k-mer in graph-node order, one edge per
//                                      right neighbour (bases ascending), weight = -ln((cov_from + cov_to) / 2)
//   * reference src/djfind.rs:124-154  find_anchor_nodes: starts_with(start) / ends_with(end) on the K-mer strings
//   * reference src/djfind.rs:157-247  find_shortest_path: petgraph::algo::dijkstra per start node, then the custom back-trace
//                                      (|d[n] + w - d[cur]| < 1e-9, smallest d[n] wins, at most 1000 steps)
//   * reference src/djfind.rs:257-304  assemble_with_path_finding; :60-74 concatenate_path_sequences
//   * reference src/djfind.rs:309-492  ShortestPathAuto: endpoint candidates, pair scoring
//   * petgraph 0.7.1 algo::dijkstra + Graph adjacency order and Rust's std BinaryHeap -- restated from the
//     published algorithms: all rogtk edge weights are <= 0 and nodes are finalised on their FIRST pop, so the answer is
//     whatever this exact procedure yields, including how the heap orders equal scores.  Hence the heap below is Rust's
//     (sift_up on push, sift_down_to_bottom + sift_up on pop), not std::priority_queue.
//
// Node order.  The reference numbers graph nodes in boomphf slot order, which this repo does not reproduce: the caller
// may pass `rank` (rank[i] = reference node id of the i-th k-mer in ascending order); without it ids are the ascending order.
// Everything order-dependent (edge insertion, anchor lists, tie-breaks) is derived from the ids, so a Rust host that passes the
// boomphf order gets the reference's tie behaviour.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "pathfind_core.h"

namespace rogtk_host {

struct GroupGraph {   // one group's pruned node set, ascending k-mer order (what rogtk_gpu_count_prune_batch returns)
    int K = 0;
    int32_t n = 0;
    const uint64_t *lo = nullptr, *hi = nullptr;   // hi == nullptr unless K > 32
    const uint8_t *exts = nullptr;
    const uint16_t *counts = nullptr;
    const int64_t *row_ptr = nullptr;              // CSR rows of this group's nodes: successors of node i are col[row_ptr[i] .. row_ptr[i+1])
    const int32_t *col = nullptr;                  // successor index inside the group, right-ext bases ascending
    const int32_t *rank = nullptr;                 // optional reference node ids (a permutation of 0..n-1)
};

inline char base_char(unsigned c) { return "ACGT"[c & 3u]; }
inline unsigned kmer_base(const GroupGraph &g, int32_t i, int p)   // base p (0 = first) of node i
{
    const int sh = 2 * (g.K - 1 - p);
    return sh >= 64 ? (unsigned)((g.hi[i] >> (sh - 64)) & 3u) : (unsigned)((g.lo[i] >> sh) & 3u);
}
inline std::string kmer_string(const GroupGraph &g, int32_t i)
{
    std::string s((size_t)g.K, 'A');
    for (int p = 0; p < g.K; p++) s[(size_t)p] = base_char(kmer_base(g, i, p));
    return s;
}
inline bool kmer_starts_with(const GroupGraph &g, int32_t i, const std::string &a)
{
    if ((int)a.size() > g.K) return false;
    for (size_t p = 0; p < a.size(); p++)
        if (a[p] != base_char(kmer_base(g, i, (int)p))) return false;
    return true;
}
inline bool kmer_ends_with(const GroupGraph &g, int32_t i, const std::string &a)
{
    if ((int)a.size() > g.K) return false;
    const int off = g.K - (int)a.size();
    for (size_t p = 0; p < a.size(); p++)
        if (a[p] != base_char(kmer_base(g, i, off + (int)p))) return false;
    return true;
}

// petgraph::Graph as convert_to_petgraph builds it; node ids are reference ids ("pid"), edges in insertion order
struct PetGraph {
    int32_t n = 0;
    std::vector<int32_t> node_of_pid, pid_of_node;   // pid -> ascending index and back
    std::vector<int32_t> e_from, e_to;               // pids
    std::vector<double> e_w;
    // petgraph keeps per node a singly linked list of its outgoing / incoming edges, newest first
    std::vector<int32_t> out_head, in_head, out_next, in_next;
};

// -ln((cov_from + cov_to) / 2) for every possible sum of two u16 counts: the same libm call the per-edge expression makes (the sum
// of two integers below 2^17 halves exactly in f64), computed once per process -- std::log was a third of a group's search time
inline const double *edge_weight_table()
{
    static std::vector<double> tab;
    static std::once_flag once;
    std::call_once(once, [] {
        tab.resize(2 * 65535 + 1);
        for (size_t sum = 0; sum < tab.size(); sum++) {
            const double mean_cov = (double)sum / 2.0;
            tab[sum] = -std::log(mean_cov);
        }
    });
    return tab.data();
}

inline void build_petgraph(const GroupGraph &g, PetGraph &pg)
{
    const double *wtab = edge_weight_table();
    const int32_t n = g.n;
    pg.n = n;
    pg.node_of_pid.assign((size_t)n, 0);
    pg.pid_of_node.assign((size_t)n, 0);
    for (int32_t i = 0; i < n; i++) {
        const int32_t pid = g.rank ? g.rank[i] : i;
        pg.pid_of_node[(size_t)i] = pid;
        pg.node_of_pid[(size_t)pid] = i;
    }
    pg.e_from.clear(); pg.e_to.clear(); pg.e_w.clear();
    pg.out_head.assign((size_t)n, -1);
    pg.in_head.assign((size_t)n, -1);
    pg.out_next.clear(); pg.in_next.clear();
    for (int32_t pid = 0; pid < n; pid++) {   // djfind.rs:92-118: from nodes in id order, right edges in base order
        const int32_t i = pg.node_of_pid[(size_t)pid];
        for (int64_t e = g.row_ptr[i]; e < g.row_ptr[i + 1]; e++) {
            const int32_t j = g.col[e];
            const double w = wtab[(size_t)g.counts[i] + (size_t)g.counts[j]];   // -ln((from_cov + to_cov) / 2.0), djfind.rs:103-106
            const int32_t to_pid = pg.pid_of_node[(size_t)j];
            const int32_t id = (int32_t)pg.e_from.size();
            pg.e_from.push_back(pid);
            pg.e_to.push_back(to_pid);
            pg.e_w.push_back(w);
            pg.out_next.push_back(pg.out_head[(size_t)pid]);
            pg.out_head[(size_t)pid] = id;
            pg.in_next.push_back(pg.in_head[(size_t)to_pid]);
            pg.in_head[(size_t)to_pid] = id;
        }
    }
}

// Rust std::collections::BinaryHeap<MinScored<f64, NodeIndex>>: a max-heap under MinScored's reversed order, i.e. the smallest
// score on top; equal scores compare Equal (the node does not take part).
struct MinScoredHeap {
    struct Item {
        double score;
        int32_t node;
    };
    std::vector<Item> d;
    // MinScored::cmp(a, b): Equal if a == b, Greater if a < b, Less if a > b (NaN never occurs: counts are >= 1)
    static bool le(const Item &a, const Item &b) { return !(a.score < b.score); }   // a <= b  <=>  cmp != Greater
    void sift_up(size_t start, size_t pos)
    {
        const Item el = d[pos];
        while (pos > start) {
            const size_t parent = (pos - 1) / 2;
            if (le(el, d[parent])) break;
            d[pos] = d[parent];
            pos = parent;
        }
        d[pos] = el;
    }
    void push(double s, int32_t v)
    {
        d.push_back(Item{s, v});
        sift_up(0, d.size() - 1);
    }
    bool pop(Item &out)
    {
        if (d.empty()) return false;
        Item item = d.back();
        d.pop_back();
        if (!d.empty()) {
            std::swap(item, d[0]);
            // sift_down_to_bottom(0)
            const size_t end = d.size();
            size_t pos = 0;
            const Item el = d[0];
            size_t child = 1;
            while (child <= (end >= 2 ? end - 2 : 0) && end >= 2) {
                if (le(d[child], d[child + 1])) child++;
                d[pos] = d[child];
                pos = child;
                child = 2 * pos + 1;
            }
            if (child == end - 1) {
                d[pos] = d[child];
                pos = child;
            }
            d[pos] = el;
            sift_up(0, pos);
        }
        out = item;
        return true;
    }
};

// petgraph::algo::dijkstra(graph, start, None, |e| *e.weight()): scores of every node that was reached
inline void petgraph_dijkstra(const PetGraph &pg, int32_t start, std::vector<double> &score, std::vector<uint8_t> &has)
{
    const double none = std::numeric_limits<double>::quiet_NaN();
    score.assign((size_t)pg.n, none);
    has.assign((size_t)pg.n, 0);
    static thread_local std::vector<uint8_t> tl_visited;   // (one TLS lookup per call: the loop below works on references)
    static thread_local MinScoredHeap tl_heap;
    std::vector<uint8_t> &visited = tl_visited;
    MinScoredHeap &heap = tl_heap;
    visited.assign((size_t)pg.n, 0);
    heap.d.clear();
    score[(size_t)start] = 0.0;
    has[(size_t)start] = 1;
    heap.push(0.0, start);
    MinScoredHeap::Item it;
    while (heap.pop(it)) {
        const int32_t node = it.node;
        if (visited[(size_t)node]) continue;
        for (int32_t e = pg.out_head[(size_t)node]; e >= 0; e = pg.out_next[(size_t)e]) {   // graph.edges(node): newest edge first
            const int32_t next = pg.e_to[(size_t)e];
            if (visited[(size_t)next]) continue;
            const double next_score = it.score + pg.e_w[(size_t)e];
            if (has[(size_t)next]) {
                if (next_score < score[(size_t)next]) {
                    score[(size_t)next] = next_score;
                    heap.push(next_score, next);
                }
            } else {
                score[(size_t)next] = next_score;
                has[(size_t)next] = 1;
                heap.push(next_score, next);
            }
        }
        visited[(size_t)node] = 1;
    }
}

// the custom back-trace of djfind.rs:176-228 from `end` to `start` over the scores of one dijkstra run; false when it breaks off
inline bool trace_back(const PetGraph &pg, int32_t start, int32_t end, const std::vector<double> &dist, const std::vector<uint8_t> &has,
                       std::vector<int32_t> &path)
{
    const size_t MAX_ITERATIONS = 1000;
    const double FLOAT_EPSILON = 1e-9;
    path.assign(1, end);
    int32_t current = end;
    bool path_valid = false;
    size_t iterations = 0;
    while (current != start) {
        iterations++;
        if (iterations > MAX_ITERATIONS) break;
        int32_t best_prev = -1;
        double best_dist = std::numeric_limits<double>::infinity();
        const double current_dist = dist[(size_t)current];
        for (int32_t e = pg.in_head[(size_t)current]; e >= 0; e = pg.in_next[(size_t)e]) {   // neighbors_directed(Incoming)
            const int32_t neighbor = pg.e_from[(size_t)e];
            if (!has[(size_t)neighbor]) continue;
            // graph.find_edge(neighbor, current): the first match in neighbor's outgoing list (k-mer graphs have no
            // parallel edges, so it is this edge's weight)
            double edge_weight = pg.e_w[(size_t)e];
            for (int32_t f = pg.out_head[(size_t)neighbor]; f >= 0; f = pg.out_next[(size_t)f])
                if (pg.e_to[(size_t)f] == current) {
                    edge_weight = pg.e_w[(size_t)f];
                    break;
                }
            const double neighbor_dist = dist[(size_t)neighbor];
            const double expected_dist = neighbor_dist + edge_weight;
            if (std::fabs(expected_dist - current_dist) < FLOAT_EPSILON) {
                if (neighbor_dist < best_dist) {
                    best_dist = neighbor_dist;
                    best_prev = neighbor;
                }
            }
        }
        if (best_prev < 0) break;
        path.push_back(best_prev);
        current = best_prev;
        if (current == start) path_valid = true;
    }
    return path_valid;
}

// djfind.rs:157-247.  start_nodes / end_nodes are pids in node-index order.  Returns false when no valid path exists.
inline bool find_shortest_path(const PetGraph &pg, const std::vector<int32_t> &start_nodes, const std::vector<int32_t> &end_nodes,
                               std::vector<int32_t> &best_path, double &best_weight)
{
    bool found = false;
    double min_total_weight = std::numeric_limits<double>::infinity();
    static thread_local std::vector<double> tl_dist;
    static thread_local std::vector<uint8_t> tl_has;
    static thread_local std::vector<int32_t> tl_path;
    std::vector<double> &dist = tl_dist;
    std::vector<uint8_t> &has = tl_has;
    std::vector<int32_t> &path = tl_path;
    for (const int32_t start : start_nodes) {
        petgraph_dijkstra(pg, start, dist, has);
        for (const int32_t end : end_nodes) {
            if (!has[(size_t)end]) continue;
            const double total_weight = dist[(size_t)end];
            if (!(total_weight < min_total_weight)) continue;
            if (trace_back(pg, start, end, dist, has, path)) {
                best_path.assign(path.rbegin(), path.rend());
                best_weight = total_weight;
                min_total_weight = total_weight;
                found = true;
            }
        }
    }
    return found;
}

// djfind.rs:60-74: first K-mer, then the last base of every further node
inline std::string path_sequence(const GroupGraph &g, const PetGraph &pg, const std::vector<int32_t> &path)
{
    if (path.empty()) return std::string();
    std::string s = kmer_string(g, pg.node_of_pid[(size_t)path[0]]);
    for (size_t t = 1; t < path.size(); t++) s.push_back(base_char(kmer_base(g, pg.node_of_pid[(size_t)path[t]], g.K - 1)));
    return s;
}

struct PathResult {
    bool ok = false;
    std::string sequence;
    double total_weight = 0.0;
    size_t n_nodes = 0;
};

// djfind.rs:257-304 on a graph that is already converted
inline PathResult path_between(const GroupGraph &g, const PetGraph &pg, const std::string &start_anchor, const std::string &end_anchor)
{
    PathResult r;
    std::vector<int32_t> starts, ends;
    for (int32_t pid = 0; pid < pg.n; pid++) {
        const int32_t i = pg.node_of_pid[(size_t)pid];
        if (kmer_starts_with(g, i, start_anchor)) starts.push_back(pid);
        if (kmer_ends_with(g, i, end_anchor)) ends.push_back(pid);
    }
    if (starts.empty() || ends.empty()) return r;
    std::vector<int32_t> path;
    double w = 0.0;
    if (!find_shortest_path(pg, starts, ends, path, w)) return r;
    r.ok = true;
    r.sequence = path_sequence(g, pg, path);
    r.total_weight = w;
    r.n_nodes = path.size();
    return r;
}

// Without a boomphf rank the node ids are the ascending order and the container-free procedure of pathfind_core.h applies -- the
// code the GPU runs per group (k_pathfind); with a rank the petgraph restatement above is used.
struct CoreScratch {
    std::vector<int32_t> path, best_path, list_a, list_b;
    std::vector<rogtk_pf::NodeRec> node;
    std::vector<rogtk_pf::EdgeRec> edge;
    std::vector<rogtk_pf::HeapItem> heap;
    std::vector<char> out;
    rogtk_pf::Scratch bind(int32_t n, int32_t m, int K)
    {
        const size_t pn = (size_t)std::min<int32_t>(n, rogtk_pf::PF_MAX_ITERATIONS + 1) + 1;
        node.resize((size_t)n); edge.resize((size_t)m + 1); heap.resize((size_t)m + 2);
        path.resize(pn); best_path.resize(pn); list_a.resize((size_t)n); list_b.resize((size_t)n);
        out.resize((size_t)n + (size_t)K + 8);
        return rogtk_pf::Scratch{node.data(), edge.data(), heap.data(), path.data(), best_path.data(), list_a.data(), list_b.data()};
    }
};
inline std::string core_contig(const GroupGraph &g, const std::string *start_anchor, const std::string *end_anchor)
{
    static thread_local CoreScratch tl;
    CoreScratch &cs = tl;
    rogtk_pf::View v{g.K, g.n, g.lo, g.hi, g.counts, g.row_ptr, g.col, edge_weight_table()};
    const rogtk_pf::Scratch s = cs.bind(g.n, rogtk_pf::edge_count(v), g.K);
    const int32_t len = start_anchor ? rogtk_pf::anchored_contig(v, s, start_anchor->data(), (int)start_anchor->size(), end_anchor->data(),
                                                                 (int)end_anchor->size(), cs.out.data())
                                     : rogtk_pf::auto_contig(v, s, cs.out.data());
    return std::string(cs.out.data(), (size_t)len);
}

// AssemblyMethod::ShortestPath (reference src/fracture.rs:385-416): the contig, or "" when anchors / path are missing
inline std::string shortest_path_contig(const GroupGraph &g, const std::string &start_anchor, const std::string &end_anchor)
{
    if (g.n == 0) return std::string();   // (the reference fails earlier: no k-mers, empty result either way)
    if (!g.rank) return core_contig(g, &start_anchor, &end_anchor);
    static thread_local PetGraph tl_pg;   // (vectors keep their capacity from group to group)
    PetGraph &pg = tl_pg;
    build_petgraph(g, pg);
    const PathResult r = path_between(g, pg, start_anchor, end_anchor);
    return r.ok ? r.sequence : std::string();
}

// AssemblyMethod::ShortestPathAuto (reference src/djfind.rs:309-492, src/fracture.rs:419-461)
inline std::string shortest_path_auto_contig(const GroupGraph &g)
{
    if (g.n == 0) return std::string();
    if (!g.rank) return core_contig(g, nullptr, nullptr);
    static thread_local PetGraph tl_pg;
    PetGraph &pg = tl_pg;
    build_petgraph(g, pg);
    // find_endpoint_candidates
    double sum = 0.0;
    for (int32_t pid = 0; pid < pg.n; pid++) sum += (double)g.counts[pg.node_of_pid[(size_t)pid]];
    const double avg = sum / (double)pg.n;
    const double thr_f = std::fmax(avg * 0.1, 1.0);
    const uint16_t thr = thr_f >= 65535.0 ? 65535 : (uint16_t)thr_f;   // `as u16` truncates (and saturates)
    std::vector<int32_t> in_deg((size_t)pg.n, 0);
    for (size_t e = 0; e < pg.e_to.size(); e++) in_deg[(size_t)pg.e_to[e]]++;
    std::vector<int32_t> start_c, end_c;
    for (int32_t pid = 0; pid < pg.n; pid++) {
        const int32_t i = pg.node_of_pid[(size_t)pid];
        if (g.counts[i] < thr) continue;
        const int32_t in_degree = in_deg[(size_t)pid];
        const int32_t out_degree = (int32_t)(g.row_ptr[i + 1] - g.row_ptr[i]);
        if (in_degree == 0 && out_degree > 0) start_c.push_back(pid);
        if (out_degree == 0 && in_degree > 0) end_c.push_back(pid);
    }
    if (start_c.empty() || end_c.empty()) return std::string();
    if (start_c.size() == 1 && end_c.size() == 1) {
        const PathResult r = path_between(g, pg, kmer_string(g, pg.node_of_pid[(size_t)start_c[0]]),
                                          kmer_string(g, pg.node_of_pid[(size_t)end_c[0]]));
        return r.ok ? r.sequence : std::string();
    }
    // find_best_endpoint_pair: the first 100 pairs, best score wins (strictly greater replaces)
    const size_t MAX_PAIRS = 100;
    size_t evaluated = 0;
    bool have = false;
    double best_score = 0.0;
    std::string best_seq;
    std::vector<double> dist;
    std::vector<uint8_t> has;
    std::vector<int32_t> rpath, path;
    for (const int32_t s : start_c) {
        bool have_dist = false;   // (the reference runs dijkstra(s) again for every t; the scores are the same)
        for (const int32_t t : end_c) {
            if (evaluated >= MAX_PAIRS) break;
            evaluated++;
            // (the reference matches the two full K-mer strings with contains(): only the candidates themselves qualify)
            if (!have_dist) {
                petgraph_dijkstra(pg, s, dist, has);
                have_dist = true;
            }
            if (!has[(size_t)t]) continue;
            const double w = dist[(size_t)t];
            if (!(w < std::numeric_limits<double>::infinity())) continue;   // find_shortest_path: total_weight < min_total_weight (inf)
            if (!trace_back(pg, s, t, dist, has, rpath)) continue;
            path.assign(rpath.rbegin(), rpath.rend());
            // score_path
            const double path_length = (double)(path.size() * (size_t)g.K);
            const double mean_coverage = 1.0 / (w / (double)path.size());
            const double normalized_length = std::fmin(path_length / 5000.0, 1.0);
            const double normalized_coverage = std::fmin(mean_coverage / 100.0, 1.0);
            const double score = 0.6 * normalized_length + 0.4 * normalized_coverage;
            if (!have || score > best_score) {
                have = true;
                best_score = score;
                best_seq = path_sequence(g, pg, path);
            }
        }
    }
    return have ? best_seq : std::string();
}

}  // namespace rogtk_host
