// =============================================================================
// gv_graph.cpp -- host-side Graph and AliasTable builder of libgv_b200.
//
// Mirrors graphvite::Graph<uint32> (reference include/instance/graph.cuh:62-277,
// include/core/graph.h:87-101) and AliasTable::build
// (include/base/alias_table.cuh:84-128) behind the C ABI of include/gv_b200.h.
// Storage differs from the reference (append-only edge log + CSR instead of a
// vector of adjacency vectors) but every observable -- first-seen vertex ids,
// adjacency order, float accumulation order of the weights, flatten() order --
// is identical.
// =============================================================================
#include "gv_host.h"

#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <deque>
#include <sstream>
#include <thread>

namespace gv {

// ---- AliasTable::build, base/alias_table.cuh:84-128 ---------------------------------
// Vose's method with two FIFO queues; the order in which entries are paired decides the
// table, so it is kept exactly (the queues are flat arrays walked by a cursor, which is
// the same FIFO order as std::queue).
template<class I>
void build_alias(const float *weights, size_t count, float *prob, I *alias) {
    if (count == 0)
        throw std::runtime_error("Invalid sampling distribution");
    double norm = 0;  // accumulated in double: alias_table.cuh:92
    for (size_t i = 0; i < count; i++)
        norm += weights[i];
    norm = norm / count;
    size_t num_little = 0;
    for (size_t i = 0; i < count; i++) {
        prob[i] = float(double(weights[i]) / norm);
        num_little += prob[i] < 1;
    }
    // One of the queues starts empty (always the case for uniform weights, i.e. every table of an unweighted
    // graph): the pairing loop never runs and every entry is a leftover that aliases to itself.
    if (num_little == 0 || num_little == count) {
        for (size_t i = 0; i < count; i++)
            alias[i] = I(i);
        return;
    }
    std::vector<I> little, large;
    little.reserve(count);
    large.reserve(count);
    for (size_t i = 0; i < count; i++) {
        if (prob[i] < 1)
            little.push_back(I(i));
        else
            large.push_back(I(i));
    }
    // Every pop from `large` is followed by exactly one push to one of the queues, so the
    // two cursors never run past entries that have not been written yet.
    size_t little_head = 0, large_head = 0;
    while (little_head < little.size() && large_head < large.size()) {
        const I i = little[little_head++], j = large[large_head++];
        alias[i] = j;
        const float sum = prob[i] + prob[j];
        prob[j] = sum - 1;
        if (prob[j] < 1)
            little.push_back(j);
        else
            large.push_back(j);
    }
    // leftovers alias to themselves ("suppress some truncation error", :117-127)
    for (; little_head < little.size(); little_head++)
        alias[little[little_head]] = little[little_head];
    for (; large_head < large.size(); large_head++)
        alias[large[large_head]] = large[large_head];
}

template void build_alias<uint32_t>(const float *, size_t, float *, uint32_t *);
template void build_alias<uint64_t>(const float *, size_t, float *, uint64_t *);

// ---- Graph -----------------------------------------------------------------------------
void Graph::clear() {
    *this = Graph();
}

uint32_t Graph::intern(const char *name, size_t length, uint64_t hash) {
    bool created;
    const uint32_t id = names.intern(name, length, hash, id2name, created);
    if (created) {
        vertex_weights.push_back(0.f);
        degrees.push_back(0);
    }
    return id;
}

uint32_t Graph::intern(const std::string &name) {
    return intern(name.data(), name.size(), NameTable::hash(name.data(), name.size()));
}

// Graph::add_edge, instance/graph.cuh:124-153
void Graph::add_edge(const std::string &u_name, const std::string &v_name, float w) {
    const uint32_t u = intern(u_name);
    const uint32_t v = intern(v_name);
    add_edge_ids(u, v, w);
}

void Graph::add_edge_ids(uint32_t u, uint32_t v, float w) {
    log_u.push_back(u);
    log_v.push_back(v);
    log_w.push_back(w);
    degrees[u]++;
    vertex_weights[u] += w;
    if (as_undirected && u != v) {
        log_u.push_back(v);
        log_v.push_back(u);
        log_w.push_back(w);
        degrees[v]++;
        vertex_weights[v] += w;
    }
    num_edge++;  // input lines, not directed edges (:152)
    flattened = false;
}

// GraphMixin::flatten, core/graph.h:87-101: edges grouped by source vertex, insertion order
// kept inside a vertex -- a stable counting sort of the edge log.
void Graph::flatten() {
    if (flattened)
        return;
    const size_t n = id2name.size(), m = log_u.size();
    offsets.assign(n + 1, 0);
    for (size_t v = 0; v < n; v++)
        offsets[v + 1] = offsets[v] + degrees[v];
    edge_u.resize(m);
    edge_v.resize(m);
    edge_w.resize(m);
    // A stable counting sort of the edge log by source vertex.  The scatter is random over the whole output, so big
    // logs are split by vertex range: thread t scans the log and places the edges whose source lies in its range
    // (ranges of about equal edge count) -- sequential reads, writes confined to 1/T of the arrays, same result.
    const size_t threads = m < (size_t(1) << 22) ? 1 : std::min<size_t>(8, std::max(1u, std::thread::hardware_concurrency()));
    std::vector<uint64_t> cursor(offsets.begin(), offsets.end() - 1);
    auto place = [&](uint32_t first, uint32_t last) {  // sources in [first, last)
        for (size_t e = 0; e < m; e++) {
            const uint32_t u = log_u[e];
            if (u < first || u >= last)
                continue;
            const uint64_t slot = cursor[u]++;
            edge_u[slot] = u;
            edge_v[slot] = log_v[e];
            edge_w[slot] = log_w[e];
        }
    };
    if (threads == 1)
        place(0, uint32_t(n));
    else {
        std::vector<uint32_t> bound(threads + 1, uint32_t(n));
        bound[0] = 0;
        for (size_t t = 1; t < threads; t++)
            bound[t] = uint32_t(std::lower_bound(offsets.begin(), offsets.end(), uint64_t(m) * t / threads) - offsets.begin());
        std::vector<std::thread> pool;
        for (size_t t = 1; t < threads; t++)
            pool.emplace_back(place, bound[t], bound[t + 1]);
        place(bound[0], bound[1]);
        for (auto &thread : pool)
            thread.join();
    }
    flattened = true;
    uniform_cache = -1;
}

// Graph::normalize, instance/graph.cuh:103-121
void Graph::normalize() {
    flatten();
    const size_t n = id2name.size();
    std::vector<float> context_weights(n, 0.f);
    for (size_t e = 0; e < edge_u.size(); e++)
        context_weights[edge_v[e]] += edge_w[e];
    for (size_t u = 0; u < n; u++) {
        float weight = 0;
        for (uint64_t e = offsets[u]; e < offsets[u + 1]; e++) {
            edge_w[e] /= std::sqrt(vertex_weights[u] * context_weights[edge_v[e]]);
            weight += edge_w[e];
        }
        vertex_weights[u] = weight;
    }
    uniform_cache = -1;
    // the edge log is no longer consulted once flattened
}

// Graph::load_file, instance/graph.cuh:163-201.  Same lines, tokens, ids and errors as the fgets / strtok-style loop
// it replaces; the file is read in large chunks, a batch of lines is tokenised and its names hashed (prefetching their
// table slots) before the batch is resolved in order.
void Graph::load_file(const char *file_name, bool undirected, bool normalized, const char *delimiters,
                      const char *comment) {
    clear();
    as_undirected = undirected;
    normalization = normalized;
    FILE *fin = fopen(file_name, "rb");
    if (!fin)
        throw std::runtime_error(std::string("File `") + file_name + "` doesn't exist");
    struct {  // GV_LOG=2: phase timings on stderr
        std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
        bool on = getenv("GV_LOG") != nullptr && atoi(getenv("GV_LOG")) >= 2;
        void mark(const char *what) {
            if (on) {
                const auto now = std::chrono::steady_clock::now();
                fprintf(stderr, "[gv] %-28s %8.3f s\n", what, std::chrono::duration<double>(now - last).count());
                last = now;
            }
        }
    } phase;
    bool is_delimiter[256] = {false};
    for (const char *d = delimiters; *d; d++)
        is_delimiter[uint8_t(*d)] = true;
    const size_t comment_length = strlen(comment);
    struct Pending {
        const char *u, *v;
        uint32_t u_length, v_length;
        uint64_t u_hash, v_hash;
        float w;
    };
    constexpr size_t kBatch = 64;
    Pending pending[kBatch];
    size_t num_pending = 0;
    uint32_t ids[kBatch][2];
    auto resolve = [&]() {
        // ids in order of first appearance, then the per-vertex counters (scattered too: fetched ahead of their use)
        for (size_t i = 0; i < num_pending; i++) {
            const Pending &edge = pending[i];
            ids[i][0] = intern(edge.u, edge.u_length, edge.u_hash);
            ids[i][1] = intern(edge.v, edge.v_length, edge.v_hash);
            __builtin_prefetch(&degrees[ids[i][0]], 1);
            __builtin_prefetch(&vertex_weights[ids[i][0]], 1);
            if (as_undirected) {
                __builtin_prefetch(&degrees[ids[i][1]], 1);
                __builtin_prefetch(&vertex_weights[ids[i][1]], 1);
            }
        }
        for (size_t i = 0; i < num_pending; i++)
            add_edge_ids(ids[i][0], ids[i][1], pending[i].w);
        num_pending = 0;
    };
    {  // room for the edge log: ~12 bytes per line in a typical edge list
        struct stat status;
        if (fstat(fileno(fin), &status) == 0 && status.st_size > 0) {
            const size_t lines = size_t(status.st_size) / 10, directed = lines * (undirected ? 2 : 1);
            log_u.reserve(directed);
            log_v.reserve(directed);
            log_w.reserve(directed);
        }
    }
    // bytes read at a time (GV_LOAD_CHUNK: a small value makes every line straddle a chunk boundary -- tests)
    const size_t kChunk = getenv("GV_LOAD_CHUNK") ? std::max<size_t>(16, strtoull(getenv("GV_LOAD_CHUNK"), nullptr, 10))
                                                  : size_t(64) << 20;
    std::vector<char> buffer(kChunk + 1);
    size_t held = 0, line_no = 0;  // bytes of an incomplete line carried over from the previous chunk
    bool at_end = false;
    while (!at_end) {
        if (held == buffer.size() - 1)  // a line longer than the buffer
            buffer.resize(buffer.size() * 2);
        const size_t got = fread(buffer.data() + held, 1, buffer.size() - 1 - held, fin);
        at_end = got == 0;
        const size_t filled = held + got;
        const char *cursor = buffer.data(), *limit = buffer.data() + filled;
        while (cursor < limit) {
            const char *newline = static_cast<const char *>(memchr(cursor, '\n', size_t(limit - cursor)));
            if (!newline && !at_end)
                break;  // the rest of this line is in the next chunk
            const char *line_end = newline ? newline + 1 : limit;  // fgets keeps the newline: it is part of the line
            line_no++;
            const char *end = line_end;
            if (const void *nul = memchr(cursor, 0, size_t(end - cursor)))  // a C string ends at its first NUL
                end = static_cast<const char *>(nul);
            if (comment_length) {
                const void *cut = memmem(cursor, size_t(end - cursor), comment, comment_length);
                if (cut)
                    end = static_cast<const char *>(cut);
            }
            int num_token = 0;
            bool bad = false;
            Pending edge;
            edge.w = 1;
            for (const char *c = cursor; c < end;) {
                while (c < end && is_delimiter[uint8_t(*c)])
                    c++;
                if (c == end)
                    break;
                const char *token = c;
                while (c < end && !is_delimiter[uint8_t(*c)])
                    c++;
                if (num_token == 0)
                    edge.u = token, edge.u_length = uint32_t(c - token);
                else if (num_token == 1)
                    edge.v = token, edge.v_length = uint32_t(c - token);
                else if (num_token == 2)
                    edge.w = float(atof(std::string(token, size_t(c - token)).c_str()));
                else
                    bad = true;
                num_token++;
            }
            cursor = line_end;
            if (num_token == 0)
                continue;
            if (num_token < 2 || bad) {
                fclose(fin);
                throw std::runtime_error("Invalid format at line " + std::to_string(line_no));
            }
            edge.u_hash = NameTable::hash(edge.u, edge.u_length);
            edge.v_hash = NameTable::hash(edge.v, edge.v_length);
            names.prefetch(edge.u_hash);
            names.prefetch(edge.v_hash);
            pending[num_pending++] = edge;
            if (num_pending == kBatch)
                resolve();
        }
        resolve();  // the names point into the buffer, which is about to move
        held = size_t(limit - cursor);
        memmove(buffer.data(), cursor, held);
    }
    fclose(fin);
    phase.mark("graph file: parse + intern");
    flatten();
    phase.mark("graph file: flatten");
    if (normalization)
        normalize();
}

// WordGraph::load_file_compact, instance/word_graph.cuh:75-166: a word co-occurrence graph from a corpus.
// Pass 1 counts the words (ids in order of first appearance) and drops those rarer than min_count; pass 2 counts
// every pair of kept words at distance <= window inside a line, in both directions.  The out-edges of a vertex
// are emitted in the iteration order of the reference's std::unordered_map<Index, float> -- the very same
// container is filled with the very same insertion sequence here, so the order (libstdc++'s) is reproduced.
void Graph::load_corpus(const char *file_name, int window, int min_count, bool normalized, const char *delimiters,
                        const char *comment) {
    clear();
    as_undirected = true;
    normalization = normalized;
    FILE *fin = fopen(file_name, "r");
    if (!fin)
        throw std::runtime_error(std::string("File `") + file_name + "` doesn't exist");
    const size_t kMaxLineLength = size_t(1) << 22;  // util/common.h:30
    std::vector<char> line(kMaxLineLength);
    auto for_each_word = [&](const std::function<void(const std::string &)> &visit) {
        char *cut = strstr(line.data(), comment);  // an empty prefix cuts the whole line, as in the reference
        if (cut)
            *cut = 0;
        std::string word;
        for (char *cursor = line.data(); *cursor;) {
            cursor += strspn(cursor, delimiters);
            if (!*cursor)
                break;
            const size_t length = strcspn(cursor, delimiters);
            word.assign(cursor, length);
            visit(word);
            cursor += length;
        }
    };
    std::vector<uint32_t> frequency;
    std::vector<std::string> words;
    std::unordered_map<std::string, uint32_t> word2id;
    while (fgets(line.data(), int(kMaxLineLength), fin))
        for_each_word([&](const std::string &word) {
            auto found = word2id.find(word);
            if (found != word2id.end())
                frequency[found->second]++;
            else {
                word2id.emplace(word, uint32_t(words.size()));
                words.push_back(word);
                frequency.push_back(1);
            }
        });
    for (size_t i = 0; i < words.size(); i++)
        if (int64_t(frequency[i]) >= int64_t(min_count))
            intern(words[i]);
    const uint32_t n = num_vertex();
    fseek(fin, 0, SEEK_SET);
    std::vector<std::unordered_map<uint32_t, float>> edge_map(n);
    std::vector<uint32_t> sentence;
    while (fgets(line.data(), int(kMaxLineLength), fin)) {
        sentence.clear();
        for_each_word([&](const std::string &word) {
            const uint32_t found = names.find(word, id2name);
            if (found != NameTable::kNone)
                sentence.push_back(found);
        });
        for (size_t i = 0; i < sentence.size(); i++)
            for (int j = 1; j <= window && i + j < sentence.size(); j++) {
                const uint32_t u = sentence[i], v = sentence[i + j];
                auto edge = edge_map[u].find(v);
                if (edge == edge_map[u].end())
                    edge_map[u][v] = 1;
                else
                    edge->second++;
                edge = edge_map[v].find(u);
                if (edge == edge_map[v].end())
                    edge_map[v][u] = 1;
                else
                    edge->second++;
                vertex_weights[u]++;
                vertex_weights[v]++;
            }
    }
    fclose(fin);
    for (uint32_t u = 0; u < n; u++) {
        for (const auto &edge : edge_map[u]) {
            log_u.push_back(u);
            log_v.push_back(edge.first);
            log_w.push_back(edge.second);
        }
        degrees[u] = uint32_t(edge_map[u].size());
        num_edge += edge_map[u].size();
    }
    flatten();
    if (normalization)
        normalize();
}

// Graph::load_edge_list / load_weighted_edge_list, instance/graph.cuh:209-252
void Graph::load_edges(const char *const *u_names, const char *const *v_names, const float *weights,
                       uint64_t count, bool undirected, bool normalized) {
    clear();
    as_undirected = undirected;
    normalization = normalized;
    for (uint64_t i = 0; i < count; i++)
        add_edge(u_names[i], v_names[i], weights ? weights[i] : 1.f);
    flatten();
    if (normalization)
        normalize();
}

// Binary edge arrays: the graph load_edges() would build from the edge list [(str(u[i]), str(v[i]))] -- internal ids
// in order of first appearance, the same edge log, degrees, weights and line count (instance/graph.cuh:124-153,
// 209-252) -- without creating or hashing 2 * count strings: a Friendster-sized edge list (1.8e9 lines) is read
// from two uint32 arrays in one pass.  name2id stays empty (gv_graph_name2id answers from id_of_original).
void Graph::load_id_edges(const uint32_t *u, const uint32_t *v, const float *weights, uint64_t count, bool undirected,
                          bool normalized) {
    clear();
    as_undirected = undirected;
    normalization = normalized;
    uint32_t bound = 0;
    for (uint64_t i = 0; i < count; i++)
        bound = std::max(bound, std::max(u[i], v[i]));
    id_of_original.assign(count ? size_t(bound) + 1 : 0, -1);
    std::vector<uint32_t> original_of;
    auto intern_id = [&](uint32_t original) {
        int64_t &slot = id_of_original[original];
        if (slot < 0) {
            slot = int64_t(original_of.size());
            original_of.push_back(original);
        }
        return uint32_t(slot);
    };
    const uint64_t directed = undirected ? 2 * count : count;
    log_u.reserve(directed);
    log_v.reserve(directed);
    log_w.reserve(directed);
    for (uint64_t i = 0; i < count; i++) {
        const uint32_t a = intern_id(u[i]), b = intern_id(v[i]);
        const float w = weights ? weights[i] : 1.f;
        log_u.push_back(a);
        log_v.push_back(b);
        log_w.push_back(w);
        if (undirected && a != b) {
            log_u.push_back(b);
            log_v.push_back(a);
            log_w.push_back(w);
        }
    }
    num_edge = count;
    const size_t n = original_of.size();
    degrees.assign(n, 0);
    vertex_weights.assign(n, 0.f);
    // weighted degrees accumulate in edge order per vertex, like add_edge's `vertex_weights[u] += w`
    for (size_t e = 0; e < log_u.size(); e++) {
        degrees[log_u[e]]++;
        vertex_weights[log_u[e]] += log_w[e];
    }
    id2name.resize(n);
    for (size_t i = 0; i < n; i++)
        id2name[i] = std::to_string(original_of[i]);
    flatten();
    if (normalization)
        normalize();
}

// Graph::save, instance/graph.cuh:260-277
void Graph::save(const char *file_name, bool weighted, bool anonymous) {
    flatten();
    FILE *fout = fopen(file_name, "w");
    if (!fout)
        throw std::runtime_error(std::string("Can't open `") + file_name + "` for writing");
    for (size_t e = 0; e < edge_u.size(); e++) {
        if (anonymous)
            fprintf(fout, "%llu\t%llu", (unsigned long long)edge_u[e], (unsigned long long)edge_v[e]);
        else
            fprintf(fout, "%s\t%s", id2name[edge_u[e]].c_str(), id2name[edge_v[e]].c_str());
        if (weighted)
            fprintf(fout, "\t%f", edge_w[e]);
        fputc('\n', fout);
    }
    fclose(fout);
}

bool Graph::uniform_edge_table(float &probability) {
    flatten();
    if (uniform_cache < 0) {
        const size_t m = edge_w.size();
        bool uniform = m > 0;
        for (size_t e = 1; e < m && uniform; e++)
            uniform = edge_w[e] == edge_w[0];
        if (uniform) {
            double norm = 0;  // AliasTable::build's own normaliser (base/alias_table.cuh:92), same order of additions
            for (size_t e = 0; e < m; e++)
                norm += edge_w[e];
            norm = norm / m;
            uniform_probability = float(double(edge_w[0]) / norm);
        }
        uniform_cache = uniform ? 1 : 0;
    }
    probability = uniform_probability;
    return uniform_cache == 1;
}

bool Graph::has_dead_end() const {
    for (size_t v = 0; v < degrees.size(); v++)
        if (degrees[v] == 0)
            return true;
    return false;
}

// Graph::info, instance/graph.cuh:88-101 + core/graph.h:110-123
std::string Graph::info() const {
    std::stringstream ss;
    ss << "Graph<uint32>" << std::endl;
    ss << "------------------ Graph -------------------" << std::endl;
    ss << "#vertex: " << num_vertex() << ", #edge: " << num_edge << std::endl;
    ss << "as undirected: " << (as_undirected ? "yes" : "no")
       << ", normalization: " << (normalization ? "yes" : "no");
    return ss.str();
}

}  // namespace gv

// =============================================================================
// C ABI
// =============================================================================
using gv::Graph;

struct gv_graph {
    Graph graph;
};

#define GV_TRY try {
#define GV_CATCH(ret)                  \
    }                                  \
    catch (const std::exception &e) {  \
        gv::set_error(e.what());       \
        return ret;                    \
    }

extern "C" {

gv_graph_t *gv_graph_create(void) {
    return new gv_graph();
}

void gv_graph_destroy(gv_graph_t *graph) {
    delete graph;
}

int gv_graph_load_file(gv_graph_t *graph, const char *file_name, int as_undirected, int normalization,
                       const char *delimiters, const char *comment) {
    GV_TRY
    graph->graph.load_file(file_name, as_undirected != 0, normalization != 0, delimiters ? delimiters : " \t\r\n",
                           comment ? comment : "#");
    return 0;
    GV_CATCH(-1)
}

int gv_graph_load_corpus(gv_graph_t *graph, const char *file_name, int window, int min_count, int normalization,
                         const char *delimiters, const char *comment) {
    GV_TRY
    graph->graph.load_corpus(file_name, window, min_count, normalization != 0, delimiters ? delimiters : " \t\r\n",
                             comment ? comment : "#");
    return 0;
    GV_CATCH(-1)
}

int gv_graph_load_edges(gv_graph_t *graph, const char *const *u_names, const char *const *v_names,
                        const float *weights, uint64_t num_edge, int as_undirected, int normalization) {
    GV_TRY
    graph->graph.load_edges(u_names, v_names, weights, num_edge, as_undirected != 0, normalization != 0);
    return 0;
    GV_CATCH(-1)
}

int gv_graph_save(gv_graph_t *graph, const char *file_name, int weighted, int anonymous) {
    GV_TRY
    graph->graph.save(file_name, weighted != 0, anonymous != 0);
    return 0;
    GV_CATCH(-1)
}

uint64_t gv_graph_num_vertex(const gv_graph_t *graph) {
    return graph->graph.num_vertex();
}

uint64_t gv_graph_num_edge(const gv_graph_t *graph) {
    return graph->graph.num_edge;
}

int gv_graph_as_undirected(const gv_graph_t *graph) {
    return graph->graph.as_undirected;
}

int gv_graph_normalization(const gv_graph_t *graph) {
    return graph->graph.normalization;
}

const char *gv_graph_id2name(const gv_graph_t *graph, uint64_t id) {
    if (id >= graph->graph.id2name.size())
        return nullptr;
    return graph->graph.id2name[id].c_str();
}

int64_t gv_graph_name2id(const gv_graph_t *graph, const char *name) {
    const Graph &g = graph->graph;
    if (!g.id_of_original.empty()) {  // loaded from id arrays: the name is the decimal id
        char *end = nullptr;
        const unsigned long long original = strtoull(name, &end, 10);
        if (!name[0] || *end || original >= g.id_of_original.size() || std::to_string(original) != name)
            return -1;
        return g.id_of_original[original];
    }
    const uint32_t found = g.names.find(name, strlen(name), gv::NameTable::hash(name, strlen(name)), g.id2name);
    return found == gv::NameTable::kNone ? -1 : int64_t(found);
}

int gv_graph_load_id_edges(gv_graph_t *graph, const uint32_t *u, const uint32_t *v, const float *weights,
                           uint64_t num_edge, int as_undirected, int normalization) {
    GV_TRY
    if (num_edge && (!u || !v))
        throw std::runtime_error("gv_graph_load_id_edges: null edge array");
    graph->graph.load_id_edges(u, v, weights, num_edge, as_undirected != 0, normalization != 0);
    return 0;
    GV_CATCH(-1)
}

uint64_t gv_graph_flatten(gv_graph_t *graph, uint32_t *u, uint32_t *v, float *w, uint64_t *flat_offsets,
                          float *vertex_weights) {
    Graph &g = graph->graph;
    g.flatten();
    const size_t m = g.edge_u.size(), n = g.num_vertex();
    if (u)
        memcpy(u, g.edge_u.data(), m * sizeof(uint32_t));
    if (v)
        memcpy(v, g.edge_v.data(), m * sizeof(uint32_t));
    if (w)
        memcpy(w, g.edge_w.data(), m * sizeof(float));
    if (flat_offsets)
        memcpy(flat_offsets, g.offsets.data(), n * sizeof(uint64_t));
    if (vertex_weights)
        memcpy(vertex_weights, g.vertex_weights.data(), n * sizeof(float));
    return m;
}

int gv_graph_info(const gv_graph_t *graph, char *buffer, size_t capacity) {
    const std::string info = graph->graph.info();
    if (buffer && capacity) {
        strncpy(buffer, info.c_str(), capacity - 1);
        buffer[capacity - 1] = 0;
    }
    return int(info.size());
}

int gv_alias_build(const float *weights, uint64_t n, float *prob, uint64_t *alias) {
    GV_TRY
    gv::build_alias<uint64_t>(weights, n, prob, alias);
    return 0;
    GV_CATCH(-1)
}

}  // extern "C"

gv::Graph &gv_graph_ref(gv_graph_t *graph) {
    return graph->graph;
}
