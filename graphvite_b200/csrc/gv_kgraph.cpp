// =============================================================================
// gv_kgraph.cpp -- host-side KnowledgeGraph of libgv_b200.
//
// Mirrors graphvite::KnowledgeGraph<uint32> (reference include/instance/knowledge_graph.cuh:67-284
// over include/core/graph.h:45-101) behind the C ABI of include/gv_b200.h.  Storage is an
// append-only triplet log plus a CSR built on demand; every observable -- first-seen entity /
// relation ids (head, then relation, then tail of a line), adjacency order, float accumulation
// order of the weights, flatten() order -- is identical to the reference's vector-of-vectors.
// =============================================================================
#include "gv_host.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace gv {

void KnowledgeGraph::clear() {
    *this = KnowledgeGraph();
}

uint32_t KnowledgeGraph::intern_entity(const std::string &name) {
    auto found = entity2id.find(name);
    if (found != entity2id.end())
        return found->second;
    const uint32_t id = uint32_t(id2entity.size());
    entity2id.emplace(name, id);
    id2entity.push_back(name);
    vertex_weights.push_back(0.f);
    degrees.push_back(0);
    return id;
}

uint32_t KnowledgeGraph::intern_relation(const std::string &name) {
    auto found = relation2id.find(name);
    if (found != relation2id.end())
        return found->second;
    const uint32_t id = uint32_t(id2relation.size());
    relation2id.emplace(name, id);
    id2relation.push_back(name);
    return id;
}

// KnowledgeGraph::add_edge, instance/knowledge_graph.cuh:135-168: ids are handed out in the order
// head, relation, tail; only the head's weight grows (triplets are directed).
void KnowledgeGraph::add_edge(const std::string &h_name, const std::string &r_name, const std::string &t_name,
                              float w) {
    const uint32_t h = intern_entity(h_name);
    const uint32_t r = intern_relation(r_name);
    const uint32_t t = intern_entity(t_name);
    log_h.push_back(h);
    log_t.push_back(t);
    log_r.push_back(r);
    log_w.push_back(w);
    degrees[h]++;
    vertex_weights[h] += w;
    num_edge++;
    flattened = false;
}

// GraphMixin::flatten, core/graph.h:87-101: a stable counting sort of the log by head entity
void KnowledgeGraph::flatten() {
    if (flattened)
        return;
    const size_t n = id2entity.size(), m = log_h.size();
    offsets.assign(n + 1, 0);
    for (size_t v = 0; v < n; v++)
        offsets[v + 1] = offsets[v] + degrees[v];
    std::vector<uint64_t> cursor(offsets.begin(), offsets.end() - 1);
    edge_h.resize(m);
    edge_t.resize(m);
    edge_r.resize(m);
    edge_w.resize(m);
    for (size_t e = 0; e < m; e++) {
        const uint64_t slot = cursor[log_h[e]]++;
        edge_h[slot] = log_h[e];
        edge_t[slot] = log_t[e];
        edge_r[slot] = log_r[e];
        edge_w[slot] = log_w[e];
    }
    flattened = true;
}

// KnowledgeGraph::normalize, instance/knowledge_graph.cuh:95-121: w /= sqrt(out-weight of (h, r) *
// in-weight of (t, r)), both accumulated in float in adjacency order.
void KnowledgeGraph::normalize() {
    flatten();
    const size_t n = id2entity.size();
    std::vector<std::unordered_map<uint32_t, float>> head_weights(n), tail_weights(n);
    for (size_t e = 0; e < edge_h.size(); e++) {
        head_weights[edge_h[e]][edge_r[e]] += edge_w[e];  // operator[] value-initialises to 0
        tail_weights[edge_t[e]][edge_r[e]] += edge_w[e];
    }
    for (size_t h = 0; h < n; h++) {
        float weight = 0;
        for (uint64_t e = offsets[h]; e < offsets[h + 1]; e++) {
            edge_w[e] /= std::sqrt(head_weights[h][edge_r[e]] * tail_weights[edge_t[e]][edge_r[e]]);
            weight += edge_w[e];
        }
        vertex_weights[h] = weight;
    }
}

// KnowledgeGraph::load_file, instance/knowledge_graph.cuh:177-213: `head relation tail [weight]`
void KnowledgeGraph::load_file(const char *file_name, bool normalized, const char *delimiters, const char *comment) {
    clear();
    normalization = normalized;
    FILE *fin = fopen(file_name, "r");
    if (!fin)
        throw std::runtime_error(std::string("File `") + file_name + "` doesn't exist");
    const size_t kMaxLineLength = size_t(1) << 22;  // util/common.h:30
    std::vector<char> line(kMaxLineLength);
    const size_t comment_length = strlen(comment);
    std::string names[3];
    for (size_t line_no = 1; fgets(line.data(), int(kMaxLineLength), fin); line_no++) {
        if (comment_length) {
            char *cut = strstr(line.data(), comment);
            if (cut)
                *cut = 0;
        }
        int num_token = 0;
        float w = 1;
        for (char *cursor = line.data(); *cursor;) {
            cursor += strspn(cursor, delimiters);
            if (!*cursor)
                break;
            const size_t length = strcspn(cursor, delimiters);
            if (num_token < 3)
                names[num_token].assign(cursor, length);
            else if (num_token == 3)
                w = float(atof(std::string(cursor, length).c_str()));
            num_token++;
            cursor += length;
        }
        if (num_token == 0)
            continue;
        if (num_token < 3 || num_token > 4) {
            fclose(fin);
            throw std::runtime_error("Invalid format at line " + std::to_string(line_no));
        }
        add_edge(names[0], names[1], names[2], w);
    }
    fclose(fin);
    flatten();
    if (normalization)
        normalize();
}

// load_triplet_list / load_weighted_triplet_list, instance/knowledge_graph.cuh:220-260
void KnowledgeGraph::load_triplets(const char *const *h_names, const char *const *r_names, const char *const *t_names,
                                   const float *weights, uint64_t count, bool normalized) {
    clear();
    normalization = normalized;
    for (uint64_t i = 0; i < count; i++)
        add_edge(h_names[i], r_names[i], t_names[i], weights ? weights[i] : 1.f);
    flatten();
    if (normalization)
        normalize();
}

// KnowledgeGraph::save, instance/knowledge_graph.cuh:267-283.  The reference writes the columns
// `head<TAB>tail<TAB>relation` (not the order load() reads) and takes the third column from the
// edge WEIGHT converted to an integer (std::get<1> instead of std::get<2>, :275) -- an out-of-range
// read of id2relation for most graphs.  We keep the column order and write the relation itself;
// the two agree exactly when every weight equals its relation id, which is what the parity test uses.
void KnowledgeGraph::save(const char *file_name, bool anonymous) {
    flatten();
    FILE *fout = fopen(file_name, "w");
    if (!fout)
        throw std::runtime_error(std::string("Can't open `") + file_name + "` for writing");
    for (size_t e = 0; e < edge_h.size(); e++) {
        if (anonymous)
            fprintf(fout, "%llu\t%llu\t%llu\n", (unsigned long long)edge_h[e], (unsigned long long)edge_t[e],
                    (unsigned long long)edge_r[e]);
        else
            fprintf(fout, "%s\t%s\t%s\n", id2entity[edge_h[e]].c_str(), id2entity[edge_t[e]].c_str(),
                    id2relation[edge_r[e]].c_str());
    }
    fclose(fout);
}

// KnowledgeGraph::name / graph_info, instance/knowledge_graph.cuh:123-133 + core/graph.h:117-123
std::string KnowledgeGraph::info() const {
    std::stringstream ss;
    ss << "KnowledgeGraph<uint32>" << std::endl;
    ss << "------------------ Graph -------------------" << std::endl;
    ss << "#entity: " << num_vertex() << ", #relation: " << num_relation() << std::endl;
    ss << "#triplet: " << num_edge << ", normalization: " << (normalization ? "yes" : "no");
    return ss.str();
}

}  // namespace gv

// =============================================================================
// C ABI
// =============================================================================
using gv::KnowledgeGraph;

struct gv_kgraph {
    KnowledgeGraph graph;
};

#define GV_TRY try {
#define GV_CATCH(ret)                  \
    }                                  \
    catch (const std::exception &e) {  \
        gv::set_error(e.what());       \
        return ret;                    \
    }

extern "C" {

gv_kgraph_t *gv_kgraph_create(void) {
    return new gv_kgraph();
}

void gv_kgraph_destroy(gv_kgraph_t *graph) {
    delete graph;
}

int gv_kgraph_load_file(gv_kgraph_t *graph, const char *file_name, int normalization, const char *delimiters,
                        const char *comment) {
    GV_TRY
    graph->graph.load_file(file_name, normalization != 0, delimiters ? delimiters : " \t\r\n", comment ? comment : "#");
    return 0;
    GV_CATCH(-1)
}

int gv_kgraph_load_triplets(gv_kgraph_t *graph, const char *const *h_names, const char *const *r_names,
                            const char *const *t_names, const float *weights, uint64_t num_triplet,
                            int normalization) {
    GV_TRY
    graph->graph.load_triplets(h_names, r_names, t_names, weights, num_triplet, normalization != 0);
    return 0;
    GV_CATCH(-1)
}

int gv_kgraph_save(gv_kgraph_t *graph, const char *file_name, int anonymous) {
    GV_TRY
    graph->graph.save(file_name, anonymous != 0);
    return 0;
    GV_CATCH(-1)
}

uint64_t gv_kgraph_num_vertex(const gv_kgraph_t *graph) {
    return graph->graph.num_vertex();
}

uint64_t gv_kgraph_num_edge(const gv_kgraph_t *graph) {
    return graph->graph.num_edge;
}

uint64_t gv_kgraph_num_relation(const gv_kgraph_t *graph) {
    return graph->graph.num_relation();
}

int gv_kgraph_normalization(const gv_kgraph_t *graph) {
    return graph->graph.normalization;
}

const char *gv_kgraph_id2entity(const gv_kgraph_t *graph, uint64_t id) {
    return id < graph->graph.id2entity.size() ? graph->graph.id2entity[id].c_str() : nullptr;
}

const char *gv_kgraph_id2relation(const gv_kgraph_t *graph, uint64_t id) {
    return id < graph->graph.id2relation.size() ? graph->graph.id2relation[id].c_str() : nullptr;
}

int64_t gv_kgraph_entity2id(const gv_kgraph_t *graph, const char *name) {
    auto found = graph->graph.entity2id.find(name);
    return found == graph->graph.entity2id.end() ? -1 : int64_t(found->second);
}

int64_t gv_kgraph_relation2id(const gv_kgraph_t *graph, const char *name) {
    auto found = graph->graph.relation2id.find(name);
    return found == graph->graph.relation2id.end() ? -1 : int64_t(found->second);
}

uint64_t gv_kgraph_flatten(gv_kgraph_t *graph, uint32_t *h, uint32_t *t, uint32_t *r, float *w,
                           uint64_t *flat_offsets, float *vertex_weights) {
    KnowledgeGraph &g = graph->graph;
    g.flatten();
    const size_t m = g.edge_h.size(), n = g.num_vertex();
    if (h)
        memcpy(h, g.edge_h.data(), m * sizeof(uint32_t));
    if (t)
        memcpy(t, g.edge_t.data(), m * sizeof(uint32_t));
    if (r)
        memcpy(r, g.edge_r.data(), m * sizeof(uint32_t));
    if (w)
        memcpy(w, g.edge_w.data(), m * sizeof(float));
    if (flat_offsets)
        memcpy(flat_offsets, g.offsets.data(), n * sizeof(uint64_t));
    if (vertex_weights)
        memcpy(vertex_weights, g.vertex_weights.data(), n * sizeof(float));
    return m;
}

int gv_kgraph_info(const gv_kgraph_t *graph, char *buffer, size_t capacity) {
    const std::string info = graph->graph.info();
    if (buffer && capacity) {
        strncpy(buffer, info.c_str(), capacity - 1);
        buffer[capacity - 1] = 0;
    }
    return int(info.size());
}

}  // extern "C"

gv::KnowledgeGraph &gv_kgraph_ref(gv_kgraph_t *graph) {
    return graph->graph;
}
