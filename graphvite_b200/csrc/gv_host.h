// Internal C++ declarations shared by the host runtime of libgv_b200 (gv_graph.cpp, gv_solver.cpp).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "gv_common.h"

namespace gv {

// AliasTable::build (reference include/base/alias_table.cuh:84-128)
template<class I>
void build_alias(const float *weights, size_t count, float *prob, I *alias);

// graphvite::Graph<uint32> (reference include/instance/graph.cuh:62-277)
struct Graph {
    std::unordered_map<std::string, uint32_t> name2id;
    std::vector<std::string> id2name;
    std::vector<float> vertex_weights;
    std::vector<uint32_t> degrees;  // out-degree = vertex_edges[v].size()
    uint64_t num_edge = 0;          // input lines
    bool as_undirected = true, normalization = false;

    // append-only log of directed edges in insertion order
    std::vector<uint32_t> log_u, log_v;
    std::vector<float> log_w;
    // flatten(): CSR in vertex order, insertion order inside a vertex
    bool flattened = false;
    std::vector<uint64_t> offsets;  // [num_vertex + 1]
    std::vector<uint32_t> edge_u, edge_v;
    std::vector<float> edge_w;

    uint32_t num_vertex() const { return uint32_t(id2name.size()); }
    void clear();
    uint32_t intern(const std::string &name);
    void add_edge(const std::string &u_name, const std::string &v_name, float w);
    void flatten();
    void normalize();
    void load_file(const char *file_name, bool undirected, bool normalized, const char *delimiters,
                   const char *comment);
    void load_corpus(const char *file_name, int window, int min_count, bool normalized, const char *delimiters,
                     const char *comment);  // WordGraph, instance/word_graph.cuh:75-166
    void load_edges(const char *const *u_names, const char *const *v_names, const float *weights, uint64_t count,
                    bool undirected, bool normalized);
    // binary edge arrays (no text, no name hashing): vertex i is named by its decimal id
    void load_id_edges(const uint32_t *u, const uint32_t *v, const float *weights, uint64_t count, bool undirected,
                       bool normalized);
    std::vector<int64_t> id_of_original;  // load_id_edges only: original id -> internal id (first-seen order), -1 = absent
    void save(const char *file_name, bool weighted, bool anonymous);
    bool has_dead_end() const;
    std::string info() const;
};

// graphvite::KnowledgeGraph<uint32> (reference include/instance/knowledge_graph.cuh:67-284)
struct KnowledgeGraph {
    std::unordered_map<std::string, uint32_t> entity2id, relation2id;
    std::vector<std::string> id2entity, id2relation;
    std::vector<float> vertex_weights;  // summed weight of the triplets an entity heads
    std::vector<uint32_t> degrees;      // number of triplets an entity heads
    uint64_t num_edge = 0;
    bool normalization = false;

    // append-only triplet log in insertion order
    std::vector<uint32_t> log_h, log_t, log_r;
    std::vector<float> log_w;
    // flatten(): CSR by head entity, insertion order inside an entity
    bool flattened = false;
    std::vector<uint64_t> offsets;  // [num_vertex + 1]
    std::vector<uint32_t> edge_h, edge_t, edge_r;
    std::vector<float> edge_w;

    uint32_t num_vertex() const { return uint32_t(id2entity.size()); }
    uint32_t num_relation() const { return uint32_t(id2relation.size()); }
    void clear();
    uint32_t intern_entity(const std::string &name);
    uint32_t intern_relation(const std::string &name);
    void add_edge(const std::string &h_name, const std::string &r_name, const std::string &t_name, float w);
    void flatten();
    void normalize();
    void load_file(const char *file_name, bool normalized, const char *delimiters, const char *comment);
    void load_triplets(const char *const *h_names, const char *const *r_names, const char *const *t_names,
                       const float *weights, uint64_t count, bool normalized);
    void save(const char *file_name, bool anonymous);
    std::string info() const;
};

}  // namespace gv

gv::Graph &gv_graph_ref(gv_graph_t *graph);
gv::KnowledgeGraph &gv_kgraph_ref(gv_kgraph_t *graph);
