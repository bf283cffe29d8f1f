// Internal C++ declarations shared by the host runtime of libgv_b200 (gv_graph.cpp, gv_solver.cpp).
#pragma once

#include <sys/mman.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "gv_common.h"

namespace gv {

// AliasTable::build (reference include/base/alias_table.cuh:84-128)
template<class I>
void build_alias(const float *weights, size_t count, float *prob, I *alias);

// name -> id of a graph's vertices: an open-addressing table (linear probing, power-of-two capacity, load <= 1/2) whose
// 16-byte slots hold the first 11 bytes of the name inline, so that a lookup of the short names edge lists consist of
// is ONE cache line, and the slot can be prefetched from the hash alone -- the loader hashes a batch of lines ahead of
// resolving them.  The reference's std::unordered_map<std::string, Index> (instance/graph.cuh:124-153) costs 2.7 of
// the 4.5 s a Youtube-sized edge list (4.9e6 lines, 1.1e6 names) takes to load: a dependent chain of cache misses per
// lookup.  Ids are handed out in order of first appearance exactly as before; names live in `id2name`.
class NameTable {
public:
    static constexpr uint32_t kNone = 0xFFFFFFFFu;
    static uint64_t hash(const char *name, size_t length) {
        uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t(length) * 0xFF51AFD7ED558CCDull);
        while (length >= 8) {
            uint64_t word;
            memcpy(&word, name, 8);
            h = (h ^ word) * 0xC4CEB9FE1A85EC53ull;
            h ^= h >> 29;
            name += 8, length -= 8;
        }
        uint64_t tail = 0;
        memcpy(&tail, name, length);
        h = (h ^ tail) * 0xFF51AFD7ED558CCDull;
        return h ^ (h >> 32);
    }
    NameTable() {}
    NameTable(const NameTable &other) { *this = other; }
    NameTable &operator=(const NameTable &other) {
        if (this != &other) {
            slots.reset(other.slots.size());
            if (other.slots.size())
                memcpy(slots.data(), other.slots.data(), other.slots.size() * sizeof(Slot));
            count = other.count;
        }
        return *this;
    }
    void clear() {
        slots.reset(0);
        count = 0;
    }
    size_t size() const { return count; }
    void prefetch(uint64_t h) const {
        if (!slots.empty())
            __builtin_prefetch(&slots[h & (slots.size() - 1)]);
    }
    // id of `name`, or kNone
    uint32_t find(const char *name, size_t length, uint64_t h, const std::vector<std::string> &id2name) const {
        if (slots.empty())
            return kNone;
        const size_t mask = slots.size() - 1;
        for (size_t i = h & mask;; i = (i + 1) & mask) {
            const Slot &slot = slots[i];
            if (slot.id_plus_1 == 0)
                return kNone;
            if (matches(slot, name, length, id2name))
                return slot.id_plus_1 - 1;
        }
    }
    uint32_t find(const std::string &name, const std::vector<std::string> &id2name) const {
        return find(name.data(), name.size(), hash(name.data(), name.size()), id2name);
    }
    // id of `name`; a new name gets id = id2name.size() and is appended to id2name (created = true)
    uint32_t intern(const char *name, size_t length, uint64_t h, std::vector<std::string> &id2name, bool &created) {
        if ((count + 1) * 2 > slots.size())
            grow(id2name);
        const size_t mask = slots.size() - 1;
        for (size_t i = h & mask;; i = (i + 1) & mask) {
            Slot &slot = slots[i];
            if (slot.id_plus_1 == 0) {
                const uint32_t id = uint32_t(id2name.size());
                fill(slot, name, length, id);
                id2name.emplace_back(name, length);
                count++;
                created = true;
                return id;
            }
            if (matches(slot, name, length, id2name)) {
                created = false;
                return slot.id_plus_1 - 1;
            }
        }
    }

private:
    struct Slot {
        char key[11];       // the first min(length, 11) bytes
        uint8_t length;     // min(length, 255)
        uint32_t id_plus_1; // 0 = empty
    };
    static_assert(sizeof(Slot) == 16, "one slot = a quarter of a cache line");
    // zeroed slots on 2-MB-aligned memory with MADV_HUGEPAGE: the table of a million names is 64 MB, and with 4-KB
    // pages every probe (and every prefetch) is a TLB miss first
    struct SlotArray {
        Slot *base = nullptr;
        size_t length = 0;
        ~SlotArray() { free(base); }
        SlotArray() {}
        SlotArray(const SlotArray &) = delete;
        SlotArray &operator=(const SlotArray &) = delete;
        void reset(size_t n) {
            free(base);
            base = nullptr;
            length = 0;
            if (n == 0)
                return;
            const size_t huge = size_t(2) << 20, bytes = (n * sizeof(Slot) + huge - 1) / huge * huge;
            void *memory = nullptr;
            if (posix_memalign(&memory, n * sizeof(Slot) >= huge ? huge : 64, bytes) != 0)
                throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
            if (n * sizeof(Slot) >= huge)
                madvise(memory, bytes, MADV_HUGEPAGE);
#endif
            memset(memory, 0, n * sizeof(Slot));
            base = static_cast<Slot *>(memory);
            length = n;
        }
        void swap(SlotArray &other) {
            std::swap(base, other.base);
            std::swap(length, other.length);
        }
        bool empty() const { return length == 0; }
        size_t size() const { return length; }
        Slot *data() { return base; }
        const Slot *data() const { return base; }
        Slot &operator[](size_t i) { return base[i]; }
        const Slot &operator[](size_t i) const { return base[i]; }
    };
    SlotArray slots;
    size_t count = 0;

    static void fill(Slot &slot, const char *name, size_t length, uint32_t id) {
        memset(slot.key, 0, sizeof(slot.key));
        memcpy(slot.key, name, length < 11 ? length : 11);
        slot.length = uint8_t(length < 255 ? length : 255);
        slot.id_plus_1 = id + 1;
    }
    static bool matches(const Slot &slot, const char *name, size_t length, const std::vector<std::string> &id2name) {
        if (slot.length != uint8_t(length < 255 ? length : 255) || memcmp(slot.key, name, length < 11 ? length : 11) != 0)
            return false;
        if (length <= 11)
            return true;
        const std::string &full = id2name[slot.id_plus_1 - 1];
        return full.size() == length && memcmp(full.data(), name, length) == 0;
    }
    void grow(const std::vector<std::string> &id2name) {
        SlotArray old;
        old.swap(slots);
        slots.reset(old.empty() ? size_t(1) << 10 : old.size() * 2);
        const size_t mask = slots.size() - 1;
        for (size_t o = 0; o < old.size(); o++) {
            const Slot &slot = old[o];
            if (slot.id_plus_1) {
                const std::string &name = id2name[slot.id_plus_1 - 1];
                size_t i = hash(name.data(), name.size()) & mask;
                while (slots[i].id_plus_1)
                    i = (i + 1) & mask;
                slots[i] = slot;
            }
        }
    }
};

// graphvite::Graph<uint32> (reference include/instance/graph.cuh:62-277)
struct Graph {
    NameTable names;  // name -> id (the reference's name2id)
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
    uint32_t intern(const char *name, size_t length, uint64_t hash);
    void add_edge(const std::string &u_name, const std::string &v_name, float w);
    void add_edge_ids(uint32_t u, uint32_t v, float w);
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
    // every edge weighs the same: AliasTable::build over the edges then takes its trivial branch (alias = identity, one
    // probability for all, returned in `probability`) -- lets the solver build the table on the device
    bool uniform_edge_table(float &probability);
    int uniform_cache = -1;  // -1 unknown, 0 no, 1 yes (reset by flatten)
    float uniform_probability = 1;
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
