// =============================================================================
// oracle/gv_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A sequential CPU restatement of the node-embedding hot path of
// DeepGraphLearning/graphvite v0.2.2 (the "reference", /root/reference).
// It exists only so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg can CHECK the CUDA product path (graphvite_b200/csrc).
// Nothing under graphvite_b200/ may include, link or call this file.
//
// Parity status: pinned against outputs of the reference itself -- golden
// vectors produced by oracle/ref_harness.cu (which drives the UNMODIFIED
// reference headers on a GPU box) and committed under tests/golden/.
// The reference ships no tests of its own.
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference/include).  Integer results (alias tables, partitions, pools,
// negative indices) are bit-exact restatements; floating-point kernels follow
// the reference's evaluation order (lane-strided partial dot products + the
// shfl_down tree) but use host libm, so they agree with the reference's device
// code to ~1e-6 relative, not bit-for-bit.
//
// Third-party arithmetic:
//   * cuRAND XORWOW (CURAND_RNG_PSEUDO_DEFAULT) uniform doubles -- the HOST
//     generator of the same libcurand reproduces the device generator's stream
//     (position n = subsequence n%4096, draw n/4096; verified in tests).
//   * libstdc++ std::mt19937 / uniform_int_distribution / uniform_real_distribution /
//     std::sort / std::pow -- we call the very same library functions.
// =============================================================================
#include "gv_oracle_common.h"

namespace oracle {

// -----------------------------------------------------------------------------
// R5: Graph, instance/graph.cuh:62-277 and core/graph.h:87-101
// -----------------------------------------------------------------------------
struct Graph {
    std::unordered_map<std::string, Index> name2id;
    std::vector<std::string> id2name;
    std::vector<std::vector<std::pair<Index, float>>> vertex_edges;
    std::vector<float> vertex_weights;
    Index num_vertex = 0;
    size_t num_edge = 0;
    bool as_undirected = true, normalization = false;
    // flatten()
    std::vector<Index> edge_u, edge_v;
    std::vector<float> edge_weights;
    std::vector<size_t> flat_offsets;

    Index vertex_id(const std::string &name) {
        auto it = name2id.find(name);
        if (it != name2id.end())
            return it->second;
        Index id = num_vertex++;
        name2id[name] = id;
        id2name.push_back(name);
        vertex_edges.emplace_back();
        vertex_weights.push_back(0);
        return id;
    }

    // instance/graph.cuh:124-153
    void add_edge(const std::string &u_name, const std::string &v_name, float w) {
        Index u = vertex_id(u_name);
        Index v = vertex_id(v_name);
        vertex_edges[u].push_back({v, w});
        vertex_weights[u] += w;
        if (as_undirected && u != v) {
            vertex_edges[v].push_back({u, w});
            vertex_weights[v] += w;
        }
        num_edge++;  // counts input lines, not directed edges (:152)
    }

    // instance/graph.cuh:103-121
    void normalize() {
        std::vector<float> context_weights(num_vertex);
        for (Index u = 0; u < num_vertex; u++)
            for (auto &e : vertex_edges[u])
                context_weights[e.first] += e.second;
        for (Index u = 0; u < num_vertex; u++) {
            float weight = 0;
            for (auto &e : vertex_edges[u]) {
                e.second /= std::sqrt(vertex_weights[u] * context_weights[e.first]);  // float sqrt
                weight += e.second;
            }
            vertex_weights[u] = weight;
        }
    }

    // instance/graph.cuh:163-201 (strstr comment cut, strtok tokenisation, atof)
    void load_file(const char *file_name, bool undirected, bool normalized, const char *delimiters,
                   const char *comment) {
        *this = Graph();
        as_undirected = undirected;
        normalization = normalized;
        FILE *fin = fopen(file_name, "r");
        if (!fin)
            fail(std::string("File `") + file_name + "` doesn't exist");
        std::vector<char> line(1 << 22);
        for (size_t line_no = 1; fgets(line.data(), line.size(), fin); line_no++) {
            char *cut = strstr(line.data(), comment);
            if (cut)
                *cut = 0;
            std::vector<std::string> tokens;
            char *p = line.data();
            while (*p) {
                while (*p && strchr(delimiters, *p))
                    p++;
                if (!*p)
                    break;
                char *q = p;
                while (*q && !strchr(delimiters, *q))
                    q++;
                tokens.emplace_back(p, q);
                p = q;
            }
            if (tokens.empty())
                continue;
            if (tokens.size() < 2 || tokens.size() > 3) {
                fclose(fin);
                fail("Invalid format at line " + std::to_string(line_no));
            }
            float w = tokens.size() == 3 ? float(atof(tokens[2].c_str())) : 1.0f;
            add_edge(tokens[0], tokens[1], w);
        }
        fclose(fin);
        if (normalization)
            normalize();
    }

    // core/graph.h:87-101
    void flatten() {
        if (!edge_u.empty())
            return;
        size_t offset = 0;
        flat_offsets.resize(num_vertex);
        for (Index u = 0; u < num_vertex; u++) {
            for (auto &e : vertex_edges[u]) {
                edge_u.push_back(u);
                edge_v.push_back(e.first);
                edge_weights.push_back(e.second);
            }
            flat_offsets[u] = offset;
            offset += vertex_edges[u].size();
        }
    }
};

// -----------------------------------------------------------------------------
// R7: get_schedule, core/solver.h:519-575 (non-tied branch; GraphSolver never ties)
// -----------------------------------------------------------------------------
static std::vector<std::vector<std::pair<int, int>>> get_schedule(int num_partition, int num_worker) {
    std::vector<std::vector<std::pair<int, int>>> schedule;
    std::vector<std::pair<int, int>> assignment(num_worker);
    if (num_partition == 1)
        return {{{0, 0}}};
    for (int x = 0; x < num_partition; x += num_worker)
        for (int y = 0; y < num_partition; y += num_worker)
            for (int offset = 0; offset < num_worker; offset++) {
                for (int i = 0; i < num_worker; i++)
                    assignment[i] = {x + (i + offset) % num_worker, y + i};
                schedule.push_back(assignment);
            }
    return schedule;
}

// -----------------------------------------------------------------------------
// R19: LINE::forward, instance/model/graph.h:40-45 + util/gpu.cuh:24-65.
// Lane l accumulates elements l, l+32, ... with fused multiply-add (nvcc -fmad=true),
// then the shfl_down tree (deltas 1,2,4,8,16); lane 0's value is broadcast.
// -----------------------------------------------------------------------------
static float warp_dot(const float *v, const float *c, int dim) {
    float lane[32];
    for (int l = 0; l < 32; l++) {
        float acc = 0;
        for (int i = l; i < dim; i += 32)
            acc = fmaf(v[i], c[i], acc);
        lane[l] = acc;
    }
    for (int delta = 1; delta < 32; delta *= 2)
        for (int l = 0; l + delta < 32; l++)  // lanes with l+delta>=32 never feed lane 0
            lane[l] = lane[l] + lane[l + delta];
    return lane[0];
}

// -----------------------------------------------------------------------------
// R17/R18: gpu::graph::train / train_1_moment / train_2_moment for ONE positive sample,
// instance/gpu/graph.cuh:54-94 (negatives first, positive last; vertex row staged in a
// buffer and written back at the end; loss normalised by 1 + k * negative_weight).
// `targets` holds the k negative ids followed by the positive tail id.
// -----------------------------------------------------------------------------
struct Matrices {
    int dim = 0;
    float *vertex = nullptr, *context = nullptr;
    float *vertex_m1 = nullptr, *context_m1 = nullptr;
    float *vertex_m2 = nullptr, *context_m2 = nullptr;
};

static float train_sample(const Matrices &m, const Optimizer &opt, Index head_id, const Index *targets,
                          int num_negative, float negative_weight) {
    const int dim = m.dim;
    std::vector<float> vertex_buffer(m.vertex + size_t(head_id) * dim, m.vertex + size_t(head_id + 1) * dim);
    float *vm1 = m.vertex_m1 ? m.vertex_m1 + size_t(head_id) * dim : nullptr;
    float *vm2 = m.vertex_m2 ? m.vertex_m2 + size_t(head_id) * dim : nullptr;
    float sample_loss = 0;
    for (int s = 0; s <= num_negative; s++) {
        Index tail_id = targets[s];
        int label = s < num_negative ? 0 : 1;
        float *context = m.context + size_t(tail_id) * dim;
        float *cm1 = m.context_m1 ? m.context_m1 + size_t(tail_id) * dim : nullptr;
        float *cm2 = m.context_m2 ? m.context_m2 + size_t(tail_id) * dim : nullptr;
        float logit = warp_dot(vertex_buffer.data(), context, dim);
        float prob = sigmoid(logit);
        float gradient, weight;
        if (label) {
            gradient = prob - 1;
            weight = 1;
            sample_loss += weight * -logf(prob + kEpsilon);
        } else {
            gradient = prob;
            weight = negative_weight;
            sample_loss += weight * -logf(1 - prob + kEpsilon);
        }
        // instance/model/graph.h:47-85: both updates read the pre-update v and c
        for (int i = 0; i < dim; i++) {
            float v = vertex_buffer[i];
            float c = context[i];
            vertex_buffer[i] -= opt.update(v, gradient * c, vm1 ? vm1 + i : nullptr, vm2 ? vm2 + i : nullptr, weight);
            context[i] -= opt.update(c, gradient * v, cm1 ? cm1 + i : nullptr, cm2 ? cm2 + i : nullptr, weight);
        }
    }
    memcpy(m.vertex + size_t(head_id) * dim, vertex_buffer.data(), dim * sizeof(float));
    return sample_loss / (1 + num_negative * negative_weight);
}

// -----------------------------------------------------------------------------
// Solver = SolverMixin + GraphSolver + samplers + workers, restated sequentially.
// -----------------------------------------------------------------------------
struct Solver {
    Graph *graph = nullptr;
    int dim = 128;
    int num_worker = 1, num_sampler = 1;
    int num_partition = 1, num_negative = 1, batch_size = 100000, episode_size = 0;
    Optimizer optimizer;
    // train() parameters
    std::string model;  // model of the PREVIOUS train call until Base::train assigns it (graph.cuh:785)
    int num_epoch = 0, augmentation_step = 0, random_walk_length = 40, random_walk_batch_size = 100;
    int shuffle_base = 0, positive_reuse = 1, log_frequency = 1000;
    float p = 1, q = 1, negative_sample_exponent = 0.75f, negative_weight = 5;
    bool resume = false;
    int batch_id = 0, num_batch = 0, pool_id = 0;

    std::vector<unsigned long long> sampler_seeds, worker_seeds;
    std::vector<std::unique_ptr<RandomStream>> sampler_streams, worker_streams;
    std::vector<std::vector<double>> sampler_random;  // host view of each sampler's buffer

    std::vector<std::vector<Index>> partitions;           // head_partitions == tail_partitions
    std::vector<std::pair<int, Index>> locations;         // head_locations == tail_locations
    Index partition_size = 0;

    AliasTable<size_t> edge_table;
    std::vector<AliasTable<Index>> vertex_edge_tables;
    std::vector<AliasTable<Index>> edge_edge_tables;
    // sample_pools[2][P][P], each episode_size*batch_size pairs stored {tail, head}
    std::vector<std::vector<std::vector<std::vector<Index>>>> sample_pools;

    std::vector<float> vertex_embeddings, context_embeddings;
    std::vector<float> vertex_m1, context_m1, vertex_m2, context_m2;

    std::vector<float> last_loss;               // per-sample loss of the last trained batch
    // each worker's loss buffer lives as long as the worker and is never cleared between blocks
    // (core/solver.h:1326,1541-1549): the loss logged at a block's first batch is the previous block's last
    std::vector<std::vector<float>> worker_loss;
    std::vector<Index> last_negative_batch;     // negatives of the last trained batch (local ids)
    std::vector<float> logged_loss;             // what the reference would LOG at each log point
    int sample_mode = 0;                        // 0 edge, 1 random walk, 2 biased random walk

    // core/solver.h:184-213: samplers are constructed first (each draws a cuRAND seed from the
    // global engine), then workers.
    Solver(int _dim, int _num_worker, int _num_sampler_per_worker) : dim(_dim), num_worker(_num_worker) {
        num_sampler = _num_sampler_per_worker * num_worker;
        std::uniform_int_distribution<unsigned long long> random_seed(0, ULLONG_MAX);
        for (int i = 0; i < num_sampler; i++)
            sampler_seeds.push_back(random_seed(global_engine()));
        for (int i = 0; i < num_worker; i++)
            worker_seeds.push_back(random_seed(global_engine()));
        for (auto s : sampler_seeds)
            sampler_streams.emplace_back(new RandomStream(s));
        for (auto s : worker_seeds)
            worker_streams.emplace_back(new RandomStream(s));
    }

    // core/solver.h:287-466 (GraphSolver: two in-place partitioned matrices, tail-partition
    // negative sampler; GPU memory budgeting is not restated -- num_partition=auto means
    // the minimum, num_worker, which is what fits on a 180 GB part).
    void build(Graph *_graph, const Optimizer &_optimizer, int _num_partition, int _num_negative, int _batch_size,
               int _episode_size) {
        graph = _graph;
        optimizer = _optimizer;
        num_partition = _num_partition;
        num_negative = _num_negative;
        batch_size = _batch_size;
        batch_id = 0;
        int min_partition = num_worker;
        if (num_partition == 0)
            num_partition = min_partition;
        if (num_partition < min_partition)
            fail("#partition should be no less than #worker");
        partitions = partition(graph->vertex_weights, num_partition);
        partition_size = 0;
        for (auto &part : partitions)
            partition_size = std::max<Index>(partition_size, part.size());
        locations.resize(graph->num_vertex);
        for (int i = 0; i < num_partition; i++)
            for (Index j = 0; j < partitions[i].size(); j++)
                locations[partitions[i][j]] = {i, j};
        // core/solver.h:426-436
        int expected_size = _episode_size;
        if (expected_size == 0) {
            expected_size = float(graph->num_vertex * kSamplePerVertex) / num_partition / batch_size;
            expected_size = std::max(expected_size, 1);
            if (num_partition == 1)
                expected_size = std::max(expected_size, kMinEpisodeSample / batch_size);
        }
        episode_size = expected_size;
        sample_pools.assign(2, {});
        for (auto &pool : sample_pools) {
            pool.resize(num_partition);
            for (auto &row : pool) {
                row.resize(num_partition);
                for (auto &block : row)
                    block.assign(size_t(episode_size) * batch_size * 2, 0);
            }
        }
        vertex_embeddings.assign(size_t(graph->num_vertex) * dim, 0);
        context_embeddings.assign(size_t(graph->num_vertex) * dim, 0);
        int nm = optimizer.num_moment();
        vertex_m1.assign(nm >= 1 ? vertex_embeddings.size() : 0, 0);
        context_m1.assign(nm >= 1 ? vertex_embeddings.size() : 0, 0);
        vertex_m2.assign(nm >= 2 ? vertex_embeddings.size() : 0, 0);
        context_m2.assign(nm >= 2 ? vertex_embeddings.size() : 0, 0);
        // core/solver.h:960-967: each sampler generates its first buffer in build()
        worker_loss.clear();
        sampler_random.resize(num_sampler);
        pool_id = 0;
    }

    // One refill of a sampler's host buffer.  The reference copies the buffer generated one
    // call earlier and immediately regenerates (core/solver.h:1015-1016); consumption is
    // therefore consecutive 5e6-blocks of the generator's stream.
    void refill(int sampler_id) {
        sampler_random[sampler_id].resize(kRandBatchSize);
        sampler_streams[sampler_id]->generate(sampler_random[sampler_id].data(), kRandBatchSize);
    }

    void store(std::vector<Index> &block, size_t offset, Index head_local, Index tail_local) {
        block[offset * 2] = tail_local;  // std::tuple<Index,Index> stores its members reversed
        block[offset * 2 + 1] = head_local;
    }

    // R8: SamplerMixin::sample, core/solver.h:1011-1055
    void sample_edges(int sampler_id, int start, int end) {
        refill(sampler_id);
        const std::vector<double> &random = sampler_random[sampler_id];
        auto &sample_pool = sample_pools[pool_id ^ 1];
        std::vector<std::vector<int>> offsets(num_partition, std::vector<int>(num_partition, start));
        int num_complete = 0, rand_id = 0;
        int sample_batch_size = random_walk_length * random_walk_batch_size;
        std::vector<std::pair<int, Index>> heads(sample_batch_size), tails(sample_batch_size);
        while (num_complete < num_partition * num_partition) {
            for (int i = 0; i < sample_batch_size; i++) {
                if (rand_id > kRandBatchSize - 2) {
                    refill(sampler_id);
                    rand_id = 0;
                }
                // gcc evaluates the two random[rand_id++] arguments right to left (appendix A.2)
                double rand2 = random[rand_id++];
                double rand1 = random[rand_id++];
                size_t edge_id = edge_table.sample(rand1, rand2);
                heads[i] = locations[graph->edge_u[edge_id]];
                tails[i] = locations[graph->edge_v[edge_id]];
            }
            for (int i = 0; i < sample_batch_size; i++) {
                int &offset = offsets[heads[i].first][tails[i].first];
                if (offset < end) {
                    store(sample_pool[heads[i].first][tails[i].first], offset, heads[i].second, tails[i].second);
                    if (++offset == end)
                        num_complete++;
                }
            }
        }
    }

    // R9 / R10: GraphSampler::sample_random_walk (instance/graph.cuh:376-450) and
    // sample_biased_random_walk (:298-373); they differ only in the step table.
    void sample_walks(int sampler_id, int start, int end, bool biased) {
        int pool_size = episode_size * batch_size;
        if (pool_size % shuffle_base != 0)
            fail("Can't perform pseudo shuffle: episode size must be a multiple of the shuffle base");
        refill(sampler_id);
        const std::vector<double> &random = sampler_random[sampler_id];
        auto &sample_pool = sample_pools[pool_id ^ 1];
        std::vector<std::vector<int>> offsets(num_partition, std::vector<int>(num_partition, start));
        const int L = random_walk_length;
        std::vector<std::vector<Index>> chains(random_walk_batch_size, std::vector<Index>(L + 1));
        std::vector<int> sample_lengths(random_walk_batch_size);
        int num_complete = 0, rand_id = 0;
        while (num_complete < num_partition * num_partition) {
            for (int i = 0; i < random_walk_batch_size; i++) {
                if (rand_id > kRandBatchSize - L * 2) {
                    refill(sampler_id);
                    rand_id = 0;
                }
                double rand2 = random[rand_id++];
                double rand1 = random[rand_id++];
                size_t edge_id = edge_table.sample(rand1, rand2);
                Index current = graph->edge_u[edge_id];
                chains[i][0] = current;
                current = graph->edge_v[edge_id];
                chains[i][1] = current;
                sample_lengths[i] = L;
                for (int j = 2; j <= L; j++) {
                    if (!graph->vertex_edges[current].empty()) {
                        rand2 = random[rand_id++];
                        rand1 = random[rand_id++];
                        Index neighbor_id;
                        if (biased) {
                            neighbor_id = edge_edge_tables[edge_id].sample(rand1, rand2);
                            edge_id = graph->flat_offsets[current] + neighbor_id;
                        } else
                            neighbor_id = vertex_edge_tables[current].sample(rand1, rand2);
                        current = graph->vertex_edges[current][neighbor_id].first;
                        chains[i][j] = current;
                    } else {
                        sample_lengths[i] = j - 1;
                        break;
                    }
                }
            }
            for (int i = 0; i < random_walk_batch_size; i++)
                for (int j = 0; j < sample_lengths[i]; j++)
                    for (int k = 1; k <= augmentation_step; k++) {
                        if (j + k > sample_lengths[i])
                            break;
                        auto head = locations[chains[i][j]];
                        auto tail = locations[chains[i][j + k]];
                        int &offset = offsets[head.first][tail.first];
                        if (offset < end) {
                            // pseudo shuffle, instance/graph.cuh:440-441
                            int shuffled = offset % shuffle_base * (pool_size / shuffle_base) + offset / shuffle_base;
                            store(sample_pool[head.first][tail.first], shuffled, head.second, tail.second);
                            if (++offset == end)
                                num_complete++;
                        }
                    }
        }
    }

    // R11: GraphSolver::get_sample_function, instance/graph.cuh:680-721 (+ :645-677)
    void prepare_sampling() {
        graph->flatten();
        edge_table.build(graph->edge_weights);
        if (augmentation_step == 1) {
            sample_mode = 0;
            return;
        }
        if (model == "DeepWalk" || model == "LINE") {
            vertex_edge_tables.assign(graph->num_vertex, AliasTable<Index>());
            for (Index i = 0; i < graph->num_vertex; i++) {
                std::vector<float> w;
                for (auto &e : graph->vertex_edges[i])
                    w.push_back(e.second);
                if (!w.empty())
                    vertex_edge_tables[i].build(w);
            }
            sample_mode = 1;
            return;
        }
        if (model == "node2vec") {
            std::vector<std::unordered_set<Index>> neighbors(graph->num_vertex);
            for (Index u = 0; u < graph->num_vertex; u++)
                for (auto &e : graph->vertex_edges[u])
                    neighbors[u].insert(e.first);
            size_t num_directed_edge = graph->edge_u.size();
            edge_edge_tables.assign(num_directed_edge, AliasTable<Index>());
            for (size_t i = 0; i < num_directed_edge; i++) {
                Index u = graph->edge_u[i], v = graph->edge_v[i];
                std::vector<float> w;
                for (auto &e : graph->vertex_edges[v]) {
                    Index x = e.first;
                    if (x == u)
                        w.push_back(e.second / p);
                    else if (neighbors[x].find(u) == neighbors[x].end())
                        w.push_back(e.second / q);
                    else
                        w.push_back(e.second);
                }
                if (!w.empty())
                    edge_edge_tables[i].build(w);
            }
            sample_mode = 2;
            return;
        }
        fail("Invalid model `" + model + "`");
    }

    // fill sample_pools[pool_id ^ 1] with all samplers (core/solver.h:614-628)
    void fill_pool() {
        int num_sample = episode_size * batch_size;
        int work_load = (num_sample + num_sampler - 1) / num_sampler;
        for (int i = 0; i < num_sampler; i++) {
            int start = work_load * i, end = std::min(work_load * (i + 1), num_sample);
            if (sample_mode == 0)
                sample_edges(i, start, end);
            else
                sample_walks(i, start, end, sample_mode == 2);
        }
    }

    // R12: init_embeddings, instance/graph.cuh:724-731
    void init_embeddings() {
        std::uniform_real_distribution<float> init(-0.5 / dim, 0.5 / dim);
        for (auto &x : vertex_embeddings)
            x = init(global_engine());
        std::fill(context_embeddings.begin(), context_embeddings.end(), 0.0f);
    }

    // GraphSolver::train prologue, instance/graph.cuh:770-793, then SolverMixin::train set-up,
    // core/solver.h:588-628 (everything up to and including the first pool fill).
    void train_begin(const std::string &_model, int _num_epoch, bool _resume, int _augmentation_step,
                     int _random_walk_length, int _random_walk_batch_size, int _shuffle_base, float _p, float _q,
                     int _positive_reuse, float _negative_sample_exponent, float _negative_weight,
                     int _log_frequency) {
        augmentation_step = _augmentation_step;
        random_walk_length = _random_walk_length;
        random_walk_batch_size = _random_walk_batch_size;
        shuffle_base = _shuffle_base;
        p = _p;
        q = _q;
        if (augmentation_step == 0)
            augmentation_step = std::log(double(kExpectedDegree)) / std::log(float(graph->num_edge) / graph->num_vertex);
        if (shuffle_base == 0)
            shuffle_base = augmentation_step;
        if (model == "DeepWalk" || model == "node2vec")  // tests the PREVIOUS call's model (appendix A.6)
            shuffle_base = 1;
        if (augmentation_step < 1)
            fail("`augmentation_step` should be a positive integer");
        if (augmentation_step > random_walk_length)
            fail("`random_walk_length` should be no less than `augmentation_step`");
        model = _model;
        if (model != "DeepWalk" && model != "LINE" && model != "node2vec")
            fail("Invalid model `" + model + "`");
        num_epoch = _num_epoch;
        resume = _resume;
        positive_reuse = _positive_reuse;
        negative_sample_exponent = _negative_sample_exponent;
        negative_weight = _negative_weight;
        log_frequency = _log_frequency;
        if (!resume) {
            init_embeddings();
            std::fill(vertex_m1.begin(), vertex_m1.end(), 0.0f);
            std::fill(context_m1.begin(), context_m1.end(), 0.0f);
            std::fill(vertex_m2.begin(), vertex_m2.end(), 0.0f);
            std::fill(context_m2.begin(), context_m2.end(), 0.0f);
            batch_id = 0;
        }
        num_batch = batch_id + size_t(num_epoch) * graph->num_edge / batch_size;
        prepare_sampling();
        fill_pool();
    }

    // R14 + R4: negative table of one tail partition (core/solver.h:1264-1278) and the
    // gpu::Sample kernel with its double->float narrowing (base/alias_table.cuh:175-183).
    AliasTable<Index> build_negative_sampler(int tail_partition) const {
        std::vector<float> weights;
        for (auto g : partitions[tail_partition])
            weights.push_back(std::pow(graph->vertex_weights[g], negative_sample_exponent));
        AliasTable<Index> table;
        table.build(weights);
        return table;
    }

    static Index device_sample(const AliasTable<Index> &table, double random1, double random2) {
        float rand1 = float(random1), rand2 = float(random2);
        return table.sample(double(rand1), double(rand2));
    }

    // R15 + R17/18: WorkerMixin::train for one (head, tail) block, core/solver.h:1511-1557,
    // samples processed sequentially (the reference races them Hogwild-style).
    // batch ids: this_batch = first_batch_id + j * batch_stride (deterministic stand-in for the
    // shared atomic counter; section 8e of SURVEY.md).
    void train_block(int worker_id, int head_partition, int tail_partition, int first_batch_id, int batch_stride,
                     const AliasTable<Index> &negative_sampler) {
        // gather the block's rows (load_partition / load_embedding, core/solver.h:1349-1386)
        const std::vector<Index> &head_ids = partitions[head_partition], &tail_ids = partitions[tail_partition];
        int nm = optimizer.num_moment();
        auto gather = [&](const std::vector<float> &global, const std::vector<Index> &ids) {
            std::vector<float> local(ids.size() * size_t(dim));
            for (size_t i = 0; i < ids.size(); i++)
                memcpy(&local[i * dim], &global[size_t(ids[i]) * dim], dim * sizeof(float));
            return local;
        };
        auto scatter = [&](const std::vector<float> &local, std::vector<float> &global, const std::vector<Index> &ids) {
            for (size_t i = 0; i < ids.size(); i++)
                memcpy(&global[size_t(ids[i]) * dim], &local[i * dim], dim * sizeof(float));
        };
        std::vector<float> v = gather(vertex_embeddings, head_ids), c = gather(context_embeddings, tail_ids);
        std::vector<float> v1, c1, v2, c2;
        if (nm >= 1) {
            v1 = gather(vertex_m1, head_ids);
            c1 = gather(context_m1, tail_ids);
        }
        if (nm >= 2) {
            v2 = gather(vertex_m2, head_ids);
            c2 = gather(context_m2, tail_ids);
        }
        Matrices m;
        m.dim = dim;
        m.vertex = v.data();
        m.context = c.data();
        m.vertex_m1 = nm >= 1 ? v1.data() : nullptr;
        m.context_m1 = nm >= 1 ? c1.data() : nullptr;
        m.vertex_m2 = nm >= 2 ? v2.data() : nullptr;
        m.context_m2 = nm >= 2 ? c2.data() : nullptr;

        const std::vector<Index> &samples = sample_pools[pool_id][head_partition][tail_partition];
        std::vector<double> random(size_t(batch_size) * num_negative * 2);
        std::vector<Index> targets(num_negative + 1);
        Optimizer opt = optimizer;
        if ((int)worker_loss.size() != num_worker)
            worker_loss.assign(num_worker, std::vector<float>());
        if ((int)worker_loss[worker_id].size() != batch_size)
            worker_loss[worker_id].assign(batch_size, 0);
        last_loss = worker_loss[worker_id];
        last_negative_batch.assign(size_t(batch_size) * num_negative, 0);
        for (int reuse = 0; reuse < positive_reuse; reuse++)
            for (int j = 0; j < episode_size; j++) {
                int this_batch = first_batch_id + (reuse * episode_size + j) * batch_stride;
                const Index *batch = &samples[size_t(j) * batch_size * 2];
                worker_streams[worker_id]->generate(random.data(), random.size());
                for (size_t t = 0; t < last_negative_batch.size(); t++)
                    last_negative_batch[t] = device_sample(negative_sampler, random[t * 2], random[t * 2 + 1]);
                // the loss logged at batch b is the buffer left by batch b-1 (appendix A.8)
                if (this_batch % log_frequency == 0) {
                    float batch_loss = 0;
                    for (int i = 0; i < batch_size; i++)
                        batch_loss += last_loss[i];
                    logged_loss.push_back(batch_loss / batch_size);
                }
                opt.apply_schedule(this_batch, num_batch);
                for (int i = 0; i < batch_size; i++) {
                    for (int s = 0; s < num_negative; s++)
                        targets[s] = last_negative_batch[size_t(i) * num_negative + s];
                    targets[num_negative] = batch[i * 2];
                    last_loss[i] = train_sample(m, opt, batch[i * 2 + 1], targets.data(), num_negative,
                                                negative_weight);
                }
            }
        worker_loss[worker_id] = last_loss;
        scatter(v, vertex_embeddings, head_ids);
        scatter(c, context_embeddings, tail_ids);
        if (nm >= 1) {
            scatter(v1, vertex_m1, head_ids);
            scatter(c1, context_m1, tail_ids);
        }
        if (nm >= 2) {
            scatter(v2, vertex_m2, head_ids);
            scatter(c2, context_m2, tail_ids);
        }
    }

    // one pass of the episode loop, core/solver.h:629-649; returns false when training is over
    bool train_episode() {
        if (batch_id >= num_batch)
            return false;
        pool_id ^= 1;
        auto schedule = get_schedule(num_partition, num_worker);
        // workers train on sample_pools[pool_id] while samplers fill the other pool; the two
        // touch disjoint state, so run the workers first, then the samplers.
        int per_block = positive_reuse * episode_size;
        for (auto &assignment : schedule) {
            for (int i = 0; i < (int)assignment.size(); i++) {
                AliasTable<Index> negative_sampler = build_negative_sampler(assignment[i].second);
                train_block(i, assignment[i].first, assignment[i].second, batch_id + i, (int)assignment.size(),
                            negative_sampler);
            }
            batch_id += per_block * (int)assignment.size();
        }
        fill_pool();
        return true;
    }

    // R22: predict, core/solver.h:729-802 + gpu/graph.cuh:250-279 (input rows are (v, c))
    void predict(const Index *pairs, size_t n, float *out) const {
        for (size_t i = 0; i < n; i++)
            out[i] = warp_dot(&vertex_embeddings[size_t(pairs[i * 2]) * dim],
                              &context_embeddings[size_t(pairs[i * 2 + 1]) * dim], dim);
    }
};

}  // namespace oracle

// =============================================================================
// C interface for ctypes (tests/, bench.py cpu_baseline, __graft_entry__.smoke)
// =============================================================================
using namespace oracle;

static thread_local std::string g_error;
void og_set_error(const std::string &message) { g_error = message; }  // used by gv_oracle_kg.cpp
#define ORACLE_TRY try {
#define ORACLE_CATCH(ret)            \
    }                                \
    catch (const std::exception &e) { \
        g_error = e.what();          \
        return ret;                  \
    }

extern "C" {

const char *og_last_error() { return g_error.c_str(); }

void og_reset_global_engine() { global_engine() = std::mt19937(); }

// ---- alias table -------------------------------------------------------------
int og_alias_build(const float *weights, uint64_t n, float *prob, uint64_t *alias) {
    ORACLE_TRY
    AliasTable<size_t> t;
    t.build(std::vector<float>(weights, weights + n));
    memcpy(prob, t.prob.data(), n * sizeof(float));
    for (uint64_t i = 0; i < n; i++)
        alias[i] = t.alias[i];
    return 0;
    ORACLE_CATCH(-1)
}

// cpu path: rand1 = random[2t+1], rand2 = random[2t]; gpu path: narrowed, rand1 = random[2t]
int og_alias_sample(const float *prob, const uint64_t *alias, uint64_t n, const double *random, uint64_t num_sample,
                    int gpu_path, uint64_t *out) {
    ORACLE_TRY
    AliasTable<size_t> t;
    t.count = n;
    t.prob.assign(prob, prob + n);
    t.alias.assign(alias, alias + n);
    for (uint64_t i = 0; i < num_sample; i++) {
        if (gpu_path)
            out[i] = t.sample(double(float(random[2 * i])), double(float(random[2 * i + 1])));
        else
            out[i] = t.sample(random[2 * i + 1], random[2 * i]);
    }
    return 0;
    ORACLE_CATCH(-1)
}

// ---- random stream -----------------------------------------------------------
int og_curand_uniform_double(uint64_t seed, const uint64_t *chunks, int num_chunk, double *out) {
    ORACLE_TRY
    RandomStream s(seed);
    for (int i = 0; i < num_chunk; i++) {
        s.generate(out, chunks[i]);
        out += chunks[i];
    }
    return 0;
    ORACLE_CATCH(-1)
}

// ---- graph -------------------------------------------------------------------
void *og_graph_load(const char *file_name, int as_undirected, int normalization, const char *delimiters,
                    const char *comment) {
    ORACLE_TRY
    Graph *g = new Graph();
    try {
        g->load_file(file_name, as_undirected, normalization, delimiters, comment);
    } catch (...) {
        delete g;
        throw;
    }
    return g;
    ORACLE_CATCH(nullptr)
}
void og_graph_free(void *g) { delete (Graph *)g; }
uint64_t og_graph_num_vertex(void *g) { return ((Graph *)g)->num_vertex; }
uint64_t og_graph_num_edge(void *g) { return ((Graph *)g)->num_edge; }
uint64_t og_graph_num_directed_edge(void *g) {
    ((Graph *)g)->flatten();
    return ((Graph *)g)->edge_u.size();
}
const char *og_graph_id2name(void *g, uint64_t id) { return ((Graph *)g)->id2name[id].c_str(); }
void og_graph_vertex_weights(void *g, float *out) {
    Graph *G = (Graph *)g;
    memcpy(out, G->vertex_weights.data(), G->num_vertex * sizeof(float));
}
void og_graph_flat(void *g, uint32_t *u, uint32_t *v, float *w, uint64_t *offsets) {
    Graph *G = (Graph *)g;
    G->flatten();
    size_t n = G->edge_u.size();
    memcpy(u, G->edge_u.data(), n * 4);
    memcpy(v, G->edge_v.data(), n * 4);
    memcpy(w, G->edge_weights.data(), n * 4);
    for (Index i = 0; i < G->num_vertex; i++)
        offsets[i] = G->flat_offsets[i];
}

// ---- partition / schedule ------------------------------------------------------
// part_of[v], local_of[v]
int og_partition(const float *weights, uint64_t n, int num_partition, int *part_of, uint32_t *local_of) {
    ORACLE_TRY
    auto parts = partition(std::vector<float>(weights, weights + n), num_partition);
    for (int i = 0; i < num_partition; i++)
        for (Index j = 0; j < parts[i].size(); j++) {
            part_of[parts[i][j]] = i;
            local_of[parts[i][j]] = j;
        }
    return 0;
    ORACLE_CATCH(-1)
}
// out: [num_step][num_worker][2]; returns num_step
int og_schedule(int num_partition, int num_worker, int *out, int capacity) {
    auto schedule = get_schedule(num_partition, num_worker);
    int n = 0;
    for (auto &a : schedule)
        for (auto &pr : a) {
            if (n + 2 > capacity)
                return -1;
            out[n++] = pr.first;
            out[n++] = pr.second;
        }
    return (int)schedule.size();
}

float og_lr(int schedule, float init_lr, int batch_id, int num_batch) {
    Optimizer o;
    o.schedule = schedule;
    o.init_lr = init_lr;
    o.apply_schedule(batch_id, num_batch);
    return o.lr;
}

// ---- kernel restatement on caller-provided matrices ----------------------------
// opt: {type, schedule, lr, weight_decay, a, b, epsilon}; moments may be null.
// batch = pairs {tail, head}; negatives = [n][k]; loss = [n] or null. Sequential.
int og_train_batch(int dim, float *vertex, float *context, float *vm1, float *cm1, float *vm2, float *cm2,
                   const uint32_t *batch, const uint32_t *negatives, uint64_t n, int num_negative, int opt_type,
                   float lr, float weight_decay, float a, float b, float epsilon, float negative_weight,
                   float *loss) {
    ORACLE_TRY
    Matrices m;
    m.dim = dim;
    m.vertex = vertex;
    m.context = context;
    m.vertex_m1 = vm1;
    m.context_m1 = cm1;
    m.vertex_m2 = vm2;
    m.context_m2 = cm2;
    Optimizer o;
    o.type = opt_type;
    o.lr = lr;
    o.weight_decay = weight_decay;
    o.a = a;
    o.b = b;
    o.epsilon = epsilon;
    std::vector<Index> targets(num_negative + 1);
    for (uint64_t i = 0; i < n; i++) {
        for (int s = 0; s < num_negative; s++)
            targets[s] = negatives[i * num_negative + s];
        targets[num_negative] = batch[i * 2];
        float l = train_sample(m, o, batch[i * 2 + 1], targets.data(), num_negative, negative_weight);
        if (loss)
            loss[i] = l;
    }
    return 0;
    ORACLE_CATCH(-1)
}

// pairs = {tail, head} like the device batch
void og_predict_batch(int dim, const float *vertex, const float *context, const uint32_t *batch, uint64_t n,
                      float *logits) {
    for (uint64_t i = 0; i < n; i++)
        logits[i] = warp_dot(vertex + size_t(batch[i * 2 + 1]) * dim, context + size_t(batch[i * 2]) * dim, dim);
}

// ---- solver ------------------------------------------------------------------
void *og_solver_create(int dim, int num_worker, int num_sampler_per_worker) {
    ORACLE_TRY
    return new Solver(dim, num_worker, num_sampler_per_worker);
    ORACLE_CATCH(nullptr)
}
void og_solver_free(void *s) { delete (Solver *)s; }
void og_solver_seeds(void *s, uint64_t *sampler_seeds, uint64_t *worker_seeds) {
    Solver *S = (Solver *)s;
    for (size_t i = 0; i < S->sampler_seeds.size(); i++)
        sampler_seeds[i] = S->sampler_seeds[i];
    for (size_t i = 0; i < S->worker_seeds.size(); i++)
        worker_seeds[i] = S->worker_seeds[i];
}
int og_solver_build(void *s, void *graph, int opt_type, int schedule, float lr, float weight_decay, float a, float b,
                    float epsilon, int num_partition, int num_negative, int batch_size, int episode_size) {
    ORACLE_TRY
    Optimizer o;
    o.type = opt_type;
    o.schedule = schedule;
    o.init_lr = o.lr = lr;
    o.weight_decay = weight_decay;
    o.a = a;
    o.b = b;
    o.epsilon = epsilon;
    ((Solver *)s)->build((Graph *)graph, o, num_partition, num_negative, batch_size, episode_size);
    return 0;
    ORACLE_CATCH(-1)
}
int og_solver_train_begin(void *s, const char *model, int num_epoch, int resume, int augmentation_step,
                          int random_walk_length, int random_walk_batch_size, int shuffle_base, float p, float q,
                          int positive_reuse, float negative_sample_exponent, float negative_weight,
                          int log_frequency) {
    ORACLE_TRY
    ((Solver *)s)->train_begin(model, num_epoch, resume, augmentation_step, random_walk_length,
                               random_walk_batch_size, shuffle_base, p, q, positive_reuse, negative_sample_exponent,
                               negative_weight, log_frequency);
    return 0;
    ORACLE_CATCH(-1)
}
// returns 1 if an episode was trained, 0 if training is complete, -1 on error
int og_solver_train_episode(void *s) {
    ORACLE_TRY
    return ((Solver *)s)->train_episode() ? 1 : 0;
    ORACLE_CATCH(-1)
}
int og_solver_fill_pool(void *s) {
    ORACLE_TRY
    ((Solver *)s)->fill_pool();
    return 0;
    ORACLE_CATCH(-1)
}
int og_solver_info(void *s, int *out) {
    Solver *S = (Solver *)s;
    out[0] = S->num_partition;
    out[1] = S->episode_size;
    out[2] = S->batch_size;
    out[3] = S->augmentation_step;
    out[4] = S->shuffle_base;
    out[5] = S->num_batch;
    out[6] = S->batch_id;
    out[7] = S->pool_id;
    out[8] = S->num_sampler;
    out[9] = (int)S->partition_size;
    return 0;
}
// which: 0 = pool being trained next (pool_id ^ 1 after a fill), given explicitly by caller
const uint32_t *og_solver_pool(void *s, int pool, int head_partition, int tail_partition) {
    return ((Solver *)s)->sample_pools[pool][head_partition][tail_partition].data();
}
void og_solver_locations(void *s, int *part_of, uint32_t *local_of) {
    Solver *S = (Solver *)s;
    for (size_t v = 0; v < S->locations.size(); v++) {
        part_of[v] = S->locations[v].first;
        local_of[v] = S->locations[v].second;
    }
}
float *og_solver_embeddings(void *s, int which) {
    Solver *S = (Solver *)s;
    return which == 0 ? S->vertex_embeddings.data() : S->context_embeddings.data();
}
float *og_solver_moments(void *s, int which, int order) {
    Solver *S = (Solver *)s;
    if (order == 1)
        return which == 0 ? S->vertex_m1.data() : S->context_m1.data();
    return which == 0 ? S->vertex_m2.data() : S->context_m2.data();
}
// negative alias table of one tail partition (prob, alias as uint32), returns its size
int64_t og_solver_negative_table(void *s, int tail_partition, float *prob, uint32_t *alias) {
    ORACLE_TRY
    auto t = ((Solver *)s)->build_negative_sampler(tail_partition);
    if (prob)
        memcpy(prob, t.prob.data(), t.prob.size() * 4);
    if (alias)
        memcpy(alias, t.alias.data(), t.alias.size() * 4);
    return (int64_t)t.prob.size();
    ORACLE_CATCH(-1)
}
const uint32_t *og_solver_last_negatives(void *s) { return ((Solver *)s)->last_negative_batch.data(); }
const float *og_solver_last_loss(void *s) { return ((Solver *)s)->last_loss.data(); }
int og_solver_logged_loss(void *s, float *out, int capacity) {
    Solver *S = (Solver *)s;
    int n = std::min<int>(capacity, S->logged_loss.size());
    for (int i = 0; i < n; i++)
        out[i] = S->logged_loss[i];
    return (int)S->logged_loss.size();
}
int og_solver_predict(void *s, const uint32_t *pairs, uint64_t n, float *out) {
    ORACLE_TRY
    ((Solver *)s)->predict(pairs, n, out);
    return 0;
    ORACLE_CATCH(-1)
}
// edge / vertex-edge alias tables for checking the product's host-side builders
int og_solver_edge_table(void *s, float *prob, uint64_t *alias) {
    Solver *S = (Solver *)s;
    memcpy(prob, S->edge_table.prob.data(), S->edge_table.prob.size() * 4);
    for (size_t i = 0; i < S->edge_table.alias.size(); i++)
        alias[i] = S->edge_table.alias[i];
    return 0;
}
// flattened per-vertex tables laid out at flat_offsets (CSR order)
int og_solver_vertex_edge_tables(void *s, float *prob, uint32_t *alias) {
    Solver *S = (Solver *)s;
    for (Index v = 0; v < S->graph->num_vertex; v++) {
        if (v >= S->vertex_edge_tables.size())
            break;
        auto &t = S->vertex_edge_tables[v];
        size_t off = S->graph->flat_offsets[v];
        for (size_t i = 0; i < t.prob.size(); i++) {
            prob[off + i] = t.prob[i];
            alias[off + i] = t.alias[i];
        }
    }
    return 0;
}

}  // extern "C"
