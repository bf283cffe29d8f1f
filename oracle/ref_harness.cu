// =============================================================================
// oracle/ref_harness.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin C driver around the UNMODIFIED reference headers (included from where
// they lie, /root/reference/include, via oracle/Makefile).  It lets
// oracle/make_golden.py run the reference's own classes on a GPU box and record
// golden vectors (tests/golden/), which pin oracle/gv_oracle.cpp and, through
// it, the CUDA product path.  All reference members used here are public.
// Built into oracle/_ref/libref_harness.so (git-ignored), loaded with ctypes
// from a Python process so that the pybind11 symbols the reference headers pull
// in resolve against the running interpreter.
// =============================================================================
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "instance/graph.cuh"

using graphvite::AliasTable;
using graphvite::CudaCheck;
using graphvite::CurandCheck;
using graphvite::Memory;

typedef unsigned int Index;
typedef graphvite::Graph<Index> RefGraph;

namespace {

graphvite::Optimizer make_optimizer(int type, int schedule, float lr, float wd, float a, float b, float eps) {
    graphvite::LRSchedule sched(schedule == 1 ? "linear" : "constant");
    switch (type) {
        case 0: return graphvite::SGD(lr, wd, sched);
        case 1: return graphvite::Momentum(lr, wd, a, sched);
        case 2: return graphvite::AdaGrad(lr, wd, eps, sched);
        case 3: return graphvite::RMSprop(lr, wd, a, eps, sched);
        default: return graphvite::Adam(lr, wd, a, b, eps, sched);
    }
}

struct SolverBase {
    virtual ~SolverBase() {}
    virtual void build(RefGraph *graph, const graphvite::Optimizer &opt, int P, int k, int B, int E) = 0;
    virtual void train(const char *model, int epochs, int resume, int aug, int L, int wb, int shuffle_base, float p,
                       float q, int reuse, float exponent, float neg_weight, int log_frequency) = 0;
    virtual void info(int *out) = 0;
    virtual void locations(int *part_of, Index *local_of) = 0;
    virtual void pool(int pool, int hp, int tp, Index *out) = 0;
    virtual void embeddings(int which, float *out) = 0;
    virtual void last_negatives(Index *out) = 0;
    virtual void last_loss(float *out) = 0;
    virtual void edge_table(float *prob, uint64_t *alias) = 0;
    virtual void negative_table(float *prob, Index *alias) = 0;
    virtual void predict(const Index *pairs, size_t n, float *out) = 0;
};

template<size_t dim>
struct SolverImpl : SolverBase {
    typedef graphvite::GraphSolver<dim, float, Index> Solver;
    Solver solver;
    SolverImpl(std::vector<int> devices, int spw, size_t limit) : solver(devices, spw, limit) {}

    void build(RefGraph *graph, const graphvite::Optimizer &opt, int P, int k, int B, int E) override {
        solver.build(*graph, opt, P, k, B, E);
    }
    void train(const char *model, int epochs, int resume, int aug, int L, int wb, int shuffle_base, float p, float q,
               int reuse, float exponent, float neg_weight, int log_frequency) override {
        solver.train(model, epochs, resume, aug, L, wb, shuffle_base, p, q, reuse, exponent, neg_weight,
                     log_frequency);
    }
    void info(int *out) override {
        out[0] = solver.num_partition;
        out[1] = solver.episode_size;
        out[2] = solver.batch_size;
        out[3] = solver.augmentation_step;
        out[4] = solver.shuffle_base;
        out[5] = solver.num_batch;
        out[6] = solver.batch_id;
        out[7] = solver.pool_id;
        out[8] = solver.num_sampler;
        out[9] = solver.head_partition_size;
    }
    void locations(int *part_of, Index *local_of) override {
        for (size_t v = 0; v < solver.head_locations.size(); v++) {
            part_of[v] = solver.head_locations[v].first;
            local_of[v] = solver.head_locations[v].second;
        }
    }
    void pool(int pool, int hp, int tp, Index *out) override {
        auto &block = solver.sample_pools[pool][hp][tp];
        memcpy(out, block.data(), block.size() * sizeof(block[0]));
    }
    void embeddings(int which, float *out) override {
        auto &e = which == 0 ? *solver.vertex_embeddings : *solver.context_embeddings;
        memcpy(out, e.data(), e.size() * sizeof(e[0]));
    }
    void last_negatives(Index *out) override {
        auto &m = solver.workers[0]->negative_batch;
        m.to_host();
        memcpy(out, m.host_ptr, m.count * sizeof(Index));
    }
    void last_loss(float *out) override {
        auto &m = solver.workers[0]->loss;
        m.to_host();
        memcpy(out, m.host_ptr, m.count * sizeof(float));
    }
    void edge_table(float *prob, uint64_t *alias) override {
        auto &t = solver.edge_table;
        memcpy(prob, t.prob_table.host_ptr, t.count * sizeof(float));
        for (size_t i = 0; i < t.count; i++)
            alias[i] = t.alias_table.host_ptr[i];
    }
    void negative_table(float *prob, Index *alias) override {
        auto &t = solver.workers[0]->negative_sampler;
        memcpy(prob, t.prob_table.host_ptr, t.count * sizeof(float));
        memcpy(alias, t.alias_table.host_ptr, t.count * sizeof(Index));
    }
    void predict(const Index *pairs, size_t n, float *out) override {
        std::vector<typename Solver::EdgeSample> samples(n);
        for (size_t i = 0; i < n; i++)
            samples[i] = std::make_tuple(pairs[i * 2], pairs[i * 2 + 1]);
        std::vector<float> r = solver.predict(samples);
        memcpy(out, r.data(), n * sizeof(float));
    }
};

template<size_t dim>
void run_kernel(int opt_type, const graphvite::Optimizer &opt, size_t num_vertex, size_t num_context, float *vertex,
                float *context, float *vm1, float *cm1, float *vm2, float *cm2, const Index *batch,
                const Index *negatives, int n, int k, float negative_weight, float *loss) {
    using namespace graphvite;
    typedef Vector<dim, float> Vec;
    auto upload = [](Memory<Vec, Index> &m, const float *src, size_t rows) {
        m.resize(rows);
        if (src)
            memcpy(m.host_ptr, src, rows * sizeof(Vec));
        m.to_device();
    };
    auto download = [](Memory<Vec, Index> &m, float *dst) {
        if (!dst)
            return;
        m.to_host();
        memcpy(dst, m.host_ptr, m.count * sizeof(Vec));
    };
    Memory<Vec, Index> v(0), c(0), v1(0), c1(0), v2(0), c2(0);
    upload(v, vertex, num_vertex);
    upload(c, context, num_context);
    Memory<Index, int> b(0), nb(0);
    Memory<float, int> l(0);
    b.resize(n * 2);
    memcpy(b.host_ptr, batch, n * 2 * sizeof(Index));
    b.to_device();
    nb.resize(n * k);
    memcpy(nb.host_ptr, negatives, size_t(n) * k * sizeof(Index));
    nb.to_device();
    l.resize(n);
    const int grid = gpu::kBlockPerGrid, block = gpu::kThreadPerBlock;
    if (opt_type == 0) {
        gpu::graph::train<Vec, Index, LINE, kSGD><<<grid, block>>>(v, c, b, nb, l, opt, negative_weight);
    } else if (opt_type <= 3) {
        upload(v1, vm1, num_vertex);
        upload(c1, cm1, num_context);
        if (opt_type == 1)
            gpu::graph::train_1_moment<Vec, Index, LINE, kMomentum><<<grid, block>>>(v, c, v1, c1, b, nb, l, opt,
                                                                                     negative_weight);
        if (opt_type == 2)
            gpu::graph::train_1_moment<Vec, Index, LINE, kAdaGrad><<<grid, block>>>(v, c, v1, c1, b, nb, l, opt,
                                                                                    negative_weight);
        if (opt_type == 3)
            gpu::graph::train_1_moment<Vec, Index, LINE, kRMSprop><<<grid, block>>>(v, c, v1, c1, b, nb, l, opt,
                                                                                    negative_weight);
    } else {
        upload(v1, vm1, num_vertex);
        upload(c1, cm1, num_context);
        upload(v2, vm2, num_vertex);
        upload(c2, cm2, num_context);
        gpu::graph::train_2_moment<Vec, Index, LINE, kAdam><<<grid, block>>>(v, c, v1, c1, v2, c2, b, nb, l, opt,
                                                                             negative_weight);
    }
    CUDA_CHECK(cudaDeviceSynchronize());
    download(v, vertex);
    download(c, context);
    if (opt_type >= 1) {
        download(v1, vm1);
        download(c1, cm1);
    }
    if (opt_type == 4) {
        download(v2, vm2);
        download(c2, cm2);
    }
    l.to_host();
    memcpy(loss, l.host_ptr, n * sizeof(float));
}

}  // namespace

extern "C" {

// ---- AliasTable (CPU build, CPU sample, GPU Sample kernel) ---------------------
void ref_alias_build(const float *weights, uint64_t n, float *prob, uint64_t *alias) {
    AliasTable<float, size_t> t(-1);
    t.build(std::vector<float>(weights, weights + n));
    memcpy(prob, t.prob_table.host_ptr, n * sizeof(float));
    for (uint64_t i = 0; i < n; i++)
        alias[i] = t.alias_table.host_ptr[i];
}

// the reference's CPU call shape: table.sample(random[r++], random[r++])
void ref_alias_sample_cpu(const float *weights, uint64_t n, const double *random, uint64_t num, uint64_t *out) {
    AliasTable<float, size_t> t(-1);
    t.build(std::vector<float>(weights, weights + n));
    std::vector<double> r(random, random + 2 * num);
    int rand_id = 0;
    for (uint64_t i = 0; i < num; i++)
        out[i] = t.sample(r[rand_id++], r[rand_id++]);
}

void ref_alias_sample_gpu(const float *weights, uint32_t n, const double *random, int num, uint32_t *out) {
    AliasTable<float, Index> t(0);
    t.build(std::vector<float>(weights, weights + n));
    t.to_device();
    Memory<double, int> r(0);
    r.resize(num * 2);
    memcpy(r.host_ptr, random, size_t(num) * 2 * sizeof(double));
    r.to_device();
    Memory<Index, int> result(0);
    result.resize(num);
    t.device_sample(r, &result);
    result.to_host();
    memcpy(out, result.host_ptr, num * sizeof(Index));
}

// ---- cuRAND device generator, seeded and called like core/solver.h:950-953,966 ----
void ref_curand(uint64_t seed, const uint64_t *chunks, int num_chunk, double *out) {
    curandGenerator_t generator;
    CURAND_CHECK(curandCreateGenerator(&generator, CURAND_RNG_PSEUDO_DEFAULT));
    CURAND_CHECK(curandSetPseudoRandomGeneratorSeed(generator, seed));
    for (int i = 0; i < num_chunk; i++) {
        Memory<double, int> r(0);
        r.resize(chunks[i]);
        CURAND_CHECK(curandGenerateUniformDouble(generator, r.device_ptr, chunks[i]));
        r.to_host();
        memcpy(out, r.host_ptr, chunks[i] * sizeof(double));
        out += chunks[i];
    }
    CURAND_CHECK(curandDestroyGenerator(generator));
}

// the process-wide engine, core/solver.h:50 -- draws what a sampler/worker ctor would draw
uint64_t ref_draw_seed() {
    std::uniform_int_distribution<unsigned long long> random_seed(0, ULLONG_MAX);
    return random_seed(graphvite::seed);
}
void ref_reset_engine() { graphvite::seed = std::mt19937(); }

// ---- Graph --------------------------------------------------------------------
void *ref_graph_load(const char *file, int undirected, int normalization) {
    RefGraph *g = new RefGraph();
    g->load_file(file, undirected, normalization);
    return g;
}
uint64_t ref_graph_num_vertex(void *g) { return ((RefGraph *)g)->num_vertex; }
uint64_t ref_graph_num_edge(void *g) { return ((RefGraph *)g)->num_edge; }
uint64_t ref_graph_num_directed_edge(void *g) {
    ((RefGraph *)g)->flatten();
    return ((RefGraph *)g)->edges.size();
}
void ref_graph_flat(void *g, uint32_t *u, uint32_t *v, float *w, float *vertex_weights) {
    RefGraph *G = (RefGraph *)g;
    G->flatten();
    for (size_t i = 0; i < G->edges.size(); i++) {
        u[i] = std::get<0>(G->edges[i]);
        v[i] = std::get<1>(G->edges[i]);
        w[i] = std::get<2>(G->edges[i]);
    }
    memcpy(vertex_weights, G->vertex_weights.data(), G->num_vertex * sizeof(float));
}

// ---- GraphSolver ----------------------------------------------------------------
void *ref_solver_new(int dim, int num_gpu, int samplers_per_worker, uint64_t memory_limit) {
    std::vector<int> devices;
    for (int i = 0; i < num_gpu; i++)
        devices.push_back(i);
    if (dim == 128)
        return new SolverImpl<128>(devices, samplers_per_worker, memory_limit);
    if (dim == 32)
        return new SolverImpl<32>(devices, samplers_per_worker, memory_limit);
    return nullptr;
}
void ref_solver_free(void *s) { delete (SolverBase *)s; }
void ref_solver_build(void *s, void *graph, int opt_type, int schedule, float lr, float wd, float a, float b,
                      float eps, int P, int k, int B, int E) {
    ((SolverBase *)s)->build((RefGraph *)graph, make_optimizer(opt_type, schedule, lr, wd, a, b, eps), P, k, B, E);
}
void ref_solver_train(void *s, const char *model, int epochs, int resume, int aug, int L, int wb, int shuffle_base,
                      float p, float q, int reuse, float exponent, float neg_weight, int log_frequency) {
    ((SolverBase *)s)->train(model, epochs, resume, aug, L, wb, shuffle_base, p, q, reuse, exponent, neg_weight,
                             log_frequency);
}
void ref_solver_info(void *s, int *out) { ((SolverBase *)s)->info(out); }
void ref_solver_locations(void *s, int *part_of, uint32_t *local_of) { ((SolverBase *)s)->locations(part_of, local_of); }
void ref_solver_pool(void *s, int pool, int hp, int tp, uint32_t *out) { ((SolverBase *)s)->pool(pool, hp, tp, out); }
void ref_solver_embeddings(void *s, int which, float *out) { ((SolverBase *)s)->embeddings(which, out); }
void ref_solver_last_negatives(void *s, uint32_t *out) { ((SolverBase *)s)->last_negatives(out); }
void ref_solver_last_loss(void *s, float *out) { ((SolverBase *)s)->last_loss(out); }
void ref_solver_edge_table(void *s, float *prob, uint64_t *alias) { ((SolverBase *)s)->edge_table(prob, alias); }
void ref_solver_negative_table(void *s, float *prob, uint32_t *alias) {
    ((SolverBase *)s)->negative_table(prob, alias);
}
void ref_solver_predict(void *s, const uint32_t *pairs, uint64_t n, float *out) {
    ((SolverBase *)s)->predict(pairs, n, out);
}

// ---- the reference kernels on caller-provided matrices -----------------------------
void ref_kernel_train(int dim, int opt_type, float lr, float wd, float a, float b, float eps, uint64_t num_vertex,
                      uint64_t num_context, float *vertex, float *context, float *vm1, float *cm1, float *vm2,
                      float *cm2, const uint32_t *batch, const uint32_t *negatives, int n, int k,
                      float negative_weight, float *loss) {
    graphvite::Optimizer opt = make_optimizer(opt_type, 0, lr, wd, a, b, eps);
    if (dim == 128)
        run_kernel<128>(opt_type, opt, num_vertex, num_context, vertex, context, vm1, cm1, vm2, cm2, batch,
                        negatives, n, k, negative_weight, loss);
    else if (dim == 32)
        run_kernel<32>(opt_type, opt, num_vertex, num_context, vertex, context, vm1, cm1, vm2, cm2, batch, negatives,
                       n, k, negative_weight, loss);
}

}  // extern "C"
