// =============================================================================
// oracle/ref_harness_kg.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The knowledge-graph counterpart of oracle/ref_harness.cu: a thin C driver around the UNMODIFIED
// reference headers (instance/knowledge_graph.cuh, included from /root/reference/include via
// oracle/Makefile).  oracle/make_golden.py (kg_* cases) runs it on a GPU box to record golden
// vectors that pin oracle/gv_oracle_kg.cpp.  All reference members used here are public.
// Built into oracle/_ref/libref_harness_kg.so (git-ignored).
// =============================================================================
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "instance/knowledge_graph.cuh"

using graphvite::Memory;

typedef unsigned int Index;
typedef graphvite::KnowledgeGraph<Index> RefKG;

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

struct KGSolverBase {
    virtual ~KGSolverBase() {}
    virtual void build(RefKG *graph, const graphvite::Optimizer &opt, int P, int k, int B, int E) = 0;
    virtual void train(const char *model, int epochs, int resume, float relation_lr_multiplier, float margin, float l3,
                       int sample_batch_size, int reuse, float temperature, int log_frequency) = 0;
    virtual void info(int *out) = 0;
    virtual void locations(int *part_of, Index *local_of) = 0;
    virtual void pool(int pool, int hp, int tp, Index *out) = 0;
    virtual void matrix(int which, int order, float *out) = 0;
    virtual void last_negatives(Index *out) = 0;
    virtual void last_loss(float *out) = 0;
    virtual int negative_table(float *prob, Index *alias) = 0;
    virtual int schedule(int *out) = 0;
    virtual void predict(const Index *triplets, size_t n, float *out) = 0;
};

template<size_t dim>
struct KGSolverImpl : KGSolverBase {
    typedef graphvite::KnowledgeGraphSolver<dim, float, Index> Solver;
    Solver solver;
    KGSolverImpl(std::vector<int> devices, int spw, size_t limit) : solver(devices, spw, limit) {}

    void build(RefKG *graph, const graphvite::Optimizer &opt, int P, int k, int B, int E) override {
        solver.build(*graph, opt, P, k, B, E);
    }
    void train(const char *model, int epochs, int resume, float relation_lr_multiplier, float margin, float l3,
               int sample_batch_size, int reuse, float temperature, int log_frequency) override {
        solver.train(model, epochs, resume, relation_lr_multiplier, margin, l3, sample_batch_size, reuse, temperature,
                     log_frequency);
    }
    void info(int *out) override {
        out[0] = solver.num_partition;
        out[1] = solver.episode_size;
        out[2] = solver.batch_size;
        out[3] = solver.num_batch;
        out[4] = solver.batch_id;
        out[5] = solver.pool_id;
        out[6] = solver.num_sampler;
        out[7] = solver.assignment_offset;
        out[8] = solver.workers[0]->negative_sampler.count;
        out[9] = solver.shuffle_partition;
        out[10] = solver.tied_weights;
        out[11] = solver.head_partition_size;
    }
    void locations(int *part_of, Index *local_of) override {
        for (size_t v = 0; v < solver.head_locations.size(); v++) {
            part_of[v] = solver.head_locations[v].first;
            local_of[v] = solver.head_locations[v].second;
        }
    }
    void pool(int pool, int hp, int tp, Index *out) override {
        auto &block = solver.sample_pools[pool][hp][tp];
        memcpy(out, block.data(), block.size() * sizeof(block[0]));  // {relation, tail, head} per sample
    }
    // which: 0 entity, 1 relation; order: 0 embeddings, 1 / 2 the solver-side moments
    void matrix(int which, int order, float *out) override {
        int id = which == 0 ? 0 : 2;
        if (order == 0) {
            auto &e = *solver.embeddings[id];
            memcpy(out, e.data(), e.size() * sizeof(e[0]));
        } else {
            auto &m = (*solver.moments[id])[order - 1];
            memcpy(out, m.data(), m.size() * sizeof(m[0]));
        }
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
    int negative_table(float *prob, Index *alias) override {
        auto &t = solver.workers[0]->negative_sampler;
        if (prob) {
            memcpy(prob, t.prob_table.host_ptr, t.count * sizeof(float));
            memcpy(alias, t.alias_table.host_ptr, t.count * sizeof(Index));
        }
        return t.count;
    }
    int schedule(int *out) override {
        auto s = solver.get_schedule();
        for (auto &step : s)
            for (auto &assignment : step) {
                *out++ = assignment.first;
                *out++ = assignment.second;
            }
        return int(s.size());
    }
    void predict(const Index *triplets, size_t n, float *out) override {
        std::vector<typename Solver::EdgeSample> samples(n);
        for (size_t i = 0; i < n; i++)
            samples[i] = std::make_tuple(triplets[i * 3], triplets[i * 3 + 1], triplets[i * 3 + 2]);
        std::vector<float> r = solver.predict(samples);
        memcpy(out, r.data(), n * sizeof(float));
    }
};

// the reference kernels on caller-provided matrices (batches chosen race-free by the caller)
template<size_t dim, template<class> class Model>
void run_kernel(int opt_type, const graphvite::Optimizer &opt, size_t num_head, size_t num_tail, size_t num_relation,
                bool shared, float *head, float *tail, float *relation, float *hm1, float *tm1, float *rm1, float *hm2,
                float *tm2, float *rm2, const Index *batch, const Index *negatives, int n, int k,
                float relation_lr_multiplier, float margin_or_l3, float temperature, float *loss) {
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
    Memory<Vec, Index> h(0), t(0), r(0), h1(0), t1(0), r1(0), h2(0), t2(0), r2(0);
    upload(h, head, num_head);
    upload(r, relation, num_relation);
    if (!shared)
        upload(t, tail, num_tail);
    if (opt_type >= 1) {
        upload(h1, hm1, num_head);
        upload(r1, rm1, num_relation);
        if (!shared)
            upload(t1, tm1, num_tail);
    }
    if (opt_type == 4) {
        upload(h2, hm2, num_head);
        upload(r2, rm2, num_relation);
        if (!shared)
            upload(t2, tm2, num_tail);
    }
    // shared: the tail block IS the head block (one partition, core/solver.h:1351-1355) -- shallow copies
    Memory<Vec, Index> &T = shared ? h : t, &T1 = shared ? h1 : t1, &T2 = shared ? h2 : t2;
    Memory<Index, int> b(0), nb(0);
    Memory<float, int> l(0);
    b.resize(n * 3);
    memcpy(b.host_ptr, batch, size_t(n) * 3 * sizeof(Index));
    b.to_device();
    nb.resize(n * k);
    memcpy(nb.host_ptr, negatives, size_t(n) * k * sizeof(Index));
    nb.to_device();
    l.resize(n);
    const int grid = gpu::kBlockPerGrid, block = gpu::kThreadPerBlock;
    namespace kg = gpu::knowledge_graph;
    switch (opt_type) {
        case 0:
            kg::train<Vec, Index, Model, kSGD><<<grid, block>>>(h, T, r, b, nb, l, opt, relation_lr_multiplier,
                                                                 margin_or_l3, temperature);
            break;
        case 1:
            kg::train_1_moment<Vec, Index, Model, kMomentum><<<grid, block>>>(h, T, r, h1, T1, r1, b, nb, l, opt,
                                                                               relation_lr_multiplier, margin_or_l3,
                                                                               temperature);
            break;
        case 2:
            kg::train_1_moment<Vec, Index, Model, kAdaGrad><<<grid, block>>>(h, T, r, h1, T1, r1, b, nb, l, opt,
                                                                              relation_lr_multiplier, margin_or_l3,
                                                                              temperature);
            break;
        case 3:
            kg::train_1_moment<Vec, Index, Model, kRMSprop><<<grid, block>>>(h, T, r, h1, T1, r1, b, nb, l, opt,
                                                                              relation_lr_multiplier, margin_or_l3,
                                                                              temperature);
            break;
        default:
            kg::train_2_moment<Vec, Index, Model, kAdam><<<grid, block>>>(h, T, r, h1, T1, r1, h2, T2, r2, b, nb, l, opt,
                                                                           relation_lr_multiplier, margin_or_l3,
                                                                           temperature);
    }
    CUDA_CHECK(cudaDeviceSynchronize());
    download(h, head);
    download(r, relation);
    if (!shared)
        download(t, tail);
    if (opt_type >= 1) {
        download(h1, hm1);
        download(r1, rm1);
        if (!shared)
            download(t1, tm1);
    }
    if (opt_type == 4) {
        download(h2, hm2);
        download(r2, rm2);
        if (!shared)
            download(t2, tm2);
    }
    l.to_host();
    memcpy(loss, l.host_ptr, n * sizeof(float));
}

template<size_t dim, template<class> class Model>
void run_predict(size_t num_entity, size_t num_relation, const float *entity, const float *relation,
                 const Index *batch, int n, float margin, float *logits) {
    using namespace graphvite;
    typedef Vector<dim, float> Vec;
    Memory<Vec, Index> e(0), r(0);
    e.resize(num_entity);
    memcpy(e.host_ptr, entity, num_entity * sizeof(Vec));
    e.to_device();
    r.resize(num_relation);
    memcpy(r.host_ptr, relation, num_relation * sizeof(Vec));
    r.to_device();
    Memory<Index, int> b(0);
    Memory<float, int> l(0);
    b.resize(n * 3);
    memcpy(b.host_ptr, batch, size_t(n) * 3 * sizeof(Index));
    b.to_device();
    l.resize(n);
    gpu::knowledge_graph::predict<Vec, Index, Model><<<gpu::kBlockPerGrid, gpu::kThreadPerBlock>>>(e, e, r, b, l,
                                                                                                   margin);
    CUDA_CHECK(cudaDeviceSynchronize());
    l.to_host();
    memcpy(logits, l.host_ptr, n * sizeof(float));
}

template<size_t dim>
int dispatch_kernel(const std::string &model, int opt_type, const graphvite::Optimizer &opt, size_t num_head,
                    size_t num_tail, size_t num_relation, bool shared, float *head, float *tail, float *relation,
                    float *hm1, float *tm1, float *rm1, float *hm2, float *tm2, float *rm2, const Index *batch,
                    const Index *negatives, int n, int k, float rlm, float margin_or_l3, float temperature,
                    float *loss) {
    using namespace graphvite;
#define RUN(M)                                                                                                    \
    run_kernel<dim, M>(opt_type, opt, num_head, num_tail, num_relation, shared, head, tail, relation, hm1, tm1, rm1, \
                       hm2, tm2, rm2, batch, negatives, n, k, rlm, margin_or_l3, temperature, loss)
    if (model == "TransE")
        RUN(TransE);
    else if (model == "DistMult")
        RUN(DistMult);
    else if (model == "ComplEx")
        RUN(ComplEx);
    else if (model == "SimplE")
        RUN(SimplE);
    else if (model == "RotatE")
        RUN(RotatE);
    else if (model == "QuatE")
        RUN(QuatE);
    else
        return -1;
#undef RUN
    return 0;
}

template<size_t dim>
int dispatch_predict(const std::string &model, size_t num_entity, size_t num_relation, const float *entity,
                     const float *relation, const Index *batch, int n, float margin, float *logits) {
    using namespace graphvite;
    if (model == "TransE")
        run_predict<dim, TransE>(num_entity, num_relation, entity, relation, batch, n, margin, logits);
    else if (model == "DistMult")
        run_predict<dim, DistMult>(num_entity, num_relation, entity, relation, batch, n, margin, logits);
    else if (model == "ComplEx")
        run_predict<dim, ComplEx>(num_entity, num_relation, entity, relation, batch, n, margin, logits);
    else if (model == "SimplE")
        run_predict<dim, SimplE>(num_entity, num_relation, entity, relation, batch, n, margin, logits);
    else if (model == "RotatE")
        run_predict<dim, RotatE>(num_entity, num_relation, entity, relation, batch, n, margin, logits);
    else if (model == "QuatE")
        run_predict<dim, QuatE>(num_entity, num_relation, entity, relation, batch, n, margin, logits);
    else
        return -1;
    return 0;
}

}  // namespace

extern "C" {

void rk_reset_engine() { graphvite::seed = std::mt19937(); }

// ---- KnowledgeGraph ---------------------------------------------------------------
void *rk_graph_load(const char *file, int normalization) {
    RefKG *g = new RefKG();
    g->load_file(file, normalization);
    return g;
}
void rk_graph_free(void *g) { delete (RefKG *)g; }
void rk_graph_sizes(void *g, uint64_t *out) {
    RefKG *G = (RefKG *)g;
    out[0] = G->num_vertex;
    out[1] = G->num_edge;
    out[2] = G->num_relation;
}
void rk_graph_flat(void *g, uint32_t *h, uint32_t *t, uint32_t *r, float *w, float *vertex_weights) {
    RefKG *G = (RefKG *)g;
    G->flatten();
    for (size_t i = 0; i < G->edges.size(); i++) {
        h[i] = std::get<0>(G->edges[i]);
        t[i] = std::get<1>(G->edges[i]);
        w[i] = std::get<2>(G->edges[i]);
        r[i] = std::get<3>(G->edges[i]);
    }
    memcpy(vertex_weights, G->vertex_weights.data(), G->num_vertex * sizeof(float));
}

// ---- KnowledgeGraphSolver -----------------------------------------------------------
void *rk_solver_new(int dim, int num_gpu, int samplers_per_worker, uint64_t memory_limit) {
    std::vector<int> devices;
    for (int i = 0; i < num_gpu; i++)
        devices.push_back(i);
    if (dim == 32)
        return new KGSolverImpl<32>(devices, samplers_per_worker, memory_limit);
    if (dim == 64)
        return new KGSolverImpl<64>(devices, samplers_per_worker, memory_limit);
    return nullptr;
}
void rk_solver_free(void *s) { delete (KGSolverBase *)s; }
void rk_solver_build(void *s, void *graph, int opt_type, int schedule, float lr, float wd, float a, float b, float eps,
                     int P, int k, int B, int E) {
    ((KGSolverBase *)s)->build((RefKG *)graph, make_optimizer(opt_type, schedule, lr, wd, a, b, eps), P, k, B, E);
}
void rk_solver_train(void *s, const char *model, int epochs, int resume, float relation_lr_multiplier, float margin,
                     float l3, int sample_batch_size, int reuse, float temperature, int log_frequency) {
    ((KGSolverBase *)s)->train(model, epochs, resume, relation_lr_multiplier, margin, l3, sample_batch_size, reuse,
                               temperature, log_frequency);
}
void rk_solver_info(void *s, int *out) { ((KGSolverBase *)s)->info(out); }
void rk_solver_locations(void *s, int *part_of, uint32_t *local_of) { ((KGSolverBase *)s)->locations(part_of, local_of); }
void rk_solver_pool(void *s, int pool, int hp, int tp, uint32_t *out) { ((KGSolverBase *)s)->pool(pool, hp, tp, out); }
void rk_solver_matrix(void *s, int which, int order, float *out) { ((KGSolverBase *)s)->matrix(which, order, out); }
void rk_solver_last_negatives(void *s, uint32_t *out) { ((KGSolverBase *)s)->last_negatives(out); }
void rk_solver_last_loss(void *s, float *out) { ((KGSolverBase *)s)->last_loss(out); }
int rk_solver_negative_table(void *s, float *prob, uint32_t *alias) {
    return ((KGSolverBase *)s)->negative_table(prob, alias);
}
int rk_solver_schedule(void *s, int *out) { return ((KGSolverBase *)s)->schedule(out); }
void rk_solver_predict(void *s, const uint32_t *triplets, uint64_t n, float *out) {
    ((KGSolverBase *)s)->predict(triplets, n, out);
}

// ---- the reference kernels on caller-provided matrices -----------------------------------
// shared != 0: one entity matrix serves as head and tail block (tail* pointers ignored)
int rk_kernel_train(const char *model, int dim, int opt_type, float lr, float wd, float a, float b, float eps,
                    uint64_t num_head, uint64_t num_tail, uint64_t num_relation, int shared, float *head, float *tail,
                    float *relation, float *hm1, float *tm1, float *rm1, float *hm2, float *tm2, float *rm2,
                    const uint32_t *batch, const uint32_t *negatives, int n, int k, float relation_lr_multiplier,
                    float margin_or_l3, float temperature, float *loss) {
    graphvite::Optimizer opt = make_optimizer(opt_type, 0, lr, wd, a, b, eps);
    if (dim == 32)
        return dispatch_kernel<32>(model, opt_type, opt, num_head, num_tail, num_relation, shared != 0, head, tail,
                                   relation, hm1, tm1, rm1, hm2, tm2, rm2, batch, negatives, n, k,
                                   relation_lr_multiplier, margin_or_l3, temperature, loss);
    if (dim == 512)
        return dispatch_kernel<512>(model, opt_type, opt, num_head, num_tail, num_relation, shared != 0, head, tail,
                                    relation, hm1, tm1, rm1, hm2, tm2, rm2, batch, negatives, n, k,
                                    relation_lr_multiplier, margin_or_l3, temperature, loss);
    return -1;
}

int rk_kernel_predict(const char *model, int dim, uint64_t num_entity, uint64_t num_relation, const float *entity,
                      const float *relation, const uint32_t *batch, int n, float margin, float *logits) {
    if (dim == 32)
        return dispatch_predict<32>(model, num_entity, num_relation, entity, relation, batch, n, margin, logits);
    if (dim == 512)
        return dispatch_predict<512>(model, num_entity, num_relation, entity, relation, batch, n, margin, logits);
    return -1;
}

}  // extern "C"
