// =============================================================================
// oracle/gv_oracle_kg.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Sequential CPU restatement of the knowledge-graph embedding path of
// DeepGraphLearning/graphvite v0.2.2: KnowledgeGraph,
// KnowledgeGraphSolver (tied head / tail entity matrix, global relation matrix, uniform negative
// sampling over head + tail partition, self-adversarial weighting) with the models TransE,
// DistMult, ComplEx, SimplE, RotatE and QuatE and all five optimizers.
//
// PARITY STATUS: pinned against the UNMODIFIED reference.  The graph loader is compared with the live reference
// object on CPU (tests/test_reference_surface.py).  Kernels and solver are compared with golden vectors
// (tests/golden/kg_kernel_*, kg_predict_*, kg_solver_*.npz) that oracle/make_golden_kg.py --emulated recorded by
// running the reference's own headers through oracle/ref_harness_kg.cu under the CUDA emulation of tests/emu
// (two-pass build like nvcc's, see oracle/Makefile target ref_emu and oracle/emulate_reference.py): train kernels
// of the 5 models x 5 optimizers and predict to rtol 5e-4; six solver runs with partition, both sample pools, last
// negatives, schedule and batch accounting bit-exact, and -- for one partition -- embeddings, loss and predict of
// the whole training run to rtol 1e-4 (same sample order; host libm / no-FMA rounding on both sides).  What this
// does NOT cover: the device's own rounding (libdevice, FMA contraction) -- rerun make_golden_kg.py on a GPU for
// that.  With several partitions the reference's partition cache can train a stale second copy of an entity
// partition (KGSolver::reference_cache below): the default restatement trains in place, as the protocol intends,
// and is compared on integer state and magnitudes there; with reference_cache switched on it reproduces those
// runs float for float as well (tests/test_oracle_kg_golden.py::test_solver_runs).
//
// Paths are relative to /root/reference/include.  Workers are emulated one after another with
// sequentially consistent entity matrices (the reference races its write-backs against the other
// workers' loads, core/solver.h:1511-1514); with one worker the emulation is exact.
// =============================================================================
#include <tuple>

#include "gv_oracle_common.h"

namespace oracle {

// core/solver.h:55
static const int kSamplePerVertexWithGlobal = 50;

// -----------------------------------------------------------------------------
// KnowledgeGraph, instance/knowledge_graph.cuh:67-284 over core/graph.h:45-101
// -----------------------------------------------------------------------------
struct KGraph {
    std::unordered_map<std::string, Index> entity2id, relation2id;
    std::vector<std::string> id2entity, id2relation;
    std::vector<std::vector<std::tuple<Index, float, Index>>> vertex_edges;  // (tail, weight, relation)
    std::vector<float> vertex_weights, edge_weights;
    std::vector<Index> edge_h, edge_t, edge_r;  // flatten(): (head, tail, weight, relation) in vertex order
    Index num_vertex = 0, num_relation = 0;
    size_t num_edge = 0;
    bool normalization = false;

    // :135-168 -- ids are handed out in the order head, relation, tail
    void add_edge(const std::string &h_name, const std::string &r_name, const std::string &t_name, float w) {
        Index h, t, r;
        auto h_iter = entity2id.find(h_name);
        if (h_iter != entity2id.end())
            h = h_iter->second;
        else {
            h = num_vertex++;
            entity2id[h_name] = h;
            id2entity.push_back(h_name);
            vertex_edges.emplace_back();
            vertex_weights.push_back(0);
        }
        auto r_iter = relation2id.find(r_name);
        if (r_iter != relation2id.end())
            r = r_iter->second;
        else {
            r = num_relation++;
            relation2id[r_name] = r;
            id2relation.push_back(r_name);
        }
        auto t_iter = entity2id.find(t_name);
        if (t_iter != entity2id.end())
            t = t_iter->second;
        else {
            t = num_vertex++;
            entity2id[t_name] = t;
            id2entity.push_back(t_name);
            vertex_edges.emplace_back();
            vertex_weights.push_back(0);
        }
        vertex_edges[h].push_back(std::make_tuple(t, w, r));
        vertex_weights[h] += w;
        num_edge++;
    }

    // :95-121
    void normalize() {
        std::vector<std::unordered_map<Index, float>> head_weights(num_vertex), tail_weights(num_vertex);
        for (Index h = 0; h < num_vertex; h++)
            for (auto &e : vertex_edges[h]) {
                Index t = std::get<0>(e), r = std::get<2>(e);
                float w = std::get<1>(e);
                head_weights[h][r] += w;
                tail_weights[t][r] += w;
            }
        for (Index h = 0; h < num_vertex; h++) {
            float weight = 0;
            for (auto &e : vertex_edges[h]) {
                Index t = std::get<0>(e), r = std::get<2>(e);
                float &w = std::get<1>(e);
                w /= sqrtf(head_weights[h][r] * tail_weights[t][r]);  // nvcc resolves sqrt(float) to the float overload
                weight += w;
            }
            vertex_weights[h] = weight;
        }
    }

    // :177-213
    void load_file(const char *file_name, bool _normalization) {
        *this = KGraph();
        normalization = _normalization;
        FILE *fin = fopen(file_name, "r");
        if (!fin)
            fail(std::string("File `") + file_name + "` doesn't exist");
        std::vector<char> line(size_t(1) << 22);
        const char *delimiters = " \t\r\n", *comment = "#";
        for (size_t i = 1; fgets(line.data(), int(line.size()), fin); i++) {
            char *comment_str = strstr(line.data(), comment);
            if (comment_str)
                *comment_str = 0;
            char *h_name = strtok(line.data(), delimiters);
            if (!h_name)
                continue;
            char *r_name = strtok(nullptr, delimiters);
            char *t_name = strtok(nullptr, delimiters);
            char *w_str = strtok(nullptr, delimiters);
            char *more = strtok(nullptr, delimiters);
            if (!t_name || more) {
                fclose(fin);
                fail("Invalid format at line " + std::to_string(i));
            }
            float w = w_str ? atof(w_str) : 1;
            add_edge(h_name, r_name, t_name, w);
        }
        fclose(fin);
        if (normalization)
            normalize();
    }

    // core/graph.h:87-101
    void flatten() {
        if (!edge_h.empty())
            return;
        for (Index u = 0; u < num_vertex; u++)
            for (auto &e : vertex_edges[u]) {
                edge_h.push_back(u);
                edge_t.push_back(std::get<0>(e));
                edge_r.push_back(std::get<2>(e));
                edge_weights.push_back(std::get<1>(e));
            }
    }
};

// -----------------------------------------------------------------------------
// util/gpu.cuh:26-29,46-65: a per-lane value reduced by the shfl_down tree, lane 0 broadcast
// -----------------------------------------------------------------------------
static float warp_sum(float *lane) {
    for (int delta = 1; delta < 32; delta *= 2)
        for (int l = 0; l + delta < 32; l++)
            lane[l] = lane[l] + lane[l + delta];
    return lane[0];
}

enum ModelType { kTransE = 0, kDistMult, kComplEx, kSimplE, kRotatE, kQuatE };

static int model_id(const std::string &name) {
    static const char *names[] = {"TransE", "DistMult", "ComplEx", "SimplE", "RotatE", "QuatE"};
    for (int i = 0; i < 6; i++)
        if (name == names[i])
            return i;
    return -1;
}

// -----------------------------------------------------------------------------
// Model::forward, instance/model/knowledge_graph.h:44-49 (TransE), :117-123 (DistMult), :208-223
// (ComplEx), :359-366 (SimplE), :453-468 (RotatE).  FOR(i, n) strides lanes over i; the products
// follow nvcc's default contraction only loosely (we do not contract), which moves results by ~1 ulp.
// -----------------------------------------------------------------------------
static float kg_forward(int model, int dim, const float *head, const float *tail, const float *relation,
                        float margin_or_l3) {
    float lane[32];
    for (int l = 0; l < 32; l++) {
        float output = 0;
        switch (model) {
            case kTransE:
                for (int i = l; i < dim; i += 32)
                    output += fabsf(head[i] + relation[i] - tail[i]);
                break;
            case kDistMult:
                for (int i = l; i < dim; i += 32)
                    output += head[i] * relation[i] * tail[i];
                break;
            case kComplEx:
                for (int i = l; i < dim / 2; i += 32) {
                    float h_re = head[i * 2], h_im = head[i * 2 + 1];
                    float t_re = tail[i * 2], t_im = tail[i * 2 + 1];
                    float r_re = relation[i * 2], r_im = relation[i * 2 + 1];
                    float product_re = h_re * r_re - h_im * r_im;
                    float product_im = h_re * r_im + h_im * r_re;
                    output += product_re * t_re + product_im * t_im;
                }
                break;
            case kSimplE:
                for (int i = l; i < dim; i += 32)
                    output += head[i] * relation[i] * tail[i ^ 1];
                break;
            case kQuatE:  // :594-618
                for (int i = l; i < dim / 4; i += 32) {
                    float h_r = head[i * 4], h_i = head[i * 4 + 1], h_j = head[i * 4 + 2], h_k = head[i * 4 + 3];
                    float r_r = relation[i * 4], r_i = relation[i * 4 + 1], r_j = relation[i * 4 + 2],
                          r_k = relation[i * 4 + 3];
                    float t_r = tail[i * 4], t_i = tail[i * 4 + 1], t_j = tail[i * 4 + 2], t_k = tail[i * 4 + 3];
                    float r_norm = sqrtf(r_r * r_r + r_i * r_i + r_j * r_j + r_k * r_k);
                    float product_r = h_r * r_r - h_i * r_i - h_j * r_j - h_k * r_k;
                    float product_i = h_r * r_i + h_i * r_r + h_j * r_k - h_k * r_j;
                    float product_j = h_r * r_j - h_i * r_k + h_j * r_r + h_k * r_i;
                    float product_k = h_r * r_k + h_i * r_j - h_j * r_i + h_k * r_r;
                    output += (product_r * t_r + product_i * t_i + product_j * t_j + product_k * t_k) / (r_norm + kEpsilon);
                }
                break;
            default:  // RotatE
                for (int i = l; i < dim / 2; i += 32) {
                    float h_re = head[i * 2], h_im = head[i * 2 + 1];
                    float t_re = tail[i * 2], t_im = tail[i * 2 + 1];
                    float phase = relation[i];
                    float r_re = cosf(phase), r_im = sinf(phase);
                    float distance_re = h_re * r_re - h_im * r_im - t_re;
                    float distance_im = h_re * r_im + h_im * r_re - t_im;
                    output += sqrtf(distance_re * distance_re + distance_im * distance_im);
                }
        }
        lane[l] = output;
    }
    float sum = warp_sum(lane);
    if (model == kTransE || model == kRotatE)
        return margin_or_l3 - sum;
    return sum;
}

// rows of one (head, tail, relation) target with their moment rows (null when the optimizer has none)
struct KGRows {
    float *head, *tail, *relation;
    float *head_m1, *tail_m1, *relation_m1;
    float *head_m2, *tail_m2, *relation_m2;
};

// -----------------------------------------------------------------------------
// Model::backward for 0 / 1 / 2 moments, instance/model/knowledge_graph.h:51-108 (TransE),
// :125-190 (DistMult), :225-340 (ComplEx), :368-433 (SimplE), :470-575 (RotatE).
// Statement order is kept, so aliased rows (head == tail entity) behave as in the reference.
// -----------------------------------------------------------------------------
static void kg_backward(int model, int dim, const KGRows &x, const Optimizer &opt, float margin_or_l3,
                        float gradient, float relation_lr_multiplier, float weight) {
    auto m1 = [](float *rows, int i) { return rows ? rows + i : nullptr; };
    auto up = [&](float parameter, float grad, float *a, float *b) { return opt.update(parameter, grad, a, b, weight); };
    if (model == kTransE) {
        for (int i = 0; i < dim; i++) {
            float h = x.head[i], t = x.tail[i], r = x.relation[i];
            float s = h + r - t > 0 ? 1 : -1;
            x.head[i] -= up(h, -gradient * s, m1(x.head_m1, i), m1(x.head_m2, i));
            x.tail[i] -= up(t, gradient * s, m1(x.tail_m1, i), m1(x.tail_m2, i));
            x.relation[i] -= relation_lr_multiplier * up(r, -gradient * s, m1(x.relation_m1, i), m1(x.relation_m2, i));
        }
    } else if (model == kDistMult || model == kSimplE) {
        float l3 = margin_or_l3 * 3;
        for (int i = 0; i < dim; i++) {
            int j = model == kSimplE ? (i ^ 1) : i;
            float h = x.head[i], t = x.tail[j], r = x.relation[i];
            x.head[i] -= up(h, gradient * r * t + l3 * fabsf(h) * h, m1(x.head_m1, i), m1(x.head_m2, i));
            x.tail[j] -= up(t, gradient * h * r + l3 * fabsf(t) * t, m1(x.tail_m1, j), m1(x.tail_m2, j));
            x.relation[i] -= relation_lr_multiplier *
                             up(r, gradient * h * t + l3 * fabsf(r) * r, m1(x.relation_m1, i), m1(x.relation_m2, i));
        }
    } else if (model == kComplEx) {
        float l3 = margin_or_l3 * 3;
        for (int i = 0; i < dim / 2; i++) {
            float h_re = x.head[i * 2], h_im = x.head[i * 2 + 1];
            float t_re = x.tail[i * 2], t_im = x.tail[i * 2 + 1];
            float r_re = x.relation[i * 2], r_im = x.relation[i * 2 + 1];
            float h_re_grad = gradient * (r_re * t_re + r_im * t_im);
            float h_im_grad = gradient * (-r_im * t_re + r_re * t_im);
            x.head[i * 2] -= up(h_re, h_re_grad + l3 * fabsf(h_re) * h_re, m1(x.head_m1, i * 2), m1(x.head_m2, i * 2));
            x.head[i * 2 + 1] -=
                up(h_im, h_im_grad + l3 * fabsf(h_im) * h_im, m1(x.head_m1, i * 2 + 1), m1(x.head_m2, i * 2 + 1));
            float t_re_grad = gradient * (h_re * r_re - h_im * r_im);
            float t_im_grad = gradient * (h_re * r_im + h_im * r_re);
            x.tail[i * 2] -= up(t_re, t_re_grad + l3 * fabsf(t_re) * t_re, m1(x.tail_m1, i * 2), m1(x.tail_m2, i * 2));
            x.tail[i * 2 + 1] -=
                up(t_im, t_im_grad + l3 * fabsf(t_im) * t_im, m1(x.tail_m1, i * 2 + 1), m1(x.tail_m2, i * 2 + 1));
            float r_re_grad = gradient * (h_re * t_re + h_im * t_im);
            float r_im_grad = gradient * (-h_im * t_re + h_re * t_im);
            // the reference indexes the relation moments by i for BOTH components (:302-306,:336-340)
            x.relation[i * 2] -= relation_lr_multiplier * up(r_re, r_re_grad + l3 * fabsf(r_re) * r_re,
                                                             m1(x.relation_m1, i), m1(x.relation_m2, i));
            x.relation[i * 2 + 1] -= relation_lr_multiplier * up(r_im, r_im_grad + l3 * fabsf(r_im) * r_im,
                                                                 m1(x.relation_m1, i), m1(x.relation_m2, i));
        }
    } else if (model == kQuatE) {  // :620-860, per-element moments for all three rows
        float l3 = margin_or_l3 * 3;
        for (int i = 0; i < dim / 4; i++) {
            int q = i * 4;
            float h_r = x.head[q], h_i = x.head[q + 1], h_j = x.head[q + 2], h_k = x.head[q + 3];
            float r_r = x.relation[q], r_i = x.relation[q + 1], r_j = x.relation[q + 2], r_k = x.relation[q + 3];
            float t_r = x.tail[q], t_i = x.tail[q + 1], t_j = x.tail[q + 2], t_k = x.tail[q + 3];
            float r_norm = sqrtf(r_r * r_r + r_i * r_i + r_j * r_j + r_k * r_k);
            float grad = gradient / (r_norm + kEpsilon);
            float h_grad[4] = {grad * (r_r * t_r + r_i * t_i + r_j * t_j + r_k * t_k),
                               grad * (-r_i * t_r + r_r * t_i - r_k * t_j + r_j * t_k),
                               grad * (-r_j * t_r + r_k * t_i + r_r * t_j - r_i * t_k),
                               grad * (-r_k * t_r - r_j * t_i + r_i * t_j + r_r * t_k)};
            float h_old[4] = {h_r, h_i, h_j, h_k};
            for (int c = 0; c < 4; c++)
                x.head[q + c] -= up(h_old[c], h_grad[c] + l3 * fabsf(h_old[c]) * h_old[c], m1(x.head_m1, q + c),
                                    m1(x.head_m2, q + c));
            float t_grad[4] = {grad * (h_r * r_r - h_i * r_i - h_j * r_j - h_k * r_k),
                               grad * (h_r * r_i + h_i * r_r + h_j * r_k - h_k * r_j),
                               grad * (h_r * r_j - h_i * r_k + h_j * r_r + h_k * r_i),
                               grad * (h_r * r_k + h_i * r_j - h_j * r_i + h_k * r_r)};
            float t_old[4] = {t_r, t_i, t_j, t_k};
            for (int c = 0; c < 4; c++)
                x.tail[q + c] -= up(t_old[c], t_grad[c] + l3 * fabsf(t_old[c]) * t_old[c], m1(x.tail_m1, q + c),
                                    m1(x.tail_m2, q + c));
            float r_grad[4] = {grad * (h_r * t_r + h_i * t_i + h_j * t_j + h_k * t_k),
                               grad * (-h_i * t_r + h_r * t_i + h_k * t_j - h_j * t_k),
                               grad * (-h_j * t_r - h_k * t_i + h_r * t_j + h_i * t_k),
                               grad * (-h_k * t_r + h_j * t_i - h_i * t_j + h_r * t_k)};
            float r_old[4] = {r_r, r_i, r_j, r_k};
            for (int c = 0; c < 4; c++)
                x.relation[q + c] -= relation_lr_multiplier * up(r_old[c], r_grad[c] + l3 * fabsf(r_old[c]) * r_old[c],
                                                                 m1(x.relation_m1, q + c), m1(x.relation_m2, q + c));
        }
    } else {  // RotatE
        for (int i = 0; i < dim / 2; i++) {
            float phase = x.relation[i];
            float r_re = cosf(phase), r_im = sinf(phase);
            float h_re = x.head[i * 2], h_im = x.head[i * 2 + 1];
            float t_re = x.tail[i * 2], t_im = x.tail[i * 2 + 1];
            float distance_re = h_re * r_re - h_im * r_im - t_re;
            float distance_im = h_re * r_im + h_im * r_re - t_im;
            float grad = gradient / (sqrtf(distance_re * distance_re + distance_im * distance_im) + kEpsilon);
            float head_re_grad = -grad * (distance_re * r_re + distance_im * r_im);
            float head_im_grad = -grad * (-distance_re * r_im + distance_im * r_re);
            x.head[i * 2] -= up(h_re, head_re_grad, m1(x.head_m1, i * 2), m1(x.head_m2, i * 2));
            x.head[i * 2 + 1] -= up(h_im, head_im_grad, m1(x.head_m1, i * 2 + 1), m1(x.head_m2, i * 2 + 1));
            x.tail[i * 2] -= up(t_re, grad * distance_re, m1(x.tail_m1, i * 2), m1(x.tail_m2, i * 2));
            x.tail[i * 2 + 1] -= up(t_im, grad * distance_im, m1(x.tail_m1, i * 2 + 1), m1(x.tail_m2, i * 2 + 1));
            float relation_grad =
                -grad * (distance_re * (h_re * -r_im + h_im * -r_re) + distance_im * (h_re * r_re + h_im * -r_im));
            x.relation[i] -=
                relation_lr_multiplier * up(phase, relation_grad, m1(x.relation_m1, i), m1(x.relation_m2, i));
        }
    }
}

// util/math.h:36-44
static float safe_exp(float x) {
    const float kLogitClip = 80;
    return expf(std::min(std::max(x, -kLogitClip), kLogitClip));
}

// matrices one kernel launch sees: head and tail entity blocks (the same memory when the two
// partitions coincide) and the worker's relation copy, each with optional moments
struct KGMatrices {
    int dim = 0;
    Index num_head = 0;  // rows of the head block: negative ids below it corrupt the head
    float *head = nullptr, *tail = nullptr, *relation = nullptr;
    float *head_m1 = nullptr, *tail_m1 = nullptr, *relation_m1 = nullptr;
    float *head_m2 = nullptr, *tail_m2 = nullptr, *relation_m2 = nullptr;
    // optional local -> row maps (null: identity); lets the solver train in place on the global matrix
    const Index *head_rows = nullptr, *tail_rows = nullptr;
};

// -----------------------------------------------------------------------------
// gpu::knowledge_graph::train / train_1_moment / train_2_moment for ONE positive sample,
// instance/gpu/knowledge_graph.cuh:38-122,132-225,234-331.  sample = {relation, tail, head}.
// -----------------------------------------------------------------------------
static float kg_train_sample(int model, const KGMatrices &m, const Optimizer &opt, const Index *sample,
                             const Index *negatives, int num_negative, float relation_lr_multiplier,
                             float margin_or_l3, float adversarial_temperature) {
    const size_t dim = m.dim;
    auto row = [dim](float *base, const Index *rows, Index id) -> float * {
        return base ? base + size_t(rows ? rows[id] : id) * dim : nullptr;
    };
    const Index relation_id = sample[0];
    KGRows x;
    x.relation = m.relation + relation_id * dim;
    x.relation_m1 = m.relation_m1 ? m.relation_m1 + relation_id * dim : nullptr;
    x.relation_m2 = m.relation_m2 ? m.relation_m2 + relation_id * dim : nullptr;
    auto corrupt = [&](int s, Index &head_id, Index &tail_id) {
        Index negative_id = negatives[s];
        if (negative_id < m.num_head)
            head_id = negative_id;
        else
            tail_id = negative_id - m.num_head;
    };
    float bias = 0, normalizer = 0;
    if (adversarial_temperature > kEpsilon)
        for (int s = 0; s < num_negative; s++) {
            Index head_id = sample[2], tail_id = sample[1];
            corrupt(s, head_id, tail_id);
            float logit = kg_forward(model, m.dim, row(m.head, m.head_rows, head_id), row(m.tail, m.tail_rows, tail_id),
                                     x.relation, margin_or_l3);
            if (s == 0)
                bias = logit;
            normalizer += safe_exp((logit - bias) / adversarial_temperature);
        }
    float sample_loss = 0;
    for (int s = 0; s <= num_negative; s++) {
        Index head_id = sample[2], tail_id = sample[1];
        int label = 1;
        if (s < num_negative) {
            corrupt(s, head_id, tail_id);
            label = 0;
        }
        x.head = row(m.head, m.head_rows, head_id);
        x.tail = row(m.tail, m.tail_rows, tail_id);
        x.head_m1 = row(m.head_m1, m.head_rows, head_id);
        x.tail_m1 = row(m.tail_m1, m.tail_rows, tail_id);
        x.head_m2 = row(m.head_m2, m.head_rows, head_id);
        x.tail_m2 = row(m.tail_m2, m.tail_rows, tail_id);
        float logit = kg_forward(model, m.dim, x.head, x.tail, x.relation, margin_or_l3);
        float prob = sigmoid(logit);
        float gradient, weight;
        if (label) {
            gradient = prob - 1;
            weight = 1;
            sample_loss += weight * -logf(prob + kEpsilon);
        } else {
            gradient = prob;
            if (adversarial_temperature > kEpsilon) {
                weight = safe_exp((logit - bias) / adversarial_temperature) / normalizer;
                weight = std::min(weight, 1.0f);
            } else
                weight = float(1.0 / num_negative);
            sample_loss += weight * -logf(1 - prob + kEpsilon);
        }
        kg_backward(model, m.dim, x, opt, margin_or_l3, gradient, relation_lr_multiplier, weight);
    }
    return sample_loss / 2;
}

// -----------------------------------------------------------------------------
// KnowledgeGraphSolver = SolverMixin (core/solver.h) + instance/knowledge_graph.cuh:531-677,
// restated sequentially.
// -----------------------------------------------------------------------------
struct KGSolver {
    KGraph *graph = nullptr;
    int dim = 512;
    int num_worker = 1, num_sampler = 1;
    int num_partition = 1, num_negative = 64, batch_size = 100000, episode_size = 0;
    Optimizer optimizer;
    std::string model;
    int num_epoch = 0, sample_batch_size = 2000, positive_reuse = 1, log_frequency = 100;
    float relation_lr_multiplier = 1, margin = 12, l3_regularization = 2e-3f, adversarial_temperature = 2;
    bool resume = false, shuffle_partition = false;
    // Switches of the MULTI-worker emulation (the reference is nondeterministic there; with one worker they
    // change nothing the product is compared on):
    //  shuffle_override >= 0  forces shuffle_partition (the product never rotates tail partitions with several
    //                         workers: the rotated steps make two reference workers train copies of one partition)
    //  synchronous_relation   all workers write their relation deltas back before any worker loads the matrix
    //                         for the next step -- one legal interleaving of the reference's concurrent
    //                         "write back, then load" per worker (core/solver.h:1436-1500), and the one an
    //                         all-reduce of the deltas implements
    int shuffle_override = -1;
    bool synchronous_relation = false;
    int batch_id = 0, num_batch = 0, pool_id = 0, assignment_offset = 0;

    std::vector<unsigned long long> sampler_seeds, worker_seeds;
    std::vector<std::unique_ptr<RandomStream>> sampler_streams, worker_streams;
    std::vector<std::vector<double>> sampler_random;

    std::vector<std::vector<Index>> partitions;    // head_partitions == tail_partitions (same weights)
    std::vector<std::pair<int, Index>> locations;
    AliasTable<size_t> edge_table;
    // sample_pools[2][P][P], each episode_size * batch_size triples stored {relation, tail, head}
    std::vector<std::vector<std::vector<std::vector<Index>>>> sample_pools;

    std::vector<float> entity_embeddings, relation_embeddings;
    std::vector<float> entity_m1, entity_m2, relation_m1, relation_m2;  // solver-side (host) moments

    // per worker: the relation copy of the block being trained, its values at load time, and the
    // worker-resident relation moments (loaded once after build(), never written back:
    // core/solver.h:1378-1385,1422-1427)
    // reference_cache: a worker's device copies of its head / tail entity partitions with the hit / miss protocol
    // of WorkerMixin::load_partition (core/solver.h:1436-1500).  The default (off) trains in place on the global
    // matrix, which is what the protocol intends; with the cache ON the restatement also reproduces its one
    // incoherent case -- after a block (q, p) a block (p, p) is a "tail hit": the trained tail copy of p stays, the
    // head copy of the SAME partition p is loaded from host memory that has not seen the tail's updates, both
    // copies are trained and the later write-back overwrites the earlier.
    struct Block {
        std::vector<float> v, m1, m2;
    };
    struct Worker {
        std::vector<float> relation, relation_loaded, relation_m1, relation_m2;
        bool has_block = false, moments_loaded = false;
        int head_id = -1, tail_id = -1;
        std::shared_ptr<Block> entity[2];  // [0] head, [1] tail; the same object when shared
    };
    bool reference_cache = false;
    std::vector<Worker> workers;

    std::vector<float> last_loss, logged_loss;
    // each worker's loss buffer lives as long as the worker and is never cleared between blocks
    // (core/solver.h:1326,1541-1549): the loss logged at a block's first batch is the previous block's last
    std::vector<std::vector<float>> worker_loss;
    std::vector<Index> last_negative_batch;
    int last_negative_count = 0;  // head + tail partition size of the last trained block

    // core/solver.h:184-213
    KGSolver(int _dim, int _num_worker, int _num_sampler_per_worker) : dim(_dim), num_worker(_num_worker) {
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

    // core/solver.h:287-466 with protocols {head | in place, tail | in place | shared, global}
    // (instance/knowledge_graph.cuh:553-555): tied weights, minimum 2W partitions when W > 1
    // (:269-276).  GPU memory budgeting is not restated: num_partition = auto means the minimum.
    void build(KGraph *_graph, const Optimizer &_optimizer, int _num_partition, int _num_negative, int _batch_size,
               int _episode_size) {
        graph = _graph;
        optimizer = _optimizer;
        num_partition = _num_partition;
        num_negative = _num_negative;
        batch_size = _batch_size;
        batch_id = 0;
        int min_partition = num_worker == 1 ? 1 : num_worker * 2;
        if (num_partition == 0)
            num_partition = min_partition;
        if (num_partition < min_partition)
            fail("#partition should be no less than " + std::to_string(min_partition));
        shuffle_partition = optimizer.num_moment() > 0;  // :385
        if (shuffle_override >= 0)
            shuffle_partition = shuffle_override != 0;
        assignment_offset = 0;
        partitions = partition(graph->vertex_weights, num_partition);
        locations.resize(graph->num_vertex);
        for (int i = 0; i < num_partition; i++)
            for (Index j = 0; j < partitions[i].size(); j++)
                locations[partitions[i][j]] = {i, j};
        // :426-436 (a global matrix is present)
        int expected_size = _episode_size;
        if (expected_size == 0) {
            expected_size = float(graph->num_vertex * kSamplePerVertexWithGlobal) / num_partition / batch_size;
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
                    block.assign(size_t(episode_size) * batch_size * 3, 0);
            }
        }
        int nm = optimizer.num_moment();
        entity_embeddings.assign(size_t(graph->num_vertex) * dim, 0);
        relation_embeddings.assign(size_t(graph->num_relation) * dim, 0);
        entity_m1.assign(nm >= 1 ? entity_embeddings.size() : 0, 0);
        entity_m2.assign(nm >= 2 ? entity_embeddings.size() : 0, 0);
        relation_m1.assign(nm >= 1 ? relation_embeddings.size() : 0, 0);
        relation_m2.assign(nm >= 2 ? relation_embeddings.size() : 0, 0);
        workers.assign(num_worker, Worker());  // Worker::build(): fresh memories, cold cache (:1281-1330)
        worker_loss.clear();
        sampler_random.resize(num_sampler);
        pool_id = 0;
    }

    void refill(int sampler_id) {
        sampler_random[sampler_id].resize(kRandBatchSize);
        sampler_streams[sampler_id]->generate(sampler_random[sampler_id].data(), kRandBatchSize);
    }

    // SamplerMixin::sample, core/solver.h:1011-1055, with the relation as attribute
    // (instance/knowledge_graph.cuh:300-302); tuple members are stored reversed: {relation, tail, head}
    void sample_edges(int sampler_id, int start, int end) {
        refill(sampler_id);
        const std::vector<double> &random = sampler_random[sampler_id];
        auto &sample_pool = sample_pools[pool_id ^ 1];
        std::vector<std::vector<int>> offsets(num_partition, std::vector<int>(num_partition, start));
        int num_complete = 0, rand_id = 0;
        std::vector<std::pair<int, Index>> heads(sample_batch_size), tails(sample_batch_size);
        std::vector<Index> attributes(sample_batch_size);
        while (num_complete < num_partition * num_partition) {
            for (int i = 0; i < sample_batch_size; i++) {
                if (rand_id > kRandBatchSize - 2) {
                    refill(sampler_id);
                    rand_id = 0;
                }
                double rand2 = random[rand_id++];  // gcc evaluates the two arguments right to left
                double rand1 = random[rand_id++];
                size_t edge_id = edge_table.sample(rand1, rand2);
                heads[i] = locations[graph->edge_h[edge_id]];
                tails[i] = locations[graph->edge_t[edge_id]];
                attributes[i] = graph->edge_r[edge_id];
            }
            for (int i = 0; i < sample_batch_size; i++) {
                int &offset = offsets[heads[i].first][tails[i].first];
                if (offset < end) {
                    std::vector<Index> &block = sample_pool[heads[i].first][tails[i].first];
                    block[size_t(offset) * 3] = attributes[i];
                    block[size_t(offset) * 3 + 1] = tails[i].second;
                    block[size_t(offset) * 3 + 2] = heads[i].second;
                    if (++offset == end)
                        num_complete++;
                }
            }
        }
    }

    void fill_pool() {
        int num_sample = episode_size * batch_size;
        int work_load = (num_sample + num_sampler - 1) / num_sampler;
        for (int i = 0; i < num_sampler; i++)
            sample_edges(i, work_load * i, std::min(work_load * (i + 1), num_sample));
    }

    // instance/knowledge_graph.cuh:567-627
    void init_embeddings() {
        static const float kPi = atan(1) * 4;
        const size_t d = dim;
        int id = model_id(model);
        if (id == kTransE) {
            std::uniform_real_distribution<float> init(-margin / d, margin / d);
            for (auto &x : entity_embeddings)
                x = init(global_engine());
            for (auto &x : relation_embeddings)
                x = init(global_engine());
        }
        if (id == kDistMult || id == kComplEx || id == kSimplE) {
            std::uniform_real_distribution<float> init(-0.5, 0.5);
            for (auto &x : entity_embeddings)
                x = init(global_engine());
            for (auto &x : relation_embeddings)
                x = init(global_engine());
        }
        if (id == kQuatE) {  // :604-626
            std::uniform_real_distribution<float> init_modulus(-1 / sqrt(d / 2), 1 / sqrt(d / 2));
            std::uniform_real_distribution<float> init_phase(-kPi, kPi);
            std::uniform_real_distribution<float> init(0, 1);
            for (auto *matrix : {&entity_embeddings, &relation_embeddings})
                for (size_t row = 0; row < matrix->size() / d; row++)
                    for (size_t i = 0; i < d / 4; i++) {
                        float modulus = init_modulus(global_engine());
                        float phase = init_phase(global_engine());
                        float v_i = init(global_engine());
                        float v_j = init(global_engine());
                        float v_k = init(global_engine());
                        float norm = sqrtf(v_i * v_i + v_j * v_j + v_k * v_k);
                        v_i /= norm + kEpsilon;
                        v_j /= norm + kEpsilon;
                        v_k /= norm + kEpsilon;
                        float *e = matrix->data() + row * d + i * 4;
                        e[0] = modulus * cosf(phase);
                        e[1] = modulus * v_i * sinf(phase);
                        e[2] = modulus * v_j * sinf(phase);
                        e[3] = modulus * v_k * sinf(phase);
                    }
        }
        if (id == kRotatE) {
            std::uniform_real_distribution<float> init(-margin * 2 / d, margin * 2 / d);
            std::uniform_real_distribution<float> init_phase(-kPi, kPi);
            for (auto &x : entity_embeddings)
                x = init(global_engine());
            for (Index r = 0; r < graph->num_relation; r++)
                for (size_t i = 0; i < d / 2; i++)
                    relation_embeddings[r * d + i] = init_phase(global_engine());
        }
    }

    // KnowledgeGraphSolver::train, instance/knowledge_graph.cuh:666-677, then SolverMixin::train up to
    // and including the first pool fill, core/solver.h:588-628
    void train_begin(const std::string &_model, int _num_epoch, bool _resume, float _relation_lr_multiplier,
                     float _margin, float _l3_regularization, int _sample_batch_size, int _positive_reuse,
                     float _adversarial_temperature, int _log_frequency) {
        relation_lr_multiplier = _relation_lr_multiplier;
        margin = _margin;
        l3_regularization = _l3_regularization;
        adversarial_temperature = _adversarial_temperature;
        if (!graph)
            fail("The model must be built on a graph first");
        model = _model;
        if (model_id(model) < 0)
            fail("Invalid model `" + model + "`");
        if ((model == "ComplEx" || model == "SimplE" || model == "RotatE") && dim % 2)
            fail("Model `" + model + "` needs an even dimension");
        if (model == "QuatE" && dim % 4)
            fail("Model `QuatE` needs a dimension divisible by 4");
        num_epoch = _num_epoch;
        resume = _resume;
        sample_batch_size = _sample_batch_size;
        positive_reuse = _positive_reuse;
        log_frequency = _log_frequency;
        if (!resume) {
            init_embeddings();
            // init_moments, core/solver.h:247-256: the shared (tail) matrix is skipped, all others zeroed
            for (auto *m : {&entity_m1, &entity_m2, &relation_m1, &relation_m2})
                std::fill(m->begin(), m->end(), 0.0f);
            batch_id = 0;
        }
        num_batch = batch_id + size_t(num_epoch) * graph->num_edge / batch_size;
        graph->flatten();
        edge_table.build(graph->edge_weights);
        fill_pool();
    }

    // get_schedule, core/solver.h:519-561 (tied weights)
    std::vector<std::vector<std::pair<int, int>>> get_schedule() const {
        std::vector<std::vector<std::pair<int, int>>> schedule;
        std::vector<std::pair<int, int>> assignment(num_worker);
        if (num_partition == 1)
            return {{{0, 0}}};
        for (int x = 0; x < num_partition; x += num_worker * 2)
            for (int y = 0; y < num_partition; y += num_worker * 2) {
                for (int i = 0; i < num_worker; i++)
                    assignment[i] = {x + i, y + i};
                schedule.push_back(assignment);
                for (int i = 0; i < num_worker; i++)
                    assignment[i] = {x + num_worker + i, y + num_worker + i};
                schedule.push_back(assignment);
                for (int group_size = 1; group_size <= num_worker; group_size *= 2)
                    for (int offset = 0; offset < group_size; offset++) {
                        for (int i = 0; i < num_worker; i++) {
                            int head_partition_id = x + (i / group_size * 2) * group_size + i % group_size;
                            int tail_partition_id = y + (i / group_size * 2 + 1) * group_size + (i + offset) % group_size;
                            assignment[i] = {head_partition_id, tail_partition_id};
                        }
                        schedule.push_back(assignment);
                        for (int i = 0; i < num_worker; i++)
                            std::swap(assignment[i].first, assignment[i].second);
                        schedule.push_back(assignment);
                    }
            }
        return schedule;
    }

    // write_embedding of the global relation matrix, core/solver.h:1413-1420:
    // global -= (values at load time - trained values)
    void write_relation(Worker &w) {
        if (!w.has_block)
            return;
        for (size_t i = 0; i < relation_embeddings.size(); i++)
            relation_embeddings[i] -= w.relation_loaded[i] - w.relation[i];
        w.has_block = false;
    }

    static Index device_sample(const AliasTable<Index> &table, double random1, double random2) {
        float rand1 = float(random1), rand2 = float(random2);
        return table.sample(double(rand1), double(rand2));
    }

    // WorkerMixin::train for one block (core/solver.h:1511-1557) around the KG kernels
    // write_embedding / load_embedding for the entity partitions (in place protocol: scatter / gather by the
    // partition's global ids), core/solver.h:1349-1428
    void write_entity(Worker &w, int i) {
        if (i > 0 && w.entity[i] == w.entity[i - 1])
            return;
        const std::vector<Index> &ids = partitions[i == 0 ? w.head_id : w.tail_id];
        Block &b = *w.entity[i];
        const size_t d = dim;
        for (size_t r = 0; r < ids.size(); r++) {
            std::copy(b.v.begin() + r * d, b.v.begin() + (r + 1) * d, entity_embeddings.begin() + size_t(ids[r]) * d);
            if (!b.m1.empty())
                std::copy(b.m1.begin() + r * d, b.m1.begin() + (r + 1) * d, entity_m1.begin() + size_t(ids[r]) * d);
            if (!b.m2.empty())
                std::copy(b.m2.begin() + r * d, b.m2.begin() + (r + 1) * d, entity_m2.begin() + size_t(ids[r]) * d);
        }
    }
    void load_entity(Worker &w, int i) {
        if (i == 1 && w.head_id == w.tail_id) {
            w.entity[1] = w.entity[0];
            return;
        }
        if (!w.entity[i] || (i > 0 && w.entity[i] == w.entity[i - 1]))
            w.entity[i] = std::make_shared<Block>();
        const std::vector<Index> &ids = partitions[i == 0 ? w.head_id : w.tail_id];
        Block &b = *w.entity[i];
        const size_t d = dim;
        const int nm = optimizer.num_moment();
        b.v.resize(ids.size() * d);
        b.m1.resize(nm >= 1 ? ids.size() * d : 0);
        b.m2.resize(nm >= 2 ? ids.size() * d : 0);
        for (size_t r = 0; r < ids.size(); r++) {
            std::copy(entity_embeddings.begin() + size_t(ids[r]) * d, entity_embeddings.begin() + size_t(ids[r] + 1) * d,
                      b.v.begin() + r * d);
            if (nm >= 1)
                std::copy(entity_m1.begin() + size_t(ids[r]) * d, entity_m1.begin() + size_t(ids[r] + 1) * d,
                          b.m1.begin() + r * d);
            if (nm >= 2)
                std::copy(entity_m2.begin() + size_t(ids[r]) * d, entity_m2.begin() + size_t(ids[r] + 1) * d,
                          b.m2.begin() + r * d);
        }
    }
    // the entity part of WorkerMixin::load_partition, core/solver.h:1436-1500
    void load_entity_partitions(Worker &w, int head_partition, int tail_partition) {
        const bool cold = w.head_id == -1 || w.tail_id == -1;
        bool hit[2] = {false, false};
        if (!cold) {
            hit[0] = w.head_id == head_partition;
            if (!hit[0] && w.head_id == tail_partition && w.tail_id == head_partition) {  // swap hit
                std::swap(w.entity[0], w.entity[1]);
                hit[0] = hit[1] = true;
            }
            if (!(w.entity[1] == w.entity[0] && !hit[0]))
                hit[1] = hit[1] || w.tail_id == tail_partition;
            for (int i = 0; i < 2; i++)
                if (!hit[i])
                    write_entity(w, i);
        }
        w.head_id = head_partition;
        w.tail_id = tail_partition;
        for (int i = 0; i < 2; i++)
            if (!hit[i])
                load_entity(w, i);
    }

    void train_block(int worker_id, int head_partition, int tail_partition, int first_batch_id, int batch_stride) {
        Worker &w = workers[worker_id];
        // load_partition, core/solver.h:1436-1500: write the previous block's relation delta back, load
        // the relation matrix (never a cache hit), keep the worker's relation moments
        write_relation(w);
        w.relation = relation_embeddings;
        w.relation_loaded = relation_embeddings;
        w.has_block = true;
        int nm = optimizer.num_moment();
        if (!w.moments_loaded) {
            w.relation_m1 = relation_m1;
            w.relation_m2 = relation_m2;
            w.moments_loaded = true;
        }
        const std::vector<Index> &head_ids = partitions[head_partition], &tail_ids = partitions[tail_partition];
        // build_negative_sampler, instance/knowledge_graph.cuh:316-319: uniform over head + tail rows
        AliasTable<Index> negative_sampler;
        negative_sampler.build(std::vector<float>(head_ids.size() + tail_ids.size(), 1));
        last_negative_count = int(head_ids.size() + tail_ids.size());

        KGMatrices m;
        m.dim = dim;
        m.num_head = head_ids.size();
        m.head = m.tail = entity_embeddings.data();
        m.head_rows = head_ids.data();
        m.tail_rows = tail_ids.data();
        m.relation = w.relation.data();
        if (nm >= 1) {
            m.head_m1 = m.tail_m1 = entity_m1.data();
            m.relation_m1 = w.relation_m1.data();
        }
        if (nm >= 2) {
            m.head_m2 = m.tail_m2 = entity_m2.data();
            m.relation_m2 = w.relation_m2.data();
        }
        if (reference_cache) {  // train the worker's copies instead of the global matrix
            load_entity_partitions(w, head_partition, tail_partition);
            Block &h = *w.entity[0], &t = *w.entity[1];
            m.head_rows = m.tail_rows = nullptr;
            m.head = h.v.data();
            m.tail = t.v.data();
            m.head_m1 = nm >= 1 ? h.m1.data() : nullptr;
            m.tail_m1 = nm >= 1 ? t.m1.data() : nullptr;
            m.head_m2 = nm >= 2 ? h.m2.data() : nullptr;
            m.tail_m2 = nm >= 2 ? t.m2.data() : nullptr;
        }
        int id = model_id(model);
        float margin_or_l3 = (id == kTransE || id == kRotatE) ? margin : l3_regularization;

        const std::vector<Index> &samples = sample_pools[pool_id][head_partition][tail_partition];
        std::vector<double> random(size_t(batch_size) * num_negative * 2);
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
                const Index *batch = &samples[size_t(j) * batch_size * 3];
                worker_streams[worker_id]->generate(random.data(), random.size());
                for (size_t t = 0; t < last_negative_batch.size(); t++)
                    last_negative_batch[t] = device_sample(negative_sampler, random[t * 2], random[t * 2 + 1]);
                if (this_batch % log_frequency == 0) {  // the loss buffer left by the previous batch
                    float batch_loss = 0;
                    for (int i = 0; i < batch_size; i++)
                        batch_loss += last_loss[i];
                    logged_loss.push_back(batch_loss / batch_size);
                }
                opt.apply_schedule(this_batch, num_batch);
                for (int i = 0; i < batch_size; i++)
                    last_loss[i] = kg_train_sample(id, m, opt, batch + size_t(i) * 3,
                                                   &last_negative_batch[size_t(i) * num_negative], num_negative,
                                                   relation_lr_multiplier, margin_or_l3, adversarial_temperature);
            }
        worker_loss[worker_id] = last_loss;
    }

    // one pass of the episode loop, core/solver.h:629-649
    bool train_episode() {
        if (batch_id >= num_batch)
            return false;
        pool_id ^= 1;
        if (shuffle_partition)
            assignment_offset = (assignment_offset + 1) % num_partition;
        auto schedule = get_schedule();
        int per_block = positive_reuse * episode_size;
        for (auto &assignment : schedule) {
            if (synchronous_relation)
                for (auto &w : workers)
                    write_relation(w);
            for (int i = 0; i < (int)assignment.size(); i++)
                train_block(i, assignment[i].first, (assignment[i].second + assignment_offset) % num_partition,
                            batch_id + i, (int)assignment.size());
            batch_id += per_block * (int)assignment.size();
        }
        fill_pool();
        if (batch_id >= num_batch)  // Worker::write_back of every worker, core/solver.h:650-653
            for (auto &w : workers) {
                if (reference_cache && w.head_id != -1 && w.tail_id != -1)
                    for (int i = 0; i < 2; i++)
                        write_entity(w, i);
                write_relation(w);
            }
        return true;
    }

    // SolverMixin::predict + gpu::knowledge_graph::predict (core/solver.h:729-802,
    // instance/gpu/knowledge_graph.cuh:341-366); rows are (head, tail, relation) global ids
    void predict(const Index *triplets, size_t n, float *out) const {
        int id = model_id(model);
        for (size_t i = 0; i < n; i++)
            out[i] = kg_forward(id, dim, &entity_embeddings[size_t(triplets[i * 3]) * dim],
                                &entity_embeddings[size_t(triplets[i * 3 + 1]) * dim],
                                &relation_embeddings[size_t(triplets[i * 3 + 2]) * dim], margin);
    }
};

}  // namespace oracle

// =============================================================================
// C API (ctypes, tests/oracle_lib.py)
// =============================================================================
using namespace oracle;

extern "C" {
const char *og_last_error();
}
void og_set_error(const std::string &message);

#define KG_TRY try {
#define KG_CATCH(ret)                 \
    }                                 \
    catch (const std::exception &e) { \
        og_set_error(e.what());       \
        return ret;                   \
    }

extern "C" {

void *og_kg_graph_load(const char *file_name, int normalization) {
    KG_TRY
    KGraph *g = new KGraph();
    try {
        g->load_file(file_name, normalization != 0);
    } catch (...) {
        delete g;
        throw;
    }
    return g;
    KG_CATCH(nullptr)
}

void *og_kg_graph_from_triplets(const char *const *h, const char *const *r, const char *const *t, const float *w,
                                uint64_t n, int normalization) {
    KGraph *g = new KGraph();
    g->normalization = normalization != 0;
    for (uint64_t i = 0; i < n; i++)
        g->add_edge(h[i], r[i], t[i], w ? w[i] : 1);
    if (g->normalization)
        g->normalize();
    return g;
}

void og_kg_graph_free(void *g) {
    delete (KGraph *)g;
}

void og_kg_graph_sizes(void *g, uint64_t *out) {
    KGraph *G = (KGraph *)g;
    out[0] = G->num_vertex;
    out[1] = G->num_edge;
    out[2] = G->num_relation;
}

void og_kg_graph_flat(void *g, uint32_t *h, uint32_t *t, uint32_t *r, float *w, float *vertex_weights) {
    KGraph *G = (KGraph *)g;
    G->flatten();
    memcpy(h, G->edge_h.data(), G->edge_h.size() * sizeof(Index));
    memcpy(t, G->edge_t.data(), G->edge_t.size() * sizeof(Index));
    memcpy(r, G->edge_r.data(), G->edge_r.size() * sizeof(Index));
    memcpy(w, G->edge_weights.data(), G->edge_weights.size() * sizeof(float));
    memcpy(vertex_weights, G->vertex_weights.data(), G->num_vertex * sizeof(float));
}

const char *og_kg_graph_entity(void *g, uint64_t id) {
    return ((KGraph *)g)->id2entity[id].c_str();
}

const char *og_kg_graph_relation(void *g, uint64_t id) {
    return ((KGraph *)g)->id2relation[id].c_str();
}

float og_kg_forward(const char *model, int dim, const float *head, const float *tail, const float *relation,
                    float margin_or_l3) {
    return kg_forward(model_id(model), dim, head, tail, relation, margin_or_l3);
}

// the kernels on caller-provided matrices; batch = [n][3] {relation, tail, head}, negatives = [n][k];
// tail may equal head (one shared entity matrix); moments may be null.  Sequential.
int og_kg_train_batch(const char *model, int dim, uint32_t num_head, float *head, float *tail, float *relation,
                      float *head_m1, float *tail_m1, float *relation_m1, float *head_m2, float *tail_m2,
                      float *relation_m2, const uint32_t *batch, const uint32_t *negatives, uint64_t n,
                      int num_negative, int opt_type, float lr, float weight_decay, float a, float b, float epsilon,
                      float relation_lr_multiplier, float margin_or_l3, float adversarial_temperature, float *loss) {
    KG_TRY
    int id = model_id(model);
    if (id < 0)
        fail(std::string("Invalid model `") + model + "`");
    KGMatrices m;
    m.dim = dim;
    m.num_head = num_head;
    m.head = head;
    m.tail = tail;
    m.relation = relation;
    m.head_m1 = head_m1;
    m.tail_m1 = tail_m1;
    m.relation_m1 = relation_m1;
    m.head_m2 = head_m2;
    m.tail_m2 = tail_m2;
    m.relation_m2 = relation_m2;
    Optimizer opt;
    opt.type = opt_type;
    opt.lr = opt.init_lr = lr;
    opt.weight_decay = weight_decay;
    opt.a = a;
    opt.b = b;
    opt.epsilon = epsilon;
    for (uint64_t i = 0; i < n; i++) {
        float l = kg_train_sample(id, m, opt, batch + i * 3, negatives + i * num_negative, num_negative,
                                  relation_lr_multiplier, margin_or_l3, adversarial_temperature);
        if (loss)
            loss[i] = l;
    }
    return 0;
    KG_CATCH(-1)
}

void *og_kg_solver_create(int dim, int num_worker, int num_sampler_per_worker) {
    KG_TRY
    return new KGSolver(dim, num_worker, num_sampler_per_worker);
    KG_CATCH(nullptr)
}

void og_kg_solver_free(void *s) {
    delete (KGSolver *)s;
}

int og_kg_solver_build(void *s, void *graph, int opt_type, int schedule, float lr, float weight_decay, float a,
                       float b, float epsilon, int num_partition, int num_negative, int batch_size,
                       int episode_size) {
    KG_TRY
    Optimizer opt;
    opt.type = opt_type;
    opt.schedule = schedule;
    opt.lr = opt.init_lr = lr;
    opt.weight_decay = weight_decay;
    opt.a = a;
    opt.b = b;
    opt.epsilon = epsilon;
    ((KGSolver *)s)->build((KGraph *)graph, opt, num_partition, num_negative, batch_size, episode_size);
    return 0;
    KG_CATCH(-1)
}

// see KGSolver::shuffle_override / synchronous_relation; call before build()
void og_kg_solver_set_emulation(void *s, int shuffle_override, int synchronous_relation) {
    ((KGSolver *)s)->shuffle_override = shuffle_override;
    ((KGSolver *)s)->synchronous_relation = synchronous_relation != 0;
}

// see KGSolver::reference_cache; call before build()
void og_kg_solver_set_reference_cache(void *s, int on) {
    ((KGSolver *)s)->reference_cache = on != 0;
}

int og_kg_solver_train_begin(void *s, const char *model, int num_epoch, int resume, float relation_lr_multiplier,
                             float margin, float l3_regularization, int sample_batch_size, int positive_reuse,
                             float adversarial_temperature, int log_frequency) {
    KG_TRY
    ((KGSolver *)s)->train_begin(model, num_epoch, resume != 0, relation_lr_multiplier, margin, l3_regularization,
                                 sample_batch_size, positive_reuse, adversarial_temperature, log_frequency);
    return 0;
    KG_CATCH(-1)
}

int og_kg_solver_train_episode(void *s) {
    KG_TRY
    return ((KGSolver *)s)->train_episode() ? 1 : 0;
    KG_CATCH(-1)
}

// {num_partition, episode_size, batch_size, num_batch, batch_id, pool_id, num_sampler, assignment_offset,
//  last_negative_count, shuffle_partition}
int og_kg_solver_info(void *s, int *out) {
    KGSolver *S = (KGSolver *)s;
    out[0] = S->num_partition;
    out[1] = S->episode_size;
    out[2] = S->batch_size;
    out[3] = S->num_batch;
    out[4] = S->batch_id;
    out[5] = S->pool_id;
    out[6] = S->num_sampler;
    out[7] = S->assignment_offset;
    out[8] = S->last_negative_count;
    out[9] = S->shuffle_partition;
    return 0;
}

const uint32_t *og_kg_solver_pool(void *s, int pool, int head_partition, int tail_partition) {
    return ((KGSolver *)s)->sample_pools[pool][head_partition][tail_partition].data();
}

void og_kg_solver_locations(void *s, int *part_of, uint32_t *local_of) {
    KGSolver *S = (KGSolver *)s;
    for (size_t v = 0; v < S->locations.size(); v++) {
        part_of[v] = S->locations[v].first;
        local_of[v] = S->locations[v].second;
    }
}

// which: 0 entity, 1 relation; order: 0 embeddings, 1 / 2 solver-side moments
float *og_kg_solver_matrix(void *s, int which, int order) {
    KGSolver *S = (KGSolver *)s;
    if (which == 0)
        return order == 0 ? S->entity_embeddings.data() : (order == 1 ? S->entity_m1.data() : S->entity_m2.data());
    return order == 0 ? S->relation_embeddings.data() : (order == 1 ? S->relation_m1.data() : S->relation_m2.data());
}

int64_t og_kg_solver_last_negatives(void *s, uint32_t *out) {
    KGSolver *S = (KGSolver *)s;
    if (out)
        memcpy(out, S->last_negative_batch.data(), S->last_negative_batch.size() * sizeof(Index));
    return int64_t(S->last_negative_batch.size());
}

int og_kg_solver_last_loss(void *s, float *out) {
    KGSolver *S = (KGSolver *)s;
    if (out)
        memcpy(out, S->last_loss.data(), S->last_loss.size() * sizeof(float));
    return int(S->last_loss.size());
}

int og_kg_solver_logged_loss(void *s, float *out, int capacity) {
    KGSolver *S = (KGSolver *)s;
    int n = std::min<int>(capacity, S->logged_loss.size());
    if (out)
        memcpy(out, S->logged_loss.data(), n * sizeof(float));
    return int(S->logged_loss.size());
}

int og_kg_solver_schedule(void *s, int *out, int capacity) {
    KG_TRY
    auto schedule = ((KGSolver *)s)->get_schedule();
    int width = schedule.empty() ? 0 : int(schedule[0].size());
    if (int(schedule.size()) * width * 2 > capacity)
        fail("og_kg_solver_schedule: capacity too small");
    for (size_t i = 0; i < schedule.size(); i++)
        for (int j = 0; j < width; j++) {
            *out++ = schedule[i][j].first;
            *out++ = schedule[i][j].second;
        }
    return int(schedule.size());
    KG_CATCH(-1)
}

int og_kg_solver_predict(void *s, const uint32_t *triplets, uint64_t n, float *out) {
    KG_TRY
    ((KGSolver *)s)->predict(triplets, n, out);
    return 0;
    KG_CATCH(-1)
}

}  // extern "C"
