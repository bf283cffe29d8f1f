// =============================================================================
// libgraphvite -- the reference's pybind11 module (src/graphvite.cu:28-105 over include/bind.h), re-implemented as a
// thin layer over the C ABI of libgv_b200 (include/gv_b200.h).  Same module name, same submodules, same class names
// (`graph.Graph_j`, `solver.GraphSolver_128_f_j`, `optimizer.SGD` ...; the suffixes are the Itanium typeid names the
// reference derives, bind.h:71-88), same constructor / method signatures, defaults and attribute names, so that the
// reference's unchanged Python package (python/graphvite/helper.py:83-105 assembles `graphvite.solver.GraphSolver`
// from these names) and its config/*.yaml files drive the B200 path.  No CUDA, no torch: this file sees only the C ABI.
//
// Differences a caller can observe: errors are Python exceptions (RuntimeError) instead of
// abort(); `num_sampler_per_worker = auto` means one sampler stream per GPU; a solver drives ONE GPU per process
// (several device ids: launch one process per GPU, graphvite_b200.solver).
// =============================================================================
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "gv_b200.h"

namespace py = pybind11;
using no_gil = py::call_guard<py::gil_scoped_release>;  // bind.h:44

static const int kAuto = 0;  // util/common.h: kAuto

static void check(int status) {
    if (status != 0)
        throw std::runtime_error(gv_last_error());
}

// ---- optimizers (bind.h:757-999 over core/optimizer.h:42-319) ---------------------------------------------
struct LRSchedule {
    std::string type = "constant";
    std::function<float(int, int)> schedule_function;
    LRSchedule() = default;
    explicit LRSchedule(const std::string &_type) : type(_type) {
        if (type != "constant" && type != "linear")
            throw std::invalid_argument("Invalid schedule `" + type + "`");
    }
    explicit LRSchedule(std::function<float(int, int)> function) : type("custom"), schedule_function(function) {}
    std::string info() const { return "schedule: " + type; }
};

struct Optimizer {
    std::string type = "Default";
    int type_id = -1;  // GV_OPT_*; -1 = the solver's default (core/solver.h:291-296)
    float lr = 1e-4f, weight_decay = 0;
    float a = 0, b = 0, eps = 0;
    LRSchedule schedule{"linear"};
    Optimizer() = default;
    explicit Optimizer(int _type) {
        if (_type != kAuto)
            throw std::invalid_argument("Optimizer(type): only `auto` selects the default optimizer");
        lr = 0;  // "Default": the solver picks its own learning rate too
    }
    explicit Optimizer(float _lr) : lr(_lr) {}
    virtual ~Optimizer() = default;

    static float trampoline(int batch_id, int num_batch, void *self) {
        py::gil_scoped_acquire gil;  // train() runs without the GIL; the schedule is a Python callable
        return static_cast<const Optimizer *>(self)->schedule.schedule_function(batch_id, num_batch);
    }
    gv_optimizer_t descriptor() const {
        gv_optimizer_t d{};
        d.type = type_id;
        d.lr = lr;
        d.weight_decay = weight_decay;
        d.a = a;
        d.b = b;
        d.epsilon = eps;
        d.schedule = schedule.type == "constant" ? GV_SCHEDULE_CONSTANT
                                                 : (schedule.type == "linear" ? GV_SCHEDULE_LINEAR : GV_SCHEDULE_CUSTOM);
        d.schedule_fn = d.schedule == GV_SCHEDULE_CUSTOM ? &Optimizer::trampoline : nullptr;
        d.schedule_ctx = const_cast<Optimizer *>(this);
        return d;
    }
    std::string info() const {
        std::stringstream ss;
        ss << "<optimizer " << type << "; learning rate " << lr << "; lr schedule " << schedule.type
           << "; weight decay " << weight_decay << ">";
        return ss.str();
    }
};

struct SGD : Optimizer {
    SGD(float _lr, float _wd, const LRSchedule &_schedule) {
        type = "SGD", type_id = GV_OPT_SGD, lr = _lr, weight_decay = _wd, schedule = _schedule;
    }
};
struct Momentum : Optimizer {
    float momentum;
    Momentum(float _lr, float _wd, float _momentum, const LRSchedule &_schedule) : momentum(_momentum) {
        type = "Momentum", type_id = GV_OPT_MOMENTUM, lr = _lr, weight_decay = _wd, a = _momentum, schedule = _schedule;
    }
};
struct AdaGrad : Optimizer {
    float epsilon;
    AdaGrad(float _lr, float _wd, float _epsilon, const LRSchedule &_schedule) : epsilon(_epsilon) {
        type = "AdaGrad", type_id = GV_OPT_ADAGRAD, lr = _lr, weight_decay = _wd, eps = _epsilon, schedule = _schedule;
    }
};
struct RMSprop : Optimizer {
    float alpha, epsilon;
    RMSprop(float _lr, float _wd, float _alpha, float _epsilon, const LRSchedule &_schedule)
        : alpha(_alpha), epsilon(_epsilon) {
        type = "RMSprop", type_id = GV_OPT_RMSPROP, lr = _lr, weight_decay = _wd, a = _alpha, eps = _epsilon;
        schedule = _schedule;
    }
};
struct Adam : Optimizer {
    float beta1, beta2, epsilon;
    Adam(float _lr, float _wd, float _beta1, float _beta2, float _epsilon, const LRSchedule &_schedule)
        : beta1(_beta1), beta2(_beta2), epsilon(_epsilon) {
        type = "Adam", type_id = GV_OPT_ADAM, lr = _lr, weight_decay = _wd, a = _beta1, b = _beta2, eps = _epsilon;
        schedule = _schedule;
    }
};

// ---- graphs (bind.h:109-314) --------------------------------------------------------------------------------
static std::vector<const char *> c_strings(const std::vector<std::string> &names) {
    std::vector<const char *> out(names.size());
    for (size_t i = 0; i < names.size(); i++)
        out[i] = names[i].c_str();
    return out;
}

struct Graph {
    gv_graph_t *handle = gv_graph_create();
    Graph() = default;
    Graph(const Graph &) = delete;
    virtual ~Graph() { gv_graph_destroy(handle); }

    void load_file(const char *file_name, bool as_undirected, bool normalization, const char *delimiters,
                   const char *comment) {
        check(gv_graph_load_file(handle, file_name, as_undirected, normalization, delimiters, comment));
    }
    void load_edge_list(const std::vector<std::tuple<std::string, std::string>> &edges, bool as_undirected,
                        bool normalization) {
        std::vector<std::string> u(edges.size()), v(edges.size());
        for (size_t i = 0; i < edges.size(); i++)
            std::tie(u[i], v[i]) = edges[i];
        check(gv_graph_load_edges(handle, c_strings(u).data(), c_strings(v).data(), nullptr, edges.size(),
                                  as_undirected, normalization));
    }
    void load_weighted_edge_list(const std::vector<std::tuple<std::string, std::string, float>> &edges,
                                 bool as_undirected, bool normalization) {
        std::vector<std::string> u(edges.size()), v(edges.size());
        std::vector<float> w(edges.size());
        for (size_t i = 0; i < edges.size(); i++)
            std::tie(u[i], v[i], w[i]) = edges[i];
        check(gv_graph_load_edges(handle, c_strings(u).data(), c_strings(v).data(), w.data(), edges.size(),
                                  as_undirected, normalization));
    }
    void save(const char *file_name, bool weighted, bool anonymous) {
        check(gv_graph_save(handle, file_name, weighted, anonymous));
    }
    size_t num_vertex() const { return gv_graph_num_vertex(handle); }
    size_t num_edge() const { return gv_graph_num_edge(handle); }
    std::vector<std::string> id2name() const {
        std::vector<std::string> names(num_vertex());
        for (size_t i = 0; i < names.size(); i++)
            names[i] = gv_graph_id2name(handle, i);
        return names;
    }
    std::unordered_map<std::string, unsigned> name2id() const {
        std::unordered_map<std::string, unsigned> map;
        const size_t n = num_vertex();
        map.reserve(n);
        for (size_t i = 0; i < n; i++)
            map.emplace(gv_graph_id2name(handle, i), unsigned(i));
        return map;
    }
    std::string info() const {
        char text[4096];
        gv_graph_info(handle, text, sizeof(text));
        return text;
    }
};

struct WordGraph : Graph {  // bind.h:190-234 over instance/word_graph.cuh:42-166
    void load_corpus(const char *file_name, int window, int min_count, bool normalization, const char *delimiters,
                     const char *comment) {
        check(gv_graph_load_corpus(handle, file_name, window, min_count, normalization, delimiters, comment));
    }
};

struct KnowledgeGraph {
    gv_kgraph_t *handle = gv_kgraph_create();
    KnowledgeGraph() = default;
    KnowledgeGraph(const KnowledgeGraph &) = delete;
    ~KnowledgeGraph() { gv_kgraph_destroy(handle); }
    void load_file(const char *file_name, bool normalization, const char *delimiters, const char *comment) {
        check(gv_kgraph_load_file(handle, file_name, normalization, delimiters, comment));
    }
    void load_triplet_list(const std::vector<std::tuple<std::string, std::string, std::string>> &triplets,
                           bool normalization) {
        std::vector<std::string> h(triplets.size()), r(triplets.size()), t(triplets.size());
        for (size_t i = 0; i < triplets.size(); i++)
            std::tie(h[i], r[i], t[i]) = triplets[i];
        check(gv_kgraph_load_triplets(handle, c_strings(h).data(), c_strings(r).data(), c_strings(t).data(), nullptr,
                                      triplets.size(), normalization));
    }
    void load_weighted_triplet_list(
        const std::vector<std::tuple<std::string, std::string, std::string, float>> &triplets, bool normalization) {
        std::vector<std::string> h(triplets.size()), r(triplets.size()), t(triplets.size());
        std::vector<float> w(triplets.size());
        for (size_t i = 0; i < triplets.size(); i++)
            std::tie(h[i], r[i], t[i], w[i]) = triplets[i];
        check(gv_kgraph_load_triplets(handle, c_strings(h).data(), c_strings(r).data(), c_strings(t).data(), w.data(),
                                      triplets.size(), normalization));
    }
    void save(const char *file_name, bool anonymous) { check(gv_kgraph_save(handle, file_name, anonymous)); }
    size_t num_vertex() const { return gv_kgraph_num_vertex(handle); }
    size_t num_edge() const { return gv_kgraph_num_edge(handle); }
    size_t num_relation() const { return gv_kgraph_num_relation(handle); }
    std::vector<std::string> id2entity() const {
        std::vector<std::string> names(num_vertex());
        for (size_t i = 0; i < names.size(); i++)
            names[i] = gv_kgraph_id2entity(handle, i);
        return names;
    }
    std::vector<std::string> id2relation() const {
        std::vector<std::string> names(num_relation());
        for (size_t i = 0; i < names.size(); i++)
            names[i] = gv_kgraph_id2relation(handle, i);
        return names;
    }
    std::unordered_map<std::string, unsigned> entity2id() const {
        std::unordered_map<std::string, unsigned> map;
        for (size_t i = 0; i < num_vertex(); i++)
            map.emplace(gv_kgraph_id2entity(handle, i), unsigned(i));
        return map;
    }
    std::unordered_map<std::string, unsigned> relation2id() const {
        std::unordered_map<std::string, unsigned> map;
        for (size_t i = 0; i < num_relation(); i++)
            map.emplace(gv_kgraph_id2relation(handle, i), unsigned(i));
        return map;
    }
    std::string info() const {
        char text[4096];
        gv_kgraph_info(handle, text, sizeof(text));
        return text;
    }
};

// ---- solvers (bind.h:383-639) -------------------------------------------------------------------------------
static std::map<std::string, std::string> parse_attributes(const char *text) {
    std::map<std::string, std::string> out;
    std::stringstream ss(text);
    std::string line;
    while (std::getline(ss, line)) {
        const size_t eq = line.find('=');
        if (eq != std::string::npos)
            out[line.substr(0, eq)] = line.substr(eq + 1);
    }
    return out;
}

static py::array_t<float> numpy_view(float *data, uint64_t rows, int dim, py::handle owner) {  // bind.h:90-106
    if (!data || rows == 0)
        return py::array_t<float>(std::vector<py::ssize_t>{0, py::ssize_t(dim)});
    return py::array_t<float>({py::ssize_t(rows), py::ssize_t(dim)}, {py::ssize_t(sizeof(float) * dim),
                                                                      py::ssize_t(sizeof(float))},
                              data, owner);  // `owner` keeps the solver alive while the view exists; no copy
}

template<size_t dim>
struct GraphSolver {
    gv_solver_t *handle;
    std::shared_ptr<Optimizer> optimizer;  // keeps a custom schedule's callable alive across train()
    py::object graph;                      // the solver borrows the graph (core/solver.h:289)
    GraphSolver(std::vector<int> device_ids, int num_sampler_per_worker, size_t gpu_memory_limit) {
        handle = gv_solver_create(int(dim), device_ids.data(), int(device_ids.size()), num_sampler_per_worker,
                                  gpu_memory_limit, 0, 1);
        if (!handle)
            throw std::runtime_error(gv_last_error());
    }
    GraphSolver(const GraphSolver &) = delete;
    ~GraphSolver() { gv_solver_destroy(handle); }
    std::map<std::string, std::string> attributes() const {
        char text[4096];
        gv_solver_attributes(handle, text, sizeof(text));
        return parse_attributes(text);
    }
    long integer(const char *name) const { return std::stol(attributes().at(name)); }
    double real(const char *name) const { return std::stod(attributes().at(name)); }
};

template<size_t dim>
struct KnowledgeGraphSolver {
    gv_kg_solver_t *handle;
    std::shared_ptr<Optimizer> optimizer;
    py::object graph;
    KnowledgeGraphSolver(std::vector<int> device_ids, int num_sampler_per_worker, size_t gpu_memory_limit) {
        handle = gv_kg_solver_create(int(dim), device_ids.data(), int(device_ids.size()), num_sampler_per_worker,
                                     gpu_memory_limit, 0, 1);
        if (!handle)
            throw std::runtime_error(gv_last_error());
    }
    KnowledgeGraphSolver(const KnowledgeGraphSolver &) = delete;
    ~KnowledgeGraphSolver() { gv_kg_solver_destroy(handle); }
    std::map<std::string, std::string> attributes() const {
        char text[4096];
        gv_kg_solver_attributes(handle, text, sizeof(text));
        return parse_attributes(text);
    }
    long integer(const char *name) const { return std::stol(attributes().at(name)); }
    double real(const char *name) const { return std::stod(attributes().at(name)); }
};

template<size_t dim>
static void bind_graph_solver(py::module &m) {
    using S = GraphSolver<dim>;
    const std::string name = "GraphSolver_" + std::to_string(dim) + "_f_j";  // signature(), bind.h:82-88
    py::class_<S> cls(m, name.c_str());
    cls.attr("__doc__") = "GraphSolver(dim, float_type=dtype.float32, index_type=dtype.uint32, device_ids=[], "
                          "num_sampler_per_worker=auto, gpu_memory_limit=auto)";
    cls.def(py::init<std::vector<int>, int, size_t>(), no_gil(), py::arg("device_ids") = std::vector<int>(),
            py::arg("num_sampler_per_worker") = kAuto, py::arg("gpu_memory_limit") = kAuto);
    cls.def("build",
            [](S &s, py::object graph, std::shared_ptr<Optimizer> optimizer, int num_partition, int num_negative,
               int batch_size, int episode_size) {
                Graph &g = graph.cast<Graph &>();
                const gv_optimizer_t descriptor = optimizer->descriptor();
                {
                    py::gil_scoped_release release;
                    check(gv_solver_build(s.handle, g.handle, &descriptor, num_partition, num_negative, batch_size,
                                          episode_size));
                }
                s.graph = graph;
                s.optimizer = optimizer;
            },
            py::arg("graph"), py::arg("optimizer") = std::make_shared<Optimizer>(kAuto),
            py::arg("num_partition") = kAuto, py::arg("num_negative") = 1, py::arg("batch_size") = 100000,
            py::arg("episode_size") = kAuto,
            "build(graph, optimizer=auto, num_partition=auto, num_negative=1, batch_size=100000, episode_size=auto)");
    cls.def("train",
            [](S &s, const std::string &model, int num_epoch, bool resume, int augmentation_step,
               int random_walk_length, int random_walk_batch_size, int shuffle_base, float p, float q,
               int positive_reuse, float negative_sample_exponent, float negative_weight, int log_frequency) {
                check(gv_solver_train(s.handle, model.c_str(), num_epoch, resume, augmentation_step, random_walk_length,
                                      random_walk_batch_size, shuffle_base, p, q, positive_reuse,
                                      negative_sample_exponent, negative_weight, log_frequency));
            },
            no_gil(), py::arg("model") = "LINE", py::arg("num_epoch") = 2000, py::arg("resume") = false,
            py::arg("augmentation_step") = kAuto, py::arg("random_walk_length") = 40,
            py::arg("random_walk_batch_size") = 100, py::arg("shuffle_base") = kAuto, py::arg("p") = 1,
            py::arg("q") = 1, py::arg("positive_reuse") = 1, py::arg("negative_sample_exponent") = 0.75,
            py::arg("negative_weight") = 5, py::arg("log_frequency") = 1000,
            "train(model='LINE', num_epoch=2000, resume=False, augmentation_step=auto, random_walk_length=40, "
            "random_walk_batch_size=100, shuffle_base=auto, p=1, q=1, positive_reuse=1, "
            "negative_sample_exponent=0.75, negative_weight=5, log_frequency=1000)");
    cls.def("predict",
            [](S &s, py::array_t<unsigned, py::array::c_style | py::array::forcecast> samples) {
                if (samples.ndim() != 2 || samples.shape(1) != 2)
                    throw std::invalid_argument("Expect an array with shape (?, 2)");
                py::array_t<float> logits(samples.shape(0));
                check(gv_solver_predict(s.handle, samples.data(), uint64_t(samples.shape(0)), logits.mutable_data()));
                return logits;
            },
            py::arg("samples"), "predict(samples)");
    cls.def("clear", [](S &s) { check(gv_solver_clear(s.handle)); }, no_gil(), "clear()");
    cls.def("__repr__", [](S &s) {
        char text[8192];
        gv_solver_info(s.handle, text, sizeof(text));
        return std::string(text);
    });
    auto view = [](int which) {
        return [which](py::object self) {
            S &s = self.cast<S &>();
            uint64_t rows = 0;
            int d = 0;
            float *data = gv_solver_embeddings(s.handle, which, &rows, &d);
            return numpy_view(data, rows, int(dim), self);
        };
    };
    cls.def_property_readonly("vertex_embeddings", view(0), "Vertex node embeddings (2D numpy view).");
    cls.def_property_readonly("context_embeddings", view(1), "Context node embeddings (2D numpy view).");
    cls.def_property_readonly("optimizer", [](S &s) { return s.optimizer; });
    cls.def_property_readonly("model", [](S &s) { return s.attributes().at("model"); });
    cls.def_property_readonly("resume", [](S &s) { return s.integer("resume") != 0; });
    for (const char *key : {"num_partition", "num_negative", "num_epoch", "episode_size", "batch_size",
                            "augmentation_step", "random_walk_length", "random_walk_batch_size", "shuffle_base",
                            "positive_reuse", "log_frequency", "num_worker", "num_sampler", "gpu_memory_limit",
                            "gpu_memory_cost"})
        cls.def_property_readonly(key, [key](S &s) { return s.integer(key); });
    for (const char *key : {"negative_sample_exponent", "negative_weight", "p", "q"})
        cls.def_property_readonly(key, [key](S &s) { return s.real(key); });
}

template<size_t dim>
static void bind_knowledge_graph_solver(py::module &m) {
    using S = KnowledgeGraphSolver<dim>;
    const std::string name = "KnowledgeGraphSolver_" + std::to_string(dim) + "_f_j";
    py::class_<S> cls(m, name.c_str());
    cls.attr("__doc__") = "KnowledgeGraphSolver(dim, float_type=dtype.float32, index_type=dtype.uint32, device_ids=[], "
                          "num_sampler_per_worker=auto, gpu_memory_limit=auto)";
    cls.def(py::init<std::vector<int>, int, size_t>(), no_gil(), py::arg("device_ids") = std::vector<int>(),
            py::arg("num_sampler_per_worker") = kAuto, py::arg("gpu_memory_limit") = kAuto);
    cls.def("build",
            [](S &s, py::object graph, std::shared_ptr<Optimizer> optimizer, int num_partition, int num_negative,
               int batch_size, int episode_size) {
                KnowledgeGraph &g = graph.cast<KnowledgeGraph &>();
                const gv_optimizer_t descriptor = optimizer->descriptor();
                {
                    py::gil_scoped_release release;
                    check(gv_kg_solver_build(s.handle, g.handle, &descriptor, num_partition, num_negative, batch_size,
                                             episode_size));
                }
                s.graph = graph;
                s.optimizer = optimizer;
            },
            py::arg("graph"), py::arg("optimizer") = std::make_shared<Optimizer>(kAuto),
            py::arg("num_partition") = kAuto, py::arg("num_negative") = 64, py::arg("batch_size") = 100000,
            py::arg("episode_size") = kAuto,
            "build(graph, optimizer=auto, num_partition=auto, num_negative=64, batch_size=100000, episode_size=auto)");
    cls.def("train",
            [](S &s, const std::string &model, int num_epoch, bool resume, float relation_lr_multiplier, float margin,
               float l3_regularization, int sample_batch_size, int positive_reuse, float adversarial_temperature,
               int log_frequency) {
                check(gv_kg_solver_train(s.handle, model.c_str(), num_epoch, resume, relation_lr_multiplier, margin,
                                         l3_regularization, sample_batch_size, positive_reuse, adversarial_temperature,
                                         log_frequency));
            },
            no_gil(), py::arg("model") = "RotatE", py::arg("num_epoch") = 2000, py::arg("resume") = false,
            py::arg("relation_lr_multiplier") = 1, py::arg("margin") = 12, py::arg("l3_regularization") = 2e-3,
            py::arg("sample_batch_size") = 2000, py::arg("positive_reuse") = 1,
            py::arg("adversarial_temperature") = 2, py::arg("log_frequency") = 100,
            "train(model='RotatE', num_epoch=2000, resume=False, relation_lr_multiplier=1, margin=12, "
            "l3_regularization=2e-3, sample_batch_size=2000, positive_reuse=1, adversarial_temperature=2, "
            "log_frequency=100)");
    cls.def("predict",
            [](S &s, py::array_t<unsigned, py::array::c_style | py::array::forcecast> samples) {
                if (samples.ndim() != 2 || samples.shape(1) != 3)
                    throw std::invalid_argument("Expect an array with shape (?, 3)");
                py::array_t<float> logits(samples.shape(0));
                check(gv_kg_solver_predict(s.handle, samples.data(), uint64_t(samples.shape(0)),
                                           logits.mutable_data()));
                return logits;
            },
            py::arg("samples"), "predict(samples)");
    cls.def("clear", [](S &s) { check(gv_kg_solver_clear(s.handle)); }, no_gil(), "clear()");
    cls.def("__repr__", [](S &s) {
        char text[8192];
        gv_kg_solver_info(s.handle, text, sizeof(text));
        return std::string(text);
    });
    auto view = [](int which) {
        return [which](py::object self) {
            S &s = self.cast<S &>();
            uint64_t rows = 0;
            int d = 0;
            float *data = gv_kg_solver_embeddings(s.handle, which, &rows, &d);
            return numpy_view(data, rows, d, self);
        };
    };
    cls.def_property_readonly("entity_embeddings", view(0), "Entity embeddings (2D numpy view).");
    cls.def_property_readonly("relation_embeddings", view(1), "Relation embeddings (2D numpy view).");
    cls.def_property_readonly("optimizer", [](S &s) { return s.optimizer; });
    cls.def_property_readonly("model", [](S &s) { return s.attributes().at("model"); });
    cls.def_property_readonly("resume", [](S &s) { return s.integer("resume") != 0; });
    for (const char *key : {"num_partition", "num_negative", "sample_batch_size", "num_epoch", "episode_size",
                            "batch_size", "positive_reuse", "log_frequency", "num_worker", "num_sampler",
                            "gpu_memory_limit", "gpu_memory_cost"})
        cls.def_property_readonly(key, [key](S &s) { return s.integer(key); });
    for (const char *key : {"negative_sample_exponent", "relation_lr_multiplier", "margin", "l3_regularization",
                            "adversarial_temperature"})
        cls.def_property_readonly(key, [key](S &s) { return s.real(key); });
}

enum class DType { uint32, uint64, float32, float64 };  // bind.h:62-69

PYBIND11_MODULE(libgraphvite, module) {
    module.doc() = "libgraphvite on libgv_b200: the reference's pybind11 surface over the B200-native C ABI";

    // optimizers
    auto optimizer = module.def_submodule("optimizer");
    py::class_<LRSchedule>(optimizer, "LRSchedule")
        .def(py::init<std::string>(), py::arg("type") = "constant")
        .def(py::init<std::function<float(int, int)>>(), py::arg("schedule_function"))
        .def_readonly("type", &LRSchedule::type)
        .def_readonly("schedule_function", &LRSchedule::schedule_function)
        .def("__repr__", &LRSchedule::info);
    py::implicitly_convertible<std::string, LRSchedule>();
    py::implicitly_convertible<py::function, LRSchedule>();
    py::class_<Optimizer, std::shared_ptr<Optimizer>>(optimizer, "Optimizer")
        .def(py::init<int>(), py::arg("type") = kAuto)
        .def(py::init<float>(), py::arg("lr") = 1e-4)
        .def_readonly("type", &Optimizer::type)
        .def_readonly("lr", &Optimizer::lr)
        .def_readonly("weight_decay", &Optimizer::weight_decay)
        .def_readonly("schedule", &Optimizer::schedule)
        .def("__repr__", &Optimizer::info);
    py::implicitly_convertible<int, Optimizer>();
    py::implicitly_convertible<float, Optimizer>();
    py::class_<SGD, Optimizer, std::shared_ptr<SGD>>(optimizer, "SGD")
        .def(py::init<float, float, LRSchedule>(), py::arg("lr") = 1e-4, py::arg("weight_decay") = 0,
             py::arg("schedule") = "linear");
    py::class_<Momentum, Optimizer, std::shared_ptr<Momentum>>(optimizer, "Momentum")
        .def(py::init<float, float, float, LRSchedule>(), py::arg("lr") = 1e-4, py::arg("weight_decay") = 0,
             py::arg("momentum") = 0.999, py::arg("schedule") = "linear")
        .def_readonly("momentum", &Momentum::momentum);
    py::class_<AdaGrad, Optimizer, std::shared_ptr<AdaGrad>>(optimizer, "AdaGrad")
        .def(py::init<float, float, float, LRSchedule>(), py::arg("lr") = 1e-4, py::arg("weight_decay") = 0,
             py::arg("epsilon") = 1e-10, py::arg("schedule") = "linear")
        .def_readonly("epsilon", &AdaGrad::epsilon);
    py::class_<RMSprop, Optimizer, std::shared_ptr<RMSprop>>(optimizer, "RMSprop")
        .def(py::init<float, float, float, float, LRSchedule>(), py::arg("lr") = 1e-4, py::arg("weight_decay") = 0,
             py::arg("alpha") = 0.999, py::arg("epsilon") = 1e-8, py::arg("schedule") = "linear")
        .def_readonly("alpha", &RMSprop::alpha)
        .def_readonly("epsilon", &RMSprop::epsilon);
    py::class_<Adam, Optimizer, std::shared_ptr<Adam>>(optimizer, "Adam")
        .def(py::init<float, float, float, float, float, LRSchedule>(), py::arg("lr") = 1e-4,
             py::arg("weight_decay") = 0, py::arg("beta1") = 0.999, py::arg("beta2") = 0.99999,
             py::arg("epsilon") = 1e-8, py::arg("schedule") = "linear")
        .def_readonly("beta1", &Adam::beta1)
        .def_readonly("beta2", &Adam::beta2)
        .def_readonly("epsilon", &Adam::epsilon);

    // graphs
    auto graph = module.def_submodule("graph");
    py::class_<Graph>(graph, "Graph_j")
        .def(py::init<>())
        .def_property_readonly("num_vertex", &Graph::num_vertex)
        .def_property_readonly("num_edge", &Graph::num_edge)
        .def_property_readonly("as_undirected", [](Graph &g) { return gv_graph_as_undirected(g.handle) != 0; })
        .def_property_readonly("normalization", [](Graph &g) { return gv_graph_normalization(g.handle) != 0; })
        .def_property_readonly("name2id", &Graph::name2id, "Map of node name to index.")
        .def_property_readonly("id2name", &Graph::id2name, "Map of node index to name.")
        .def("load", &Graph::load_file, no_gil(), py::arg("file_name"), py::arg("as_undirected") = true,
             py::arg("normalization") = false, py::arg("delimiters") = " \t\r\n", py::arg("comment") = "#",
             "load(*args, **kwargs)")
        .def("load", &Graph::load_edge_list, no_gil(), py::arg("edge_list"), py::arg("as_undirected") = true,
             py::arg("normalization") = false)
        .def("load", &Graph::load_weighted_edge_list, no_gil(), py::arg("weighted_edge_list"),
             py::arg("as_undirected") = true, py::arg("normalization") = false)
        .def("save", &Graph::save, no_gil(), py::arg("file_name"), py::arg("weighted") = true,
             py::arg("anonymous") = false, "save(file_name, weighted=True, anonymous=False)")
        .def("__repr__", &Graph::info);
    py::class_<WordGraph, Graph>(graph, "WordGraph_j")
        .def(py::init<>())
        .def("load", &WordGraph::load_corpus, no_gil(), py::arg("file_name"), py::arg("window") = 5,
             py::arg("min_count") = 5, py::arg("normalization") = false, py::arg("delimiters") = " \t\r\n",
             py::arg("comment") = "#",
             "load(file_name, window=5, min_count=5, normalization=False, delimiters=' \\t\\r\\n', comment='#')");
    py::class_<KnowledgeGraph>(graph, "KnowledgeGraph_j")
        .def(py::init<>())
        .def_property_readonly("num_vertex", &KnowledgeGraph::num_vertex)
        .def_property_readonly("num_edge", &KnowledgeGraph::num_edge)
        .def_property_readonly("num_relation", &KnowledgeGraph::num_relation)
        .def_property_readonly("normalization", [](KnowledgeGraph &g) { return gv_kgraph_normalization(g.handle) != 0; })
        .def_property_readonly("entity2id", &KnowledgeGraph::entity2id)
        .def_property_readonly("relation2id", &KnowledgeGraph::relation2id)
        .def_property_readonly("id2entity", &KnowledgeGraph::id2entity)
        .def_property_readonly("id2relation", &KnowledgeGraph::id2relation)
        .def("load", &KnowledgeGraph::load_file, no_gil(), py::arg("file_name"), py::arg("normalization") = false,
             py::arg("delimiters") = " \t\r\n", py::arg("comment") = "#", "load(*args, **kwargs)")
        .def("load", &KnowledgeGraph::load_triplet_list, no_gil(), py::arg("triplet_list"),
             py::arg("normalization") = false)
        .def("load", &KnowledgeGraph::load_weighted_triplet_list, no_gil(), py::arg("weighted_triplet_list"),
             py::arg("normalization") = false)
        .def("save", &KnowledgeGraph::save, no_gil(), py::arg("file_name"), py::arg("anonymous") = false,
             "save(file_name, anonymous=False)")
        .def("__repr__", &KnowledgeGraph::info);

    // solvers: the dimensions src/graphvite.cu:52-70 instantiates
    auto solver = module.def_submodule("solver");
    bind_graph_solver<32>(solver);
    bind_graph_solver<64>(solver);
    bind_graph_solver<96>(solver);
    bind_graph_solver<128>(solver);
    bind_graph_solver<256>(solver);
    bind_graph_solver<512>(solver);
    bind_knowledge_graph_solver<32>(solver);
    bind_knowledge_graph_solver<64>(solver);
    bind_knowledge_graph_solver<96>(solver);
    bind_knowledge_graph_solver<128>(solver);
    bind_knowledge_graph_solver<256>(solver);
    bind_knowledge_graph_solver<512>(solver);
    bind_knowledge_graph_solver<1024>(solver);
    bind_knowledge_graph_solver<2048>(solver);

    // interface (src/graphvite.cu:75-104)
    py::enum_<DType>(module, "dtype")
        .value("uint32", DType::uint32)
        .value("uint64", DType::uint64)
        .value("float32", DType::float32)
        .value("float64", DType::float64);
    py::dict dtype2name;
    dtype2name[py::cast(DType::uint32)] = "j";  // Itanium typeid names, bind.h:71-76
    dtype2name[py::cast(DType::uint64)] = "m";
    dtype2name[py::cast(DType::float32)] = "f";
    dtype2name[py::cast(DType::float64)] = "d";
    module.attr("dtype2name") = dtype2name;
    module.def("init_logging", [](int, const std::string &, bool) {}, py::arg("threshhold") = 0, py::arg("dir") = "",
               py::arg("verbose") = false, "glog is not used by libgv_b200: errors are exceptions, GV_LOG=1 logs");
    module.attr("INFO") = 0;
    module.attr("WARNING") = 1;
    module.attr("ERROR") = 2;
    module.attr("FATAL") = 3;
    module.attr("auto") = kAuto;
    module.def("KiB", [](size_t size) { return size << 10; }, py::arg("size"));
    module.def("MiB", [](size_t size) { return size << 20; }, py::arg("size"));
    module.def("GiB", [](size_t size) { return size << 30; }, py::arg("size"));
    module.attr("__version__") = gv_version();
    module.attr("__backend__") = "libgv_b200";
}
