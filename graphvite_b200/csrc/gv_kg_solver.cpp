// =============================================================================
// gv_kg_solver.cpp -- host runtime of the knowledge-graph embedding solver (C++ over the CUDA C ABI).
//
// Mirrors graphvite::KnowledgeGraphSolver<dim, float, uint32> (reference
// include/instance/knowledge_graph.cuh:286-677) on the SolverMixin / SamplerMixin / WorkerMixin machinery of
// include/core/solver.h, re-designed for B200 like gv_solver.cpp:
//   * the tied entity matrix lives in HBM as P partition blocks (with their moments); a worker trains a
//     (head block, tail block) pair in place -- no gather / scatter through host memory per block;
//   * the edge sampler runs on the device (gv_kg_sampler.cu draw + gv_sampler.cu stable partition with the
//     relation as attribute); streams, consumption and pool layout are the reference's, bit for bit;
//   * every worker trains a private copy of the global relation matrix; after each schedule step the
//     copies are reconciled by summing their deltas (one all-reduce over NVLink for world_size > 1),
//     which is what the reference's per-worker `global -= loaded - trained` amounts to;
//   * between steps entity blocks move GPU-to-GPU through the caller's exchange (NCCL P2P).
// One process drives one GPU; world_size processes form the reference's num_worker.
// =============================================================================
#include <atomic>
#include <climits>
#include <cmath>
#include <cstring>
#include <deque>
#include <memory>
#include <thread>

#include "gv_runtime.h"

namespace gv {

static const int kSamplePerVertexWithGlobal = 50;  // core/solver.h:55
static const char *kModelNames[] = {"TransE", "DistMult", "ComplEx", "SimplE", "RotatE", "QuatE"};

static int kg_model_id(const std::string &name) {
    for (int i = 0; i < 6; i++)
        if (name == kModelNames[i])
            return i;
    return -1;
}

struct KgAssignment {
    int head, tail;
};

// SolverMixin::get_schedule for tied weights, core/solver.h:519-561: within a group of 2W partitions first the
// two diagonals, then for group sizes 1, 2, 4 .. W the off-diagonal pairings and their transposes; every
// step touches 2W distinct partitions (W on a diagonal), so concurrent workers never share an entity row.
static std::vector<std::vector<KgAssignment>> make_tied_schedule(int num_partition, int num_worker) {
    std::vector<std::vector<KgAssignment>> schedule;
    if (num_partition == 1) {
        schedule.push_back({{0, 0}});
        return schedule;
    }
    std::vector<KgAssignment> assignment(num_worker);
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
                        const int head = x + (i / group_size * 2) * group_size + i % group_size;
                        const int tail = y + (i / group_size * 2 + 1) * group_size + (i + offset) % group_size;
                        assignment[i] = {head, tail};
                    }
                    schedule.push_back(assignment);
                    for (int i = 0; i < num_worker; i++)
                        std::swap(assignment[i].head, assignment[i].tail);
                    schedule.push_back(assignment);
                }
        }
    return schedule;
}

struct KgSolver {
    // ---- construction (SolverMixin ctor, core/solver.h:184-213) ----
    int dim, device = 0, rank, world_size;
    int num_worker, num_sampler;
    uint64_t gpu_memory_limit, gpu_memory_cost = 0;
    std::vector<unsigned long long> sampler_seeds, worker_seeds;
    cudaStream_t work_stream = nullptr, sample_stream = nullptr, random_stream = nullptr;
    std::vector<gv_rng_t *> sampler_generators;
    gv_rng_t *worker_generator = nullptr;
    DeviceArray d_rng_snapshot;
    gv_exchange_fn exchange_fn = nullptr;
    void *exchange_ctx = nullptr;
    gv_allreduce_fn allreduce_fn = nullptr;
    void *allreduce_ctx = nullptr;

    // ---- build ----
    KnowledgeGraph *graph = nullptr;
    HostOptimizer optimizer;
    int num_partition = 0, num_negative = 64, batch_size = 100000, episode_size = 0;
    bool shuffle_partition = false;
    int shuffle_override = -1;  // test hook: -1 = the reference's rule
    int assignment_offset = 0;
    std::vector<std::vector<uint32_t>> partitions;
    std::vector<gv_location_t> locations;
    uint32_t partition_size = 0;
    bool built = false;
    std::vector<float> entity_host, entity_m1_host, entity_m2_host, relation_host;  // numpy views + resume state

    // ---- train parameters (readonly attributes, bind.h:547-566) ----
    std::string model;
    int model_id = -1;
    int num_epoch = 0, sample_batch_size = 2000, positive_reuse = 1, log_frequency = 100;
    float relation_lr_multiplier = 1, margin = 12, l3_regularization = 2e-3f, adversarial_temperature = 2;
    bool resume = false;
    int batch_id = 0, num_batch = 0, pool_id = 0;
    bool training = false;

    // ---- device state ----
    size_t block_floats = 0, relation_floats = 0;
    int num_state = 1;                        // 1 + num_moment matrices per block
    std::deque<DeviceArray> entity_slots;     // each num_state * block_floats floats
    std::vector<int> slot_of_part, owner;     // partition -> local slot (-1: elsewhere), partition -> rank
    std::vector<int> free_slots;
    std::vector<DeviceArray> partition_ids;   // [P] global ids of each partition
    DeviceArray d_relation_global, d_relation_work, d_relation_delta;  // work: values then the worker's moments
    DeviceArray pool_arena, pool_pointers[2];
    DeviceArray d_edge_h, d_edge_t, d_edge_r, d_edge_prob, d_edge_alias, d_locations;
    gv_device_kgraph_t device_graph;
    bool sampling_ready = false;
    DeviceArray d_sampler_random, d_chains, d_relations, d_fill, d_last_walk, d_fill_scratch;
    uint64_t draw_chunk = 1 << 20;
    static const int kRandomBuffers = 4;
    DeviceArray d_random[kRandomBuffers], d_lr, d_loss, d_negatives_out;
    cudaEvent_t random_ready[kRandomBuffers] = {}, random_free[kRandomBuffers] = {};
    int chunk_batches = 1;
    bool capture_negatives = false;
    int train_num_groups = 0;  // 0 = fill the device; 1 = one thread group (sequential, reproducible; tests)
    std::vector<uint32_t> last_negatives;
    std::vector<float> logged_loss;
    float previous_batch_loss = 0;  // the worker's loss buffer outlives blocks and train() calls (solver.h:1326)
    double stat_positive = 0, stat_kernel_seconds = 0, stat_train_seconds = 0, stat_sample_seconds = 0;
    std::atomic<unsigned long long> stat_launches{0};

    KgSolver(int _dim, const int *device_ids, int num_device, int num_sampler_per_worker, uint64_t memory_limit,
             int _rank, int _world_size)
        : dim(_dim), rank(_rank), world_size(_world_size), gpu_memory_limit(memory_limit) {
        require(dim >= 2 && dim % 2 == 0 && dim <= 2048, "unsupported embedding dimension " + std::to_string(dim));
        require(world_size >= 1 && rank >= 0 && rank < world_size, "invalid rank / world_size");
        require(num_device <= 1, "one process drives one GPU: launch one process per GPU (torchrun) and pass "
                                 "rank / world_size instead of several device ids");
        if (num_device == 1)
            device = device_ids[0];
        else {
            int count = 0;
            GV_CHECK_CUDA(cudaGetDeviceCount(&count));
            require(count > 0, "No GPU devices found");
            device = world_size > 1 ? rank % count : 0;
        }
        num_worker = world_size;
        if (num_sampler_per_worker == 0)  // samplers are device streams here, not CPU threads: auto = 1
            num_sampler_per_worker = 1;
        require(num_sampler_per_worker > 0, "num_sampler_per_worker must be positive");
        num_sampler = num_sampler_per_worker * num_worker;
        GV_CHECK_CUDA(cudaSetDevice(device));
        if (gpu_memory_limit == 0) {
            size_t free_bytes = 0, total_bytes = 0;
            GV_CHECK_CUDA(cudaMemGetInfo(&free_bytes, &total_bytes));
            gpu_memory_limit = free_bytes;
        }
        // seeds: samplers first, then workers, from the process-wide engine (core/solver.h:208-212)
        std::uniform_int_distribution<unsigned long long> random_seed(0, ULLONG_MAX);
        for (int i = 0; i < num_sampler; i++)
            sampler_seeds.push_back(random_seed(g_engine));
        for (int i = 0; i < num_worker; i++)
            worker_seeds.push_back(random_seed(g_engine));
        GV_CHECK_CUDA(cudaStreamCreateWithFlags(&work_stream, cudaStreamNonBlocking));
        int least = 0, greatest = 0;
        GV_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        GV_CHECK_CUDA(cudaStreamCreateWithPriority(&sample_stream, cudaStreamNonBlocking, greatest));
        GV_CHECK_CUDA(cudaStreamCreateWithPriority(&random_stream, cudaStreamNonBlocking, greatest));
        for (int i = 0; i < num_sampler; i++) {
            gv_rng_t *generator = gv_rng_create(sampler_seeds[i], sample_stream);
            require(generator != nullptr, gv_last_error());
            sampler_generators.push_back(generator);
        }
        worker_generator = gv_rng_create(worker_seeds[rank], random_stream);
        require(worker_generator != nullptr, gv_last_error());
        d_rng_snapshot.allocate(gv_rng_state_bytes());
        for (int i = 0; i < kRandomBuffers; i++) {
            GV_CHECK_CUDA(cudaEventCreateWithFlags(&random_ready[i], cudaEventDisableTiming));
            GV_CHECK_CUDA(cudaEventCreateWithFlags(&random_free[i], cudaEventDisableTiming));
        }
        memset(&device_graph, 0, sizeof(device_graph));
    }

    ~KgSolver() {
        cudaSetDevice(device);
        if (sampler_thread.joinable())
            sampler_thread.join();
        for (auto g : sampler_generators)
            gv_rng_destroy(g);
        gv_rng_destroy(worker_generator);
        for (int i = 0; i < kRandomBuffers; i++) {
            if (random_ready[i])
                cudaEventDestroy(random_ready[i]);
            if (random_free[i])
                cudaEventDestroy(random_free[i]);
        }
        if (work_stream)
            cudaStreamDestroy(work_stream);
        if (sample_stream)
            cudaStreamDestroy(sample_stream);
        if (random_stream)
            cudaStreamDestroy(random_stream);
    }

    int num_moment() const { return optimizer.num_moment(); }
    uint64_t pool_size() const { return uint64_t(episode_size) * batch_size; }
    uint64_t pool_block_bytes() const { return pool_size() * 3 * sizeof(uint32_t); }
    uint32_t *pool_block(int side, int head, int tail) const {
        return reinterpret_cast<uint32_t *>(static_cast<char *>(pool_arena.ptr) +
                                            ((uint64_t(side) * num_partition + head) * num_partition + tail) *
                                                pool_block_bytes());
    }

    // bytes this rank keeps resident (our memory model; every rank samples all P * P pool blocks)
    uint64_t memory_demand(int P, int episode) const {
        const uint64_t rows = (graph->num_vertex() + P - 1) / P;
        const uint64_t block = rows * dim * sizeof(float) * (1 + optimizer.num_moment());
        const int resident = num_worker > 1 ? std::min(P, 4) : P;  // blocks a rank typically holds
        uint64_t demand = block * resident;
        demand += uint64_t(graph->num_relation()) * dim * sizeof(float) * (3 + optimizer.num_moment());
        demand += uint64_t(2) * P * P * episode * batch_size * 12;                       // both sample pools
        demand += uint64_t(graph->log_h.size()) * (4 + 4 + 4 + 4 + 8);                  // triplets + edge table
        demand += uint64_t(graph->num_vertex()) * 8;                                    // locations
        demand += uint64_t(kSpanBuffers) * kRandBatchSize * 8;                          // samplers' refill buffers
        demand += uint64_t(kRandomBuffers) * batch_size * std::max(1, num_negative) * 16;  // negatives' randoms
        demand += uint64_t(3) * 64 * 1024 * 1024;                                       // draws + fill scratch
        demand += uint64_t(graph->num_vertex()) * dim * sizeof(float);                  // staging for load / write-back
        return demand;
    }

    // ---- SolverMixin::build, core/solver.h:287-466 ----
    void build(KnowledgeGraph *_graph, const gv_optimizer_t *_optimizer, int _num_partition, int _num_negative,
               int _batch_size, int _episode_size) {
        require(!training, "build() during training");
        GV_CHECK_CUDA(cudaSetDevice(device));
        graph = _graph;
        require(graph->num_vertex() > 0 && graph->num_relation() > 0, "The graph is empty");
        optimizer.desc = *_optimizer;
        if (optimizer.desc.type < 0) {  // get_default_optimizer: Adam(5e-5, 0), knowledge_graph.cuh:557-559
            const float lr = optimizer.desc.lr;
            optimizer.desc.type = GV_OPT_ADAM;
            optimizer.desc.lr = lr > 0 ? lr : 5e-5f;
            optimizer.desc.weight_decay = 0;
            optimizer.desc.a = 0.999f;  // core/optimizer.h:312
            optimizer.desc.b = 0.99999f;
            optimizer.desc.epsilon = 1e-8f;
            optimizer.desc.schedule = GV_SCHEDULE_LINEAR;
        }
        optimizer.init_lr = optimizer.desc.lr;
        num_negative = _num_negative;
        batch_size = _batch_size;
        require(num_negative >= 1 && batch_size > 0, "invalid num_negative / batch_size");
        batch_id = 0;
        graph->flatten();

        // tied weights: a step keeps 2 partitions per worker busy (core/solver.h:266-277)
        const int min_partition = num_worker == 1 ? 1 : num_worker * 2;
        auto auto_episode = [&](int P) {  // core/solver.h:426-436, a global matrix is present
            int expected = float(uint64_t(graph->num_vertex()) * kSamplePerVertexWithGlobal) / P / batch_size;
            expected = std::max(expected, 1);
            if (P == 1)
                expected = std::max(expected, kMinEpisodeSample / batch_size);
            return expected;
        };
        // the tied schedule pairs partitions inside groups of 2 * #worker (core/solver.h:532-561): any other
        // count > 1 makes the reference index partitions that do not exist, so it is refused here
        auto valid = [&](int P) { return P == 1 ? num_worker == 1 : P % (num_worker * 2) == 0; };
        num_partition = _num_partition;
        if (num_partition == 0) {
            for (num_partition = min_partition; num_partition < kMaxPartition; num_partition += min_partition)
                if (valid(num_partition) &&
                    memory_demand(num_partition, _episode_size ? _episode_size : auto_episode(num_partition)) <
                        gpu_memory_limit)
                    break;
        } else
            require(num_partition >= min_partition,
                    "#partition should be no less than " + std::to_string(min_partition));
        require(valid(num_partition), "with tied weights #partition must be 1 (single worker) or a multiple of " +
                                          std::to_string(num_worker * 2));
        episode_size = _episode_size ? _episode_size : auto_episode(num_partition);
        while (episode_size > 1 && memory_demand(num_partition, episode_size) >= gpu_memory_limit)
            episode_size /= 2;
        gpu_memory_cost = memory_demand(num_partition, episode_size);
        require(gpu_memory_cost < gpu_memory_limit, "Can't satisfy the specified GPU memory limit");
        // core/solver.h:385: the tail partitions are rotated every episode when a global matrix has moments.
        // With several workers the rotated steps would make two workers train private copies of one entity
        // partition (the later write-back wins in the reference); blocks have ONE owner here, so the rotation
        // is only applied with a single worker.
        shuffle_partition = optimizer.num_moment() > 0 && num_worker == 1;
        if (shuffle_override >= 0)
            shuffle_partition = shuffle_override != 0;
        require(!(shuffle_partition && num_worker > 1), "shuffle_partition needs a single worker");
        assignment_offset = 0;

        partitions = partition_vertices(graph->vertex_weights, num_partition);
        partition_size = 0;
        for (auto &part : partitions)
            partition_size = std::max<uint32_t>(partition_size, part.size());
        locations.resize(graph->num_vertex());
        for (int i = 0; i < num_partition; i++)
            for (uint32_t j = 0; j < partitions[i].size(); j++)
                locations[partitions[i][j]] = {uint32_t(i), j};

        const int nm = num_moment();
        const size_t total = size_t(graph->num_vertex()) * dim;
        entity_host.assign(total, 0.f);
        entity_m1_host.assign(nm >= 1 ? total : 0, 0.f);
        entity_m2_host.assign(nm >= 2 ? total : 0, 0.f);
        relation_floats = size_t(graph->num_relation()) * dim;
        relation_host.assign(relation_floats, 0.f);

        // ---- device residency ----
        block_floats = size_t(partition_size) * dim;
        num_state = 1 + nm;
        entity_slots.clear();
        free_slots.clear();
        slot_of_part.assign(num_partition, -1);
        owner.assign(num_partition, 0);
        partition_ids = std::vector<DeviceArray>(num_partition);
        for (int i = 0; i < num_partition; i++)
            partition_ids[i].upload(partitions[i], work_stream);
        d_locations.upload(locations, work_stream);
        d_relation_global.allocate(relation_floats * sizeof(float));
        d_relation_work.allocate(relation_floats * num_state * sizeof(float));
        d_relation_delta.allocate(relation_floats * sizeof(float));
        // the worker's relation moments are loaded once (zeros) and then live on the device: they are
        // neither written back nor re-initialised by train(resume=False) (core/solver.h:1378-1385,1422-1427)
        GV_CHECK_CUDA(cudaMemsetAsync(d_relation_work.ptr, 0, d_relation_work.bytes, work_stream));
        pool_arena.allocate(uint64_t(2) * num_partition * num_partition * pool_block_bytes());
        for (int side = 0; side < 2; side++) {
            std::vector<uint32_t *> pointers(size_t(num_partition) * num_partition);
            for (int h = 0; h < num_partition; h++)
                for (int t = 0; t < num_partition; t++)
                    pointers[size_t(h) * num_partition + t] = pool_block(side, h, t);
            pool_pointers[side].upload(pointers, work_stream);
        }
        const uint64_t per_batch_random = uint64_t(batch_size) * num_negative * 2 * sizeof(double);
        chunk_batches = int(std::max<uint64_t>(1, std::min<uint64_t>(std::min(episode_size, 16),
                                                                    (uint64_t(192) << 20) / per_batch_random)));
        for (int i = 0; i < kRandomBuffers; i++)
            d_random[i].allocate(per_batch_random * chunk_batches);
        d_lr.allocate(size_t(episode_size) * sizeof(float));
        d_loss.allocate(size_t(episode_size) * sizeof(float));
        d_fill.allocate(size_t(num_partition) * num_partition * sizeof(unsigned long long));
        d_last_walk.allocate(sizeof(unsigned long long));
        GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
        sampling_ready = false;
        pool_id = 0;
        previous_batch_loss = 0;
        logged_loss.clear();
        built = true;
    }

    // ---- device triplets + edge alias table (edge_table.build(graph->edge_weights), core/solver.h:255-256) ----
    void prepare_sampling() {
        if (sampling_ready)
            return;
        const size_t m = graph->edge_h.size();
        require(m > 0, "The graph has no edges");
        std::vector<float> edge_prob(m);
        std::vector<uint64_t> edge_alias(m);
        build_alias<uint64_t>(graph->edge_w.data(), m, edge_prob.data(), edge_alias.data());
        d_edge_h.upload(graph->edge_h, sample_stream);
        d_edge_t.upload(graph->edge_t, sample_stream);
        d_edge_r.upload(graph->edge_r, sample_stream);
        d_edge_prob.upload(edge_prob, sample_stream);
        d_edge_alias.upload(edge_alias, sample_stream);
        device_graph.num_edge = m;
        device_graph.edge_h = d_edge_h.as<uint32_t>();
        device_graph.edge_t = d_edge_t.as<uint32_t>();
        device_graph.edge_r = d_edge_r.as<uint32_t>();
        device_graph.edge_prob = d_edge_prob.as<float>();
        device_graph.edge_alias = d_edge_alias.as<uint64_t>();
        device_graph.locations = d_locations.as<gv_location_t>();
        draw_chunk = std::min<uint64_t>(uint64_t(1) << 20, (uint64_t(64) << 20) / (uint64_t(num_partition) * num_partition * 4));
        draw_chunk = std::max<uint64_t>(draw_chunk, 1024);
        d_chains.allocate(draw_chunk * 2 * sizeof(gv_location_t));
        d_relations.allocate(draw_chunk * sizeof(uint32_t));
        d_fill_scratch.allocate(gv_cuda_fill_scratch_bytes(uint32_t(draw_chunk), num_partition));
        d_sampler_random.allocate(size_t(kSpanBuffers) * kRandBatchSize * sizeof(double));
        sampling_ready = true;
    }

    // ---- one sampler's share of a pool (SamplerMixin::sample, core/solver.h:1011-1055) ----
    void run_sampler(int sampler_id, int side, uint64_t start, uint64_t end) {
        const uint64_t slice = end - start;
        if (slice == 0)
            return;
        // a refill buffer of 5e6 doubles holds exactly 2.5e6 draws; termination is checked once per
        // sample_batch_size draws (:1026-1054)
        const uint64_t draws_per_buffer = kRandBatchSize / 2, draw_batch = uint64_t(sample_batch_size);
        const int num_block = num_partition * num_partition;
        gv_fill_params_t params;
        params.num_partition = num_partition;
        params.walk_length = 1;
        params.augmentation_step = 1;
        params.shuffle_base = 1;
        params.pool_size = pool_size();
        params.start = start;
        params.end = end;
        params.attributes = d_relations.as<uint32_t>();

        GV_CHECK_CUDA(cudaMemsetAsync(d_fill.ptr, 0, d_fill.bytes, sample_stream));
        GV_CHECK_CUDA(cudaMemsetAsync(d_last_walk.ptr, 0, sizeof(unsigned long long), sample_stream));
        GV_CHECK_ABI(gv_rng_save(sampler_generators[sampler_id], d_rng_snapshot.ptr, sample_stream));
        std::vector<unsigned long long> fill(num_block, 0);
        uint64_t buffers = 0, draws_done = 0;
        bool complete = false;
        unsigned long long last_draw = 0;
        const uint64_t span_capacity = d_sampler_random.bytes / (size_t(kRandBatchSize) * sizeof(double));
        auto missing_draws = [&]() {
            uint64_t missing = 0;
            for (int b = 0; b < num_block; b++)
                missing = std::max<uint64_t>(missing, slice - std::min<uint64_t>(slice, fill[b]));
            return double(missing) * num_block;  // the emptiest block receives at most 1 / num_block of the draws
        };
        while (!complete) {
            const uint64_t span = std::max<uint64_t>(
                1, std::min<uint64_t>(span_capacity, uint64_t(missing_draws() * 0.97 / draws_per_buffer)));
            GV_CHECK_ABI(gv_rng_generate(sampler_generators[sampler_id], d_sampler_random.as<double>(),
                                         span * uint64_t(kRandBatchSize), sample_stream));
            stat_launches++;
            buffers += span;
            const uint64_t in_span = span * draws_per_buffer;
            uint64_t done_in_span = 0;
            while (done_in_span < in_span && !complete) {
                uint64_t want = in_span - done_in_span;
                if (span == 1) {  // finishing: stop as soon as possible (checked per batch of draws)
                    want = uint64_t(missing_draws() * 1.02) + 2 * draw_batch;
                    want = (want + draw_batch - 1) / draw_batch * draw_batch;
                }
                const uint32_t count = uint32_t(std::min<uint64_t>(std::min<uint64_t>(want, draw_chunk), in_span - done_in_span));
                GV_CHECK_ABI(gv_cuda_kg_draw(&device_graph, d_sampler_random.as<double>() + 2 * done_in_span, count,
                                             d_chains.as<gv_location_t>(), d_relations.as<uint32_t>(), sample_stream));
                GV_CHECK_ABI(gv_cuda_fill_pool(&params, d_chains.as<gv_location_t>(), count, draws_done,
                                               pool_pointers[side].as<uint32_t *>(), d_fill.as<unsigned long long>(),
                                               d_last_walk.as<unsigned long long>(), d_fill_scratch.ptr, sample_stream));
                stat_launches += num_partition == 1 ? 3 : 4;
                GV_CHECK_CUDA(cudaMemcpyAsync(fill.data(), d_fill.ptr, num_block * sizeof(unsigned long long),
                                              cudaMemcpyDeviceToHost, sample_stream));
                GV_CHECK_CUDA(cudaMemcpyAsync(&last_draw, d_last_walk.ptr, sizeof(last_draw), cudaMemcpyDeviceToHost,
                                              sample_stream));
                GV_CHECK_CUDA(cudaStreamSynchronize(sample_stream));
                done_in_span += count;
                draws_done += count;
                complete = true;
                for (int b = 0; b < num_block; b++)
                    complete = complete && fill[b] >= slice;
            }
        }
        // The reference stops at the end of the batch of draws that completed the last block and has pulled
        // one refill per started 2.5e6 draws: keep the generator in step (rewind if a span overshot).
        const uint64_t executed = (last_draw / draw_batch + 1) * draw_batch;
        const uint64_t needed_buffers = (executed - 1) / draws_per_buffer + 1;
        for (; buffers < needed_buffers; buffers++)
            GV_CHECK_ABI(gv_rng_generate(sampler_generators[sampler_id], d_sampler_random.as<double>(), kRandBatchSize,
                                         sample_stream));
        if (buffers > needed_buffers) {
            GV_CHECK_ABI(gv_rng_restore(sampler_generators[sampler_id], d_rng_snapshot.ptr, sample_stream));
            for (uint64_t j = 0; j < needed_buffers; j++)
                GV_CHECK_ABI(gv_rng_generate(sampler_generators[sampler_id], d_sampler_random.as<double>(),
                                             kRandBatchSize, sample_stream));
        }
    }

    // fill one side of the sample pools with all samplers (core/solver.h:614-628); every rank fills all blocks
    void fill_pool(int side) {
        GV_CHECK_CUDA(cudaSetDevice(device));
        const double begin = now_seconds();
        const uint64_t num_sample = pool_size();
        const uint64_t work_load = (num_sample + num_sampler - 1) / num_sampler;
        for (int i = 0; i < num_sampler; i++)
            run_sampler(i, side, std::min(num_sample, work_load * i), std::min(num_sample, work_load * (i + 1)));
        stat_sample_seconds += now_seconds() - begin;
    }

    // ---- entity blocks ----
    int acquire_slot() {
        if (!free_slots.empty()) {
            const int slot = free_slots.back();
            free_slots.pop_back();
            return slot;
        }
        entity_slots.emplace_back();
        entity_slots.back().allocate(block_floats * num_state * sizeof(float));
        return int(entity_slots.size()) - 1;
    }

    std::vector<float> *entity_state(int s) { return s == 0 ? &entity_host : (s == 1 ? &entity_m1_host : &entity_m2_host); }

    // load the blocks this rank owns initially (partition p lives on rank p % W) and the relation matrix
    void load_blocks() {
        const size_t total = size_t(graph->num_vertex()) * dim * sizeof(float);
        DeviceArray staging;
        staging.allocate(total);
        for (int slot = 0; slot < int(entity_slots.size()); slot++)
            if (std::find(free_slots.begin(), free_slots.end(), slot) == free_slots.end())
                free_slots.push_back(slot);
        for (int p = 0; p < num_partition; p++) {
            owner[p] = p % num_worker;
            slot_of_part[p] = owner[p] == rank ? acquire_slot() : -1;
        }
        for (int s = 0; s < num_state; s++) {
            GV_CHECK_CUDA(cudaMemcpyAsync(staging.ptr, entity_state(s)->data(), total, cudaMemcpyHostToDevice, work_stream));
            for (int p = 0; p < num_partition; p++)
                if (slot_of_part[p] >= 0)
                    GV_CHECK_ABI(gv_cuda_move_rows(entity_slots[slot_of_part[p]].as<float>() + s * block_floats,
                                                   staging.as<float>(), partition_ids[p].as<uint32_t>(),
                                                   partitions[p].size(), dim, 1, work_stream));
            GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
        }
        GV_CHECK_CUDA(cudaMemcpyAsync(d_relation_global.ptr, relation_host.data(), relation_floats * sizeof(float),
                                      cudaMemcpyHostToDevice, work_stream));
        GV_CHECK_CUDA(cudaMemcpyAsync(d_relation_work.ptr, relation_host.data(), relation_floats * sizeof(float),
                                      cudaMemcpyHostToDevice, work_stream));
        GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
    }

    void exchange(const void *send, int dst, void *recv, int src, uint64_t bytes) {
        require(exchange_fn != nullptr, "world_size > 1 needs gv_kg_solver_set_exchange()");
        if (exchange_fn(send, dst, recv, src, bytes, work_stream, exchange_ctx) != 0)
            throw std::runtime_error("block exchange failed");
    }

    // Bring the entity blocks of this step onto their workers.  Every rank derives the same transfer list
    // (ordered by worker, head before tail) and takes part in the transfers that name it, in list order.
    void move_blocks(const std::vector<KgAssignment> &step) {
        const uint64_t bytes = block_floats * num_state * sizeof(float);
        std::vector<int> claimed(num_partition, -1);
        for (int i = 0; i < int(step.size()); i++)
            for (int part : {step[i].head, step[i].tail}) {
                require(claimed[part] < 0 || claimed[part] == i,
                        "internal error: two workers need entity partition " + std::to_string(part) + " in one step");
                if (claimed[part] == i)
                    continue;
                claimed[part] = i;
                const int from = owner[part];
                if (from == i)
                    continue;
                if (from == rank) {
                    exchange(entity_slots[slot_of_part[part]].ptr, i, nullptr, -1, bytes);
                    // the slot may be reused by a later receive of this step: wait until the block has left
                    GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
                    free_slots.push_back(slot_of_part[part]);
                    slot_of_part[part] = -1;
                } else if (i == rank) {
                    const int slot = acquire_slot();
                    exchange(nullptr, -1, entity_slots[slot].ptr, from, bytes);
                    slot_of_part[part] = slot;
                }
                owner[part] = i;
            }
    }

    // write_embedding of the global relation matrix for every worker, then load_embedding
    // (core/solver.h:1413-1420,1436-1500): global -= sum over the workers of (loaded - trained)
    void sync_relation() {
        GV_CHECK_ABI(gv_cuda_kg_relation_delta(d_relation_global.as<float>(), d_relation_work.as<float>(),
                                               d_relation_delta.as<float>(), relation_floats, work_stream));
        if (num_worker > 1) {
            require(allreduce_fn != nullptr, "world_size > 1 needs gv_kg_solver_set_allreduce()");
            if (allreduce_fn(d_relation_delta.ptr, relation_floats, work_stream, allreduce_ctx) != 0)
                throw std::runtime_error("relation all-reduce failed");
        }
        GV_CHECK_ABI(gv_cuda_kg_relation_apply(d_relation_global.as<float>(), d_relation_work.as<float>(),
                                               d_relation_delta.as<float>(), relation_floats, work_stream));
        stat_launches += 2;
    }

    // write everything back into the host matrices (write_back, core/solver.h:650-653,1498-1504); with several
    // ranks every block is passed to every rank so that all of them end up with complete numpy views
    void write_back() {
        const size_t total = size_t(graph->num_vertex()) * dim * sizeof(float);
        const uint64_t block_bytes = block_floats * num_state * sizeof(float);
        DeviceArray staging, incoming;
        staging.allocate(total * num_state);
        if (num_worker > 1)
            incoming.allocate(block_bytes);
        for (int p = 0; p < num_partition; p++) {
            const float *block;
            if (owner[p] == rank) {
                block = entity_slots[slot_of_part[p]].as<float>();
                for (int dst = 0; dst < num_worker; dst++)
                    if (dst != rank)
                        exchange(block, dst, nullptr, -1, block_bytes);
            } else {
                exchange(nullptr, -1, incoming.ptr, owner[p], block_bytes);
                block = incoming.as<float>();
            }
            for (int s = 0; s < num_state; s++)
                GV_CHECK_ABI(gv_cuda_move_rows(staging.as<float>() + s * (total / sizeof(float)), block + s * block_floats,
                                               partition_ids[p].as<uint32_t>(), partitions[p].size(), dim, 0,
                                               work_stream));
            GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
        }
        for (int s = 0; s < num_state; s++)
            GV_CHECK_CUDA(cudaMemcpyAsync(entity_state(s)->data(), staging.as<char>() + s * total, total,
                                          cudaMemcpyDeviceToHost, work_stream));
        GV_CHECK_CUDA(cudaMemcpyAsync(relation_host.data(), d_relation_global.ptr, relation_floats * sizeof(float),
                                      cudaMemcpyDeviceToHost, work_stream));
        GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
    }

    // KnowledgeGraphSolver::init_embeddings, knowledge_graph.cuh:567-627
    void init_embeddings() {
        static const float kPi = atan(1) * 4;
        const size_t d = dim;
        if (model == "TransE") {  // every element in order: drawn in bulk (gv_engine.h), same values
            std::uniform_real_distribution<float> init(-margin / d, margin / d);
            g_engine.fill_uniform(entity_host.data(), entity_host.size(), init.a(), init.b());
            g_engine.fill_uniform(relation_host.data(), relation_host.size(), init.a(), init.b());
        }
        if (model == "DistMult" || model == "ComplEx" || model == "SimplE") {
            std::uniform_real_distribution<float> init(-0.5, 0.5);
            g_engine.fill_uniform(entity_host.data(), entity_host.size(), init.a(), init.b());
            g_engine.fill_uniform(relation_host.data(), relation_host.size(), init.a(), init.b());
        }
        if (model == "QuatE") {
            std::uniform_real_distribution<float> init_modulus(-1 / sqrt(d / 2), 1 / sqrt(d / 2));  // he init
            std::uniform_real_distribution<float> init_phase(-kPi, kPi);
            std::uniform_real_distribution<float> init(0, 1);
            for (auto *matrix : {&entity_host, &relation_host})
                for (size_t row = 0; row < matrix->size() / d; row++)
                    for (size_t i = 0; i < d / 4; i++) {
                        const float modulus = init_modulus(g_engine);
                        const float phase = init_phase(g_engine);
                        float v_i = init(g_engine);
                        float v_j = init(g_engine);
                        float v_k = init(g_engine);
                        const float norm = sqrtf(v_i * v_i + v_j * v_j + v_k * v_k);
                        v_i /= norm + 1e-15f;
                        v_j /= norm + 1e-15f;
                        v_k /= norm + 1e-15f;
                        float *e = matrix->data() + row * d + i * 4;
                        e[0] = modulus * cosf(phase);
                        e[1] = modulus * v_i * sinf(phase);
                        e[2] = modulus * v_j * sinf(phase);
                        e[3] = modulus * v_k * sinf(phase);
                    }
        }
        if (model == "RotatE") {
            std::uniform_real_distribution<float> init(-margin * 2 / d, margin * 2 / d);
            std::uniform_real_distribution<float> init_phase(-kPi, kPi);
            g_engine.fill_uniform(entity_host.data(), entity_host.size(), init.a(), init.b());
            for (uint32_t r = 0; r < graph->num_relation(); r++)
                for (size_t i = 0; i < d / 2; i++)
                    relation_host[r * d + i] = init_phase(g_engine);
        }
    }

    // ---- KnowledgeGraphSolver::train prologue + SolverMixin::train up to the first pool fill ----
    void train_begin(const std::string &_model, int _num_epoch, bool _resume, float _relation_lr_multiplier,
                     float _margin, float _l3_regularization, int _sample_batch_size, int _positive_reuse,
                     float _adversarial_temperature, int _log_frequency) {
        require(built, "The model must be built on a graph first");
        require(!training, "train() is already running");
        GV_CHECK_CUDA(cudaSetDevice(device));
        relation_lr_multiplier = _relation_lr_multiplier;
        margin = _margin;
        l3_regularization = _l3_regularization;
        adversarial_temperature = _adversarial_temperature;
        require(kg_model_id(_model) >= 0, "Invalid model `" + _model + "`");
        require(_model != "QuatE" || dim % 4 == 0, "Model `QuatE` needs a dimension divisible by 4");
        model = _model;
        model_id = kg_model_id(model);
        num_epoch = _num_epoch;
        resume = _resume;
        sample_batch_size = _sample_batch_size;
        positive_reuse = _positive_reuse;
        log_frequency = std::max(1, _log_frequency);
        require(sample_batch_size >= 1 && positive_reuse >= 1, "invalid sample batch size / positive reuse");
        if (log_enabled())
            fprintf(stderr, "%s\n", info().c_str());
        if (!resume) {
            init_embeddings();
            // init_moments, core/solver.h:247-256 (the workers' relation moments stay where they are)
            std::fill(entity_m1_host.begin(), entity_m1_host.end(), 0.f);
            std::fill(entity_m2_host.begin(), entity_m2_host.end(), 0.f);
            batch_id = 0;
        }
        num_batch = int(batch_id + uint64_t(num_epoch) * graph->num_edge / batch_size);
        prepare_sampling();
        load_blocks();
        if (capture_negatives)
            d_negatives_out.allocate(uint64_t(chunk_batches) * batch_size * num_negative * 4);
        stat_positive = stat_kernel_seconds = stat_train_seconds = stat_sample_seconds = 0;
        stat_launches = 0;
        training = true;
        step_in_episode = 0;
        fill_pool(pool_id ^ 1);
    }

    // WorkerMixin::train for one block, core/solver.h:1511-1557, around the KG kernel
    void train_block(int head, int tail, int first_batch, int batch_stride) {
        float *head_block = entity_slots[slot_of_part[head]].as<float>();
        float *tail_block = entity_slots[slot_of_part[tail]].as<float>();
        float *relation = d_relation_work.as<float>();
        gv_kg_matrices_t matrices;
        matrices.dim = dim;
        matrices.num_head = uint32_t(partitions[head].size());
        matrices.head = head_block;
        matrices.tail = tail_block;
        matrices.relation = relation;
        matrices.head_m1 = num_state >= 2 ? head_block + block_floats : nullptr;
        matrices.tail_m1 = num_state >= 2 ? tail_block + block_floats : nullptr;
        matrices.relation_m1 = num_state >= 2 ? relation + relation_floats : nullptr;
        matrices.head_m2 = num_state >= 3 ? head_block + 2 * block_floats : nullptr;
        matrices.tail_m2 = num_state >= 3 ? tail_block + 2 * block_floats : nullptr;
        matrices.relation_m2 = num_state >= 3 ? relation + 2 * relation_floats : nullptr;
        gv_device_optimizer_t device_optimizer = {optimizer.desc.type, optimizer.desc.weight_decay, optimizer.desc.a,
                                                  optimizer.desc.b, optimizer.desc.epsilon};
        // build_negative_sampler, knowledge_graph.cuh:316-319: uniform over head rows then tail rows
        const uint32_t negative_count = uint32_t(partitions[head].size() + partitions[tail].size());
        const float margin_or_l3 = (model_id == GV_KG_TRANSE || model_id == GV_KG_ROTATE) ? margin : l3_regularization;
        const uint32_t *pool = pool_block(pool_id, head, tail);
        const uint64_t per_batch_random = uint64_t(batch_size) * num_negative * 2;
        std::vector<float> lr(episode_size), loss(episode_size);
        std::vector<cudaEvent_t> timers;
        int buffer = 0;
        for (int reuse = 0; reuse < positive_reuse; reuse++) {
            for (int j = 0; j < episode_size; j++)
                lr[j] = optimizer.lr_at(first_batch + (reuse * episode_size + j) * batch_stride, num_batch);
            GV_CHECK_CUDA(cudaMemcpyAsync(d_lr.ptr, lr.data(), episode_size * sizeof(float), cudaMemcpyHostToDevice,
                                          work_stream));
            GV_CHECK_CUDA(cudaMemsetAsync(d_loss.ptr, 0, episode_size * sizeof(float), work_stream));
            for (int j0 = 0; j0 < episode_size; j0 += chunk_batches, buffer = (buffer + 1) % kRandomBuffers) {
                const int count = std::min(chunk_batches, episode_size - j0);
                // negatives: curandGenerateUniformDouble(2 * B * k) per batch like train_batch (solver.h:1536);
                // the stream is positional, so one call per chunk of batches is the same stream
                GV_CHECK_CUDA(cudaStreamWaitEvent(random_stream, random_free[buffer], 0));
                GV_CHECK_ABI(gv_rng_generate(worker_generator, d_random[buffer].as<double>(),
                                             uint64_t(count) * per_batch_random, random_stream));
                GV_CHECK_CUDA(cudaEventRecord(random_ready[buffer], random_stream));
                GV_CHECK_CUDA(cudaStreamWaitEvent(work_stream, random_ready[buffer], 0));
                cudaEvent_t begin, end;
                GV_CHECK_CUDA(cudaEventCreate(&begin));
                GV_CHECK_CUDA(cudaEventCreate(&end));
                GV_CHECK_CUDA(cudaEventRecord(begin, work_stream));
                GV_CHECK_ABI(gv_cuda_kg_train_block(
                    &matrices, model_id, pool + uint64_t(j0) * batch_size * 3, uint64_t(count) * batch_size, num_negative,
                    nullptr, d_random[buffer].as<double>(), negative_count,
                    capture_negatives ? d_negatives_out.as<uint32_t>() : nullptr, &device_optimizer,
                    d_lr.as<float>() + j0, batch_size, relation_lr_multiplier, margin_or_l3, adversarial_temperature,
                    nullptr, d_loss.as<float>() + j0, train_num_groups, work_stream));
                stat_launches += 2;
                GV_CHECK_CUDA(cudaEventRecord(end, work_stream));
                GV_CHECK_CUDA(cudaEventRecord(random_free[buffer], work_stream));
                timers.push_back(begin);
                timers.push_back(end);
                if (capture_negatives && reuse == positive_reuse - 1 && j0 + count == episode_size) {
                    last_negatives.resize(size_t(batch_size) * num_negative);
                    GV_CHECK_CUDA(cudaMemcpyAsync(last_negatives.data(),
                                                  d_negatives_out.as<uint32_t>() +
                                                      size_t(count - 1) * batch_size * num_negative,
                                                  last_negatives.size() * 4, cudaMemcpyDeviceToHost, work_stream));
                }
            }
            GV_CHECK_CUDA(cudaMemcpyAsync(loss.data(), d_loss.ptr, episode_size * sizeof(float),
                                          cudaMemcpyDeviceToHost, work_stream));
            GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
            // at batch b the reference logs the loss buffer the worker's previous batch left (solver.h:1541-1549)
            for (int j = 0; j < episode_size; j++) {
                const int this_batch = first_batch + (reuse * episode_size + j) * batch_stride;
                if (this_batch % log_frequency == 0) {
                    logged_loss.push_back(previous_batch_loss);
                    if (log_enabled())
                        fprintf(stderr, "Batch id: %d / %d\nloss = %g\n", this_batch, num_batch, previous_batch_loss);
                }
                previous_batch_loss = loss[j] / batch_size;
            }
        }
        for (size_t i = 0; i < timers.size(); i += 2) {
            float ms = 0;
            GV_CHECK_CUDA(cudaEventElapsedTime(&ms, timers[i], timers[i + 1]));
            stat_kernel_seconds += ms * 1e-3;
            cudaEventDestroy(timers[i]);
            cudaEventDestroy(timers[i + 1]);
        }
        stat_positive += double(positive_reuse) * episode_size * batch_size;
    }

    // ---- the episode loop, core/solver.h:629-649, cut into schedule steps ----
    std::vector<std::vector<KgAssignment>> schedule;
    size_t step_in_episode = 0;
    std::thread sampler_thread;
    std::exception_ptr sampler_error;

    void finish_sampler() {
        if (sampler_thread.joinable())
            sampler_thread.join();
        if (sampler_error) {
            std::exception_ptr error = sampler_error;
            sampler_error = nullptr;
            std::rethrow_exception(error);
        }
    }

    bool train_step() {
        require(training, "train_begin() has not been called");
        GV_CHECK_CUDA(cudaSetDevice(device));
        if (step_in_episode == 0) {
            if (batch_id >= num_batch)
                return false;
            pool_id ^= 1;
            if (shuffle_partition)
                assignment_offset = (assignment_offset + 1) % num_partition;
            schedule = make_tied_schedule(num_partition, num_worker);
            for (auto &step : schedule)
                for (auto &assignment : step)
                    assignment.tail = (assignment.tail + assignment_offset) % num_partition;
            const int side = pool_id ^ 1;
            sampler_thread = std::thread([this, side]() {
                try {
                    fill_pool(side);
                } catch (...) {
                    sampler_error = std::current_exception();
                }
            });
        }
        try {
            const auto &step = schedule[step_in_episode];
            const int width = int(step.size());
            const double begin = now_seconds();
            if (num_worker > 1)
                move_blocks(step);
            // batch ids: the deterministic interleaving first + j * width of lock-step workers (the reference's
            // workers share an atomic counter, solver.h:1520)
            train_block(step[rank].head, step[rank].tail, batch_id + rank, width);
            batch_id += positive_reuse * episode_size * width;
            sync_relation();
            GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
            stat_train_seconds += now_seconds() - begin;
        } catch (...) {
            if (sampler_thread.joinable())
                sampler_thread.join();
            step_in_episode = 0;
            throw;
        }
        if (++step_in_episode == schedule.size()) {
            step_in_episode = 0;
            finish_sampler();
        }
        return true;
    }

    bool train_episode() {
        if (!train_step())
            return false;
        while (step_in_episode != 0)
            train_step();
        return true;
    }

    void train_end() {
        require(training, "train_begin() has not been called");
        GV_CHECK_CUDA(cudaSetDevice(device));
        while (step_in_episode != 0)
            train_step();
        write_back();
        training = false;
    }

    void train(const std::string &_model, int _num_epoch, bool _resume, float _relation_lr_multiplier, float _margin,
               float _l3_regularization, int _sample_batch_size, int _positive_reuse, float _adversarial_temperature,
               int _log_frequency) {
        train_begin(_model, _num_epoch, _resume, _relation_lr_multiplier, _margin, _l3_regularization,
                    _sample_batch_size, _positive_reuse, _adversarial_temperature, _log_frequency);
        try {
            while (train_episode())
                ;
        } catch (...) {
            training = false;
            throw;
        }
        train_end();
    }

    // SolverMixin::predict_numpy (core/solver.h:729-802) + gpu::knowledge_graph::predict: triplets (h, t, r)
    void predict(const uint32_t *triplets, uint64_t num, float *logits) {
        require(built, "The model must be built on a graph first");
        require(model_id >= 0, "predict() needs a trained model");
        GV_CHECK_CUDA(cudaSetDevice(device));
        const uint32_t n = graph->num_vertex(), r = graph->num_relation();
        std::vector<uint32_t> batch(num * 3);
        for (uint64_t i = 0; i < num; i++) {
            require(triplets[i * 3] < n && triplets[i * 3 + 1] < n && triplets[i * 3 + 2] < r,
                    "predict: entity / relation id out of range");
            batch[i * 3] = triplets[i * 3 + 2];      // device layout {relation, tail, head}
            batch[i * 3 + 1] = triplets[i * 3 + 1];
            batch[i * 3 + 2] = triplets[i * 3];
        }
        DeviceArray d_entity, d_relation, d_batch, d_logits;
        d_entity.upload(entity_host, work_stream);
        d_relation.upload(relation_host, work_stream);
        d_batch.upload(batch, work_stream);
        d_logits.allocate(std::max<uint64_t>(1, num) * sizeof(float));
        gv_kg_matrices_t matrices;
        memset(&matrices, 0, sizeof(matrices));
        matrices.dim = dim;
        matrices.num_head = n;
        matrices.head = matrices.tail = d_entity.as<float>();
        matrices.relation = d_relation.as<float>();
        GV_CHECK_ABI(gv_cuda_kg_predict(&matrices, model_id, d_batch.as<uint32_t>(), num, margin, d_logits.as<float>(),
                                        work_stream));
        GV_CHECK_CUDA(cudaMemcpyAsync(logits, d_logits.ptr, num * sizeof(float), cudaMemcpyDeviceToHost, work_stream));
        GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
    }

    void clear() {
        require(!training, "clear() during training");
        cudaSetDevice(device);
        entity_slots.clear();
        free_slots.clear();
        partition_ids.clear();
        pool_arena.release();
        for (int side = 0; side < 2; side++)
            pool_pointers[side].release();
        for (auto *a : {&d_relation_global, &d_relation_work, &d_relation_delta, &d_edge_h, &d_edge_t, &d_edge_r,
                        &d_edge_prob, &d_edge_alias, &d_locations, &d_sampler_random, &d_chains, &d_relations, &d_fill,
                        &d_last_walk, &d_fill_scratch, &d_random[0], &d_random[1], &d_random[2], &d_random[3], &d_lr,
                        &d_loss, &d_negatives_out})
            a->release();
        std::vector<float>().swap(entity_m1_host);
        std::vector<float>().swap(entity_m2_host);
        partitions.clear();
        sampling_ready = false;
        built = false;
    }

    // SolverMixin::info (core/solver.h:468-516) with the overrides of knowledge_graph.cuh:599-629
    std::string info() const {
        auto yes_no = [](bool x) { return x ? "yes" : "no"; };
        std::stringstream ss;
        ss << "KnowledgeGraphSolver<" << dim << ", float32, uint32>" << std::endl;
        ss << "----------------- Resource -----------------" << std::endl;
        ss << "#worker: " << num_worker << ", #sampler: " << num_sampler << ", #partition: " << num_partition
           << std::endl;
        ss << "tied weights: yes, episode size: " << episode_size << std::endl;
        ss << "gpu memory limit: " << size_string(gpu_memory_limit) << std::endl;
        ss << "gpu memory cost: " << size_string(gpu_memory_cost) << std::endl;
        ss << "----------------- Sampling -----------------" << std::endl;
        ss << "positive sample batch size: " << sample_batch_size << std::endl;
        ss << "#negative: " << num_negative << std::endl;
        ss << "----------------- Training -----------------" << std::endl;
        ss << "model: " << model << std::endl;
        ss << optimizer.info() << std::endl;
        ss << "#epoch: " << num_epoch << ", batch size: " << batch_size << std::endl;
        ss << "resume: " << yes_no(resume) << ", relation lr multiplier: " << relation_lr_multiplier << std::endl;
        if (model == "TransE" || model == "RotatE")
            ss << "margin: " << margin << ", positive reuse: " << positive_reuse << std::endl;
        if (model == "DistMult" || model == "ComplEx" || model == "SimplE" || model == "QuatE")
            ss << "l3 regularization: " << l3_regularization << ", positive reuse: " << positive_reuse << std::endl;
        ss << "adversarial temperature: " << adversarial_temperature;
        return ss.str();
    }

    std::string attributes() const {
        std::stringstream ss;
        ss.precision(9);
        ss << "dim=" << dim << "\nnum_partition=" << num_partition << "\nnum_negative=" << num_negative
           << "\nsample_batch_size=" << sample_batch_size << "\nnegative_sample_exponent=0"
           << "\nmodel=" << model << "\nnum_epoch=" << num_epoch << "\nresume=" << int(resume)
           << "\nrelation_lr_multiplier=" << relation_lr_multiplier << "\nepisode_size=" << episode_size
           << "\nbatch_size=" << batch_size << "\nmargin=" << margin << "\nl3_regularization=" << l3_regularization
           << "\nadversarial_temperature=" << adversarial_temperature << "\npositive_reuse=" << positive_reuse
           << "\nlog_frequency=" << log_frequency << "\nnum_worker=" << num_worker << "\nnum_sampler=" << num_sampler
           << "\ngpu_memory_limit=" << gpu_memory_limit << "\ngpu_memory_cost=" << gpu_memory_cost
           << "\nnum_batch=" << num_batch << "\nbatch_id=" << batch_id << "\npool_id=" << pool_id
           << "\npartition_size=" << partition_size << "\nrank=" << rank
           << "\nassignment_offset=" << assignment_offset << "\nshuffle_partition=" << int(shuffle_partition)
           << "\noptimizer_type=" << optimizer.type_name() << "\noptimizer_lr=" << optimizer.init_lr
           << "\noptimizer_weight_decay=" << optimizer.desc.weight_decay << "\n";
        return ss.str();
    }
};

}  // namespace gv

// =============================================================================
// C ABI
// =============================================================================
using gv::KgSolver;

struct gv_kg_solver {
    std::unique_ptr<KgSolver> solver;
};

#define GV_TRY try {
#define GV_CATCH(ret)                  \
    }                                  \
    catch (const std::exception &e) {  \
        gv::set_error(e.what());       \
        return ret;                    \
    }

static int copy_text(const std::string &text, char *buffer, size_t capacity) {
    if (buffer && capacity) {
        strncpy(buffer, text.c_str(), capacity - 1);
        buffer[capacity - 1] = 0;
    }
    return int(text.size());
}

extern "C" {

gv_kg_solver_t *gv_kg_solver_create(int dim, const int *device_ids, int num_device, int num_sampler_per_worker,
                                    uint64_t gpu_memory_limit, int rank, int world_size) {
    GV_TRY
    std::unique_ptr<KgSolver> solver(
        new KgSolver(dim, device_ids, num_device, num_sampler_per_worker, gpu_memory_limit, rank, world_size));
    gv_kg_solver *handle = new gv_kg_solver();
    handle->solver = std::move(solver);
    return handle;
    GV_CATCH(nullptr)
}

void gv_kg_solver_destroy(gv_kg_solver_t *solver) {
    delete solver;
}

int gv_kg_solver_set_exchange(gv_kg_solver_t *solver, gv_exchange_fn fn, void *ctx) {
    solver->solver->exchange_fn = fn;
    solver->solver->exchange_ctx = ctx;
    return 0;
}

int gv_kg_solver_set_allreduce(gv_kg_solver_t *solver, gv_allreduce_fn fn, void *ctx) {
    solver->solver->allreduce_fn = fn;
    solver->solver->allreduce_ctx = ctx;
    return 0;
}

int gv_kg_solver_set_option(gv_kg_solver_t *solver, const char *name, int value) {
    GV_TRY
    const std::string option(name);
    if (option == "capture_negatives")
        solver->solver->capture_negatives = value != 0;
    else if (option == "train_num_groups")
        solver->solver->train_num_groups = value;
    else if (option == "shuffle_partition")
        solver->solver->shuffle_override = value;
    else
        throw std::runtime_error("unknown option `" + option + "`");
    return 0;
    GV_CATCH(-1)
}

int gv_kg_solver_build(gv_kg_solver_t *solver, gv_kgraph_t *graph, const gv_optimizer_t *optimizer, int num_partition,
                       int num_negative, int batch_size, int episode_size) {
    GV_TRY
    solver->solver->build(&gv_kgraph_ref(graph), optimizer, num_partition, num_negative, batch_size, episode_size);
    return 0;
    GV_CATCH(-1)
}

int gv_kg_solver_train(gv_kg_solver_t *solver, const char *model, int num_epoch, int resume,
                       float relation_lr_multiplier, float margin, float l3_regularization, int sample_batch_size,
                       int positive_reuse, float adversarial_temperature, int log_frequency) {
    GV_TRY
    solver->solver->train(model, num_epoch, resume != 0, relation_lr_multiplier, margin, l3_regularization,
                          sample_batch_size, positive_reuse, adversarial_temperature, log_frequency);
    return 0;
    GV_CATCH(-1)
}

int gv_kg_solver_train_begin(gv_kg_solver_t *solver, const char *model, int num_epoch, int resume,
                             float relation_lr_multiplier, float margin, float l3_regularization,
                             int sample_batch_size, int positive_reuse, float adversarial_temperature,
                             int log_frequency) {
    GV_TRY
    solver->solver->train_begin(model, num_epoch, resume != 0, relation_lr_multiplier, margin, l3_regularization,
                                sample_batch_size, positive_reuse, adversarial_temperature, log_frequency);
    return 0;
    GV_CATCH(-1)
}

int gv_kg_solver_train_episode(gv_kg_solver_t *solver) {
    GV_TRY
    return solver->solver->train_episode() ? 1 : 0;
    GV_CATCH(-1)
}

int gv_kg_solver_train_end(gv_kg_solver_t *solver) {
    GV_TRY
    solver->solver->train_end();
    return 0;
    GV_CATCH(-1)
}

int gv_kg_solver_predict(gv_kg_solver_t *solver, const uint32_t *triplets, uint64_t num, float *logits) {
    GV_TRY
    solver->solver->predict(triplets, num, logits);
    return 0;
    GV_CATCH(-1)
}

int gv_kg_solver_clear(gv_kg_solver_t *solver) {
    GV_TRY
    solver->solver->clear();
    return 0;
    GV_CATCH(-1)
}

float *gv_kg_solver_embeddings(gv_kg_solver_t *solver, int which, uint64_t *rows, int *dim) {
    KgSolver &s = *solver->solver;
    if (rows)
        *rows = which == 0 ? s.entity_host.size() / s.dim : s.relation_host.size() / s.dim;
    if (dim)
        *dim = s.dim;
    return which == 0 ? s.entity_host.data() : s.relation_host.data();
}

int gv_kg_solver_info(const gv_kg_solver_t *solver, char *buffer, size_t capacity) {
    return copy_text(solver->solver->info(), buffer, capacity);
}

int gv_kg_solver_attributes(const gv_kg_solver_t *solver, char *buffer, size_t capacity) {
    return copy_text(solver->solver->attributes(), buffer, capacity);
}

int gv_kg_solver_logged_loss(const gv_kg_solver_t *solver, float *out, int capacity) {
    const auto &loss = solver->solver->logged_loss;
    for (int i = 0; i < capacity && i < int(loss.size()); i++)
        out[i] = loss[i];
    return int(loss.size());
}

int gv_kg_solver_stats(const gv_kg_solver_t *solver, double *out, int capacity) {
    const KgSolver &s = *solver->solver;
    const double values[] = {s.stat_positive, s.stat_kernel_seconds, s.stat_train_seconds, s.stat_sample_seconds,
                             double(s.stat_launches.load())};
    for (int i = 0; i < capacity && i < 5; i++)
        out[i] = values[i];
    return 5;
}

int gv_kg_solver_locations(const gv_kg_solver_t *solver, uint32_t *part_of, uint32_t *local_of) {
    const auto &locations = solver->solver->locations;
    for (size_t v = 0; v < locations.size(); v++) {
        part_of[v] = locations[v].part;
        local_of[v] = locations[v].local;
    }
    return 0;
}

int64_t gv_kg_solver_pool(gv_kg_solver_t *solver, int pool, int head_partition, int tail_partition, uint32_t *out) {
    GV_TRY
    KgSolver &s = *solver->solver;
    if (!s.built || pool < 0 || pool > 1 || head_partition < 0 || head_partition >= s.num_partition ||
        tail_partition < 0 || tail_partition >= s.num_partition)
        throw std::runtime_error("gv_kg_solver_pool: index out of range");
    cudaSetDevice(s.device);
    if (out && cudaMemcpy(out, s.pool_block(pool, head_partition, tail_partition), s.pool_block_bytes(),
                          cudaMemcpyDeviceToHost) != cudaSuccess)
        throw std::runtime_error("gv_kg_solver_pool: copy failed");
    return int64_t(s.pool_size());
    GV_CATCH(-1)
}

int gv_kg_solver_last_negatives(gv_kg_solver_t *solver, uint32_t *out) {
    const auto &negatives = solver->solver->last_negatives;
    if (out && !negatives.empty())
        memcpy(out, negatives.data(), negatives.size() * sizeof(uint32_t));
    return int(negatives.size());
}

int gv_kg_schedule(int num_partition, int num_worker, int *out, int capacity) {
    GV_TRY
    const int min_partition = num_worker == 1 ? 1 : num_worker * 2;
    if (num_worker < 1 || num_partition < min_partition || num_partition % min_partition != 0)
        throw std::runtime_error("#partition must be a positive multiple of " + std::to_string(min_partition));
    const auto schedule = gv::make_tied_schedule(num_partition, num_worker);
    const int width = int(schedule[0].size());
    if (int(schedule.size()) * width * 2 > capacity)
        throw std::runtime_error("gv_kg_schedule: capacity too small");
    for (const auto &step : schedule)
        for (const auto &assignment : step) {
            *out++ = assignment.head;
            *out++ = assignment.tail;
        }
    return int(schedule.size());
    GV_CATCH(-1)
}

}  // extern "C"
