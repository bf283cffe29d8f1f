// =============================================================================
// gv_solver.cpp -- host runtime of the node-embedding solver (C++ over the CUDA C ABI).
//
// Mirrors graphvite::GraphSolver<dim, float, uint32> and the SolverMixin / SamplerMixin /
// WorkerMixin machinery it inherits (reference include/instance/graph.cuh:587-813,
// include/core/solver.h:100-1624), re-designed for B200:
//   * every embedding block, both sample pools, the graph CSR and all alias tables are
//     RESIDENT in HBM; host memory only holds the matrices behind the numpy views;
//   * the sampler "threads" are cuRAND streams consumed by device kernels (gv_sampler.cu);
//     their seeds, stream consumption and pool layout are the reference's, bit for bit;
//   * one launch of the train kernel consumes a whole chunk of batches (gv_train.cu);
//   * between sub-episodes vertex blocks move GPU-to-GPU through a caller-provided
//     exchange (NCCL P2P over NVLink) instead of D2H -> CPU scatter/gather -> H2D.
// One process drives one GPU; world_size processes form the reference's num_worker.
// =============================================================================
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <set>
#include <sstream>
#include <thread>

#include "gv_runtime.h"


namespace gv {

Mt19937 g_engine;  // core/solver.h:50: one engine per process (default seed 5489), never re-seedable from Python

struct Assignment {
    int head, tail;
};

// SolverMixin::get_schedule, core/solver.h:519-575 (GraphSolver: partitioned, weights not tied)
static std::vector<std::vector<Assignment>> make_schedule(int num_partition, int num_worker) {
    std::vector<std::vector<Assignment>> schedule;
    if (num_partition == 1) {
        schedule.push_back({{0, 0}});
        return schedule;
    }
    for (int x = 0; x < num_partition; x += num_worker)
        for (int y = 0; y < num_partition; y += num_worker)
            for (int offset = 0; offset < num_worker; offset++) {
                std::vector<Assignment> step(num_worker);
                for (int i = 0; i < num_worker; i++)
                    step[i] = {x + (i + offset) % num_worker, y + i};
                schedule.push_back(step);
            }
    return schedule;
}

// Which rank holds which vertex (head) block.  Context blocks never move (tail t lives on rank
// t % W); inside one group of W head blocks every schedule step assigns the blocks to the ranks by
// a permutation, so each rank gives away at most one block and receives at most one per step.
struct Transfer {
    int need, source;        // block this rank trains next and the rank holding it now
    int give, destination;   // block another rank needs from this rank (-1: none) and that rank
};

struct BlockDirectory {
    int num_partition = 0, num_worker = 1;
    std::vector<int> owner;  // head block -> rank

    void reset(int P, int W) {
        num_partition = P;
        num_worker = W;
        owner.resize(P);
        for (int h = 0; h < P; h++)
            owner[h] = h % W;
    }
    Transfer plan(int rank, const std::vector<Assignment> &step) const {
        Transfer t = {step[rank].head, owner[step[rank].head], -1, -1};
        for (int i = 0; i < num_worker; i++)
            if (i != rank && owner[step[i].head] == rank) {
                t.give = step[i].head;
                t.destination = i;
            }
        return t;
    }
    void commit(const std::vector<Assignment> &step) {
        for (int i = 0; i < num_worker; i++)
            owner[step[i].head] = i;
    }
};

struct Solver {
    // ---- construction (SolverMixin ctor, core/solver.h:184-213) ----
    int dim, device, rank, world_size;
    int num_worker, num_sampler;
    uint64_t gpu_memory_limit, gpu_memory_cost = 0;
    std::vector<unsigned long long> sampler_seeds, worker_seeds;
    cudaStream_t work_stream = nullptr, sample_stream = nullptr, random_stream = nullptr;
    std::vector<gv_rng_t *> sampler_generators;  // one XORWOW stream per sampler (gv_rng.cu)
    std::vector<uint64_t> sampler_buffers;  // refill buffers each sampler's stream has consumed so far
    uint64_t walk_chunk = 1 << 18;           // walks per sampler launch (bounds chains + scratch)
    gv_rng_t *worker_generator = nullptr;
    DeviceArray d_rng_snapshot;
    gv_exchange_fn exchange_fn = nullptr;
    void *exchange_ctx = nullptr;
    gv_host_allgather_fn host_allgather_fn = nullptr;
    void *host_allgather_ctx = nullptr;

    // ---- build ----
    Graph *graph = nullptr;
    HostOptimizer optimizer;
    int num_partition = 0, num_negative = 1, batch_size = 100000, episode_size = 0;
    int num_group = 1;  // num_partition / num_worker
    std::vector<std::vector<uint32_t>> partitions;
    std::vector<gv_location_t> locations;
    uint32_t partition_size = 0;
    bool built = false;
    HostMatrix vertex_host, context_host;                               // numpy views
    HostMatrix vertex_m1_host, context_m1_host, vertex_m2_host, context_m2_host;

    // ---- train parameters (readonly attributes, bind.h:415-436) ----
    std::string model;
    int num_epoch = 0, augmentation_step = 0, random_walk_length = 40, random_walk_batch_size = 100;
    int shuffle_base = 0, positive_reuse = 1, log_frequency = 1000;
    float p = 1, q = 1, negative_sample_exponent = 0.75f, negative_weight = 5;
    bool resume = false;
    int batch_id = 0, num_batch = 0, pool_id = 0;
    bool training = false;

    // ---- device state ----
    size_t block_floats = 0;                       // partition_size * dim
    int num_state = 1;                             // 1 + num_moment matrices per block
    std::vector<DeviceArray> vertex_slots;         // each num_state * block_floats floats
    std::vector<int> slot_of_head;                 // head block -> local slot or -1
    BlockDirectory directory;                      // head block -> rank
    std::vector<int> free_slots;
    std::vector<DeviceArray> context_blocks;       // [num_group]
    std::vector<DeviceArray> negative_tables;      // [num_group] gv_alias_entry_t
    std::vector<uint32_t> negative_counts;
    std::vector<DeviceArray> partition_ids;        // [num_partition] global ids of each partition
    // sample pools: one arena [side][head][group] of pool_size pairs + the peer control region.
    // With partitioned sampling the arenas of all ranks are mapped into every rank (CUDA IPC).
    DeviceArray pool_arena;
    DeviceArray pool_pointers[2];                  // [P*P] block pointers (NULL: not reachable from here)
    bool peer_pools = false;
    std::vector<void *> peer_arenas;               // [W] mapped arenas (own entry = pool_arena.ptr)
    DeviceArray d_peer_controls, d_totals, d_bases;
    uint64_t peer_round = 0;
    // device graph
    DeviceArray d_offsets, d_edge_u, d_edge_v, d_edge_prob, d_edge_alias, d_vertex_tables, d_locations;
    DeviceArray d_edge_tables, d_table_offsets;  // node2vec: one alias table per directed edge
    // With partitioned sampling the tables are split over the ranks by entry count (d_edge_tables is then this rank's
    // shard) and every rank maps the others' shards with CUDA IPC: Sigma deg^2 x 8 B does not fit one GPU beyond
    // small graphs (209 GB on the Youtube-shaped graph), but 1 / W of it does.
    bool staged_scatter = false;  // partitioned sampling: stage + forward the pairs of peer-owned blocks
    DeviceArray d_stage, d_stage_offsets, d_remote_blocks;
    gv_table_shards_t table_shards;
    std::vector<void *> peer_tables;  // [W] mapped shards of the other ranks (own entry = nullptr)
    gv_device_graph_t device_graph;
    bool sampling_ready = false;
    int sample_mode = 0, tables_mode = -1;
    float tables_p = 0, tables_q = 0;
    // sampler scratch
    DeviceArray d_sampler_random, d_chains, d_fill, d_last_walk, d_fill_scratch;
    // worker scratch
    // negatives' randoms are generated kRandomBuffers chunks ahead on a high-priority stream, so the
    // (tiny) generator kernel slips into SM slots freed between train launches instead of delaying one
    static const int kRandomBuffers = 4;
    DeviceArray d_random[kRandomBuffers], d_lr, d_loss, d_negatives_out;
    DeviceArray d_random_step;  // the worker's randoms of a whole sub-episode, when they fit (see train_block)
    DeviceArray d_negatives_step;  // ... and the negatives drawn from them
    cudaEvent_t step_timer[2] = {nullptr, nullptr};
    cudaEvent_t random_ready[kRandomBuffers] = {}, random_free[kRandomBuffers] = {};
    int chunk_batches = 1;
    bool capture_negatives = false;
    int train_num_warps = 0;  // 0 = persistent grid; 1 = single warp (sequential, reproducible; tests)
    std::vector<uint32_t> last_negatives;
    std::vector<float> logged_loss;
    float previous_batch_loss = 0;  // mean loss of the batch trained before (what the reference logs)
    // stats
    double stat_positive = 0, stat_kernel_seconds = 0, stat_train_seconds = 0, stat_sample_seconds = 0;
    std::atomic<unsigned long long> stat_launches{0};
    cudaEvent_t timer_begin = nullptr, timer_end = nullptr;

    Solver(int _dim, const int *device_ids, int num_device, int num_sampler_per_worker, uint64_t memory_limit,
           int _rank, int _world_size)
        : dim(_dim), rank(_rank), world_size(_world_size), gpu_memory_limit(memory_limit) {
        static const std::set<int> dims = {32, 64, 96, 128, 256, 512};  // src/graphvite.cu:52-59
        require(dims.count(dim) == 1, "unsupported embedding dimension " + std::to_string(dim));
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
        // The reference defaults to hardware_concurrency / num_worker - 1 CPU sampler threads
        // (core/solver.h:193-195).  Samplers are cuRAND streams here, not CPU threads: auto = 1.
        if (num_sampler_per_worker == 0)
            num_sampler_per_worker = 1;
        require(num_sampler_per_worker > 0, "num_sampler_per_worker must be positive");
        num_sampler = num_sampler_per_worker * num_worker;
        GV_CHECK_CUDA(cudaSetDevice(device));
        if (gpu_memory_limit == 0) {
            size_t free_bytes = 0, total_bytes = 0;
            GV_CHECK_CUDA(cudaMemGetInfo(&free_bytes, &total_bytes));
            gpu_memory_limit = free_bytes;  // the reference caps at 32 GiB (V100 era), core/solver.h:199-207
        }
        // seeds: samplers first, then workers, from the process-wide engine (core/solver.h:208-212,
        // :950-952, :1247-1249).  Every rank draws the full sequence and keeps its own worker seed.
        std::uniform_int_distribution<unsigned long long> random_seed(0, ULLONG_MAX);
        for (int i = 0; i < num_sampler; i++)
            sampler_seeds.push_back(random_seed(g_engine));
        for (int i = 0; i < num_worker; i++)
            worker_seeds.push_back(random_seed(g_engine));
        GV_CHECK_CUDA(cudaStreamCreateWithFlags(&work_stream, cudaStreamNonBlocking));
        {
            // sampler kernels are short: let them run ahead of the queued train launches
            int least = 0, greatest = 0;
            GV_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            GV_CHECK_CUDA(cudaStreamCreateWithPriority(&sample_stream, cudaStreamNonBlocking, greatest));
            GV_CHECK_CUDA(cudaStreamCreateWithPriority(&random_stream, cudaStreamNonBlocking, greatest));
        }

        // the reference's cuRAND XORWOW streams (solver.h:950-953, 1247-1250), from our own generator
        for (int i = 0; i < num_sampler; i++) {
            gv_rng_t *generator = gv_rng_create(sampler_seeds[i], sample_stream);
            require(generator != nullptr, gv_last_error());
            sampler_generators.push_back(generator);
            sampler_buffers.push_back(0);
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

    ~Solver() {
        cudaSetDevice(device);
        if (sampler_thread.joinable())
            sampler_thread.join();
        close_peers();
        unpin_host();
        for (auto g : sampler_generators)
            gv_rng_destroy(g);
        gv_rng_destroy(worker_generator);
        for (cudaEvent_t event : step_timer)
            if (event)
                cudaEventDestroy(event);
        for (int i = 0; i < kRandomBuffers; i++) {
            if (random_ready[i])
                cudaEventDestroy(random_ready[i]);
            if (random_free[i])
                cudaEventDestroy(random_free[i]);
        }
        if (sampler_thread.joinable())
            sampler_thread.join();
        if (work_stream)
            cudaStreamDestroy(work_stream);
        if (sample_stream)
            cudaStreamDestroy(sample_stream);
        if (random_stream)
            cudaStreamDestroy(random_stream);
    }

    StagedUploader uploader;  // pageable graph arrays -> device
    std::vector<void *> pinned_host;
    void unpin_host() {
        for (void *pointer : pinned_host)
            cudaHostUnregister(pointer);
        pinned_host.clear();
    }

    int num_moment() const { return optimizer.num_moment(); }
    uint64_t pool_size() const { return uint64_t(episode_size) * batch_size; }
    bool owns_tail(int tail) const { return tail % num_worker == rank; }
    uint64_t pool_block_bytes() const { return pool_size() * 2 * sizeof(uint32_t); }
    uint64_t pool_block_offset(int side, int head, int group) const {
        return ((uint64_t(side) * num_partition + head) * num_group + group) * pool_block_bytes();
    }
    uint64_t control_offset() const { return pool_block_offset(2, 0, 0); }
    uint32_t *pool_block(int side, int head, int group) const {
        return reinterpret_cast<uint32_t *>(static_cast<char *>(pool_arena.ptr) + pool_block_offset(side, head, group));
    }
    void close_peer_tables() {
        for (void *mapped : peer_tables)
            if (mapped)
                cudaIpcCloseMemHandle(mapped);
        if (!peer_tables.empty()) {
            peer_tables.clear();
            sampling_ready = false;  // the shard table referenced the peers: rebuild before sampling again
        }
    }

    // Free this rank's table shard.  If the peers map it, every importer has to close it first (CUDA IPC rule):
    // all ranks get here together (prepare_sampling is collective), close their imports and meet at a host barrier.
    void release_table_shard() {
        if (!peer_tables.empty()) {
            close_peer_tables();
            int token = 0, all[256] = {0};
            require(host_allgather_fn && host_allgather_fn(&token, all, sizeof(int), host_allgather_ctx) == 0,
                    "host barrier before freeing the node2vec table shard failed");
        }
        d_edge_tables.release();
        d_table_offsets.release();
    }

    void close_peers() {
        close_peer_tables();
        for (int r = 0; r < int(peer_arenas.size()); r++)
            if (r != rank && peer_arenas[r])
                cudaIpcCloseMemHandle(peer_arenas[r]);
        peer_arenas.clear();
        if (peer_pools)
            built = false;  // the pool pointer tables referenced the peers: build() again before training
        peer_pools = false;
    }

    // bytes this rank keeps resident for a given partition count (our memory model; the
    // reference's gpu_memory_demand, core/solver.h:829-866, budgets ONE cached block pair)
    uint64_t memory_demand(int P, int episode) const {
        const uint64_t rows = (graph->num_vertex() + P - 1) / P;
        const uint64_t block = rows * dim * sizeof(float) * (1 + optimizer.num_moment());
        const int groups = P / num_worker;
        const int slots = num_worker > 1 ? groups + 1 : P;
        uint64_t demand = block * (slots + groups);
        demand += uint64_t(2) * P * groups * episode * batch_size * 8;             // both sample pools
        demand += uint64_t(graph->log_u.size()) * (4 + 4 + 4 + 8 + 8);              // CSR + edge/vertex tables
        demand += uint64_t(graph->num_vertex()) * (8 + 8 + 8);                     // offsets, locations, negatives
        demand += uint64_t(kSpanBuffers) * kRandBatchSize * 8;                     // samplers' refill buffers
        demand += uint64_t(kRandomBuffers) * 16 * batch_size * std::max(1, num_negative) * 16;  // negatives' randoms
        demand += uint64_t(2) * 256 * 1024 * 1024;                                 // walk chains + fill scratch
        demand += uint64_t(graph->num_vertex()) * dim * sizeof(float);             // staging for load / write-back
        return demand;
    }

    // ---- SolverMixin::build, core/solver.h:287-466 ----
    void build(Graph *_graph, const gv_optimizer_t *_optimizer, int _num_partition, int _num_negative,
               int _batch_size, int _episode_size) {
        require(!training, "build() during training");
        GV_CHECK_CUDA(cudaSetDevice(device));
        graph = _graph;
        require(graph->num_vertex() > 0, "The graph is empty");
        optimizer.desc = *_optimizer;
        if (optimizer.desc.type < 0) {  // "Default": SGD(0.025, 5e-3) unless a learning rate was given
            const float lr = optimizer.desc.lr;
            optimizer.desc.type = GV_OPT_SGD;  // GraphSolver::get_default_optimizer, graph.cuh:634-636
            optimizer.desc.lr = lr > 0 ? lr : 0.025f;
            optimizer.desc.weight_decay = 5e-3f;
            optimizer.desc.schedule = GV_SCHEDULE_LINEAR;
        }
        optimizer.init_lr = optimizer.desc.lr;
        num_negative = _num_negative;
        batch_size = _batch_size;
        require(num_negative >= 0 && batch_size > 0, "invalid num_negative / batch_size");
        if (batch_size < kMinBatchSize && log_enabled())
            fprintf(stderr, "It is recommended to a minimum batch size of %d, but %d is specified\n", kMinBatchSize,
                    batch_size);
        batch_id = 0;
        graph->flatten();

        const int min_partition = num_worker;  // get_min_partition, core/solver.h:266-277 (weights never tied)
        auto auto_episode = [&](int P) {       // core/solver.h:426-436
            int expected = float(uint64_t(graph->num_vertex()) * kSamplePerVertex) / P / batch_size;
            expected = std::max(expected, 1);
            if (P == 1)
                expected = std::max(expected, kMinEpisodeSample / batch_size);
            return expected;
        };
        num_partition = _num_partition;
        if (num_partition == 0) {
            // With one worker every block stays resident in HBM (no host paging, section 3 of DESIGN.md), so more
            // partitions never lower the demand of a single GPU: P = 1, and only the pools can shrink below.
            const int last = num_worker == 1 ? min_partition + 1 : kMaxPartition;
            for (num_partition = min_partition; num_partition < last; num_partition += min_partition)
                if (memory_demand(num_partition, _episode_size ? _episode_size : auto_episode(num_partition)) <
                    gpu_memory_limit)
                    break;
            if (num_worker == 1)
                num_partition = min_partition;
        } else
            require(num_partition >= min_partition,
                    "#partition should be no less than " + std::to_string(min_partition));
        require(num_partition % num_worker == 0, "#partition must be a multiple of #worker");
        num_group = num_partition / num_worker;
        episode_size = _episode_size ? _episode_size : auto_episode(num_partition);
        // the pools are the most elastic part: halve the episode until everything fits (solver.h:437-462)
        while (episode_size > 1 && memory_demand(num_partition, episode_size) >= gpu_memory_limit)
            episode_size /= 2;
        gpu_memory_cost = memory_demand(num_partition, episode_size);
        require(gpu_memory_cost < gpu_memory_limit,
                "Can't satisfy the specified GPU memory limit: " + std::to_string(gpu_memory_cost >> 20) + " MiB needed (" +
                    std::to_string(num_partition) + " partition(s), every block resident on its GPU), " +
                    std::to_string(gpu_memory_limit >> 20) + " MiB allowed" +
                    (num_worker == 1 ? "; embeddings larger than one GPU need more GPUs (one process per GPU)" : ""));

        partitions = partition_vertices(graph->vertex_weights, num_partition);
        partition_size = 0;
        for (auto &part : partitions)
            partition_size = std::max<uint32_t>(partition_size, part.size());
        locations.resize(graph->num_vertex());
        for (int i = 0; i < num_partition; i++)
            for (uint32_t j = 0; j < partitions[i].size(); j++)
                locations[partitions[i][j]] = {uint32_t(i), j};

        const size_t total = size_t(graph->num_vertex()) * dim;
        unpin_host();
        vertex_host.assign(total, 0.f);
        context_host.assign(total, 0.f);
        // page-lock the matrices behind the numpy views: upload and write-back run at PCIe speed
        for (auto *m : {&vertex_host, &context_host})
            if (cudaHostRegister(m->data(), m->size() * sizeof(float), cudaHostRegisterDefault) == cudaSuccess)
                pinned_host.push_back(m->data());
            else
                cudaGetLastError();
        const int nm = num_moment();
        vertex_m1_host.assign(nm >= 1 ? total : 0, 0.f);
        context_m1_host.assign(nm >= 1 ? total : 0, 0.f);
        vertex_m2_host.assign(nm >= 2 ? total : 0, 0.f);
        context_m2_host.assign(nm >= 2 ? total : 0, 0.f);

        // ---- device residency ----
        block_floats = size_t(partition_size) * dim;
        num_state = 1 + nm;
        const int num_slot = num_worker > 1 ? num_group + 1 : num_partition;
        vertex_slots = std::vector<DeviceArray>(num_slot);
        for (auto &slot : vertex_slots)
            slot.allocate(block_floats * num_state * sizeof(float));
        context_blocks = std::vector<DeviceArray>(num_group);
        for (auto &block : context_blocks)
            block.allocate(block_floats * num_state * sizeof(float));
        partition_ids = std::vector<DeviceArray>(num_partition);
        for (int i = 0; i < num_partition; i++)
            partition_ids[i].upload(partitions[i], work_stream);
        d_locations.upload(locations, work_stream);
        close_peers();
        const size_t control_bytes = gv_cuda_peer_control_bytes(num_worker, num_partition);
        pool_arena.allocate(control_offset() + control_bytes);
        GV_CHECK_CUDA(cudaMemsetAsync(static_cast<char *>(pool_arena.ptr) + control_offset(), 0, control_bytes,
                                      work_stream));
        GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
        peer_arenas.assign(num_worker, nullptr);
        peer_arenas[rank] = pool_arena.ptr;
        if (num_worker > 1 && host_allgather_fn && !getenv("GV_REPLICATED_SAMPLING")) {
            // trade CUDA IPC handles of the arenas: samplers are then partitioned over the ranks and
            // scatter their pairs straight into the owners' pools (peer stores over NVLink)
            cudaIpcMemHandle_t mine;
            std::vector<cudaIpcMemHandle_t> all(num_worker);
            bool ok = cudaIpcGetMemHandle(&mine, pool_arena.ptr) == cudaSuccess;
            if (!ok)
                memset(&mine, 0, sizeof(mine));
            require(host_allgather_fn(&mine, all.data(), sizeof(mine), host_allgather_ctx) == 0,
                    "host all-gather of the IPC handles failed");
            for (int r = 0; r < num_worker && ok; r++)
                if (r != rank)
                    ok = cudaIpcOpenMemHandle(&peer_arenas[r], all[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            // everybody must agree on the mode
            int mine_ok = ok, all_ok[256] = {0};
            require(num_worker <= 256, "too many workers");
            require(host_allgather_fn(&mine_ok, all_ok, sizeof(int), host_allgather_ctx) == 0,
                    "host all-gather failed");
            for (int r = 0; r < num_worker; r++)
                ok = ok && all_ok[r];
            cudaGetLastError();
            if (ok)
                peer_pools = true;
            else {
                for (int r = 0; r < num_worker; r++)
                    if (r != rank && peer_arenas[r]) {
                        cudaIpcCloseMemHandle(peer_arenas[r]);
                        peer_arenas[r] = nullptr;
                    }
                if (log_enabled())
                    fprintf(stderr, "CUDA IPC unavailable: falling back to replicated sampling\n");
            }
        }
        for (int side = 0; side < 2; side++) {
            std::vector<uint32_t *> pointers(size_t(num_partition) * num_partition, nullptr);
            for (int h = 0; h < num_partition; h++)
                for (int t = 0; t < num_partition; t++) {
                    const int owner = t % num_worker;
                    if (peer_arenas[owner])
                        pointers[size_t(h) * num_partition + t] = reinterpret_cast<uint32_t *>(
                            static_cast<char *>(peer_arenas[owner]) + pool_block_offset(side, h, t / num_worker));
                }
            pool_pointers[side].upload(pointers, work_stream);
        }
        if (peer_pools) {
            std::vector<unsigned long long *> controls(num_worker);
            for (int r = 0; r < num_worker; r++)
                controls[r] = reinterpret_cast<unsigned long long *>(static_cast<char *>(peer_arenas[r]) +
                                                                     control_offset());
            d_peer_controls.upload(controls, work_stream);
            d_totals.allocate(size_t(num_partition) * num_partition * sizeof(unsigned long long));
            d_bases.allocate(size_t(num_partition) * num_partition * sizeof(unsigned long long));
            peer_round = 0;
            // pairs of blocks owned by other ranks are staged locally and forwarded with coalesced peer stores
            // (gv_cuda_fill_scatter_staged); GV_DIRECT_PEER_SCATTER=1 keeps the direct 8-byte peer stores
            staged_scatter = !getenv("GV_DIRECT_PEER_SCATTER");
            std::vector<unsigned char> remote(size_t(num_partition) * num_partition, 0);
            for (int h = 0; h < num_partition; h++)
                for (int t = 0; t < num_partition; t++)
                    remote[size_t(h) * num_partition + t] = t % num_worker != rank;
            d_remote_blocks.upload(remote, work_stream);
            d_stage_offsets.allocate(remote.size() * sizeof(unsigned long long));
        }
        negative_tables = std::vector<DeviceArray>(num_group);
        negative_counts.assign(num_group, 0);
        // worker scratch
        const uint64_t per_batch_random = uint64_t(batch_size) * num_negative * 2 * sizeof(double);
        // one train launch consumes a chunk of batches: long enough to amortise the launch and the tail
        // (16 batches = 1.6e6 edges ~ 1 ms), short enough that the samplers' kernels get SMs in between
        chunk_batches = int(std::max<uint64_t>(1, std::min<uint64_t>(std::min(episode_size, 16), (uint64_t(192) << 20) /
                                                                                    std::max<uint64_t>(1, per_batch_random))));
        // SGD trains with the reference's geometry -- one warp per sample, ONE LAUNCH PER BATCH (the launch boundary
        // bounds the staleness of L1 copies exactly as in the reference, instance/graph.cuh:487); the persistent
        // kernels (other optimizers, or kernel_flags & 64) amortise their launch over a chunk of batches
        if (optimizer.desc.type == GV_OPT_SGD && !(gv_cuda_get_tunable("kernel_flags") & 64))
            chunk_batches = 1;
        if (getenv("GV_CHUNK_BATCHES"))  // experiment: launch granularity in batches
            chunk_batches = std::max(1, std::min(episode_size, atoi(getenv("GV_CHUNK_BATCHES"))));
        for (int i = 0; i < kRandomBuffers; i++)
            d_random[i].allocate(std::max<uint64_t>(16, per_batch_random * chunk_batches));
        d_lr.allocate(size_t(episode_size) * sizeof(float));
        d_loss.allocate(size_t(episode_size) * sizeof(float));
        // sampler scratch
        d_fill.allocate(size_t(num_partition) * num_partition * sizeof(unsigned long long));
        d_last_walk.allocate(sizeof(unsigned long long));
        sampling_ready = false;
        pool_id = 0;
        built = true;
        logged_loss.clear();
    }

    // batches per train launch (1 = the reference's launch granularity); the per-chunk random buffers follow
    void set_chunk_batches(int value) {
        require(built && !training, "chunk_batches can be set between build() and train()");
        GV_CHECK_CUDA(cudaSetDevice(device));
        chunk_batches = std::max(1, std::min(episode_size, value));
        const uint64_t per_batch_random = uint64_t(batch_size) * num_negative * 2 * sizeof(double);
        for (int i = 0; i < kRandomBuffers; i++)
            d_random[i].allocate(std::max<uint64_t>(16, per_batch_random * chunk_batches));
    }

    // ---- device graph + sampler tables (GraphSolver::get_sample_function, graph.cuh:680-721) ----
    void prepare_sampling() {
        require(!graph->has_dead_end() || augmentation_step == 1,
                "graphs with dead ends (vertices without out-edges) are not supported by the device walker yet");
        if (augmentation_step == 1)
            sample_mode = 0;
        else if (model == "DeepWalk" || model == "LINE")
            sample_mode = 1;
        else
            sample_mode = 2;
        const size_t m = graph->edge_u.size();
        require(m > 0, "The graph has no edges");
        // The tables depend on the graph, the kind of walk and (node2vec) p, q only: a later train() call
        // on the same build() -- e.g. train, evaluate, train(resume=True) -- reuses what is resident.
        // (The reference rebuilds them in every call, instance/graph.cuh:680-721.)
        const bool reusable = sampling_ready && tables_mode == sample_mode && (sample_mode != 2 || (tables_p == p && tables_q == q));
        if (reusable) {
            allocate_sampler_scratch();
            return;
        }
        sampling_ready = false;
        PhaseTimer phase;  // GV_LOG=2
        // edge_table.build(graph->edge_weights), core/solver.h:255-256.  When every edge weighs the same (any
        // unweighted graph) AliasTable::build takes its trivial branch -- one probability, alias = identity -- and the
        // table (12 bytes per directed edge) is written by the device instead of built on the host and uploaded.
        float uniform_probability = 1;
        const bool uniform = graph->uniform_edge_table(uniform_probability);
        std::vector<float> edge_prob(uniform ? 0 : m);
        std::vector<uint64_t> edge_alias(uniform ? 0 : m);
        // Vose over all directed edges is sequential (0.1 s for 1e7 edges): overlap it with the uploads
        // and with the per-vertex tables below
        std::exception_ptr edge_error;
        std::thread edge_builder([&]() {
            try {
                if (!uniform)
                    build_alias<uint64_t>(graph->edge_w.data(), m, edge_prob.data(), edge_alias.data());
            } catch (...) {
                edge_error = std::current_exception();
            }
        });
        struct Joiner {
            std::thread &thread;
            ~Joiner() {
                if (thread.joinable())
                    thread.join();
            }
        } joiner{edge_builder};
        // the CSR: offsets and targets travel (through page-locked staging), the source column is expanded from the
        // offsets on the device
        uploader.upload(d_offsets, graph->offsets, sample_stream);
        uploader.upload(d_edge_v, graph->edge_v, sample_stream);
        d_edge_u.allocate(m * sizeof(uint32_t));
        GV_CHECK_ABI(gv_cuda_expand_sources(d_offsets.as<uint64_t>(), graph->num_vertex(), d_edge_u.as<uint32_t>(),
                                            sample_stream));
        stat_launches++;
        phase.mark("  CSR upload");
        device_graph.num_vertex = graph->num_vertex();
        device_graph.num_edge = m;
        device_graph.offsets = d_offsets.as<uint64_t>();
        device_graph.edge_u = d_edge_u.as<uint32_t>();
        device_graph.edge_v = d_edge_v.as<uint32_t>();
        device_graph.edge_prob = d_edge_prob.as<float>();
        device_graph.edge_alias = d_edge_alias.as<uint64_t>();
        device_graph.locations = d_locations.as<gv_location_t>();
        device_graph.vertex_tables = nullptr;
        if (sample_mode == 1) {
            // build_vertex_edge, graph.cuh:645-653: one alias table per vertex over its out-edges, laid out at the
            // vertex's CSR range; built on the device (thread per vertex, the reference's pairing order)
            DeviceArray d_weights, d_little, d_large;
            if (uniform) {
                d_weights.allocate(m * sizeof(float));
                GV_CHECK_ABI(gv_cuda_fill_float(d_weights.as<float>(), m, graph->edge_w[0], sample_stream));
                stat_launches++;
            } else
                uploader.upload(d_weights, graph->edge_w, sample_stream);
            d_vertex_tables.allocate(std::max<size_t>(m, 1) * sizeof(gv_alias_entry_t));
            d_little.allocate(std::max<size_t>(m, 1) * sizeof(uint32_t));
            d_large.allocate(std::max<size_t>(m, 1) * sizeof(uint32_t));
            GV_CHECK_ABI(gv_cuda_vertex_tables_build(&device_graph, d_weights.as<float>(),
                                                     d_vertex_tables.as<gv_alias_entry_t>(), d_little.as<uint32_t>(),
                                                     d_large.as<uint32_t>(), sample_stream));
            stat_launches++;
            GV_CHECK_CUDA(cudaStreamSynchronize(sample_stream));  // the scratch arrays go out of scope
            phase.mark("  per-vertex alias tables");
            device_graph.vertex_tables = d_vertex_tables.as<gv_alias_entry_t>();
        }
        edge_builder.join();
        phase.mark("  edge alias table (rest)");
        if (edge_error)
            std::rethrow_exception(edge_error);
        if (uniform) {
            d_edge_prob.allocate(m * sizeof(float));
            d_edge_alias.allocate(m * sizeof(uint64_t));
            GV_CHECK_ABI(gv_cuda_fill_float(d_edge_prob.as<float>(), m, uniform_probability, sample_stream));
            GV_CHECK_ABI(gv_cuda_fill_identity(d_edge_alias.as<uint64_t>(), m, sample_stream));
            stat_launches += 2;
        } else {
            uploader.upload(d_edge_prob, edge_prob, sample_stream);
            uploader.upload(d_edge_alias, edge_alias, sample_stream);
        }
        device_graph.edge_prob = d_edge_prob.as<float>();
        device_graph.edge_alias = d_edge_alias.as<uint64_t>();
        if (sample_mode == 2)
            build_node2vec_tables();
        else
            release_table_shard();
        tables_mode = sample_mode;
        tables_p = p;
        tables_q = q;
        allocate_sampler_scratch();
        sampling_ready = true;
    }

    // chains / histogram scratch / refill buffers of the samplers (depend on the walk length and P)
    void allocate_sampler_scratch() {
        const int L = sample_mode == 0 ? 1 : random_walk_length;
        // per launch: at most walk_chunk walks per rank (chains <= 256 MB, histogram scratch <= 256 MB)
        walk_chunk = std::min<uint64_t>(uint64_t(1) << 20, (uint64_t(256) << 20) / (uint64_t(L + 1) * sizeof(gv_location_t)));
        walk_chunk = std::min<uint64_t>(walk_chunk, (uint64_t(256) << 20) / (uint64_t(num_partition) * num_partition * 4));
        walk_chunk = std::max<uint64_t>(walk_chunk, 1024);
        d_chains.allocate(walk_chunk * (L + 1) * sizeof(gv_location_t));
        d_fill_scratch.allocate(gv_cuda_fill_scratch_bytes(uint32_t(walk_chunk), num_partition));
        if (peer_pools && staged_scatter)
            d_stage.allocate(gv_cuda_fill_staging_bytes(uint32_t(walk_chunk), L, sample_mode == 0 ? 1 : augmentation_step));
        d_sampler_random.allocate(size_t(kSpanBuffers) * kRandBatchSize * sizeof(double));
    }

    // GraphSolver::build_edge_edge, instance/graph.cuh:656-677, on the device: one alias table per
    // directed edge (u -> v) over the out-edges of v, Sigma deg^2 entries in one flat array.
    void build_node2vec_tables() {
        const size_t m = graph->edge_u.size();
        std::vector<unsigned long long> table_offsets(m + 1, 0);
        for (size_t e = 0; e < m; e++) {
            const uint32_t v = graph->edge_v[e];
            table_offsets[e + 1] = table_offsets[e] + (graph->offsets[v + 1] - graph->offsets[v]);
        }
        const unsigned long long total = table_offsets[m];
        // shards: rank r owns the tables of the edges [first_edge[r], first_edge[r + 1]), cut where the entry count
        // passes r / W of the total (one shard = everything unless the sampling is partitioned)
        release_table_shard();
        const int num_shard = peer_pools ? num_worker : 1;
        require(num_shard <= GV_MAX_TABLE_SHARDS, "too many workers for sharded node2vec tables");
        std::vector<size_t> first_edge(num_shard + 1, m);
        first_edge[0] = 0;
        for (int r = 1; r < num_shard; r++)
            first_edge[r] = std::lower_bound(table_offsets.begin(), table_offsets.end(), total / num_shard * r) -
                            table_offsets.begin();
        memset(&table_shards, 0, sizeof(table_shards));
        table_shards.num_shard = num_shard;
        for (int r = 0; r <= num_shard; r++)
            table_shards.first_entry[r] = table_offsets[first_edge[r]];
        const int mine = peer_pools ? rank : 0;
        const unsigned long long own = table_shards.first_entry[mine + 1] - table_shards.first_entry[mine];
        size_t free_bytes = 0, total_bytes = 0;
        GV_CHECK_CUDA(cudaMemGetInfo(&free_bytes, &total_bytes));
        const unsigned long long budget = std::min<unsigned long long>(std::max<unsigned long long>(own, 1), 1ull << 27);
        const unsigned long long needed = own * sizeof(gv_alias_entry_t) + budget * 8 + m * 8 + (m + 1) * 8;
        require(needed + (1ull << 30) < free_bytes,
                "node2vec needs " + std::to_string(needed >> 20) + " MiB of device memory for its per-edge alias tables "
                "(sum of squared degrees = " + std::to_string(total) + " entries over " + std::to_string(num_shard) +
                " GPU(s)), only " + std::to_string(free_bytes >> 20) + " MiB are free");
        // neighbour lists sorted inside every vertex's CSR range, for the membership test
        std::vector<uint32_t> sorted(graph->edge_v);
        for (uint32_t v = 0; v < graph->num_vertex(); v++)
            std::sort(sorted.begin() + graph->offsets[v], sorted.begin() + graph->offsets[v + 1]);
        DeviceArray d_sorted, d_weights, d_little, d_large;
        d_sorted.upload(sorted, sample_stream);
        d_weights.upload(graph->edge_w, sample_stream);
        d_table_offsets.upload(table_offsets, sample_stream);
        d_edge_tables.allocate(std::max<unsigned long long>(own, 1) * sizeof(gv_alias_entry_t));
        d_little.allocate(budget * sizeof(uint32_t));
        d_large.allocate(budget * sizeof(uint32_t));
        // the build kernel addresses tables[table_offsets[e]]: hand it the shard's virtual origin
        gv_alias_entry_t *origin = d_edge_tables.as<gv_alias_entry_t>() - table_shards.first_entry[mine];
        for (size_t first = first_edge[mine]; first < first_edge[mine + 1];) {
            size_t last = first;
            while (last < first_edge[mine + 1] && table_offsets[last + 1] - table_offsets[first] <= budget)
                last++;
            require(last > first, "internal error: node2vec table larger than the build batch");
            GV_CHECK_ABI(gv_cuda_node2vec_build(&device_graph, d_weights.as<float>(), d_sorted.as<uint32_t>(),
                                                d_table_offsets.as<unsigned long long>(), first,
                                                uint32_t(last - first), p, q, origin, d_little.as<uint32_t>(),
                                                d_large.as<uint32_t>(), sample_stream));
            stat_launches++;
            first = last;
        }
        GV_CHECK_CUDA(cudaStreamSynchronize(sample_stream));
        table_shards.shard[mine] = d_edge_tables.as<gv_alias_entry_t>();
        if (num_shard > 1) {
            // trade the shards' IPC handles; every rank must have finished building before anybody walks
            cudaIpcMemHandle_t handle;
            std::vector<cudaIpcMemHandle_t> all(num_worker);
            require(cudaIpcGetMemHandle(&handle, d_edge_tables.ptr) == cudaSuccess,
                    "cannot export the node2vec table shard (CUDA IPC)");
            require(host_allgather_fn(&handle, all.data(), sizeof(handle), host_allgather_ctx) == 0,
                    "host all-gather of the table handles failed");
            peer_tables.assign(num_worker, nullptr);
            for (int r = 0; r < num_worker; r++)
                if (r != rank) {
                    require(cudaIpcOpenMemHandle(&peer_tables[r], all[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess,
                            "cannot map a peer's node2vec table shard (CUDA IPC)");
                    table_shards.shard[r] = static_cast<const gv_alias_entry_t *>(peer_tables[r]);
                }
            if (rank == 0 && getenv("GV_LOG"))
                fprintf(stderr, "node2vec tables: %llu entries sharded over %d ranks (%llu on rank 0)\n", total,
                        num_shard, own);
        }
    }

    // ---- one sampler's share of a pool (SamplerMixin::sample / GraphSampler::sample_random_walk) ----
    void run_sampler(int sampler_id, int side, uint64_t start, uint64_t end) {
        const int L = sample_mode == 0 ? 1 : random_walk_length;
        const int aug = sample_mode == 0 ? 1 : augmentation_step;
        const int shuffle = sample_mode == 0 ? 1 : shuffle_base;
        // termination is checked once per batch of walks: random_walk_batch_size walks, or
        // sample_batch_size = L * walk_batch edges in edge mode (graph.cuh:791, solver.h:1026)
        const uint64_t walk_batch = sample_mode == 0 ? uint64_t(random_walk_length) * random_walk_batch_size
                                                     : uint64_t(random_walk_batch_size);
        const uint64_t walks_per_buffer = uint64_t(kRandBatchSize - 2 * L) / (2 * L) + 1;
        const int num_block = num_partition * num_partition;
        uint64_t pairs_per_walk = 0;
        for (int j = 0; j < L; j++)
            pairs_per_walk += std::min(aug, L - j);

        gv_fill_params_t params;
        params.num_partition = num_partition;
        params.walk_length = L;
        params.augmentation_step = aug;
        params.shuffle_base = shuffle;
        params.pool_size = pool_size();
        params.start = start;
        params.end = end;
        params.attributes = nullptr;
        const uint64_t slice = end - start;
        if (slice == 0)
            return;

        GV_CHECK_CUDA(cudaMemsetAsync(d_fill.ptr, 0, d_fill.bytes, sample_stream));
        GV_CHECK_CUDA(cudaMemsetAsync(d_last_walk.ptr, 0, sizeof(unsigned long long), sample_stream));
        GV_CHECK_ABI(gv_rng_save(sampler_generators[sampler_id], d_rng_snapshot.ptr, sample_stream));
        std::vector<unsigned long long> fill(num_block, 0);
        uint64_t buffers = 0, walks_done = 0;
        bool complete = false;
        unsigned long long last_walk = 0;
        const uint64_t span_capacity = d_sampler_random.bytes / (size_t(kRandBatchSize) * sizeof(double));
        while (!complete) {
            // walks still needed, judged from the emptiest block: the slowest block receives at most
            // 1 / num_block of the pairs, so 97 % of this estimate is certainly needed -- that many whole
            // refill buffers are generated and walked in one go; the remainder goes buffer by buffer
            uint64_t missing = 0;
            for (int b = 0; b < num_block; b++)
                missing = std::max<uint64_t>(missing, slice - std::min<uint64_t>(slice, fill[b]));
            const double estimate = double(missing) * num_block / pairs_per_walk;
            const uint64_t span = std::max<uint64_t>(1, std::min<uint64_t>(span_capacity,
                                                                          uint64_t(estimate * 0.97 / walks_per_buffer)));
            // refill: the next kRandBatchSize doubles of this sampler's stream per buffer
            // (solver.h:1015-1016,1028-1031), same call size as the reference
            GV_CHECK_ABI(gv_rng_generate(sampler_generators[sampler_id], d_sampler_random.as<double>(),
                                         span * uint64_t(kRandBatchSize), sample_stream));
            stat_launches++;
            buffers += span;
            const uint64_t in_span = span * walks_per_buffer;
            uint64_t done_in_span = 0;
            while (done_in_span < in_span && !complete) {
                uint64_t want = in_span - done_in_span;
                if (span == 1) {  // finishing: stop as soon as possible (checked per batch of walks)
                    missing = 0;
                    for (int b = 0; b < num_block; b++)
                        missing = std::max<uint64_t>(missing, slice - std::min<uint64_t>(slice, fill[b]));
                    want = uint64_t(double(missing) * num_block / pairs_per_walk * 1.02) + 2 * walk_batch;
                    want = (want + walk_batch - 1) / walk_batch * walk_batch;
                }
                const uint32_t count = uint32_t(std::min<uint64_t>(std::min<uint64_t>(want, walk_chunk),
                                                                   in_span - done_in_span));
                if (sample_mode == 2)
                    GV_CHECK_ABI(gv_cuda_biased_walk_sharded(&device_graph, &table_shards,
                                                             d_table_offsets.as<unsigned long long>(),
                                                             d_sampler_random.as<double>(), count, L, done_in_span,
                                                             uint32_t(walks_per_buffer), kRandBatchSize,
                                                             d_chains.as<gv_location_t>(), sample_stream));
                else
                    GV_CHECK_ABI(gv_cuda_random_walk(&device_graph, d_sampler_random.as<double>(), count, L,
                                                     done_in_span, uint32_t(walks_per_buffer), kRandBatchSize,
                                                     d_chains.as<gv_location_t>(), sample_stream));
                if (!peer_pools) {
                    GV_CHECK_ABI(gv_cuda_fill_pool(&params, d_chains.as<gv_location_t>(), count, walks_done,
                                                   pool_pointers[side].as<uint32_t *>(),
                                                   d_fill.as<unsigned long long>(),
                                                   d_last_walk.as<unsigned long long>(), d_fill_scratch.ptr,
                                                   sample_stream));
                    stat_launches += num_partition == 1 ? 3 : 4;
                } else {
                    // some blocks of the slice live in a peer's pool: the same stable partition, but the pairs of
                    // peer-owned blocks are staged locally and forwarded with coalesced stores over NVLink
                    GV_CHECK_ABI(gv_cuda_fill_count(&params, d_chains.as<gv_location_t>(), count,
                                                    d_fill_scratch.ptr, d_totals.as<unsigned long long>(),
                                                    sample_stream));
                    GV_CHECK_ABI(gv_cuda_fill_advance(num_block, d_totals.as<unsigned long long>(),
                                                      d_fill.as<unsigned long long>(),
                                                      d_bases.as<unsigned long long>(), sample_stream));
                    GV_CHECK_ABI(gv_cuda_fill_scatter_staged(
                        &params, d_chains.as<gv_location_t>(), count, walks_done,
                        pool_pointers[side].as<uint32_t *>(), d_bases.as<unsigned long long>(),
                        d_last_walk.as<unsigned long long>(), d_fill_scratch.ptr,
                        staged_scatter ? d_remote_blocks.as<unsigned char>() : nullptr, d_totals.as<unsigned long long>(),
                        d_stage.ptr, d_stage_offsets.as<unsigned long long>(), sample_stream));
                    stat_launches += staged_scatter ? 8 : 6;
                }
                GV_CHECK_CUDA(cudaMemcpyAsync(fill.data(), d_fill.ptr, num_block * sizeof(unsigned long long),
                                              cudaMemcpyDeviceToHost, sample_stream));
                GV_CHECK_CUDA(cudaMemcpyAsync(&last_walk, d_last_walk.ptr, sizeof(last_walk),
                                              cudaMemcpyDeviceToHost, sample_stream));
                GV_CHECK_CUDA(cudaStreamSynchronize(sample_stream));
                done_in_span += count;
                walks_done += count;
                complete = true;
                for (int b = 0; b < num_block; b++)
                    complete = complete && fill[b] >= slice;
            }
        }
        // The reference stops at the end of the batch of walks that completed the last block; a
        // batch that runs past the current buffer pulls one more refill (only possible when
        // walks_per_buffer is not a multiple of the batch).  Keep the generator in step.
        const uint64_t executed = (last_walk / walk_batch + 1) * walk_batch;
        const uint64_t needed_buffers = (executed - 1) / walks_per_buffer + 1;
        for (; buffers < needed_buffers; buffers++)
            GV_CHECK_ABI(gv_rng_generate(sampler_generators[sampler_id], d_sampler_random.as<double>(), kRandBatchSize,
                                         sample_stream));
        if (buffers > needed_buffers) {
            // a multi-buffer span overshot (a block filled faster than its 1 / num_block share allows --
            // not expected): rewind the stream to the start of this call and replay what the reference consumed
            if (log_enabled())
                fprintf(stderr, "sampler %d: over-generated %llu refill buffers, rewinding the stream\n", sampler_id,
                        (unsigned long long)(buffers - needed_buffers));
            GV_CHECK_ABI(gv_rng_restore(sampler_generators[sampler_id], d_rng_snapshot.ptr, sample_stream));
            for (uint64_t j = 0; j < needed_buffers; j++)
                GV_CHECK_ABI(gv_rng_generate(sampler_generators[sampler_id], d_sampler_random.as<double>(),
                                             kRandBatchSize, sample_stream));
            buffers = needed_buffers;
        }
        sampler_buffers[sampler_id] += needed_buffers;
    }

    // Barrier over the ranks' sampler threads: everything this rank queued on its sample stream has completed (its
    // deliveries into the peers' pools included), then the hosts meet (gv_host_allgather_fn: a gloo group of its own,
    // graphvite_b200/distributed.py).  A host barrier on purpose: a kernel that spins on a peer's flag holds a hardware
    // queue, and a train launch or an NCCL send queued behind it on the same GPU can then wait for a peer that waits
    // for exactly that send (seen on 2 x B200 with single-warp training, where the main threads lag the samplers).
    void peer_barrier() {
        GV_CHECK_CUDA(cudaStreamSynchronize(sample_stream));
        int token = int(++peer_round), all[256] = {0};
        require(host_allgather_fn && host_allgather_fn(&token, all, sizeof(int), host_allgather_ctx) == 0,
                "host barrier of the samplers failed");
        for (int r = 0; r < num_worker; r++)
            require(all[r] == token, "the ranks' samplers are out of step (barrier " + std::to_string(token) + ")");
    }

    // fill one side of the sample pools with all samplers (core/solver.h:614-628).  Sampler i fills slice i of
    // EVERY block from its own random stream, exactly as in the reference; with several ranks sampler i runs on
    // rank i mod W and writes the blocks owned by other ranks into their pools over NVLink -- the samplers stay
    // independent of one another, so the ranks only meet at the two barriers below.
    void fill_pool(int side) {
        GV_CHECK_CUDA(cudaSetDevice(device));
        // nobody may write into a pool that some rank is still training on: every rank gets here
        // only after it finished the previous episode, so a barrier over the ranks is enough
        if (peer_pools)
            peer_barrier();
        cudaEvent_t begin, end;
        GV_CHECK_CUDA(cudaEventCreate(&begin));
        GV_CHECK_CUDA(cudaEventCreate(&end));
        GV_CHECK_CUDA(cudaEventRecord(begin, sample_stream));
        const uint64_t num_sample = pool_size();
        const uint64_t work_load = (num_sample + num_sampler - 1) / num_sampler;
        for (int i = 0; i < num_sampler; i++)
            if (!peer_pools || i % num_worker == rank)
                run_sampler(i, side, std::min(num_sample, work_load * i), std::min(num_sample, work_load * (i + 1)));
        GV_CHECK_CUDA(cudaEventRecord(end, sample_stream));
        GV_CHECK_CUDA(cudaEventSynchronize(end));
        float ms = 0;
        GV_CHECK_CUDA(cudaEventElapsedTime(&ms, begin, end));
        stat_sample_seconds += ms * 1e-3;
        cudaEventDestroy(begin);
        cudaEventDestroy(end);
        // the pools are complete once every rank's samplers have delivered
        if (peer_pools)
            peer_barrier();
    }

    // ---- host <-> device block movement (replaces Memory::gather/scatter + to_device/to_host) ----
    struct HostState {
        HostMatrix *matrix[3];
    };
    HostState vertex_state() { return {{&vertex_host, &vertex_m1_host, &vertex_m2_host}}; }
    HostState context_state() { return {{&context_host, &context_m1_host, &context_m2_host}}; }

    // load every resident block from the host matrices (load_partition/load_embedding, solver.h:1349-1495)
    // fresh: train(resume=False) has just initialised the host matrices, so the context matrix and every moment
    // matrix are known to be all zeros -- those blocks are cleared on the device instead of uploaded and gathered
    void load_blocks(bool fresh) {
        const size_t total = size_t(graph->num_vertex()) * dim * sizeof(float);
        DeviceArray staging;
        staging.allocate(total);
        slot_of_head.assign(num_partition, -1);
        directory.reset(num_partition, num_worker);
        free_slots.clear();
        int next_slot = 0;
        for (int h = 0; h < num_partition; h++)
            if (directory.owner[h] == rank)
                slot_of_head[h] = next_slot++;
        for (int s = next_slot; s < int(vertex_slots.size()); s++)
            free_slots.push_back(s);
        HostState vertex = vertex_state(), context = context_state();
        const size_t block_bytes = block_floats * sizeof(float);
        for (int s = 0; s < num_state; s++) {
            if (fresh && s > 0) {
                for (int h = 0; h < num_partition; h++)
                    if (slot_of_head[h] >= 0)
                        GV_CHECK_CUDA(cudaMemsetAsync(vertex_slots[slot_of_head[h]].as<float>() + s * block_floats, 0,
                                                      block_bytes, work_stream));
            } else {
                GV_CHECK_CUDA(cudaMemcpyAsync(staging.ptr, vertex.matrix[s]->data(), total, cudaMemcpyHostToDevice,
                                              work_stream));
                for (int h = 0; h < num_partition; h++)
                    if (slot_of_head[h] >= 0)
                        GV_CHECK_ABI(gv_cuda_move_rows(vertex_slots[slot_of_head[h]].as<float>() + s * block_floats,
                                                       staging.as<float>(), partition_ids[h].as<uint32_t>(),
                                                       partitions[h].size(), dim, 1, work_stream));
                GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
            }
            if (fresh) {
                for (int g = 0; g < num_group; g++)
                    GV_CHECK_CUDA(cudaMemsetAsync(context_blocks[g].as<float>() + s * block_floats, 0, block_bytes,
                                                  work_stream));
            } else {
                GV_CHECK_CUDA(cudaMemcpyAsync(staging.ptr, context.matrix[s]->data(), total, cudaMemcpyHostToDevice,
                                              work_stream));
                for (int g = 0; g < num_group; g++) {
                    const int t = g * num_worker + rank;
                    GV_CHECK_ABI(gv_cuda_move_rows(context_blocks[g].as<float>() + s * block_floats, staging.as<float>(),
                                                   partition_ids[t].as<uint32_t>(), partitions[t].size(), dim, 1,
                                                   work_stream));
                }
            }
            GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
        }
    }

    // exchange wrapper: both directions optional
    void exchange(const void *send, int dst, void *recv, int src, uint64_t bytes) {
        require(exchange_fn != nullptr, "world_size > 1 needs gv_solver_set_exchange()");
        if (exchange_fn(send, dst, recv, src, bytes, work_stream, exchange_ctx) != 0)
            throw std::runtime_error("block exchange failed");
    }

    // the block of group g currently held by rank `holder` (vertex blocks migrate, context blocks do not)
    int held_part(bool vertex_side, int g, int holder) const {
        if (!vertex_side)
            return g * num_worker + holder;
        for (int h = g * num_worker; h < (g + 1) * num_worker; h++)
            if (directory.owner[h] == holder)
                return h;
        throw std::runtime_error("internal error: inconsistent block ownership");
    }

    // write every block back into the host matrices (write_back, core/solver.h:1498-1504).  With
    // several ranks each group's blocks are passed around the ring so that every rank ends up
    // with complete matrices behind its numpy views.
    void write_back() {
        const size_t total = size_t(graph->num_vertex()) * dim * sizeof(float);
        DeviceArray staging, incoming;
        staging.allocate(total);
        const uint64_t block_bytes = block_floats * num_state * sizeof(float);
        if (num_worker > 1)
            incoming.allocate(block_bytes);
        for (int side = 0; side < 2; side++) {
            const bool vertex_side = side == 0;
            HostState host = vertex_side ? vertex_state() : context_state();
            for (int s = 0; s < num_state; s++) {
                for (int g = 0; g < num_group; g++) {
                    const int my_part = held_part(vertex_side, g, rank);
                    const float *mine = vertex_side ? vertex_slots[slot_of_head[my_part]].as<float>()
                                                    : context_blocks[g].as<float>();
                    for (int d = 0; d < num_worker; d++) {
                        const int from = (rank - d + num_worker) % num_worker;
                        const float *block = mine;
                        if (d > 0) {  // send mine d ranks ahead, receive the block held d ranks behind
                            exchange(mine, (rank + d) % num_worker, incoming.ptr, from, block_bytes);
                            block = incoming.as<float>();
                        }
                        const int part = held_part(vertex_side, g, from);
                        GV_CHECK_ABI(gv_cuda_move_rows(staging.as<float>(), block + s * block_floats,
                                                       partition_ids[part].as<uint32_t>(), partitions[part].size(),
                                                       dim, 0, work_stream));
                        GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
                    }
                }
                GV_CHECK_CUDA(cudaMemcpyAsync(host.matrix[s]->data(), staging.ptr, total, cudaMemcpyDeviceToHost,
                                              work_stream));
                GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
            }
        }
    }

    // WorkerMixin::build_negative_sampler, core/solver.h:1264-1278 (tail partition, pow(degree, 0.75))
    void build_negative_tables() {
        for (int g = 0; g < num_group; g++) {
            const int tail = g * num_worker + rank;
            const auto &ids = partitions[tail];
            std::vector<float> weights(ids.size());
            {  // elementwise powf, 30 ns each: a few threads for a million vertices
                const size_t threads = ids.size() < 100000 ? 1 : std::min<size_t>(8, std::max(1u, std::thread::hardware_concurrency()));
                auto range = [&](size_t t) {
                    for (size_t i = ids.size() * t / threads; i < ids.size() * (t + 1) / threads; i++)
                        weights[i] = std::pow(graph->vertex_weights[ids[i]], negative_sample_exponent);
                };
                std::vector<std::thread> pool;
                for (size_t t = 1; t < threads; t++)
                    pool.emplace_back(range, t);
                range(0);
                for (auto &thread : pool)
                    thread.join();
            }
            std::vector<float> prob(ids.size());
            std::vector<uint32_t> alias(ids.size());
            build_alias<uint32_t>(weights.data(), weights.size(), prob.data(), alias.data());
            std::vector<gv_alias_entry_t> table(ids.size());
            for (size_t i = 0; i < ids.size(); i++)
                table[i] = {prob[i], alias[i]};
            negative_tables[g].upload(table, work_stream);
            negative_counts[g] = uint32_t(ids.size());
        }
    }

    // GraphSolver::init_embeddings, instance/graph.cuh:724-731
    void init_embeddings() {
        // x = init(seed) for every element, in row-major order -- drawn in bulk (gv_engine.h), same values
        std::uniform_real_distribution<float> init(-0.5 / dim, 0.5 / dim);
        g_engine.fill_uniform(vertex_host.data(), vertex_host.size(), init.a(), init.b());
        context_host.zero();
    }

    // ---- GraphSolver::train prologue + SolverMixin::train up to the first pool fill ----
    void train_begin(const std::string &_model, int _num_epoch, bool _resume, int _augmentation_step,
                     int _random_walk_length, int _random_walk_batch_size, int _shuffle_base, float _p, float _q,
                     int _positive_reuse, float _negative_sample_exponent, float _negative_weight,
                     int _log_frequency) {
        require(built, "The model must be built on a graph first");
        require(!training, "train() is already running");
        GV_CHECK_CUDA(cudaSetDevice(device));
        // instance/graph.cuh:774-789
        augmentation_step = _augmentation_step;
        random_walk_length = _random_walk_length;
        random_walk_batch_size = _random_walk_batch_size;
        shuffle_base = _shuffle_base;
        p = _p;
        q = _q;
        if (augmentation_step == 0)
            augmentation_step = int(std::log(double(kExpectedDegree)) /
                                    std::log(float(graph->num_edge) / graph->num_vertex()));
        if (shuffle_base == 0)
            shuffle_base = augmentation_step;
        // `model` still holds the PREVIOUS call's model here (graph.cuh:785 runs before solver.h:592)
        if (model == "DeepWalk" || model == "node2vec")
            shuffle_base = 1;
        require(augmentation_step >= 1, "`augmentation_step` should be a positive integer");
        require(augmentation_step <= random_walk_length,
                "`random_walk_length` should be no less than `augmentation_step`");
        // core/solver.h:588-611
        require(_model == "DeepWalk" || _model == "LINE" || _model == "node2vec", "Invalid model `" + _model + "`");
        model = _model;
        num_epoch = _num_epoch;
        resume = _resume;
        positive_reuse = _positive_reuse;
        negative_sample_exponent = _negative_sample_exponent;
        negative_weight = _negative_weight;
        log_frequency = std::max(1, _log_frequency);
        require(random_walk_length >= 1 && random_walk_batch_size >= 1 && positive_reuse >= 1,
                "invalid random walk / positive reuse parameters");
        require(pool_size() % shuffle_base == 0 || augmentation_step == 1,
                "Can't perform pseudo shuffle on " + std::to_string(pool_size()) + " elements by a shuffle base of " +
                    std::to_string(shuffle_base) + ". Try setting the episode size to a multiple of the shuffle base");
        if (log_enabled())
            fprintf(stderr, "%s\n", info().c_str());
        PhaseTimer phase;
        // init_embeddings draws |V| * dim floats from the process-wide mt19937 (inherently sequential, 0.1-0.2 s at
        // Youtube size): it runs on its own thread while the sampler tables are built and uploaded and the first
        // pool is sampled -- none of which touches the engine or the embeddings
        std::thread initializer;
        if (!resume) {
            initializer = std::thread([this]() {
                init_embeddings();
                for (auto *m : {&vertex_m1_host, &context_m1_host, &vertex_m2_host, &context_m2_host})
                    m->zero();
            });
            batch_id = 0;
        }
        num_batch = int(batch_id + uint64_t(num_epoch) * graph->num_edge / batch_size);
        // the negative tables (powf + Vose per owned tail partition, uploaded on the work stream) are independent of
        // the samplers' tables and pools (sample stream): built next to them
        std::exception_ptr negative_error;
        std::thread negatives([this, &negative_error]() {
            try {
                GV_CHECK_CUDA(cudaSetDevice(device));
                build_negative_tables();
            } catch (...) {
                negative_error = std::current_exception();
            }
        });
        try {
            prepare_sampling();
            phase.mark("sampler tables + graph upload");
            stat_positive = stat_kernel_seconds = stat_train_seconds = stat_sample_seconds = 0;
            stat_launches = 0;
            fill_pool(pool_id ^ 1);
            phase.mark("first pool fill");
        } catch (...) {
            negatives.join();
            if (initializer.joinable())
                initializer.join();
            throw;
        }
        negatives.join();
        phase.mark("negative tables (rest)");
        if (initializer.joinable())
            initializer.join();
        phase.mark("init embeddings (rest)");
        if (negative_error)
            std::rethrow_exception(negative_error);
        load_blocks(!resume);
        phase.mark("embedding upload");
        if (capture_negatives)
            d_negatives_out.allocate(uint64_t(chunk_batches) * batch_size * std::max(1, num_negative) * 4);
        previous_batch_loss = 0;
        training = true;
        step_in_episode = 0;
    }

    // WorkerMixin::train for one block, core/solver.h:1511-1557: positive_reuse * episode_size batches
    void train_block(int head, int tail, int first_batch, int batch_stride) {
        const int g = tail / num_worker;
        float *vertex = vertex_slots[slot_of_head[head]].as<float>();
        float *context = context_blocks[g].as<float>();
        gv_matrices_t matrices;
        matrices.dim = dim;
        matrices.vertex = vertex;
        matrices.context = context;
        matrices.vertex_m1 = num_state >= 2 ? vertex + block_floats : nullptr;
        matrices.context_m1 = num_state >= 2 ? context + block_floats : nullptr;
        matrices.vertex_m2 = num_state >= 3 ? vertex + 2 * block_floats : nullptr;
        matrices.context_m2 = num_state >= 3 ? context + 2 * block_floats : nullptr;
        gv_device_optimizer_t device_optimizer = {optimizer.desc.type, optimizer.desc.weight_decay, optimizer.desc.a,
                                                  optimizer.desc.b, optimizer.desc.epsilon};
        const uint32_t *pool = pool_block(pool_id, head, g);
        const uint64_t per_batch_random = uint64_t(batch_size) * num_negative * 2;
        std::vector<float> lr(episode_size), loss(episode_size);
        int buffer = 0;
        // The stream is positional (one call for n batches == n calls for one batch), and a call is 4096 sequential
        // XORWOW streams however short it is: the randoms of the whole sub-episode are generated by ONE call when they
        // fit 2 GB (long calls are split by skip-ahead, gv_rng.cu) instead of one 4096-thread launch per chunk that the
        // next train launch has to wait for, and the negatives of the whole sub-episode are drawn from them by ONE
        // gpu::Sample launch (the reference: one per batch, solver.h:1536-1539).  GV_RNG_PER_CHUNK=1 keeps the
        // per-chunk calls.
        const uint64_t step_random = uint64_t(episode_size) * per_batch_random;
        const bool whole_step = num_negative > 0 && step_random * sizeof(double) <= (uint64_t(2) << 30) &&
                                !getenv("GV_RNG_PER_CHUNK");
        if (whole_step) {
            d_random_step.allocate(step_random * sizeof(double));
            d_negatives_step.allocate(uint64_t(episode_size) * batch_size * num_negative * sizeof(uint32_t));
        }
        if (!step_timer[0]) {
            GV_CHECK_CUDA(cudaEventCreate(&step_timer[0]));
            GV_CHECK_CUDA(cudaEventCreate(&step_timer[1]));
        }
        for (int reuse = 0; reuse < positive_reuse; reuse++) {
            for (int j = 0; j < episode_size; j++)
                lr[j] = optimizer.lr_at(first_batch + (reuse * episode_size + j) * batch_stride, num_batch);
            GV_CHECK_CUDA(cudaMemcpyAsync(d_lr.ptr, lr.data(), episode_size * sizeof(float), cudaMemcpyHostToDevice,
                                          work_stream));
            GV_CHECK_CUDA(cudaMemsetAsync(d_loss.ptr, 0, episode_size * sizeof(float), work_stream));
            if (whole_step) {  // random_free[0] / random_ready[0] guard the step buffers
                GV_CHECK_CUDA(cudaStreamWaitEvent(random_stream, random_free[0], 0));
                GV_CHECK_ABI(gv_rng_generate(worker_generator, d_random_step.as<double>(), step_random, random_stream));
                GV_CHECK_ABI(gv_cuda_sample_negatives(negative_tables[g].as<gv_alias_entry_t>(), negative_counts[g],
                                                      d_random_step.as<double>(),
                                                      uint64_t(episode_size) * batch_size * num_negative,
                                                      d_negatives_step.as<uint32_t>(), random_stream));
                stat_launches += 2;
                GV_CHECK_CUDA(cudaEventRecord(random_ready[0], random_stream));
                GV_CHECK_CUDA(cudaStreamWaitEvent(work_stream, random_ready[0], 0));
            }
            // device time of the pass's train launches: one event pair around all of them
            GV_CHECK_CUDA(cudaEventRecord(step_timer[0], work_stream));
            for (int j0 = 0; j0 < episode_size; j0 += chunk_batches, buffer = (buffer + 1) % kRandomBuffers) {
                const int count = std::min(chunk_batches, episode_size - j0);
                // negatives: one curandGenerateUniformDouble(2 * B * k) per batch, like train_batch (solver.h:1536)
                if (num_negative > 0 && !whole_step) {
                    GV_CHECK_CUDA(cudaStreamWaitEvent(random_stream, random_free[buffer], 0));
                    // (the stream is positional: one call for the chunk == one call per batch)
                    GV_CHECK_ABI(gv_rng_generate(worker_generator, d_random[buffer].as<double>(),
                                                 uint64_t(count) * per_batch_random, random_stream));
                    stat_launches++;
                    GV_CHECK_CUDA(cudaEventRecord(random_ready[buffer], random_stream));
                    GV_CHECK_CUDA(cudaStreamWaitEvent(work_stream, random_ready[buffer], 0));
                }
                // The loss is only needed for a batch whose successor logs it (one in log_frequency,
                // core/solver.h:1541-1549): those batches get their own launch of the LOSS kernel.
                for (int j = j0; j < j0 + count;) {
                    auto wants_loss = [&](int b) {
                        return (first_batch + (reuse * episode_size + b + 1) * batch_stride) % log_frequency == 0;
                    };
                    const bool with_loss = wants_loss(j);
                    int j1 = j + 1;
                    while (!with_loss && j1 < j0 + count && !wants_loss(j1))
                        j1++;
                    const uint64_t first_negative = uint64_t(j) * batch_size * num_negative;
                    GV_CHECK_ABI(gv_cuda_train_block(
                        &matrices, pool + uint64_t(j) * batch_size * 2, uint64_t(j1 - j) * batch_size, num_negative,
                        whole_step ? d_negatives_step.as<uint32_t>() + first_negative : nullptr,
                        whole_step ? nullptr : d_random[buffer].as<double>() + uint64_t(j - j0) * per_batch_random,
                        negative_tables[g].as<gv_alias_entry_t>(), negative_counts[g],
                        capture_negatives ? d_negatives_out.as<uint32_t>() + uint64_t(j - j0) * batch_size * num_negative
                                          : nullptr,
                        &device_optimizer, d_lr.as<float>() + j, batch_size, negative_weight, nullptr,
                        with_loss ? d_loss.as<float>() + j : nullptr, train_num_warps, work_stream));
                    stat_launches++;
                    j = j1;
                }
                if (!whole_step)
                    GV_CHECK_CUDA(cudaEventRecord(random_free[buffer], work_stream));
                if (capture_negatives && reuse == positive_reuse - 1 && j0 + count == episode_size) {
                    last_negatives.resize(size_t(batch_size) * num_negative);
                    GV_CHECK_CUDA(cudaMemcpyAsync(last_negatives.data(),
                                                  d_negatives_out.as<uint32_t>() +
                                                      size_t(count - 1) * batch_size * num_negative,
                                                  last_negatives.size() * 4, cudaMemcpyDeviceToHost, work_stream));
                }
            }
            GV_CHECK_CUDA(cudaEventRecord(step_timer[1], work_stream));
            if (whole_step)
                GV_CHECK_CUDA(cudaEventRecord(random_free[0], work_stream));
            GV_CHECK_CUDA(cudaMemcpyAsync(loss.data(), d_loss.ptr, episode_size * sizeof(float),
                                          cudaMemcpyDeviceToHost, work_stream));
            GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
            float ms = 0;
            GV_CHECK_CUDA(cudaEventElapsedTime(&ms, step_timer[0], step_timer[1]));
            stat_kernel_seconds += ms * 1e-3;
            // the reference logs, at batch b, the mean loss of the batch trained before it (appendix A.8)
            for (int j = 0; j < episode_size; j++) {
                const int this_batch = first_batch + (reuse * episode_size + j) * batch_stride;
                if (this_batch % log_frequency == 0) {  // previous_batch_loss was computed for exactly this
                    logged_loss.push_back(previous_batch_loss);
                    if (log_enabled())
                        fprintf(stderr, "Batch id: %d / %d\nloss = %g\n", this_batch, num_batch, previous_batch_loss);
                }
                previous_batch_loss = loss[j] / batch_size;
            }
        }
        stat_positive += double(positive_reuse) * episode_size * batch_size;
    }

    // ---- the episode loop, core/solver.h:629-649, cut into schedule steps (sub-episodes) ----
    std::vector<std::vector<Assignment>> schedule;
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

    // One sub-episode: every worker trains one (head, tail) block.  Returns false when training is over.
    bool train_step() {
        require(training, "train_begin() has not been called");
        GV_CHECK_CUDA(cudaSetDevice(device));
        if (step_in_episode == 0) {
            if (batch_id >= num_batch)
                return false;
            pool_id ^= 1;
            schedule = make_schedule(num_partition, num_worker);
            // the samplers fill the other pool while the workers train on this one
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
            cudaEvent_t begin, end;
            GV_CHECK_CUDA(cudaEventCreate(&begin));
            GV_CHECK_CUDA(cudaEventCreate(&end));
            GV_CHECK_CUDA(cudaEventRecord(begin, work_stream));
            if (num_worker > 1)
                rotate_vertex_blocks(step);
            // batch ids: the reference's workers share an atomic counter (solver.h:1520); we use the
            // interleaving first_batch + j * width, which is what lock-step workers would draw.
            train_block(step[rank].head, step[rank].tail, batch_id + rank, width);
            batch_id += positive_reuse * episode_size * width;
            GV_CHECK_CUDA(cudaEventRecord(end, work_stream));
            GV_CHECK_CUDA(cudaEventSynchronize(end));
            float ms = 0;
            GV_CHECK_CUDA(cudaEventElapsedTime(&ms, begin, end));
            stat_train_seconds += ms * 1e-3;
            cudaEventDestroy(begin);
            cudaEventDestroy(end);
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

    // Move the vertex blocks this step needs onto their workers.  Inside one group of
    // num_worker head blocks the assignment is a permutation, so each rank sends at most one
    // block and receives at most one (a ring shift for the default schedule).
    void rotate_vertex_blocks(const std::vector<Assignment> &step) {
        const Transfer t = directory.plan(rank, step);
        const uint64_t bytes = block_floats * num_state * sizeof(float);
        int incoming_slot = -1;
        if (t.source != rank) {
            require(!free_slots.empty(), "internal error: no free vertex slot");
            incoming_slot = free_slots.back();
            free_slots.pop_back();
        }
        if (t.give >= 0 || incoming_slot >= 0)
            exchange(t.give >= 0 ? vertex_slots[slot_of_head[t.give]].ptr : nullptr, t.destination,
                     incoming_slot >= 0 ? vertex_slots[incoming_slot].ptr : nullptr,
                     incoming_slot >= 0 ? t.source : -1, bytes);
        if (t.give >= 0) {
            free_slots.push_back(slot_of_head[t.give]);
            slot_of_head[t.give] = -1;
        }
        if (incoming_slot >= 0)
            slot_of_head[t.need] = incoming_slot;
        directory.commit(step);
    }

    void train_end() {
        require(training, "train_begin() has not been called");
        GV_CHECK_CUDA(cudaSetDevice(device));
        while (step_in_episode != 0)  // never stop in the middle of an episode
            train_step();
        PhaseTimer phase;  // GV_LOG=2
        GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
        phase.mark("last train launches drained");
        write_back();
        phase.mark("write-back");
        training = false;
    }

    void train(const std::string &_model, int _num_epoch, bool _resume, int _augmentation_step,
               int _random_walk_length, int _random_walk_batch_size, int _shuffle_base, float _p, float _q,
               int _positive_reuse, float _negative_sample_exponent, float _negative_weight, int _log_frequency) {
        train_begin(_model, _num_epoch, _resume, _augmentation_step, _random_walk_length, _random_walk_batch_size,
                    _shuffle_base, _p, _q, _positive_reuse, _negative_sample_exponent, _negative_weight,
                    _log_frequency);
        try {
            while (train_episode())
                ;
        } catch (...) {
            training = false;
            throw;
        }
        train_end();
    }

    // SolverMixin::predict_numpy, core/solver.h:729-802.  Both matrices are uploaded in global-id
    // order, so no bucketing by partition block is needed; rows are (v, c) like the reference.
    void predict(const uint32_t *pairs, uint64_t num, float *logits) {
        require(built, "The model must be built on a graph first");
        GV_CHECK_CUDA(cudaSetDevice(device));
        const uint32_t n = graph->num_vertex();
        std::vector<uint32_t> batch(num * 2);
        for (uint64_t i = 0; i < num; i++) {
            require(pairs[i * 2] < n && pairs[i * 2 + 1] < n, "predict: vertex id out of range");
            batch[i * 2] = pairs[i * 2 + 1];  // device layout {tail, head}
            batch[i * 2 + 1] = pairs[i * 2];
        }
        DeviceArray d_vertex, d_context, d_batch, d_logits;
        d_vertex.upload(vertex_host, work_stream);
        d_context.upload(context_host, work_stream);
        d_batch.upload(batch, work_stream);
        d_logits.allocate(std::max<uint64_t>(1, num) * sizeof(float));
        GV_CHECK_ABI(gv_cuda_predict(dim, d_vertex.as<float>(), d_context.as<float>(), d_batch.as<uint32_t>(), num,
                                     d_logits.as<float>(), work_stream));
        GV_CHECK_CUDA(cudaMemcpyAsync(logits, d_logits.ptr, num * sizeof(float), cudaMemcpyDeviceToHost, work_stream));
        GV_CHECK_CUDA(cudaStreamSynchronize(work_stream));
    }

    // SolverMixin::clear + GraphSolver::clear: free everything but the host embeddings
    void clear() {
        require(!training, "clear() during training");
        cudaSetDevice(device);
        vertex_slots.clear();
        context_blocks.clear();
        negative_tables.clear();
        partition_ids.clear();
        close_peers();
        uploader.release();
        pool_arena.release();
        for (int side = 0; side < 2; side++)
            pool_pointers[side].release();
        for (auto *a : {&d_offsets, &d_edge_u, &d_edge_v, &d_edge_prob, &d_edge_alias, &d_vertex_tables, &d_locations,
                        &d_sampler_random, &d_chains, &d_fill, &d_last_walk, &d_fill_scratch, &d_random[0],
                        &d_random[1], &d_random[2], &d_random[3], &d_random_step, &d_lr, &d_loss, &d_negatives_out, &d_peer_controls, &d_totals, &d_bases, &d_stage,
                        &d_stage_offsets, &d_remote_blocks, &d_edge_tables,
                        &d_table_offsets})
            a->release();
        for (auto *m : {&vertex_m1_host, &context_m1_host, &vertex_m2_host, &context_m2_host})
            m->clear();
        partitions.clear();
        sampling_ready = false;
        built = false;
    }

    // SolverMixin::info, core/solver.h:468-516 + GraphSolver overrides, graph.cuh:733-752
    std::string info() const {
        auto yes_no = [](bool x) { return x ? "yes" : "no"; };
        auto size_string = [](uint64_t size) {
            std::stringstream ss;
            ss.precision(3);
            if (size >= (uint64_t(1) << 40))
                ss << double(size) / (uint64_t(1) << 40) << " TiB";
            else if (size >= (uint64_t(1) << 30))
                ss << double(size) / (uint64_t(1) << 30) << " GiB";
            else if (size >= (uint64_t(1) << 20))
                ss << double(size) / (uint64_t(1) << 20) << " MiB";
            else if (size >= (uint64_t(1) << 10))
                ss << double(size) / (uint64_t(1) << 10) << " KiB";
            else
                ss << size << " B";
            return ss.str();
        };
        std::stringstream ss;
        ss << "GraphSolver<" << dim << ", float32, uint32>" << std::endl;
        ss << "----------------- Resource -----------------" << std::endl;
        ss << "#worker: " << num_worker << ", #sampler: " << num_sampler << ", #partition: " << num_partition
           << std::endl;
        ss << "tied weights: no, episode size: " << episode_size << std::endl;
        ss << "gpu memory limit: " << size_string(gpu_memory_limit) << std::endl;
        ss << "gpu memory cost: " << size_string(gpu_memory_cost) << std::endl;
        ss << "----------------- Sampling -----------------" << std::endl;
        if (model == "LINE")
            ss << "augmentation step: " << augmentation_step << ", shuffle base: " << shuffle_base << std::endl;
        if (model == "DeepWalk")
            ss << "augmentation step: " << augmentation_step << std::endl;
        if (model == "node2vec")
            ss << "augmentation step: " << augmentation_step << ", p: " << p << ", q: " << q << std::endl;
        ss << "random walk length: " << random_walk_length << std::endl;
        ss << "random walk batch size: " << random_walk_batch_size << std::endl;
        ss << "#negative: " << num_negative << ", negative sample exponent: " << negative_sample_exponent
           << std::endl;
        ss << "----------------- Training -----------------" << std::endl;
        ss << "model: " << model << std::endl;
        ss << optimizer.info() << std::endl;
        ss << "#epoch: " << num_epoch << ", batch size: " << batch_size << std::endl;
        ss << "resume: " << yes_no(resume) << std::endl;
        ss << "positive reuse: " << positive_reuse << ", negative weight: " << negative_weight;
        return ss.str();
    }

    std::string attributes() const {
        std::stringstream ss;
        ss.precision(9);
        ss << "dim=" << dim << "\nnum_partition=" << num_partition << "\nnum_negative=" << num_negative
           << "\nnegative_sample_exponent=" << negative_sample_exponent << "\nnegative_weight=" << negative_weight
           << "\nmodel=" << model << "\nnum_epoch=" << num_epoch << "\nresume=" << int(resume)
           << "\nepisode_size=" << episode_size << "\nbatch_size=" << batch_size
           << "\naugmentation_step=" << augmentation_step << "\nrandom_walk_length=" << random_walk_length
           << "\nrandom_walk_batch_size=" << random_walk_batch_size << "\nshuffle_base=" << shuffle_base
           << "\np=" << p << "\nq=" << q << "\npositive_reuse=" << positive_reuse
           << "\nlog_frequency=" << log_frequency << "\nnum_worker=" << num_worker
           << "\nnum_sampler=" << num_sampler << "\ngpu_memory_limit=" << gpu_memory_limit
           << "\ngpu_memory_cost=" << gpu_memory_cost << "\nnum_batch=" << num_batch << "\nbatch_id=" << batch_id
           << "\npool_id=" << pool_id << "\npartition_size=" << partition_size << "\nrank=" << rank
           << "\nchunk_batches=" << chunk_batches
           << "\noptimizer_type=" << optimizer.type_name() << "\noptimizer_lr=" << optimizer.init_lr
           << "\noptimizer_weight_decay=" << optimizer.desc.weight_decay << "\n";
        return ss.str();
    }
};

}  // namespace gv

// =============================================================================
// C ABI
// =============================================================================
using gv::Solver;

struct gv_solver {
    std::unique_ptr<Solver> solver;
};

#define GV_TRY try {
#define GV_CATCH(ret)                  \
    }                                  \
    catch (const std::exception &e) {  \
        gv::set_error(e.what());       \
        return ret;                    \
    }

static int copy_string(const std::string &text, char *buffer, size_t capacity) {
    if (buffer && capacity) {
        strncpy(buffer, text.c_str(), capacity - 1);
        buffer[capacity - 1] = 0;
    }
    return int(text.size());
}

extern "C" {

// The block movement of `num_episode` episodes for every rank, from the code the solver itself runs:
// out[((e * steps + s) * W + rank) * 6 + {0..5}] = head, tail, need-source, give, destination, #held
int gv_schedule_plan(int num_partition, int num_worker, int num_episode, int *out, int capacity) {
    GV_TRY
    if (num_partition < num_worker || num_partition % num_worker != 0)
        throw std::runtime_error("#partition must be a positive multiple of #worker");
    const auto schedule = gv::make_schedule(num_partition, num_worker);
    gv::BlockDirectory directory;
    directory.reset(num_partition, num_worker);
    const int width = int(schedule[0].size());
    if (int(schedule.size()) * width * 6 * num_episode > capacity)
        throw std::runtime_error("gv_schedule_plan: capacity too small");
    int n = 0;
    for (int e = 0; e < num_episode; e++)
        for (const auto &step : schedule) {
            std::vector<gv::Transfer> transfers;
            for (int rank = 0; rank < width; rank++)
                transfers.push_back(directory.plan(rank, step));
            directory.commit(step);
            for (int rank = 0; rank < width; rank++) {
                int held = 0;
                for (int owner : directory.owner)
                    held += owner == rank;
                const int values[6] = {step[rank].head, step[rank].tail, transfers[rank].source, transfers[rank].give,
                                       transfers[rank].destination, held};
                for (int v : values)
                    out[n++] = v;
            }
        }
    return int(schedule.size());
    GV_CATCH(-1)
}

void gv_reset_global_engine(uint32_t seed) {
    gv::g_engine = gv::Mt19937(seed);
}

// test hook: gv::Mt19937 against libstdc++'s std::mt19937 -- raw draws, the seeds' distribution, and
// fill_uniform() against std::uniform_real_distribution<float> for awkward sizes; 0 = identical
int gv_engine_self_check(uint32_t seed, uint64_t bulk) {
    std::mt19937 reference(seed);
    gv::Mt19937 ours(seed);
    for (int i = 0; i < 2000; i++)
        if (reference() != ours())
            return 1;
    std::uniform_int_distribution<unsigned long long> seeds(0, ULLONG_MAX);
    for (int i = 0; i < 16; i++)
        if (seeds(reference) != seeds(ours))
            return 2;
    const uint64_t sizes[] = {1, 3, 623, 624, 625, bulk, 5};
    for (uint64_t n : sizes) {
        std::uniform_real_distribution<float> distribution(-0.5 / 96, 0.5 / 96);
        std::vector<float> expected(n), got(n);
        for (auto &x : expected)
            x = distribution(reference);
        ours.fill_uniform(got.data(), n, distribution.a(), distribution.b());
        if (memcmp(expected.data(), got.data(), n * sizeof(float)) != 0)
            return 3;
        if (reference() != ours())
            return 4;
    }
    // skip_spans(5) = levels 0 and 2 = 5 * kJumpBlocks * 624 draws, from a position in the middle of a block
    if (bulk >= 1000) {
        gv::Mt19937 jumped = ours, stepped = ours;
        jumped.skip_spans(5);
        for (uint64_t i = 0; i < uint64_t(gv::kJumpBlocks) * 624 * 5; i++)
            stepped();
        for (int i = 0; i < 2000; i++)
            if (jumped() != stepped())
                return 5;
    }
    // the multi-threaded fill (spans of kJumpBlocks * 624 draws) against the sequential one, engine state included
    if (bulk >= 1000) {
        const uint64_t n = uint64_t(gv::kJumpBlocks) * 624 * 6 + bulk;
        for (int threads : {2, 5}) {
            gv::Mt19937 sequential = ours, parallel = ours;
            std::vector<float> expected(n), got(n);
            sequential.fill_uniform(expected.data(), n, -0.25f, 0.75f, 1);
            parallel.fill_uniform(got.data(), n, -0.25f, 0.75f, threads);
            if (memcmp(expected.data(), got.data(), n * sizeof(float)) != 0)
                return 6;
            for (int i = 0; i < 700; i++)
                if (sequential() != parallel())
                    return 7;
        }
    }
    return 0;
}

gv_solver_t *gv_solver_create(int dim, const int *device_ids, int num_device, int num_sampler_per_worker,
                              uint64_t gpu_memory_limit, int rank, int world_size) {
    GV_TRY
    std::unique_ptr<Solver> solver(
        new Solver(dim, device_ids, num_device, num_sampler_per_worker, gpu_memory_limit, rank, world_size));
    gv_solver *handle = new gv_solver();
    handle->solver = std::move(solver);
    return handle;
    GV_CATCH(nullptr)
}

void gv_solver_destroy(gv_solver_t *solver) {
    delete solver;
}

int gv_solver_set_exchange(gv_solver_t *solver, gv_exchange_fn fn, void *ctx) {
    solver->solver->exchange_fn = fn;
    solver->solver->exchange_ctx = ctx;
    return 0;
}

int gv_solver_set_host_allgather(gv_solver_t *solver, gv_host_allgather_fn fn, void *ctx) {
    solver->solver->host_allgather_fn = fn;
    solver->solver->host_allgather_ctx = ctx;
    return 0;
}

// Unmap the other ranks' pool arenas.  CUDA requires every importer of an IPC allocation to close it
// before the exporter frees it, so a multi-rank teardown is: release_peers on every rank, a barrier,
// then destroy / clear / rebuild.
int gv_solver_release_peers(gv_solver_t *solver) {
    GV_TRY
    Solver &s = *solver->solver;
    if (s.training)
        throw std::runtime_error("release_peers() during training");
    cudaSetDevice(s.device);
    s.close_peers();
    return 0;
    GV_CATCH(-1)
}

int gv_solver_set_option(gv_solver_t *solver, const char *name, int value) {
    GV_TRY
    if (std::string(name) == "capture_negatives")
        solver->solver->capture_negatives = value != 0;
    else if (std::string(name) == "train_num_warps")
        solver->solver->train_num_warps = value;
    else if (std::string(name) == "chunk_batches")
        solver->solver->set_chunk_batches(value);
    else
        throw std::runtime_error(std::string("unknown option `") + name + "`");
    return 0;
    GV_CATCH(-1)
}

int gv_solver_build(gv_solver_t *solver, gv_graph_t *graph, const gv_optimizer_t *optimizer, int num_partition,
                    int num_negative, int batch_size, int episode_size) {
    GV_TRY
    solver->solver->build(&gv_graph_ref(graph), optimizer, num_partition, num_negative, batch_size, episode_size);
    return 0;
    GV_CATCH(-1)
}

int gv_solver_train(gv_solver_t *solver, const char *model, int num_epoch, int resume, int augmentation_step,
                    int random_walk_length, int random_walk_batch_size, int shuffle_base, float p, float q,
                    int positive_reuse, float negative_sample_exponent, float negative_weight, int log_frequency) {
    GV_TRY
    solver->solver->train(model, num_epoch, resume != 0, augmentation_step, random_walk_length,
                          random_walk_batch_size, shuffle_base, p, q, positive_reuse, negative_sample_exponent,
                          negative_weight, log_frequency);
    return 0;
    GV_CATCH(-1)
}

int gv_solver_train_begin(gv_solver_t *solver, const char *model, int num_epoch, int resume, int augmentation_step,
                          int random_walk_length, int random_walk_batch_size, int shuffle_base, float p, float q,
                          int positive_reuse, float negative_sample_exponent, float negative_weight,
                          int log_frequency) {
    GV_TRY
    solver->solver->train_begin(model, num_epoch, resume != 0, augmentation_step, random_walk_length,
                                random_walk_batch_size, shuffle_base, p, q, positive_reuse, negative_sample_exponent,
                                negative_weight, log_frequency);
    return 0;
    GV_CATCH(-1)
}

int gv_solver_train_episode(gv_solver_t *solver) {
    GV_TRY
    return solver->solver->train_episode() ? 1 : 0;
    GV_CATCH(-1)
}

int gv_solver_train_step(gv_solver_t *solver) {
    GV_TRY
    return solver->solver->train_step() ? 1 : 0;
    GV_CATCH(-1)
}

double gv_solver_device_timer(gv_solver_t *solver, int stop) {
    GV_TRY
    Solver &s = *solver->solver;
    cudaSetDevice(s.device);
    if (!s.timer_begin) {
        cudaEventCreate(&s.timer_begin);
        cudaEventCreate(&s.timer_end);
    }
    if (!stop) {
        if (cudaEventRecord(s.timer_begin, s.work_stream) != cudaSuccess)
            throw std::runtime_error("cudaEventRecord failed");
        return 0;
    }
    float ms = 0;
    if (cudaEventRecord(s.timer_end, s.work_stream) != cudaSuccess || cudaEventSynchronize(s.timer_end) != cudaSuccess ||
        cudaEventElapsedTime(&ms, s.timer_begin, s.timer_end) != cudaSuccess)
        throw std::runtime_error("device timer failed");
    return double(ms) * 1e-3;
    GV_CATCH(-1.0)
}

int gv_solver_train_end(gv_solver_t *solver) {
    GV_TRY
    solver->solver->train_end();
    return 0;
    GV_CATCH(-1)
}

int gv_solver_predict(gv_solver_t *solver, const uint32_t *pairs, uint64_t num, float *logits) {
    GV_TRY
    solver->solver->predict(pairs, num, logits);
    return 0;
    GV_CATCH(-1)
}

int gv_solver_clear(gv_solver_t *solver) {
    GV_TRY
    solver->solver->clear();
    return 0;
    GV_CATCH(-1)
}

float *gv_solver_embeddings(gv_solver_t *solver, int which, uint64_t *rows, int *dim) {
    Solver &s = *solver->solver;
    if (rows)
        *rows = s.graph ? s.vertex_host.size() / s.dim : 0;
    if (dim)
        *dim = s.dim;
    return which == 0 ? s.vertex_host.data() : s.context_host.data();
}

int gv_solver_info(const gv_solver_t *solver, char *buffer, size_t capacity) {
    return copy_string(solver->solver->info(), buffer, capacity);
}

int gv_solver_attributes(const gv_solver_t *solver, char *buffer, size_t capacity) {
    return copy_string(solver->solver->attributes(), buffer, capacity);
}

int gv_solver_logged_loss(const gv_solver_t *solver, float *out, int capacity) {
    const auto &loss = solver->solver->logged_loss;
    for (int i = 0; i < capacity && i < int(loss.size()); i++)
        out[i] = loss[i];
    return int(loss.size());
}

int gv_solver_stats(const gv_solver_t *solver, double *out, int capacity) {
    const Solver &s = *solver->solver;
    const double values[] = {s.stat_positive, s.stat_kernel_seconds, s.stat_train_seconds, s.stat_sample_seconds,
                             double(s.stat_launches.load())};
    for (int i = 0; i < capacity && i < 5; i++)
        out[i] = values[i];
    return 5;
}

int gv_solver_locations(const gv_solver_t *solver, uint32_t *part_of, uint32_t *local_of) {
    const auto &locations = solver->solver->locations;
    for (size_t v = 0; v < locations.size(); v++) {
        part_of[v] = locations[v].part;
        local_of[v] = locations[v].local;
    }
    return 0;
}

int64_t gv_solver_pool(gv_solver_t *solver, int pool, int head_partition, int tail_partition, uint32_t *out) {
    GV_TRY
    Solver &s = *solver->solver;
    if (pool < 0 || pool > 1 || head_partition < 0 || head_partition >= s.num_partition || tail_partition < 0 ||
        tail_partition >= s.num_partition)
        throw std::runtime_error("gv_solver_pool: index out of range");
    if (!s.owns_tail(tail_partition))
        return 0;
    cudaSetDevice(s.device);
    if (out) {
        if (cudaMemcpy(out, s.pool_block(pool, head_partition, tail_partition / s.num_worker), s.pool_block_bytes(),
                       cudaMemcpyDeviceToHost) != cudaSuccess)
            throw std::runtime_error("gv_solver_pool: copy failed");
    }
    return int64_t(s.pool_size());
    GV_CATCH(-1)
}

int gv_solver_last_negatives(gv_solver_t *solver, uint32_t *out) {
    const auto &negatives = solver->solver->last_negatives;
    if (out && !negatives.empty())
        memcpy(out, negatives.data(), negatives.size() * sizeof(uint32_t));
    return int(negatives.size());
}

}  // extern "C"
