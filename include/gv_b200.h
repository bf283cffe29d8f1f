/*
 * gv_b200.h -- C ABI of the B200-native node-embedding trainer (libgv_b200.so).
 *
 * This is the drop-in boundary for the node-embedding hot path of
 * DeepGraphLearning/graphvite v0.2.2.  Every entry point names the reference
 * interface it replaces (paths relative to the reference's include/).  All
 * signatures are plain C: pointers, sizes and PODs; no C++ or torch types.
 *
 * Layers
 *   gv_cuda_*   device layer: hand-written sm_100a kernels + launchers.  Pointers are
 *               DEVICE pointers unless stated; `stream` is a cudaStream_t passed as void*.
 *   gv_graph_*, gv_optimizer_*, gv_solver_*
 *               host runtime (C++): what bind.h exposes as graphvite.graph.Graph,
 *               graphvite.optimizer.* and graphvite.solver.GraphSolver.
 *
 * Error model: functions return 0 on success and a non-zero code on failure;
 * gv_last_error() returns a thread-local message.  (The reference aborts the
 * process through glog CHECK/LOG(FATAL), util/debug.h:27-38; a C ABI must not.)
 */
#ifndef GV_B200_H_
#define GV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GV_VERSION "0.1.0"

/* ---- enums / PODs -------------------------------------------------------- */

/* core/optimizer.h:28-36 OptimizerType */
enum { GV_OPT_SGD = 0, GV_OPT_MOMENTUM = 1, GV_OPT_ADAGRAD = 2, GV_OPT_RMSPROP = 3, GV_OPT_ADAM = 4 };
/* core/optimizer.h:42-85 LRSchedule::type */
enum { GV_SCHEDULE_CONSTANT = 0, GV_SCHEDULE_LINEAR = 1, GV_SCHEDULE_CUSTOM = 2 };
/* instance/graph.cuh:620-622 available models */
enum { GV_MODEL_DEEPWALK = 0, GV_MODEL_LINE = 1, GV_MODEL_NODE2VEC = 2 };
/* instance/knowledge_graph.cuh:575-577 available models (all six are implemented) */
enum { GV_KG_TRANSE = 0, GV_KG_DISTMULT = 1, GV_KG_COMPLEX = 2, GV_KG_SIMPLE = 3, GV_KG_ROTATE = 4, GV_KG_QUATE = 5 };

/* Device-side view of core/optimizer.h Optimizer (the reference passes the whole C++
 * object, std::string included, by value into its kernels; appendix A.14 of SURVEY.md). */
typedef struct {
    int type;            /* GV_OPT_* */
    float weight_decay;
    float a;             /* momentum | alpha | beta1 */
    float b;             /* beta2 */
    float epsilon;
} gv_device_optimizer_t;

/* Host-side optimizer description (core/optimizer.h:272-319 helper classes). */
typedef struct {
    int type;            /* GV_OPT_*; -1 = "Default" (solver default, core/solver.h:291-296) */
    float lr;
    float weight_decay;
    float a, b, epsilon;
    int schedule;        /* GV_SCHEDULE_* */
    /* GV_SCHEDULE_CUSTOM: factor = fn(batch_id, num_batch, ctx) (LRSchedule::ScheduleFunction) */
    float (*schedule_fn)(int batch_id, int num_batch, void *ctx);
    void *schedule_ctx;
} gv_optimizer_t;

/* Row-major [rows][dim] fp32 matrices of one (head block, tail block) pair resident in HBM
 * (base/vector.h Vector<dim,float>, base/memory.h Memory<Vector,Index>::device_ptr). */
typedef struct {
    int dim;
    float *vertex, *context;           /* embeddings[0], embeddings[1] */
    float *vertex_m1, *context_m1;     /* first moments  (NULL unless num_moment >= 1) */
    float *vertex_m2, *context_m2;     /* second moments (NULL unless num_moment == 2) */
} gv_matrices_t;

/* One alias-table entry as laid out on the device: prob and alias of
 * base/alias_table.cuh AliasTable<float, uint32> interleaved so one 8-byte load fetches both. */
typedef struct {
    float prob;
    uint32_t alias;
} gv_alias_entry_t;

const char *gv_last_error(void);
const char *gv_version(void);

/* ---- device layer -------------------------------------------------------- */

/* Replaces gpu::graph::train / train_1_moment / train_2_moment
 * (instance/gpu/graph.cuh:36-242) + the per-batch loop of WorkerMixin::train
 * (core/solver.h:1511-1557): ONE launch consumes `num_sample` positive samples
 * (= any number of reference batches) from a device-resident pool block.
 *
 *   pool           [num_sample] pairs {tail_local, head_local} (std::tuple layout, appendix A.1)
 *   negatives      [num_sample][num_negative] local tail ids, or NULL to draw them in-kernel:
 *   random         [num_sample][num_negative][2] cuRAND doubles (core/solver.h:1536) and
 *   negative_table the tail partition's alias table (replaces gpu::Sample,
 *                  base/alias_table.cuh:175-183, including its double->float narrowing)
 *   negatives_out  optional [num_sample][num_negative]: the ids the kernel drew (test hook)
 *   lr_per_batch   [ceil(num_sample / batch_size)] learning rates (Optimizer::apply_schedule)
 *   loss_per_sample optional [num_sample] (the reference's `loss` buffer, gpu/graph.cuh:91-92)
 *   loss_per_batch  optional [num batches], ACCUMULATED (+=) sums of the per-sample loss
 *   num_warps      0 = persistent grid sized to the device; 1 = single-warp deterministic order
 */
int gv_cuda_train_block(const gv_matrices_t *matrices, const uint32_t *pool, uint64_t num_sample, int num_negative,
                        const uint32_t *negatives, const double *random, const gv_alias_entry_t *negative_table,
                        uint32_t negative_count, uint32_t *negatives_out, const gv_device_optimizer_t *optimizer,
                        const float *lr_per_batch, uint32_t batch_size, float negative_weight,
                        float *loss_per_sample, float *loss_per_batch, int num_warps, void *stream);

/* The reference's random stream: cuRAND XORWOW uniform doubles exactly as
 * curandCreateGenerator(CURAND_RNG_PSEUDO_DEFAULT) + curandSetPseudoRandomGeneratorSeed(seed) +
 * curandGenerateUniformDouble produce them (core/solver.h:950-953,966,1247-1250,1536), generated by
 * our own kernel: stream position n = the (n / 4096)-th double of XORWOW subsequence n % 4096.
 * Consecutive gv_rng_generate calls continue the stream; save / restore snapshot it on the device. */
typedef struct gv_rng gv_rng_t;
gv_rng_t *gv_rng_create(unsigned long long seed, void *stream);
void gv_rng_destroy(gv_rng_t *rng);
int gv_rng_generate(gv_rng_t *rng, double *out, uint64_t n, void *stream);
uint64_t gv_rng_position(const gv_rng_t *rng);
size_t gv_rng_state_bytes(void);
int gv_rng_save(const gv_rng_t *rng, void *snapshot, void *stream);
int gv_rng_restore(gv_rng_t *rng, const void *snapshot, void *stream);

/* Device-layer tunables: "hot_rows" = rows with a local id below this value are read through L1
 * (default 128; rows are in degree order, so these are the hubs), "kernel_flags" = experiment bits. */
int gv_cuda_set_tunable(const char *name, long value);
/* current value of a tunable, or -1 (and gv_last_error) for an unknown name */
long gv_cuda_get_tunable(const char *name);

/* gpu::Sample (base/alias_table.cuh:175-183): out[t] = table.sample(float(random[2t]), float(random[2t+1])). */
int gv_cuda_sample_negatives(const gv_alias_entry_t *table, uint32_t count, const double *random, uint64_t num,
                             uint32_t *out, void *stream);

/* gpu::graph::predict (instance/gpu/graph.cuh:250-279): logits[i] = <vertex[head], context[tail]>,
 * batch = pairs {tail, head}. */
int gv_cuda_predict(int dim, const float *vertex, const float *context, const uint32_t *batch, uint64_t num,
                    float *logits, void *stream);

/* One chain position as the samplers store it: head/tail_locations[v] (core/solver.h:399-410). */
typedef struct {
    uint32_t part;    /* partition id */
    uint32_t local;   /* row inside the partition block */
} gv_location_t;

/* Device CSR of the flattened graph plus the sampler tables
 * (core/graph.h:87-101 flatten(); instance/graph.cuh:645-653 vertex_edge_tables;
 *  core/solver.h:123 edge_table).  All arrays are device pointers. */
typedef struct {
    uint32_t num_vertex;
    uint64_t num_edge;                       /* directed edges after flatten() */
    const uint64_t *offsets;                 /* [num_vertex + 1] flat_offsets (+ end) */
    const uint32_t *edge_u, *edge_v;         /* [num_edge] endpoints in flatten() order */
    const float *edge_prob;                  /* [num_edge] edge_table.prob_table */
    const uint64_t *edge_alias;              /* [num_edge] edge_table.alias_table (Index = size_t) */
    const gv_alias_entry_t *vertex_tables;   /* [num_edge] per-vertex alias tables at offsets[v]; may be NULL */
    const gv_location_t *locations;          /* [num_vertex] head_locations == tail_locations */
} gv_device_graph_t;

/* Walk part of GraphSampler::sample_random_walk (instance/graph.cuh:400-425) for `num_walk`
 * independent walks.  `random` holds consecutive refill buffers of `buffer_doubles` cuRAND doubles
 * (kRandBatchSize = 5e6 in the reference, core/solver.h:52), each serving `walks_per_buffer` walks
 * (the reference refills when rand_id > 5e6 - 2L, instance/graph.cuh:401).  Walk first_walk + w reads
 * the 2*L doubles of its slot in the reference's (right-to-left) argument order and writes the
 * locations of its L+1 vertices to chains[j * num_walk + w], j = 0..L.  With walk_length == 1 this is
 * the draw part of SamplerMixin::sample (core/solver.h:1026-1040): one alias-sampled edge per "walk".
 * Walks must not hit dead ends (every vertex has an out-edge), see DESIGN.md. */
int gv_cuda_random_walk(const gv_device_graph_t *graph, const double *random, uint32_t num_walk, int walk_length,
                        uint64_t first_walk, uint32_t walks_per_buffer, uint64_t buffer_doubles,
                        gv_location_t *chains, void *stream);

/* GraphSolver::build_vertex_edge (instance/graph.cuh:645-653) on the device: the alias table of vertex v over its
 * out-edges, built from edge_weights[offsets[v] ..] into tables[offsets[v] ..] with the reference's FIFO pairing
 * order (include/base/alias_table.cuh:84-128; bit-identical tables).  scratch_little / scratch_large hold num_edge
 * entries each; they are only touched by tables whose weights are not uniform. */
int gv_cuda_vertex_tables_build(const gv_device_graph_t *graph, const float *edge_weights, gv_alias_entry_t *tables,
                                uint32_t *scratch_little, uint32_t *scratch_large, void *stream);

/* node2vec.  GraphSolver::build_edge_edge (instance/graph.cuh:656-677): the alias table of directed
 * edge e = (u -> v) covers the out-edges (v -> x) with weight w/p if x == u, w/q if u is not a
 * neighbour of x, w otherwise; it has deg(v) entries and lives at tables[table_offsets[e]].  One
 * launch builds the tables of edges [first_edge, first_edge + num_table) with the reference's FIFO
 * pairing order (bit-identical tables); scratch_little / scratch_large hold
 * table_offsets[first_edge + num_table] - table_offsets[first_edge] entries each.
 * sorted_neighbors = edge_v sorted inside every vertex's CSR range (membership test). */
int gv_cuda_node2vec_build(const gv_device_graph_t *graph, const float *edge_weights, const uint32_t *sorted_neighbors,
                           const unsigned long long *table_offsets, uint64_t first_edge, uint32_t num_table, float p,
                           float q, gv_alias_entry_t *tables, uint32_t *scratch_little, uint32_t *scratch_large,
                           void *stream);
/* The flat array of per-edge tables split over the ranks of a multi-GPU run (Sigma deg^2 entries do not fit one GPU
 * beyond small graphs): shard r holds the entries [first_entry[r], first_entry[r + 1]) -- whole tables -- and may be
 * peer memory mapped with CUDA IPC. */
#define GV_MAX_TABLE_SHARDS 16
typedef struct {
    int num_shard;
    const gv_alias_entry_t *shard[GV_MAX_TABLE_SHARDS];
    unsigned long long first_entry[GV_MAX_TABLE_SHARDS + 1];
} gv_table_shards_t;
/* Walk part of GraphSampler::sample_biased_random_walk (instance/graph.cuh:321-349); arguments as
 * gv_cuda_random_walk, steps drawn from the table of the edge the walk arrived by. */
int gv_cuda_biased_walk_sharded(const gv_device_graph_t *graph, const gv_table_shards_t *tables,
                                const unsigned long long *table_offsets, const double *random, uint32_t num_walk,
                                int walk_length, uint64_t first_walk, uint32_t walks_per_buffer,
                                uint64_t buffer_doubles, gv_location_t *chains, void *stream);
int gv_cuda_biased_walk(const gv_device_graph_t *graph, const gv_alias_entry_t *tables,
                        const unsigned long long *table_offsets, const double *random, uint32_t num_walk,
                        int walk_length, uint64_t first_walk, uint32_t walks_per_buffer, uint64_t buffer_doubles,
                        gv_location_t *chains, void *stream);

/* Pool fill (instance/graph.cuh:427-447 / core/solver.h:1041-1053): expands the chains into positive
 * pairs in stream order (walk-major, then j, then k = 1..augmentation_step) and appends each pair
 * to block [head part][tail part] of one sampler's slice [start, end) until that slice is full,
 * through the pseudo-shuffle map.  Stable: the r-th pair of a block in stream order lands at slice
 * offset r.
 *   pool_blocks   device array [P*P] of device pointers to the blocks (pairs {tail_local,
 *                 head_local}); a NULL block is counted but not written (owned by another rank)
 *   fill          device [P*P], in/out: pairs of this slice each block has been offered so far
 *   last_walk     device scalar, in/out: max (first_walk + w) over walks that completed a block
 *   scratch       device scratch of gv_cuda_fill_scratch_bytes(num_walk, P) bytes
 */
typedef struct {
    int num_partition;
    int walk_length;          /* chain length - 1; 1 for edge sampling */
    int augmentation_step;
    int shuffle_base;
    uint64_t pool_size;       /* episode_size * batch_size */
    uint64_t start, end;      /* this sampler's slice */
    /* knowledge graphs (SamplerMixin::sample with KnowledgeGraphSampler::get_attributes,
     * instance/knowledge_graph.cuh:300-302): device [num_walk] relation of every sampled edge, or NULL.
     * When set (walk_length must be 1) pool entries are 12-byte triplets {relation, tail_local, head_local}
     * -- the reference's std::tuple<Index, Index, Index> byte order -- instead of 8-byte pairs. */
    const uint32_t *attributes;
} gv_fill_params_t;

size_t gv_cuda_fill_scratch_bytes(uint32_t num_walk, int num_partition);
int gv_cuda_fill_pool(const gv_fill_params_t *params, const gv_location_t *chains, uint32_t num_walk,
                      uint64_t first_walk, uint32_t *const *pool_blocks, unsigned long long *fill,
                      unsigned long long *last_walk, void *scratch, void *stream);

/* The same pool fill split in steps, for pools spread over several GPUs (sampler i of the reference runs on
 * rank i mod W and writes slice i of every block, core/solver.h:617-625, into the pool of the rank that owns the
 * block): gv_cuda_fill_count leaves the per-walk histogram in `scratch` and the round's per-block totals in
 * `totals` [P*P]; gv_cuda_fill_advance turns them into the slice offsets `bases` [P*P] at which the round's pairs
 * start (bases = fill, fill += totals); gv_cuda_fill_scatter[_staged] emits the pairs.  pool_blocks may point into
 * peer GPUs' memory.  (gv_cuda_peer_exchange can stand in for gv_cuda_fill_advance when ONE sampler's walks are
 * split over the ranks: it also adds the lower ranks' totals.) */
int gv_cuda_fill_count(const gv_fill_params_t *params, const gv_location_t *chains, uint32_t num_walk, void *scratch,
                       unsigned long long *totals, void *stream);
int gv_cuda_fill_advance(int num_block, const unsigned long long *totals, unsigned long long *fill,
                         unsigned long long *bases, void *stream);
int gv_cuda_fill_scatter(const gv_fill_params_t *params, const gv_location_t *chains, uint32_t num_walk,
                         uint64_t first_walk, uint32_t *const *pool_blocks, const unsigned long long *bases,
                         unsigned long long *last_walk, void *scratch, void *stream);
/* gv_cuda_fill_scatter for pools that live on peer GPUs: pairs of the blocks flagged in remote_blocks [P*P] (device)
 * are first written to `staging` (gv_cuda_fill_staging_bytes; local memory) in append order and then moved to their
 * shuffled positions with coalesced stores -- scattered 8-byte stores over NVLink are one request each.  `totals` =
 * this rank's output of gv_cuda_fill_count, stage_offsets = scratch of P*P entries.  remote_blocks == NULL is
 * gv_cuda_fill_scatter.  Pairs only (no attributes). */
size_t gv_cuda_fill_staging_bytes(uint32_t num_walk, int walk_length, int augmentation_step);
int gv_cuda_fill_scatter_staged(const gv_fill_params_t *params, const gv_location_t *chains, uint32_t num_walk,
                                uint64_t first_walk, uint32_t *const *pool_blocks, const unsigned long long *bases,
                                unsigned long long *last_walk, void *scratch, const unsigned char *remote_blocks,
                                const unsigned long long *totals, void *staging, unsigned long long *stage_offsets,
                                void *stream);
/* All-gather of the per-block totals through NVLink peer memory, fused with the prefix over ranks:
 * publishes `totals` (NULL = zeros) and *last_walk into every rank's control region (`controls` [W]
 * device pointers, peers mapped with CUDA IPC), waits until all W ranks published round `round_id`
 * (strictly increasing), then bases[b] = fill[b] + sum_{r<rank} totals_r[b], fill[b] += sum_r totals_r[b],
 * *last_walk = max over ranks.  Control regions are gv_cuda_peer_control_bytes() long and start zeroed. */
size_t gv_cuda_peer_control_bytes(int world_size, int num_partition);
int gv_cuda_peer_exchange(int rank, int world_size, int num_partition, uint64_t round_id,
                          const unsigned long long *totals, unsigned long long *const *controls,
                          unsigned long long *control, unsigned long long *fill, unsigned long long *bases,
                          unsigned long long *last_walk, void *stream);

/* Knowledge-graph train / predict kernels (instance/gpu/knowledge_graph.cuh:38-366 with the models of
 * instance/model/knowledge_graph.h).  NOT yet validated on a GPU.
 *   matrices   head and tail entity blocks [rows][dim] (the SAME pointers when the two partitions
 *              coincide, core/solver.h:1351-1355) and the relation matrix [num_relation][dim]
 *              (RotatE uses the first dim/2 floats of a row as phases); moments as the optimizer needs
 *   num_head   rows of the head block: a negative id below it replaces the head, otherwise id - num_head
 *              replaces the tail (gpu/knowledge_graph.cuh:64-71)
 *   batch      device [n][3] {relation, tail_local, head_local} (the reference's tuple byte order)
 *   negatives  device [n][k] ids, or NULL to draw them from `random` (2 doubles per negative, consumed
 *              like WorkerMixin::train_batch + gpu::Sample) uniformly over negative_count ids
 *   loss_*     optional: per-sample loss (sample_loss / 2) and its per-batch sum (atomicAdd)
 *   num_group  0 = fill the device; 1 = one thread group, samples in order (deterministic)
 */
typedef struct {
    int dim;
    uint32_t num_head;
    float *head, *tail, *relation;
    float *head_m1, *tail_m1, *relation_m1;
    float *head_m2, *tail_m2, *relation_m2;
} gv_kg_matrices_t;

int gv_cuda_kg_train_block(const gv_kg_matrices_t *matrices, int model, const uint32_t *batch, uint64_t num_sample,
                           int num_negative, const uint32_t *negatives, const double *random, uint32_t negative_count,
                           uint32_t *negatives_out, const gv_device_optimizer_t *optimizer, const float *lr_per_batch,
                           uint32_t batch_size, float relation_lr_multiplier, float margin_or_l3,
                           float adversarial_temperature, float *loss_per_sample, float *loss_per_batch, int num_group,
                           void *stream);
/* logits[i] = Model::forward(head[batch[i].head], tail[batch[i].tail], relation[batch[i].relation], margin) */
int gv_cuda_kg_predict(const gv_kg_matrices_t *matrices, int model, const uint32_t *batch, uint64_t num_sample,
                       float margin, float *logits, void *stream);

/* Positive-sample draw of the knowledge-graph sampler (SamplerMixin::sample, core/solver.h:1036-1044, with
 * KnowledgeGraphSampler::get_attributes, instance/knowledge_graph.cuh:300-302): draw d consumes
 * random[2d] (accept, narrowed to float) and random[2d+1] (index) like edge_table.sample(random[r++],
 * random[r++]) does under gcc; chains [2][num_draw] receive the (partition, local row) of head and tail,
 * relations [num_draw] the edge's relation.  Feed both to gv_cuda_fill_pool (walk_length 1, attributes). */
typedef struct {
    uint64_t num_edge;
    const uint32_t *edge_h, *edge_t, *edge_r;   /* flatten() order */
    const float *edge_prob;                     /* AliasTable<float, size_t> over the triplet weights */
    const uint64_t *edge_alias;
    const gv_location_t *locations;             /* entity -> (partition, local row); tied head / tail */
} gv_device_kgraph_t;
int gv_cuda_kg_draw(const gv_device_kgraph_t *graph, const double *random, uint32_t num_draw, gv_location_t *chains,
                    uint32_t *relations, void *stream);
/* Write-back of the global relation matrix (WorkerMixin::write_embedding for kGlobal, core/solver.h:1413-1420:
 * global -= loaded - trained), split so that several workers' deltas can be summed in between:
 * delta = global - work;  [all-reduce(sum) of delta over the ranks];  global -= delta, work = global. */
int gv_cuda_kg_relation_delta(const float *global, const float *work, float *delta, uint64_t n, void *stream);
int gv_cuda_kg_relation_apply(float *global, float *work, const float *delta, uint64_t n, void *stream);

/* Memory::gather / Memory::scatter (base/memory.h:194-217) on the device: rows of `dim` floats,
 * dst[i] = src[ids[i]] when gather != 0, else dst[ids[i]] = src[i]. */
int gv_cuda_move_rows(float *dst, const float *src, const uint32_t *ids, uint64_t num_row, int dim, int gather,
                      void *stream);
/* Device-side construction of what the host would otherwise compute and upload for the samplers:
 * dst[i] = value; dst[i] = i (the alias column of AliasTable::build for uniform weights, base/alias_table.cuh:84-128:
 * every entry is a leftover that aliases to itself); edge_u[e] = v for offsets[v] <= e < offsets[v + 1] (the source
 * column of GraphMixin::flatten's edge list, core/graph.h:87-101). */
int gv_cuda_fill_float(float *dst, uint64_t n, float value, void *stream);
int gv_cuda_fill_identity(uint64_t *dst, uint64_t n, void *stream);
int gv_cuda_expand_sources(const uint64_t *offsets, uint32_t num_vertex, uint32_t *edge_u, void *stream);

/* ---- host runtime: Graph (instance/graph.cuh:62-277, bind.h:109-187) ------------------- */

typedef struct gv_graph gv_graph_t;

gv_graph_t *gv_graph_create(void);
void gv_graph_destroy(gv_graph_t *graph);
/* Graph::load_file (instance/graph.cuh:163-201) */
int gv_graph_load_file(gv_graph_t *graph, const char *file_name, int as_undirected, int normalization,
                       const char *delimiters, const char *comment);
/* WordGraph::load_file_compact (instance/word_graph.cuh:75-166, bind.h:216-230): word co-occurrence graph of a
 * corpus; words rarer than min_count are dropped, pairs at distance <= window inside a line count as edges */
int gv_graph_load_corpus(gv_graph_t *graph, const char *file_name, int window, int min_count, int normalization,
                         const char *delimiters, const char *comment);
/* Graph::load_edge_list / load_weighted_edge_list (instance/graph.cuh:209-252); weights may be NULL */
int gv_graph_load_edges(gv_graph_t *graph, const char *const *u_names, const char *const *v_names,
                        const float *weights, uint64_t num_edge, int as_undirected, int normalization);
/* The same graph as gv_graph_load_edges on the edge list [(str(u[i]), str(v[i]))] -- first-seen ids, edge order,
 * line count -- from two uint32 arrays (binary edge lists: Friendster-sized inputs without 3.6e9 strings). */
int gv_graph_load_id_edges(gv_graph_t *graph, const uint32_t *u, const uint32_t *v, const float *weights,
                           uint64_t num_edge, int as_undirected, int normalization);
/* Graph::save (instance/graph.cuh:260-277) */
int gv_graph_save(gv_graph_t *graph, const char *file_name, int weighted, int anonymous);
uint64_t gv_graph_num_vertex(const gv_graph_t *graph);
uint64_t gv_graph_num_edge(const gv_graph_t *graph);      /* counts input lines (graph.cuh:152) */
int gv_graph_as_undirected(const gv_graph_t *graph);
int gv_graph_normalization(const gv_graph_t *graph);
const char *gv_graph_id2name(const gv_graph_t *graph, uint64_t id);
int64_t gv_graph_name2id(const gv_graph_t *graph, const char *name); /* -1 if absent */
/* GraphMixin::flatten (core/graph.h:87-101): returns #directed edges; arrays may be NULL */
uint64_t gv_graph_flatten(gv_graph_t *graph, uint32_t *u, uint32_t *v, float *w, uint64_t *flat_offsets,
                          float *vertex_weights);
/* Graph::info() */
int gv_graph_info(const gv_graph_t *graph, char *buffer, size_t capacity);

/* ---- host runtime: KnowledgeGraph (instance/knowledge_graph.cuh:67-284, bind.h:237-314) ---------
 * Entities and relations get ids in order of first appearance (head, relation, tail of each line). */

typedef struct gv_kgraph gv_kgraph_t;

gv_kgraph_t *gv_kgraph_create(void);
void gv_kgraph_destroy(gv_kgraph_t *graph);
/* KnowledgeGraph::load_file (instance/knowledge_graph.cuh:177-213): `head relation tail [weight]` per line */
int gv_kgraph_load_file(gv_kgraph_t *graph, const char *file_name, int normalization, const char *delimiters,
                        const char *comment);
/* load_triplet_list / load_weighted_triplet_list (instance/knowledge_graph.cuh:220-260); weights may be NULL */
int gv_kgraph_load_triplets(gv_kgraph_t *graph, const char *const *h_names, const char *const *r_names,
                            const char *const *t_names, const float *weights, uint64_t num_triplet,
                            int normalization);
/* KnowledgeGraph::save (instance/knowledge_graph.cuh:267-283): columns head, tail, relation.  The
 * reference takes the third column from the edge weight by mistake (:275); we write the relation. */
int gv_kgraph_save(gv_kgraph_t *graph, const char *file_name, int anonymous);
uint64_t gv_kgraph_num_vertex(const gv_kgraph_t *graph);
uint64_t gv_kgraph_num_edge(const gv_kgraph_t *graph);
uint64_t gv_kgraph_num_relation(const gv_kgraph_t *graph);
int gv_kgraph_normalization(const gv_kgraph_t *graph);
const char *gv_kgraph_id2entity(const gv_kgraph_t *graph, uint64_t id);
const char *gv_kgraph_id2relation(const gv_kgraph_t *graph, uint64_t id);
int64_t gv_kgraph_entity2id(const gv_kgraph_t *graph, const char *name);   /* -1 if absent */
int64_t gv_kgraph_relation2id(const gv_kgraph_t *graph, const char *name); /* -1 if absent */
/* GraphMixin::flatten (core/graph.h:87-101) with the relation attribute: returns #triplets; arrays may be NULL */
uint64_t gv_kgraph_flatten(gv_kgraph_t *graph, uint32_t *h, uint32_t *t, uint32_t *r, float *w,
                           uint64_t *flat_offsets, float *vertex_weights);
/* KnowledgeGraph::info() */
int gv_kgraph_info(const gv_kgraph_t *graph, char *buffer, size_t capacity);

/* AliasTable::build (base/alias_table.cuh:84-128); alias is uint64 (Index = size_t) */
int gv_alias_build(const float *weights, uint64_t n, float *prob, uint64_t *alias);

/* ---- host runtime: GraphSolver (instance/graph.cuh:587-813, core/solver.h, bind.h:383-513) ---- */

typedef struct gv_solver gv_solver_t;

/* GraphSolver(device_ids, num_sampler_per_worker, gpu_memory_limit) (bind.h:445-447).
 * dim in {32,64,96,128,256,512} (src/graphvite.cu:52-59); 0 for the two autos.
 * Distributed (one process per GPU): world_size > 1, `rank` in [0, world_size), exactly one
 * device id; the caller then provides the block exchange with gv_solver_set_exchange(). */
gv_solver_t *gv_solver_create(int dim, const int *device_ids, int num_device, int num_sampler_per_worker,
                              uint64_t gpu_memory_limit, int rank, int world_size);
void gv_solver_destroy(gv_solver_t *solver);

/* Block exchange for world_size > 1 (replaces WorkerMixin::write_embedding/load_embedding through
 * host memory, core/solver.h:1349-1428): send `bytes` from device pointer `send` to rank `dst` and
 * receive into `recv` from rank `src`, on CUDA stream `stream`; either side may be a no-op
 * (dst/src = -1).  Python wires this to torch.distributed NCCL P2P over NVLink. */
typedef int (*gv_exchange_fn)(const void *send, int dst, void *recv, int src, uint64_t bytes, void *stream,
                              void *ctx);
int gv_solver_set_exchange(gv_solver_t *solver, gv_exchange_fn fn, void *ctx);
/* Host all-gather used once in build() to trade CUDA IPC handles of the sample-pool arenas, so that
 * the samplers can be partitioned over the ranks and scatter pairs straight into the owner's pool
 * over NVLink.  recv holds world_size * bytes.  Without it every rank samples all blocks itself. */
typedef int (*gv_host_allgather_fn)(const void *send, void *recv, uint64_t bytes, void *ctx);
int gv_solver_set_host_allgather(gv_solver_t *solver, gv_host_allgather_fn fn, void *ctx);
/* Unmap the peers' arenas (collective teardown: every rank calls this, then a barrier, then
 * gv_solver_destroy / gv_solver_build; CUDA IPC memory must be closed by all importers before the
 * exporter frees it).  The solver needs build() again before it can train. */
int gv_solver_release_peers(gv_solver_t *solver);

/* SolverMixin::build (core/solver.h:287-466); num_partition / episode_size 0 = auto */
int gv_solver_build(gv_solver_t *solver, gv_graph_t *graph, const gv_optimizer_t *optimizer, int num_partition,
                    int num_negative, int batch_size, int episode_size);
/* GraphSolver::train (instance/graph.cuh:770-793); augmentation_step / shuffle_base 0 = auto */
int gv_solver_train(gv_solver_t *solver, const char *model, int num_epoch, int resume, int augmentation_step,
                    int random_walk_length, int random_walk_batch_size, int shuffle_base, float p, float q,
                    int positive_reuse, float negative_sample_exponent, float negative_weight, int log_frequency);
/* SolverMixin::predict_numpy (core/solver.h:729-802); pairs = (v, c) global ids, HOST memory */
int gv_solver_predict(gv_solver_t *solver, const uint32_t *pairs, uint64_t num, float *logits);
/* SolverMixin::clear (core/solver.h:805-816) */
int gv_solver_clear(gv_solver_t *solver);
/* numpy views (bind.h:90-106,439-442): solver-owned HOST matrix [num_vertex][dim]; which 0 vertex, 1 context */
float *gv_solver_embeddings(gv_solver_t *solver, int which, uint64_t *rows, int *dim);
/* read-only attributes (bind.h:415-436) as "key=value" lines, and SolverMixin::info() */
int gv_solver_info(const gv_solver_t *solver, char *buffer, size_t capacity);
int gv_solver_attributes(const gv_solver_t *solver, char *buffer, size_t capacity);
/* mean loss values the reference would LOG (core/solver.h:1541-1549), in order */
int gv_solver_logged_loss(const gv_solver_t *solver, float *out, int capacity);
/* throughput counters of the last train(): positives consumed and device seconds in the train kernels */
int gv_solver_stats(const gv_solver_t *solver, double *out, int capacity);

/* ---- host runtime: KnowledgeGraphSolver (instance/knowledge_graph.cuh:531-677, core/solver.h,
 * bind.h:516-639).  Same conventions as gv_solver_*: one process drives one GPU, world_size processes form
 * the reference's num_worker.  Entity embeddings are ONE matrix used for heads and tails (tied weights,
 * protocols kHeadPartition|kInPlace / kTailPartition|kInPlace|kSharedWithPredecessor); the relation matrix is
 * global: every worker trains a private copy and the copies are reconciled by summing their deltas after
 * every schedule step (the caller-provided all-reduce, NCCL over NVLink), which replaces the reference's
 * scatter_sub through host memory. */
typedef struct gv_kg_solver gv_kg_solver_t;

/* dim even, <= 2048 (the reference instantiates 32..2048, src/graphvite.cu:61-70); 0 for the two autos */
gv_kg_solver_t *gv_kg_solver_create(int dim, const int *device_ids, int num_device, int num_sampler_per_worker,
                                    uint64_t gpu_memory_limit, int rank, int world_size);
void gv_kg_solver_destroy(gv_kg_solver_t *solver);
/* world_size > 1: entity blocks move between the ranks through `exchange` (see gv_solver_set_exchange) and
 * the relation deltas are summed in place over all ranks by `allreduce` (count floats, device memory, on
 * `stream`) */
typedef int (*gv_allreduce_fn)(void *buffer, uint64_t count, void *stream, void *ctx);
int gv_kg_solver_set_exchange(gv_kg_solver_t *solver, gv_exchange_fn fn, void *ctx);
int gv_kg_solver_set_allreduce(gv_kg_solver_t *solver, gv_allreduce_fn fn, void *ctx);
/* SolverMixin::build; optimizer type -1 = Adam(5e-5, 0) (knowledge_graph.cuh:557-559); minimum
 * #partition is 1 for one worker, 2 * #worker otherwise (tied weights, core/solver.h:266-277) */
int gv_kg_solver_build(gv_kg_solver_t *solver, gv_kgraph_t *graph, const gv_optimizer_t *optimizer, int num_partition,
                       int num_negative, int batch_size, int episode_size);
/* KnowledgeGraphSolver::train (knowledge_graph.cuh:666-677) */
int gv_kg_solver_train(gv_kg_solver_t *solver, const char *model, int num_epoch, int resume,
                       float relation_lr_multiplier, float margin, float l3_regularization, int sample_batch_size,
                       int positive_reuse, float adversarial_temperature, int log_frequency);
/* SolverMixin::predict_numpy: triplets = HOST [n][3] global ids ordered (head, tail, relation) (bind.h:621-629) */
int gv_kg_solver_predict(gv_kg_solver_t *solver, const uint32_t *triplets, uint64_t num, float *logits);
int gv_kg_solver_clear(gv_kg_solver_t *solver);
/* numpy views (bind.h:569-572): which 0 = entity_embeddings [num_vertex][dim], 1 = relation_embeddings
 * [num_relation][dim] (RotatE uses the first dim / 2 floats of a row) */
float *gv_kg_solver_embeddings(gv_kg_solver_t *solver, int which, uint64_t *rows, int *dim);
int gv_kg_solver_info(const gv_kg_solver_t *solver, char *buffer, size_t capacity);
int gv_kg_solver_attributes(const gv_kg_solver_t *solver, char *buffer, size_t capacity);
int gv_kg_solver_logged_loss(const gv_kg_solver_t *solver, float *out, int capacity);
int gv_kg_solver_stats(const gv_kg_solver_t *solver, double *out, int capacity);
/* test hooks, as for gv_solver_*: staged training, pools ({relation, tail_local, head_local} triplets),
 * partition of the entities, the negatives drawn for the last batch, and the tied-weight schedule
 * (SolverMixin::get_schedule, core/solver.h:519-561: out[(step * W + worker) * 2 + {0, 1}] = head, tail
 * partition; returns #steps).  Options: "capture_negatives", "train_num_groups" (1 = one thread group,
 * samples in order: deterministic), "shuffle_partition" (-1 = the reference's rule, 0 / 1 forces it). */
int gv_kg_solver_train_begin(gv_kg_solver_t *solver, const char *model, int num_epoch, int resume,
                             float relation_lr_multiplier, float margin, float l3_regularization,
                             int sample_batch_size, int positive_reuse, float adversarial_temperature,
                             int log_frequency);
int gv_kg_solver_train_episode(gv_kg_solver_t *solver);
int gv_kg_solver_train_end(gv_kg_solver_t *solver);
int gv_kg_solver_set_option(gv_kg_solver_t *solver, const char *name, int value);
int gv_kg_solver_locations(const gv_kg_solver_t *solver, uint32_t *part_of, uint32_t *local_of);
int64_t gv_kg_solver_pool(gv_kg_solver_t *solver, int pool, int head_partition, int tail_partition, uint32_t *out);
int gv_kg_solver_last_negatives(gv_kg_solver_t *solver, uint32_t *out);
int gv_kg_schedule(int num_partition, int num_worker, int *out, int capacity);

/* ---- test hooks (no reference counterpart: the reference's members are simply public) ------- */
/* SolverMixin::get_schedule (core/solver.h:519-575) plus the vertex-block movement it implies:
 * for `num_episode` episodes, out[((e * steps + s) * W + rank) * 6 + i] = head, tail, source rank of
 * the head block, block given away (-1 none), its destination rank, #head blocks held afterwards.
 * Returns the number of steps per episode. */
int gv_schedule_plan(int num_partition, int num_worker, int num_episode, int *out, int capacity);
/* the process-wide engine core/solver.h:50: re-seed (5489 = default-constructed) */
void gv_reset_global_engine(uint32_t seed);
/* the bulk generator behind init_embeddings against libstdc++'s std::mt19937 + distributions (0 = identical) */
int gv_engine_self_check(uint32_t seed, uint64_t bulk);
/* head_locations (core/solver.h:399-410) */
int gv_solver_locations(const gv_solver_t *solver, uint32_t *part_of, uint32_t *local_of);
/* copy one sample-pool block (pairs {tail, head}) to host; returns #pairs */
int64_t gv_solver_pool(gv_solver_t *solver, int pool, int head_partition, int tail_partition, uint32_t *out);
/* set-up of train() up to and including the first pool fill (core/solver.h:588-628), then single episodes */
int gv_solver_train_begin(gv_solver_t *solver, const char *model, int num_epoch, int resume, int augmentation_step,
                          int random_walk_length, int random_walk_batch_size, int shuffle_base, float p, float q,
                          int positive_reuse, float negative_sample_exponent, float negative_weight,
                          int log_frequency);
int gv_solver_train_episode(gv_solver_t *solver);   /* 1 = trained one episode, 0 = done, <0 error */
/* one schedule step (sub-episode): every worker trains one (head, tail) block; 1 / 0 / <0 like above */
int gv_solver_train_step(gv_solver_t *solver);
/* CUDA-event stopwatch on the solver's work stream: stop = 0 starts it, stop = 1 returns the seconds */
double gv_solver_device_timer(gv_solver_t *solver, int stop);
int gv_solver_train_end(gv_solver_t *solver);       /* write_back (core/solver.h:650-653) */
/* options: "capture_negatives" = 1 keeps the negatives the train kernel drew for the last batch;
 * "train_num_warps" = 1 runs the train kernel on a single warp (sequential order, reproducible) */
int gv_solver_set_option(gv_solver_t *solver, const char *name, int value);
/* negatives drawn for the last trained batch of this rank's worker, [batch_size][num_negative] */
int gv_solver_last_negatives(gv_solver_t *solver, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* GV_B200_H_ */
