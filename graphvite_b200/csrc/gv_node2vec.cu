// =============================================================================
// gv_node2vec.cu -- second-order (node2vec) sampler tables and walker on the device.
//
// Replaces GraphSolver::build_edge_edge (reference include/instance/graph.cuh:656-677), which
// builds one AliasTable object per DIRECTED edge on CPU threads (Sigma deg^2 entries, > 200 GiB of
// host memory beyond Youtube, doc/source/benchmark.rst:53-54), and the walk part of
// GraphSampler::sample_biased_random_walk (instance/graph.cuh:298-373).
//
// Layout: the table of edge e = (u -> v) has deg(v) entries {prob, alias} and starts at
// table_offsets[e] (prefix sum of deg(v_e) in flatten() order) inside one flat device array.
// The per-vertex tables of the first-order walk (build_vertex_edge, graph.cuh:645-653) are built by the same code.
// Build: one thread per table runs Vose's method with the reference's FIFO pairing order
// (include/base/alias_table.cuh:84-128), so the tables -- and therefore the walks -- are
// bit-identical; the two FIFO queues live in a scratch ring of deg(v) entries each.
// =============================================================================
#include <cuda_runtime.h>

#include <algorithm>

#include <cstdint>
#include <string>

#include "gv_common.h"
#include "gv_device.cuh"

namespace gv {
namespace device {

// is `target` in the sorted neighbour list [begin, end)?
__device__ __forceinline__ bool contains(const uint32_t *sorted, unsigned long long begin, unsigned long long end,
                                         uint32_t target) {
    while (begin < end) {
        const unsigned long long middle = (begin + end) >> 1;
        const uint32_t value = __ldg(sorted + middle);
        if (value == target)
            return true;
        if (value < target)
            begin = middle + 1;
        else
            end = middle;
    }
    return false;
}

// AliasTable::build, include/base/alias_table.cuh:84-128, for one table whose entries hold the raw weights in
// .prob: normalise, then Vose's method with the reference's two FIFO queues (rings of capacity `count`: an index is
// in at most one queue).  `norm` = sum of the weights, accumulated in double in entry order (alias_table.cuh:92).
__device__ __forceinline__ void finish_alias_table(gv_alias_entry_t *table, uint32_t count, double norm,
                                                   uint32_t *little_ring, uint32_t *large_ring) {
    norm = norm / count;
    uint32_t num_little = 0;
    for (uint32_t i = 0; i < count; i++) {
        const float prob = float(double(table[i].prob) / norm);
        table[i].prob = prob;
        table[i].alias = i;
        num_little += prob < 1;
    }
    // one queue starts empty (every table of an unweighted graph): the pairing loop never runs and all entries are
    // leftovers that alias to themselves -- the rings are not touched
    if (num_little == 0 || num_little == count)
        return;
    uint32_t little_head = 0, little_tail = 0, large_head = 0, large_tail = 0;
    for (uint32_t i = 0; i < count; i++) {
        if (table[i].prob < 1)
            little_ring[little_tail++ % count] = i;
        else
            large_ring[large_tail++ % count] = i;
    }
    while (little_head != little_tail && large_head != large_tail) {
        const uint32_t i = little_ring[little_head++ % count], j = large_ring[large_head++ % count];
        table[i].alias = j;
        const float sum = table[i].prob + table[j].prob;
        const float rest = sum - 1;
        table[j].prob = rest;
        if (rest < 1)
            little_ring[little_tail++ % count] = j;
        else
            large_ring[large_tail++ % count] = j;
    }
    // leftovers keep the self alias written above ("suppress some truncation error", alias_table.cuh:117-127)
}

__global__ void __launch_bounds__(128) node2vec_build_kernel(const gv_device_graph_t g, const float *edge_w,
                                                             const uint32_t *sorted_v,
                                                             const unsigned long long *table_offsets,
                                                             unsigned long long first_edge, uint32_t num_table,
                                                             float p, float q, gv_alias_entry_t *tables,
                                                             uint32_t *little, uint32_t *large) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_table)
        return;
    const unsigned long long e = first_edge + t;
    const uint32_t u = __ldg(g.edge_u + e), v = __ldg(g.edge_v + e);
    const unsigned long long v_begin = __ldg(g.offsets + v);
    const uint32_t count = uint32_t(__ldg(g.offsets + v + 1) - v_begin);
    if (count == 0)
        return;
    gv_alias_entry_t *table = tables + table_offsets[e];
    const unsigned long long ring = table_offsets[e] - table_offsets[first_edge];

    // build_edge_edge, graph.cuh:660-670: w / p back to u, w / q to non-neighbours of u, w otherwise
    double norm = 0;
    for (uint32_t i = 0; i < count; i++) {
        const uint32_t x = __ldg(g.edge_v + v_begin + i);
        const float w = __ldg(edge_w + v_begin + i);
        float weight;
        if (x == u)
            weight = w / p;
        else if (!contains(sorted_v, __ldg(g.offsets + x), __ldg(g.offsets + x + 1), u))
            weight = w / q;
        else
            weight = w;
        table[i].prob = weight;
        norm += weight;
    }
    finish_alias_table(table, count, norm, little + ring, large + ring);
}

// GraphSolver::build_vertex_edge, instance/graph.cuh:645-653: one alias table per vertex over its out-edges, laid
// out at the vertex's CSR range (and so are its two rings).  Thread per vertex.
__global__ void __launch_bounds__(128) vertex_tables_kernel(const gv_device_graph_t g, const float *edge_w,
                                                            gv_alias_entry_t *tables, uint32_t *little,
                                                            uint32_t *large) {
    const unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.num_vertex)
        return;
    const unsigned long long begin = __ldg(g.offsets + v);
    const uint32_t count = uint32_t(__ldg(g.offsets + v + 1) - begin);
    if (count == 0)
        return;
    gv_alias_entry_t *table = tables + begin;
    double norm = 0;
    for (uint32_t i = 0; i < count; i++) {
        const float w = __ldg(edge_w + begin + i);
        table[i].prob = w;
        norm += w;
    }
    finish_alias_table(table, count, norm, little + begin, large + begin);
}

template<class Count>
__device__ __forceinline__ Count alias_slot(double rand1, Count count) {
    Count index = Count(rand1 * double(count));
    return index < count ? index : count - 1;
}

// The flat table array split over the ranks of a multi-GPU run: shard r holds the entries
// [first_entry[r], first_entry[r + 1]) (whole tables) and may live in a peer GPU's memory (CUDA IPC over NVLink).
// One shard covering everything is the single-GPU case.
__device__ __forceinline__ uint2 load_table_entry(const gv_table_shards_t &shards, unsigned long long entry) {
    int r = 0;
    while (r + 1 < shards.num_shard && entry >= shards.first_entry[r + 1])
        r++;
    const uint2 *shard = reinterpret_cast<const uint2 *>(shards.shard[r]);
    return shard[entry - shards.first_entry[r]];  // a plain load: the shard may be peer memory
}

// walk part of sample_biased_random_walk, graph.cuh:321-349
__global__ void __launch_bounds__(256) biased_walk_kernel(const gv_device_graph_t g, const gv_table_shards_t tables,
                                                          const unsigned long long *table_offsets,
                                                          const double *random, uint32_t num_walk, int walk_length,
                                                          uint64_t first_walk, uint32_t walks_per_buffer,
                                                          uint64_t buffer_doubles, gv_location_t *chains) {
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < num_walk; w += gridDim.x * blockDim.x) {  // see gv_sampler.cu
    const uint64_t walk = first_walk + w;
    const double2 *r = reinterpret_cast<const double2 *>(random + (walk / walks_per_buffer) * buffer_doubles) +
                       (walk % walks_per_buffer) * walk_length;
    uint2 *out = reinterpret_cast<uint2 *>(chains) + w;
    const uint2 *locations = reinterpret_cast<const uint2 *>(g.locations);
    double2 draw = __ldcs(r);  // .x = random[r] (accept), .y = random[r+1] (index): right-to-left evaluation
    unsigned long long index = alias_slot<unsigned long long>(draw.y, g.num_edge);
    unsigned long long edge = float(draw.x) < __ldg(g.edge_prob + index) ? index : __ldg(g.edge_alias + index);
    uint32_t current = __ldg(g.edge_u + edge);
    out[0] = __ldg(locations + current);
    current = __ldg(g.edge_v + edge);
    out[num_walk] = __ldg(locations + current);
    for (int j = 2; j <= walk_length; j++) {
        const unsigned long long begin = __ldg(g.offsets + current);
        const uint32_t degree = uint32_t(__ldg(g.offsets + current + 1) - begin);
        if (degree == 0) {  // dead end: refused by the host; never read out of bounds
            for (; j <= walk_length; j++)
                out[size_t(j) * num_walk] = __ldg(locations + current);
            break;
        }
        draw = __ldcs(r + j - 1);
        const uint32_t slot = alias_slot<uint32_t>(draw.y, degree);
        const uint2 entry = load_table_entry(tables, __ldg(table_offsets + edge) + slot);
        const uint32_t neighbor = float(draw.x) < __uint_as_float(entry.x) ? slot : entry.y;
        edge = begin + neighbor;  // edge_id = flat_offsets[current] + neighbor_id, graph.cuh:341
        current = __ldg(g.edge_v + edge);
        out[size_t(j) * num_walk] = __ldg(locations + current);
    }
    }
}

}  // namespace device
}  // namespace gv

using namespace gv;
using namespace gv::device;

extern "C" {

int gv_cuda_node2vec_build(const gv_device_graph_t *graph, const float *edge_weights, const uint32_t *sorted_neighbors,
                           const unsigned long long *table_offsets, uint64_t first_edge, uint32_t num_table, float p,
                           float q, gv_alias_entry_t *tables, uint32_t *scratch_little, uint32_t *scratch_large,
                           void *stream) {
    if (num_table == 0)
        return 0;
    if (!graph || !edge_weights || !sorted_neighbors || !table_offsets || !tables || !scratch_little || !scratch_large)
        return fail("gv_cuda_node2vec_build: null argument");
    GV_LAUNCH((num_table + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), node2vec_build_kernel)(
        *graph, edge_weights, sorted_neighbors, table_offsets, first_edge, num_table, p, q, tables, scratch_little,
        scratch_large);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gv_cuda_vertex_tables_build(const gv_device_graph_t *graph, const float *edge_weights, gv_alias_entry_t *tables,
                                uint32_t *scratch_little, uint32_t *scratch_large, void *stream) {
    if (!graph || !edge_weights || !tables || !scratch_little || !scratch_large)
        return fail("gv_cuda_vertex_tables_build: null argument");
    if (graph->num_vertex == 0)
        return 0;
    GV_LAUNCH(unsigned((uint64_t(graph->num_vertex) + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream),
              vertex_tables_kernel)(*graph, edge_weights, tables, scratch_little, scratch_large);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gv_cuda_biased_walk_sharded(const gv_device_graph_t *graph, const gv_table_shards_t *tables,
                                const unsigned long long *table_offsets, const double *random, uint32_t num_walk,
                                int walk_length, uint64_t first_walk, uint32_t walks_per_buffer,
                                uint64_t buffer_doubles, gv_location_t *chains, void *stream) {
    if (num_walk == 0)
        return 0;
    if (!graph || !tables || !table_offsets || !random || !chains || walk_length < 1 || walks_per_buffer == 0 ||
        buffer_doubles < uint64_t(walks_per_buffer) * 2 * walk_length || buffer_doubles % 2 != 0 ||
        tables->num_shard < 1 || tables->num_shard > GV_MAX_TABLE_SHARDS)
        return fail("gv_cuda_biased_walk: invalid argument");
    for (int r = 0; r < tables->num_shard; r++)
        if (!tables->shard[r] && tables->first_entry[r + 1] > tables->first_entry[r])
            return fail("gv_cuda_biased_walk: a non-empty table shard has no memory");
    uint32_t blocks = (num_walk + 255) / 256;
    if (gv::sampler_max_ctas() > 0)
        blocks = std::min<uint32_t>(blocks, uint32_t(gv::sampler_max_ctas()));
    GV_LAUNCH(blocks, 256, 0, static_cast<cudaStream_t>(stream), biased_walk_kernel)(
        *graph, *tables, table_offsets, random, num_walk, walk_length, first_walk, walks_per_buffer, buffer_doubles,
        chains);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gv_cuda_biased_walk(const gv_device_graph_t *graph, const gv_alias_entry_t *tables,
                        const unsigned long long *table_offsets, const double *random, uint32_t num_walk,
                        int walk_length, uint64_t first_walk, uint32_t walks_per_buffer, uint64_t buffer_doubles,
                        gv_location_t *chains, void *stream) {
    if (!tables)
        return fail("gv_cuda_biased_walk: invalid argument");
    gv_table_shards_t one;
    one.num_shard = 1;
    one.shard[0] = tables;
    one.first_entry[0] = 0;
    one.first_entry[1] = ~0ull;
    return gv_cuda_biased_walk_sharded(graph, &one, table_offsets, random, num_walk, walk_length, first_walk,
                                       walks_per_buffer, buffer_doubles, chains, stream);
}

}  // extern "C"
