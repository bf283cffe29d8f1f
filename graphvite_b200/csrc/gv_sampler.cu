// =============================================================================
// gv_sampler.cu -- the positive-sample path on the device.
//
// Replaces the CPU sampler threads of the reference:
//   GraphSampler::sample_random_walk   include/instance/graph.cuh:376-450
//   SamplerMixin::sample               include/core/solver.h:1011-1055
//   AliasTable::sample (CPU call shape) include/base/alias_table.cuh:148-152
//
// The reference walks sequentially, but with no dead ends every walk consumes exactly
// 2*L doubles, so walk w owns random[2*L*w, 2*L*(w+1)) and walks are independent given
// the stream: one thread per walk.  The sequential part -- appending each pair to its
// (head part, tail part) block until that block's slice is full -- is a stable
// partition, done as count / scan / scatter so that the pools are bit-identical to the
// reference's (stream order preserved inside every block, pseudo-shuffle included).
//
// Memory traffic: a walk emits pairs_per_walk pairs (190 at L = 40, a = 5), so "one thread writes its walk's pairs"
// means 32 lanes storing 8 bytes each to places ~1.5 KB apart.  The default fill kernels therefore stage a CTA's walks
// in shared memory and write runs: fill_direct_tiled_kernel (one block: the offsets are analytic),
// fill_scatter_tiled_kernel (P x P blocks: per-thread scan, pairs parked per block), fill_forward_kernel (pairs of
// blocks owned by a peer GPU: staged locally, forwarded as 256-byte warp stores over NVLink).  The thread-per-walk
// kernels remain for edge sampling (one pair per walk: already coalesced), for entries with attributes (knowledge
// graphs) and as the comparison switch `fill_per_walk`.
// =============================================================================
#include <cuda_runtime.h>

#include <algorithm>

#include <cstdint>
#include <string>

#include "gv_common.h"
#include "gv_device.cuh"

namespace gv {
namespace device {

// AliasTable::sample with the CPU call shape table.sample(random[r++], random[r++]):
// gcc evaluates the arguments right to left, so rand1 (index draw) = random[r+1] and
// rand2 (accept draw) = random[r].  rand1 is NOT narrowed here.
template<class Count>
__device__ __forceinline__ Count alias_index(double rand1, Count count) {
    Count index = Count(rand1 * double(count));
    return index < count ? index : count - 1;  // cuRAND doubles lie in (0,1]; clamp rand1 == 1
}

__global__ void __launch_bounds__(256) random_walk_kernel(const gv_device_graph_t g, const double *random,
                                                          uint32_t num_walk, int walk_length, uint64_t first_walk,
                                                          uint32_t walks_per_buffer, uint64_t buffer_doubles,
                                                          gv_location_t *chains) {
    // one walk per thread and round: the grid may be capped (sampler_max_ctas) so that the walker, which is latency
    // bound, shares the device with a resident train launch instead of displacing it
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < num_walk; w += gridDim.x * blockDim.x) {
    // walk (first_walk + w) of the span: buffer (index / walks_per_buffer), slot (index % walks_per_buffer);
    // the unused tail of a refill buffer (when 5e6 is not a multiple of 2L) is skipped like the reference does
    const uint64_t walk = first_walk + w;
    const double2 *r = reinterpret_cast<const double2 *>(random + (walk / walks_per_buffer) * buffer_doubles) +
                       (walk % walks_per_buffer) * walk_length;
    uint2 *out = reinterpret_cast<uint2 *>(chains) + w;
    const uint2 *locations = reinterpret_cast<const uint2 *>(g.locations);

    // first edge from the global edge table (AliasTable<float, size_t>)
    double2 draw = __ldcs(r);  // .x = random[r] (accept), .y = random[r+1] (index)
    unsigned long long index = alias_index<unsigned long long>(draw.y, g.num_edge);
    unsigned long long edge = float(draw.x) < __ldg(g.edge_prob + index) ? index : __ldg(g.edge_alias + index);
    uint32_t current = __ldg(g.edge_u + edge);
    out[0] = __ldg(locations + current);
    current = __ldg(g.edge_v + edge);
    out[num_walk] = __ldg(locations + current);
    // remaining steps from the per-vertex tables (AliasTable<float, Index>)
    for (int j = 2; j <= walk_length; j++) {
        const unsigned long long begin = __ldg(g.offsets + current);
        const uint32_t degree = uint32_t(__ldg(g.offsets + current + 1) - begin);
        if (degree == 0) {  // dead end: the host refuses such graphs; never read out of bounds
            for (; j <= walk_length; j++)
                out[size_t(j) * num_walk] = __ldg(locations + current);
            break;
        }
        draw = __ldcs(r + j - 1);
        const uint32_t slot = alias_index<uint32_t>(draw.y, degree);
        const uint2 entry = __ldg(reinterpret_cast<const uint2 *>(g.vertex_tables) + begin + slot);
        const uint32_t neighbor = float(draw.x) < __uint_as_float(entry.x) ? slot : entry.y;
        current = __ldg(g.edge_v + begin + neighbor);
        out[size_t(j) * num_walk] = __ldg(locations + current);
    }
    }
}

// ---- pool fill -----------------------------------------------------------------
struct FillParams {
    int num_partition, walk_length, augmentation_step, shuffle_base;
    unsigned long long pool_size, start, slice;  // slice = end - start
    const uint32_t *attributes;                  // knowledge graphs: relation of every sampled edge, else null
};

// one pool entry: a pair {tail_local, head_local}, or with attributes a triplet {relation, tail_local,
// head_local} (the byte order of the reference's std::tuple<Index, Index[, Index]>)
__device__ __forceinline__ void write_entry(const FillParams &p, uint32_t *block, unsigned long long position,
                                            uint32_t walk, uint32_t tail_local, uint32_t head_local) {
    if (p.attributes) {
        uint32_t *entry = block + position * 3;
        entry[0] = p.attributes[walk];
        entry[1] = tail_local;
        entry[2] = head_local;
    } else
        reinterpret_cast<uint2 *>(block)[position] = make_uint2(tail_local, head_local);
}

// Stable partition of the pairs by block, two levels.  A CTA owns `T` consecutive walks (one per
// thread) and keeps a histogram counters[b][thread] in shared memory (conflict-free: thread is the
// fast index).
//   fill_count_kernel   per-CTA totals per block              -> cta_counts[b][cta]
//   fill_scan_kernel    exclusive scan over the CTAs per block -> cta_counts[b][cta] = slice offset of
//                       the CTA's first pair of block b (seeded, saturating at the slice length)
//   fill_scatter_kernel recounts, scans the threads of the CTA per block and emits the pairs in order
__device__ __forceinline__ void count_walk_pairs(const FillParams &p, const uint2 *chains, uint32_t num_walk,
                                                 uint32_t w, uint32_t *counters, int T) {
    for (int j = 0; j < p.walk_length; j++) {
        const uint32_t head_part = chains[size_t(j) * num_walk + w].x;
        for (int k = 1; k <= p.augmentation_step && j + k <= p.walk_length; k++) {
            const uint32_t tail_part = chains[size_t(j + k) * num_walk + w].x;
            counters[(head_part * p.num_partition + tail_part) * T + threadIdx.x]++;
        }
    }
}

__global__ void fill_count_kernel(const FillParams p, const uint2 *chains, uint32_t num_walk, uint32_t *cta_counts) {
    GV_DYNAMIC_SHARED(uint32_t, counters);
    const int T = blockDim.x, num_block = p.num_partition * p.num_partition;
    const uint32_t w = blockIdx.x * T + threadIdx.x;
    for (int b = 0; b < num_block; b++)
        counters[b * T + threadIdx.x] = 0;
    if (w < num_walk)
        count_walk_pairs(p, chains, num_walk, w, counters, T);
    __syncthreads();
    // one warp per block id: sum the T per-thread counts
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, num_warp = T >> 5;
    for (int b = warp; b < num_block; b += num_warp) {
        uint32_t sum = 0;
        for (int t = lane; t < T; t += 32)
            sum += counters[b * T + t];
        for (int delta = 16; delta > 0; delta >>= 1)
            sum += __shfl_xor_sync(0xFFFFFFFFu, sum, delta);
        if (lane == 0)
            cta_counts[size_t(b) * gridDim.x + blockIdx.x] = sum;
    }
}

// one CTA per block id: exclusive scan of the per-CTA totals, seeded with seeds[b]
__global__ void __launch_bounds__(1024) fill_scan_kernel(const FillParams p, uint32_t num_cta, uint32_t *cta_counts,
                                                         const unsigned long long *seeds, unsigned long long *fill,
                                                         unsigned long long *totals) {
    __shared__ unsigned long long partial[1024];
    const int b = blockIdx.x;
    uint32_t *row = cta_counts + size_t(b) * num_cta;
    const uint32_t per_thread = (num_cta + blockDim.x - 1) / blockDim.x;
    const uint32_t begin = min(num_cta, threadIdx.x * per_thread), end = min(num_cta, begin + per_thread);
    unsigned long long sum = 0;
    for (uint32_t c = begin; c < end; c++)
        sum += row[c];
    partial[threadIdx.x] = sum;
    __syncthreads();
    for (int offset = 1; offset < blockDim.x; offset <<= 1) {  // Hillis-Steele inclusive scan
        unsigned long long add = threadIdx.x >= offset ? partial[threadIdx.x - offset] : 0;
        __syncthreads();
        partial[threadIdx.x] += add;
        __syncthreads();
    }
    const unsigned long long seed = seeds ? seeds[b] : 0;
    unsigned long long running = seed + partial[threadIdx.x] - sum;
    for (uint32_t c = begin; c < end; c++) {
        const uint32_t count = row[c];
        row[c] = uint32_t(min(running, p.slice));
        running += count;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) {
        if (fill)
            fill[b] = seed + partial[threadIdx.x];
        if (totals)
            totals[b] = partial[threadIdx.x];
    }
}

// ---- cross-rank stable partition over NVLink peer memory ----------------------------------------
// Every rank walks a contiguous slice of a round's walks.  To keep the pools bit-identical to the
// sequential reference, rank r's pairs of block b must land behind those of ranks < r: each rank
// publishes its per-block totals into every peer's inbox (peer stores + a system-scope fence + a
// flag), waits for all W flags of the round and derives its base offsets.  Control region of a rank:
//   inbox [2][W][num_block + 1] u64 (last slot: walk index that completed a block), flags [2][W] u64.
constexpr unsigned long long kPeerTimeoutNs = 120ull * 1000 * 1000 * 1000;  // 2 minutes

__device__ __forceinline__ unsigned long long *control_inbox(unsigned long long *control, int W, int num_block,
                                                             int parity, int rank) {
    return control + (size_t(parity) * W + rank) * (num_block + 1);
}
__device__ __forceinline__ unsigned long long *control_flag(unsigned long long *control, int W, int num_block,
                                                            int parity, int rank) {
    return control + size_t(2) * W * (num_block + 1) + size_t(parity) * W + rank;
}

__global__ void __launch_bounds__(512) peer_publish_kernel(int rank, int W, int num_block, int parity,
                                                           unsigned long long round_id,
                                                           const unsigned long long *totals,
                                                           const unsigned long long *last_walk,
                                                           unsigned long long *const *controls) {
    for (int i = threadIdx.x; i < (num_block + 1) * W; i += blockDim.x) {
        const int peer = i / (num_block + 1), slot = i % (num_block + 1);
        const unsigned long long value = slot < num_block ? (totals ? totals[slot] : 0ull) : *last_walk;
        control_inbox(controls[peer], W, num_block, parity, rank)[slot] = value;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < W) {
        volatile unsigned long long *flag = control_flag(controls[threadIdx.x], W, num_block, parity, rank);
        *flag = round_id;
    }
}

__global__ void __launch_bounds__(512) peer_gather_kernel(int rank, int W, int num_block, int parity,
                                                          unsigned long long round_id, unsigned long long *control,
                                                          unsigned long long *fill, unsigned long long *bases,
                                                          unsigned long long *last_walk) {
    // wait for every rank's flag, but never hang the GPU: after kPeerTimeoutNs the round is abandoned and
    // *last_walk is set to the all-ones marker, which the host turns into an error
    __shared__ int timed_out;
    if (threadIdx.x == 0)
        timed_out = 0;
    __syncthreads();
    if (threadIdx.x < W) {
        volatile unsigned long long *flag = control_flag(control, W, num_block, parity, threadIdx.x);
        unsigned long long begin, now;
        begin = gv_global_timer_ns();
        while (*flag != round_id) {
            __nanosleep(200);
            now = gv_global_timer_ns();
            if (now - begin > kPeerTimeoutNs) {
                timed_out = 1;
                break;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (timed_out) {
        if (threadIdx.x == 0)
            *last_walk = ~0ull;
        return;
    }
    for (int b = threadIdx.x; b <= num_block; b += blockDim.x) {
        unsigned long long before = 0, all = 0, latest = 0;
        for (int r = 0; r < W; r++) {
            const unsigned long long value =
                *(volatile unsigned long long *)(control_inbox(control, W, num_block, parity, r) + b);
            if (r < rank)
                before += value;
            all += value;
            latest = max(latest, value);
        }
        if (b < num_block) {
            bases[b] = fill[b] + before;
            fill[b] += all;
        } else
            *last_walk = max(*last_walk, latest);
    }
}

// every walk re-emits its pairs in order and writes the ones that still fit
// Staging (multi-GPU): 8-byte stores of 32 lanes to 32 unrelated places of a PEER's pool are 32 NVLink write
// requests.  Pairs of blocks flagged in stage.remote are therefore written to a local staging array, in append
// order behind the rank's first pair of the block (stage.offsets[b] + in_slice - bases[b]), and fill_forward_kernel
// moves them to their shuffled positions with consecutive lanes on consecutive addresses.
struct StageParams {
    uint2 *staging;                       // nullptr: write everything in place
    const unsigned char *remote;          // [num_block] 1 = stage and forward
    const unsigned long long *offsets;    // [num_block] first staging entry of the block
    const unsigned long long *bases;      // [num_block] slice offset of this rank's first pair
};

__global__ void fill_scatter_kernel(const FillParams p, const uint2 *chains, uint32_t num_walk,
                                    unsigned long long first_walk, const uint32_t *cta_bases,
                                    uint32_t *const *pool_blocks, unsigned long long *last_walk,
                                    const StageParams stage) {
    GV_DYNAMIC_SHARED(uint32_t, counters);
    const int T = blockDim.x, num_block = p.num_partition * p.num_partition;
    const uint32_t w = blockIdx.x * T + threadIdx.x;
    for (int b = 0; b < num_block; b++)
        counters[b * T + threadIdx.x] = 0;
    if (w < num_walk)
        count_walk_pairs(p, chains, num_walk, w, counters, T);
    __syncthreads();
    // exclusive scan over the CTA's threads, one warp per block id, 4 consecutive threads per lane step
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, num_warp = T >> 5;
    const uint32_t slice = uint32_t(p.slice);
    for (int b = warp; b < num_block; b += num_warp) {
        uint32_t running = cta_bases[size_t(b) * gridDim.x + blockIdx.x];
        for (int t0 = 0; t0 < T; t0 += 32) {
            const uint32_t count = counters[b * T + t0 + lane];
            uint32_t inclusive = count;
            for (int delta = 1; delta < 32; delta <<= 1) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inclusive, delta);
                if (lane >= delta)
                    inclusive += up;
            }
            const unsigned long long start = (unsigned long long)running + inclusive - count;
            counters[b * T + t0 + lane] = uint32_t(min(start, (unsigned long long)slice));
            const unsigned long long next = (unsigned long long)running + __shfl_sync(0xFFFFFFFFu, inclusive, 31);
            running = uint32_t(min(next, (unsigned long long)slice));
        }
    }
    __syncthreads();
    if (w >= num_walk)
        return;
    const unsigned long long shuffle_stride = p.pool_size / p.shuffle_base;
    bool completed = false;
    for (int j = 0; j < p.walk_length; j++) {
        const uint2 head = chains[size_t(j) * num_walk + w];
        for (int k = 1; k <= p.augmentation_step && j + k <= p.walk_length; k++) {
            const uint2 tail = chains[size_t(j + k) * num_walk + w];
            const int b = head.x * p.num_partition + tail.x;
            const uint32_t in_slice = counters[b * T + threadIdx.x];
            if (in_slice < slice) {
                counters[b * T + threadIdx.x] = in_slice + 1;
                const unsigned long long offset = p.start + in_slice;
                // pseudo shuffle, instance/graph.cuh:440-441
                const unsigned long long shuffled = offset % p.shuffle_base * shuffle_stride + offset / p.shuffle_base;
                uint32_t *block = pool_blocks[b];
                if (stage.staging && stage.remote[b])
                    stage.staging[stage.offsets[b] + (in_slice - stage.bases[b])] = make_uint2(tail.y, head.y);
                else if (block)
                    write_entry(p, block, shuffled, w, tail.y, head.y);
                completed |= in_slice + 1 == slice;
            }
        }
    }
    if (completed)
        atomicMax(last_walk, first_walk + w);
}

// fill_scatter_kernel with coalesced traffic.  The pairs a CTA emits for block b are a contiguous range of b's slice
// offsets, starting at cta_bases[b][cta]: every thread parks its pairs in shared memory at (start of b inside the
// CTA) + (offset - CTA base) and the CTA then writes block after block -- a staged block as one contiguous run of the
// staging array, a pool block class by class of the pseudo shuffle (offsets = c mod shuffle_base are neighbours in
// the pool), consecutive threads on consecutive addresses.  Shared memory: counters [num_block][T], block starts
// [num_block + 1], kept counts [num_block], pairs [T * pairs_per_walk].
__global__ void fill_scatter_tiled_kernel(const FillParams p, const uint2 *chains, uint32_t num_walk,
                                          unsigned long long first_walk, const uint32_t *cta_bases,
                                          uint32_t *const *pool_blocks, unsigned long long *last_walk,
                                          const StageParams stage, uint32_t pairs_per_walk) {
    GV_DYNAMIC_SHARED(uint32_t, counters);
    const int T = blockDim.x, num_block = p.num_partition * p.num_partition;
    uint32_t *block_start = counters + size_t(num_block) * T;  // [num_block + 1] first pair of block b in `pairs`
    uint32_t *block_kept = block_start + num_block + 1;        // [num_block] pairs of b below the slice end
    uint2 *pairs = reinterpret_cast<uint2 *>(block_kept + num_block + ((num_block * 2 + 1) & 1));  // 8-byte aligned
    const uint32_t w = blockIdx.x * T + threadIdx.x;
    for (int b = 0; b < num_block; b++)
        counters[b * T + threadIdx.x] = 0;
    if (w < num_walk)
        count_walk_pairs(p, chains, num_walk, w, counters, T);
    __syncthreads();
    // per block: exclusive scan over the threads (slice offsets, saturating) and the CTA's total
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, num_warp = T >> 5;
    const uint32_t slice = uint32_t(p.slice);
    for (int b = warp; b < num_block; b += num_warp) {
        const uint32_t cta_base = cta_bases[size_t(b) * gridDim.x + blockIdx.x];
        uint32_t running = cta_base, total = 0;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const uint32_t count = counters[b * T + t0 + lane];
            uint32_t inclusive = count;
            for (int delta = 1; delta < 32; delta <<= 1) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inclusive, delta);
                if (lane >= delta)
                    inclusive += up;
            }
            const unsigned long long start = (unsigned long long)running + inclusive - count;
            counters[b * T + t0 + lane] = uint32_t(min(start, (unsigned long long)slice));
            const uint32_t chunk_total = __shfl_sync(0xFFFFFFFFu, inclusive, 31);
            const unsigned long long next = (unsigned long long)running + chunk_total;
            running = uint32_t(min(next, (unsigned long long)slice));
            total += chunk_total;
        }
        if (lane == 0) {
            block_start[b + 1] = total;  // turned into a prefix sum below
            block_kept[b] = min(total, slice - min(cta_base, slice));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        block_start[0] = 0;
        for (int b = 0; b < num_block; b++)
            block_start[b + 1] += block_start[b];
    }
    __syncthreads();
    if (w < num_walk) {
        bool completed = false;
        for (int j = 0; j < p.walk_length; j++) {
            const uint2 head = chains[size_t(j) * num_walk + w];
            for (int k = 1; k <= p.augmentation_step && j + k <= p.walk_length; k++) {
                const uint2 tail = chains[size_t(j + k) * num_walk + w];
                const int b = head.x * p.num_partition + tail.x;
                const uint32_t in_slice = counters[b * T + threadIdx.x];
                if (in_slice < slice) {
                    counters[b * T + threadIdx.x] = in_slice + 1;
                    const uint32_t cta_base = cta_bases[size_t(b) * gridDim.x + blockIdx.x];
                    pairs[block_start[b] + (in_slice - cta_base)] = make_uint2(tail.y, head.y);
                    completed |= in_slice + 1 == slice;
                }
            }
        }
        if (completed)
            atomicMax(last_walk, first_walk + w);
    }
    __syncthreads();
    const unsigned long long base = p.shuffle_base, stride = p.pool_size / base;
    for (int b = 0; b < num_block; b++) {
        const uint32_t kept = block_kept[b];
        if (kept == 0)
            continue;
        const uint32_t cta_base = cta_bases[size_t(b) * gridDim.x + blockIdx.x];
        const uint2 *source = pairs + block_start[b];
        if (stage.staging && stage.remote[b]) {
            uint2 *run = stage.staging + stage.offsets[b] + (cta_base - stage.bases[b]);
            for (uint32_t i = threadIdx.x; i < kept; i += T)
                run[i] = source[i];
            continue;
        }
        uint2 *block = reinterpret_cast<uint2 *>(pool_blocks[b]);
        if (!block)
            continue;
        const unsigned long long g0 = p.start + cta_base;
        for (unsigned long long c = 0; c < base && c < kept; c++) {
            const uint32_t members = uint32_t((kept - c + base - 1) / base);
            const unsigned long long cell = (g0 + c) % base * stride + (g0 + c) / base;
            for (uint32_t q = threadIdx.x; q < members; q += T)
                block[cell + q] = source[uint32_t(c) + q * uint32_t(base)];
        }
    }
}

// staging offsets: exclusive prefix sum of the rank's totals over the staged blocks (one small CTA)
__global__ void fill_stage_offsets_kernel(int num_block, const unsigned char *remote, const unsigned long long *totals,
                                          unsigned long long *offsets) {
    if (threadIdx.x == 0) {
        unsigned long long running = 0;
        for (int b = 0; b < num_block; b++) {
            offsets[b] = running;
            if (remote[b])
                running += totals[b];
        }
    }
}

// Move the staged pairs of block blockIdx.y to their pool: entry i of the rank's run has slice offset
// g = start + base + i and lives at (g % shuffle_base) * stride + g / shuffle_base (instance/graph.cuh:440-441), so
// the entries i = c, c + shuffle_base, c + 2 shuffle_base ... are neighbours in the pool: consecutive threads take
// consecutive members of one such class -> a warp writes 256 contiguous bytes (to a peer: full NVLink packets).
__global__ void __launch_bounds__(256) fill_forward_kernel(const FillParams p, const StageParams stage,
                                                           const unsigned long long *totals,
                                                           uint32_t *const *pool_blocks) {
    const int b = blockIdx.y;
    uint2 *block = reinterpret_cast<uint2 *>(pool_blocks[b]);
    if (!stage.remote[b] || !block)
        return;
    const unsigned long long slice = p.slice;
    const unsigned long long begin = min(stage.bases[b], slice), end = min(stage.bases[b] + totals[b], slice);
    const unsigned long long n = end - begin;
    const uint2 *source = stage.staging + stage.offsets[b];
    const unsigned long long g0 = p.start + begin, base = p.shuffle_base, stride = p.pool_size / base;
    const unsigned long long threads = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long first = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned long long c = 0; c < base && c < n; c++) {
        const unsigned long long members = (n - c + base - 1) / base;
        // g % base is constant inside a class; 4 independent loads in flight per thread
        const unsigned long long row = (g0 + c) % base * stride;
#pragma unroll 4
        for (unsigned long long q = first; q < members; q += threads) {
            const unsigned long long i = c + q * base;
            block[row + (g0 + i) / base] = source[i];
        }
    }
}

// single-block fast path (num_partition == 1): the slice offset of a pair is its stream index
__global__ void __launch_bounds__(256) fill_direct_kernel(const FillParams p, const uint2 *chains, uint32_t num_walk,
                                                          unsigned long long first_walk, uint32_t pairs_per_walk,
                                                          const unsigned long long *fill,
                                                          uint32_t *const *pool_blocks,
                                                          unsigned long long *last_walk) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= num_walk)
        return;
    uint32_t *block = pool_blocks[0];
    const unsigned long long shuffle_stride = p.pool_size / p.shuffle_base;
    unsigned long long in_slice = fill[0] + (unsigned long long)w * pairs_per_walk;
    bool completed = false;
    for (int j = 0; j < p.walk_length && in_slice < p.slice; j++) {
        const uint2 head = chains[size_t(j) * num_walk + w];
        for (int k = 1; k <= p.augmentation_step && j + k <= p.walk_length; k++, in_slice++) {
            if (in_slice >= p.slice)
                break;
            const uint2 tail = chains[size_t(j + k) * num_walk + w];
            const unsigned long long offset = p.start + in_slice;
            const unsigned long long shuffled = offset % p.shuffle_base * shuffle_stride + offset / p.shuffle_base;
            if (block)
                write_entry(p, block, shuffled, w, tail.y, head.y);
            completed |= in_slice + 1 == p.slice;
        }
    }
    if (completed)
        atomicMax(last_walk, first_walk + w);
}

// The same fill with coalesced traffic (walk_length > 1, pairs only).  fill_direct_kernel gives a walk to a thread:
// the 32 stores of a warp instruction land 8 bytes each in 32 places of the pool (pairs_per_walk entries apart).
// Here a CTA owns kTileWalks consecutive walks, whose pairs are ONE contiguous range of slice offsets:
//   1. the tile's chains go to shared memory with coalesced loads (rows of kTileWalks locations);
//   2. the range is written class by class: offset g lives at (g % shuffle_base) * stride + g / shuffle_base
//      (instance/graph.cuh:440-441), so the offsets of one residue class are neighbours in the pool and consecutive
//      threads take consecutive members -- every warp store is 256 contiguous bytes;
//   3. the pair behind offset g is recovered from g alone: walk = g / pairs_per_walk, position -> (j, k).
constexpr int kTileWalks = 32;

__global__ void __launch_bounds__(256) fill_direct_tiled_kernel(const FillParams p, const uint2 *chains,
                                                                uint32_t num_walk, unsigned long long first_walk,
                                                                uint32_t pairs_per_walk,
                                                                const unsigned long long *fill,
                                                                uint32_t *const *pool_blocks,
                                                                unsigned long long *last_walk) {
    GV_DYNAMIC_SHARED(uint2, tile);  // [walk_length + 1][kTileWalks]
    const uint32_t w0 = blockIdx.x * kTileWalks;
    const uint32_t walks = min(uint32_t(kTileWalks), num_walk - w0);
    const unsigned long long first = fill[0] + (unsigned long long)w0 * pairs_per_walk;  // slice offset of the tile
    if (first >= p.slice)
        return;
    for (uint32_t e = threadIdx.x; e < uint32_t(p.walk_length + 1) * kTileWalks; e += blockDim.x) {
        const uint32_t j = e / kTileWalks, w = e % kTileWalks;
        if (w < walks)
            tile[e] = chains[size_t(j) * num_walk + w0 + w];
    }
    __syncthreads();
    uint2 *block = reinterpret_cast<uint2 *>(pool_blocks[0]);
    const unsigned long long count = min((unsigned long long)walks * pairs_per_walk, p.slice - first);
    const unsigned long long g0 = p.start + first, base = p.shuffle_base, stride = p.pool_size / base;
    // pairs (j, k) of a walk in emission order: j-major, k = 1 .. min(augmentation_step, walk_length - j); the first
    // `full` values of j have all augmentation_step pairs
    const uint32_t aug = uint32_t(p.augmentation_step), L = uint32_t(p.walk_length);
    const uint32_t full = L >= aug ? L - aug + 1 : 0;
    for (unsigned long long c = 0; c < base && c < count; c++) {
        const unsigned long long members = (count - c + base - 1) / base;
        // (g0 + c + q * base) / base = (g0 + c) / base + q: one 64-bit division per class, none per pair
        const unsigned long long cell = (g0 + c) % base * stride + (g0 + c) / base;
        for (uint32_t q = threadIdx.x; q < members; q += blockDim.x) {
            const uint32_t i = uint32_t(c) + q * uint32_t(base);  // position inside the tile's range (< 2^32)
            const uint32_t w = i / pairs_per_walk;
            uint32_t position = i % pairs_per_walk, j, k;
            if (position < full * aug) {
                j = position / aug;
                k = position % aug + 1;
            } else {  // the last walk positions emit walk_length - j < augmentation_step pairs each
                position -= full * aug;
                j = full;
                while (position >= L - j) {
                    position -= L - j;
                    j++;
                }
                k = position + 1;
            }
            const uint2 head = tile[j * kTileWalks + w], tail = tile[(j + k) * kTileWalks + w];
            if (block)
                block[cell + q] = make_uint2(tail.y, head.y);
            if (first + i + 1 == p.slice)
                atomicMax(last_walk, first_walk + w0 + w);
        }
    }
}

__global__ void fill_advance_kernel(unsigned long long *fill, unsigned long long amount) {
    fill[0] += amount;
}

static uint32_t pairs_per_walk(int walk_length, int augmentation_step) {
    uint32_t n = 0;
    for (int j = 0; j < walk_length; j++)
        for (int k = 1; k <= augmentation_step && j + k <= walk_length; k++)
            n++;
    return n;
}

}  // namespace device
}  // namespace gv

using namespace gv;
using namespace gv::device;

extern "C" {

int gv_cuda_random_walk(const gv_device_graph_t *graph, const double *random, uint32_t num_walk, int walk_length,
                        uint64_t first_walk, uint32_t walks_per_buffer, uint64_t buffer_doubles,
                        gv_location_t *chains, void *stream) {
    if (num_walk == 0)
        return 0;
    if (!graph || !random || !chains || walk_length < 1 || walks_per_buffer == 0 ||
        buffer_doubles < uint64_t(walks_per_buffer) * 2 * walk_length || buffer_doubles % 2 != 0)
        return fail("gv_cuda_random_walk: invalid argument");
    if (walk_length > 1 && !graph->vertex_tables)
        return fail("gv_cuda_random_walk: per-vertex alias tables are required for walk_length > 1");
    const int threads = 256;
    uint32_t blocks = (num_walk + threads - 1) / threads;
    if (gv::sampler_max_ctas() > 0)
        blocks = std::min<uint32_t>(blocks, uint32_t(gv::sampler_max_ctas()));
    GV_LAUNCH(blocks, threads, 0, static_cast<cudaStream_t>(stream), random_walk_kernel)(
        *graph, random, num_walk, walk_length, first_walk, walks_per_buffer, buffer_doubles, chains);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

// CTA size of the histogram kernels: as many walks per CTA as 48 KB of counters allow
static int fill_threads_untiled(int num_partition) {
    const int num_block = num_partition * num_partition;
    int threads = int((48 * 1024) / (size_t(num_block) * sizeof(uint32_t))) / 32 * 32;
    return threads > 128 ? 128 : (threads < 32 ? 32 : threads);
}

// shared memory of fill_scatter_tiled_kernel for T walks per CTA
static size_t tiled_shared_bytes(int num_partition, uint32_t per_walk, int T) {
    const size_t num_block = size_t(num_partition) * num_partition;
    const size_t words = num_block * T + (num_block + 1) + num_block + ((num_block * 2 + 1) & 1);
    return words * sizeof(uint32_t) + size_t(T) * per_walk * sizeof(uint2);
}

// The tiled scatter parks all pairs of a CTA in shared memory, which bounds the walks per CTA; count, scan and scatter
// must agree on that number.  Returns the CTA size and whether the tiled kernel is used (pairs only, and at least
// one warp of walks must fit into 200 KB).
static int fill_threads(const FillParams &p, bool &tiled) {
    int threads = fill_threads_untiled(p.num_partition);
    tiled = false;
    if (p.attributes || gv::direct_fill_per_walk())
        return threads;
    const uint32_t per_walk = pairs_per_walk(p.walk_length, p.augmentation_step);
    // <= 110 KB keeps two CTAs on an SM (228 KB); a single warp of walks may take up to 200 KB
    while (threads > 32 && tiled_shared_bytes(p.num_partition, per_walk, threads) > 110 * 1024)
        threads -= 32;
    if (tiled_shared_bytes(p.num_partition, per_walk, threads) > 200 * 1024)
        return fill_threads_untiled(p.num_partition);
    tiled = true;
    return threads;
}

size_t gv_cuda_fill_scratch_bytes(uint32_t num_walk, int num_partition) {
    if (num_partition <= 1)
        return 16;
    const int threads = 32;  // the smallest CTA any configuration uses: an upper bound of the CTA count
    const size_t num_cta = (size_t(num_walk) + threads - 1) / threads;
    return num_cta * num_partition * num_partition * sizeof(uint32_t) + 256;
}

// launches fill_scatter_kernel or its tiled variant (same arguments)
static int launch_scatter(const FillParams &p, const uint2 *chains, uint32_t num_walk, unsigned long long first_walk,
                          const uint32_t *cta_counts, uint32_t *const *pool_blocks, unsigned long long *last_walk,
                          const StageParams &stage, cudaStream_t s) {
    bool tiled;
    const int T = fill_threads(p, tiled);
    const int num_block = p.num_partition * p.num_partition;
    const uint32_t num_cta = (num_walk + T - 1) / T;
    if (!tiled) {
        GV_LAUNCH(num_cta, T, size_t(num_block) * T * sizeof(uint32_t), s, fill_scatter_kernel)(
            p, chains, num_walk, first_walk, cta_counts, pool_blocks, last_walk, stage);
        GV_CUDA_OK(cudaGetLastError());
        return 0;
    }
    const uint32_t per_walk = pairs_per_walk(p.walk_length, p.augmentation_step);
    const size_t shared = tiled_shared_bytes(p.num_partition, per_walk, T);
    if (shared > 48 * 1024)
        GV_CUDA_OK(cudaFuncSetAttribute(fill_scatter_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        int(shared)));
    GV_LAUNCH(num_cta, T, shared, s, fill_scatter_tiled_kernel)(p, chains, num_walk, first_walk, cta_counts,
                                                                pool_blocks, last_walk, stage, per_walk);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gv_cuda_fill_pool(const gv_fill_params_t *params, const gv_location_t *chains, uint32_t num_walk,
                      uint64_t first_walk, uint32_t *const *pool_blocks, unsigned long long *fill,
                      unsigned long long *last_walk, void *scratch, void *stream) {
    if (num_walk == 0)
        return 0;
    if (!params || !chains || !pool_blocks || !fill || !last_walk)
        return fail("gv_cuda_fill_pool: null argument");
    if (params->num_partition < 1 || params->num_partition > 16)
        return fail("gv_cuda_fill_pool: num_partition must be in [1, 16]");
    if (params->shuffle_base < 1 || params->pool_size % params->shuffle_base != 0)
        return fail("Can't perform pseudo shuffle: episode size must be a multiple of the shuffle base");
    if (params->end < params->start || params->end > params->pool_size)
        return fail("gv_cuda_fill_pool: invalid slice");
    FillParams p;
    p.num_partition = params->num_partition;
    p.walk_length = params->walk_length;
    p.augmentation_step = params->augmentation_step;
    p.shuffle_base = params->shuffle_base;
    p.pool_size = params->pool_size;
    p.start = params->start;
    p.slice = params->end - params->start;
    p.attributes = params->attributes;
    if (p.attributes && (p.walk_length != 1 || p.augmentation_step != 1))
        return fail("gv_cuda_fill_pool: attributes need walk_length == 1 (edge sampling)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint2 *c = reinterpret_cast<const uint2 *>(chains);
    const int threads = 256;
    const uint32_t blocks = (num_walk + threads - 1) / threads;
    if (p.num_partition == 1) {
        // every walk is full length, so pair (w, j, k) sits at stream index w * pairs_per_walk + f(j, k);
        // fill[0] is read on the device and advanced by a 1-thread kernel behind the fill (same stream)
        const uint32_t per_walk = pairs_per_walk(p.walk_length, p.augmentation_step);
        const size_t tile_bytes = size_t(p.walk_length + 1) * kTileWalks * sizeof(uint2);
        if (p.walk_length > 1 && !p.attributes && tile_bytes <= 48 * 1024 && !gv::direct_fill_per_walk())
            GV_LAUNCH((num_walk + kTileWalks - 1) / kTileWalks, 256, tile_bytes, s, fill_direct_tiled_kernel)(
                p, c, num_walk, first_walk, per_walk, fill, pool_blocks, last_walk);
        else
            GV_LAUNCH(blocks, threads, 0, s, fill_direct_kernel)(p, c, num_walk, first_walk, per_walk, fill, pool_blocks,
                                                          last_walk);
        GV_CUDA_OK(cudaGetLastError());
        GV_LAUNCH(1, 1, 0, s, fill_advance_kernel)(fill, (unsigned long long)num_walk * per_walk);
        GV_CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (!scratch)
        return fail("gv_cuda_fill_pool: scratch required for num_partition > 1");
    uint32_t *cta_counts = static_cast<uint32_t *>(scratch);
    const int num_block = p.num_partition * p.num_partition;
    bool tiled;
    const int T = fill_threads(p, tiled);
    const uint32_t num_cta = (num_walk + T - 1) / T;
    const size_t shared = size_t(num_block) * T * sizeof(uint32_t);
    GV_LAUNCH(num_cta, T, shared, s, fill_count_kernel)(p, c, num_walk, cta_counts);
    GV_CUDA_OK(cudaGetLastError());
    GV_LAUNCH(num_block, 1024, 0, s, fill_scan_kernel)(p, num_cta, cta_counts, fill, fill, nullptr);
    GV_CUDA_OK(cudaGetLastError());
    return launch_scatter(p, c, num_walk, first_walk, cta_counts, pool_blocks, last_walk,
                          StageParams{nullptr, nullptr, nullptr, nullptr}, s);
}

}  // extern "C"

extern "C" {

size_t gv_cuda_peer_control_bytes(int world_size, int num_partition) {
    const size_t num_block = size_t(num_partition) * num_partition;
    return (size_t(2) * world_size * (num_block + 1) + size_t(2) * world_size) * sizeof(unsigned long long);
}

static int make_fill_params(const gv_fill_params_t *params, FillParams &p) {
    if (!params || params->num_partition < 1 || params->num_partition > 16)
        return fail("fill: num_partition must be in [1, 16]");
    if (params->shuffle_base < 1 || params->pool_size % params->shuffle_base != 0)
        return fail("Can't perform pseudo shuffle: episode size must be a multiple of the shuffle base");
    if (params->end < params->start || params->end > params->pool_size)
        return fail("fill: invalid slice");
    p.num_partition = params->num_partition;
    p.walk_length = params->walk_length;
    p.augmentation_step = params->augmentation_step;
    p.shuffle_base = params->shuffle_base;
    p.pool_size = params->pool_size;
    p.start = params->start;
    p.slice = params->end - params->start;
    p.attributes = params->attributes;
    if (p.attributes && (p.walk_length != 1 || p.augmentation_step != 1))
        return fail("fill: attributes need walk_length == 1 (edge sampling)");
    return 0;
}

int gv_cuda_fill_count(const gv_fill_params_t *params, const gv_location_t *chains, uint32_t num_walk, void *scratch,
                       unsigned long long *totals, void *stream) {
    FillParams p;
    if (make_fill_params(params, p))
        return -1;
    if (!chains || !scratch || !totals)
        return fail("gv_cuda_fill_count: null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int num_block = p.num_partition * p.num_partition;
    if (num_walk == 0) {
        GV_CUDA_OK(cudaMemsetAsync(totals, 0, num_block * sizeof(unsigned long long), s));
        return 0;
    }
    uint32_t *cta_counts = static_cast<uint32_t *>(scratch);
    bool tiled;
    const int T = fill_threads(p, tiled);
    const uint32_t num_cta = (num_walk + T - 1) / T;
    const size_t shared = size_t(num_block) * T * sizeof(uint32_t);
    GV_LAUNCH(num_cta, T, shared, s, fill_count_kernel)(p, reinterpret_cast<const uint2 *>(chains), num_walk, cta_counts);
    GV_CUDA_OK(cudaGetLastError());
    // unseeded scan: cta_counts become offsets relative to the start of this rank's pairs; totals out
    FillParams unbounded = p;
    unbounded.slice = ~0ull;
    GV_LAUNCH(num_block, 1024, 0, s, fill_scan_kernel)(unbounded, num_cta, cta_counts, nullptr, nullptr, totals);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

// adds the rank's base offsets to the relative CTA offsets left by gv_cuda_fill_count (saturating)
__global__ void fill_rebase_kernel(unsigned long long slice, uint32_t num_cta, uint32_t *cta_counts,
                                   const unsigned long long *bases) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < num_cta) {
        const unsigned long long value = bases[blockIdx.y] + cta_counts[size_t(blockIdx.y) * num_cta + c];
        cta_counts[size_t(blockIdx.y) * num_cta + c] = uint32_t(value < slice ? value : slice);
    }
}

int gv_cuda_fill_scatter(const gv_fill_params_t *params, const gv_location_t *chains, uint32_t num_walk,
                         uint64_t first_walk, uint32_t *const *pool_blocks, const unsigned long long *bases,
                         unsigned long long *last_walk, void *scratch, void *stream) {
    return gv_cuda_fill_scatter_staged(params, chains, num_walk, first_walk, pool_blocks, bases, last_walk, scratch,
                                       nullptr, nullptr, nullptr, nullptr, stream);
}

// bases = fill, fill += totals: where this round's pairs start in every block's slice (one sampler, one rank)
__global__ void fill_advance_bases_kernel(int num_block, const unsigned long long *totals, unsigned long long *fill,
                                          unsigned long long *bases) {
    for (int b = threadIdx.x; b < num_block; b += blockDim.x) {
        const unsigned long long before = fill[b];
        bases[b] = before;
        fill[b] = before + totals[b];
    }
}

int gv_cuda_fill_advance(int num_block, const unsigned long long *totals, unsigned long long *fill,
                         unsigned long long *bases, void *stream) {
    if (num_block < 1 || !totals || !fill || !bases)
        return fail("gv_cuda_fill_advance: invalid argument");
    GV_LAUNCH(1, 256, 0, static_cast<cudaStream_t>(stream), fill_advance_bases_kernel)(num_block, totals, fill, bases);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

size_t gv_cuda_fill_staging_bytes(uint32_t num_walk, int walk_length, int augmentation_step) {
    return size_t(num_walk) * pairs_per_walk(walk_length, augmentation_step) * sizeof(uint2) + 16;
}

int gv_cuda_fill_scatter_staged(const gv_fill_params_t *params, const gv_location_t *chains, uint32_t num_walk,
                                uint64_t first_walk, uint32_t *const *pool_blocks, const unsigned long long *bases,
                                unsigned long long *last_walk, void *scratch, const unsigned char *remote_blocks,
                                const unsigned long long *totals, void *staging, unsigned long long *stage_offsets,
                                void *stream) {
    FillParams p;
    if (make_fill_params(params, p))
        return -1;
    if (num_walk == 0)
        return 0;
    if (!chains || !scratch || !bases || !pool_blocks || !last_walk)
        return fail("gv_cuda_fill_scatter: null argument");
    const bool staged = remote_blocks != nullptr;
    if (staged && (!totals || !staging || !stage_offsets))
        return fail("gv_cuda_fill_scatter_staged: totals, staging and stage_offsets are required with remote_blocks");
    if (staged && p.attributes)
        return fail("gv_cuda_fill_scatter_staged: staging moves 8-byte pairs only (no attributes)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint32_t *cta_counts = static_cast<uint32_t *>(scratch);
    const int num_block = p.num_partition * p.num_partition;
    bool tiled;
    const int T = fill_threads(p, tiled);
    const uint32_t num_cta = (num_walk + T - 1) / T;
    GV_LAUNCH(dim3((num_cta + 255) / 256, num_block), 256, 0, s, fill_rebase_kernel)(p.slice, num_cta, cta_counts, bases);
    GV_CUDA_OK(cudaGetLastError());
    StageParams stage{nullptr, nullptr, nullptr, nullptr};
    if (staged) {
        stage = StageParams{static_cast<uint2 *>(staging), remote_blocks, stage_offsets, bases};
        GV_LAUNCH(1, 32, 0, s, fill_stage_offsets_kernel)(num_block, remote_blocks, totals, stage_offsets);
        GV_CUDA_OK(cudaGetLastError());
    }
    if (launch_scatter(p, reinterpret_cast<const uint2 *>(chains), num_walk, first_walk, cta_counts, pool_blocks,
                       last_walk, stage, s))
        return -1;
    if (staged) {
        // enough CTAs per block to keep the copy engines of the NVLink path busy, few enough to stay cheap
        const uint64_t per_block = uint64_t(num_walk) * pairs_per_walk(p.walk_length, p.augmentation_step) / num_block;
        const unsigned ctas = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(128, per_block / 2048)));
        GV_LAUNCH(dim3(ctas, num_block), 256, 0, s, fill_forward_kernel)(p, stage, totals, pool_blocks);
        GV_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int gv_cuda_peer_exchange(int rank, int world_size, int num_partition, uint64_t round_id,
                          const unsigned long long *totals, unsigned long long *const *controls,
                          unsigned long long *control, unsigned long long *fill, unsigned long long *bases,
                          unsigned long long *last_walk, void *stream) {
    if (!controls || !control || !fill || !bases || !last_walk || world_size < 1 || world_size > 256)
        return fail("gv_cuda_peer_exchange: invalid argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int num_block = num_partition * num_partition, parity = int(round_id & 1);
    GV_LAUNCH(1, 512, 0, s, peer_publish_kernel)(rank, world_size, num_block, parity, round_id, totals, last_walk, controls);
    GV_CUDA_OK(cudaGetLastError());
    GV_LAUNCH(1, 512, 0, s, peer_gather_kernel)(rank, world_size, num_block, parity, round_id, control, fill, bases,
                                         last_walk);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
