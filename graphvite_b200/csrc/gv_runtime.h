// Internal C++ declarations shared by the solvers of libgv_b200 (gv_solver.cpp, gv_kg_solver.cpp): the
// process-wide engine, SolverMixin's constants, RAII device memory, the optimizer descriptor with its
// learning-rate schedule, and SolverMixin::partition.
#pragma once

#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "gv_engine.h"
#include "gv_host.h"

namespace gv {

// core/solver.h:50-57 and instance/graph.cuh:56
extern Mt19937 g_engine;  // defined in gv_solver.cpp; shared by every solver of the process (std::mt19937's sequence)
static const int kMaxPartition = 16;
static const int kRandBatchSize = 5000000;
static const int kMinBatchSize = 10000;
static const int kSamplePerVertex = 175;
static const int kMinEpisodeSample = 20000000;
static const int kExpectedDegree = 1600;
static const int kSpanBuffers = 16;  // refill buffers generated and walked per sampler round

#define GV_CHECK_CUDA(call)                                                                              \
    do {                                                                                                 \
        cudaError_t gv_e__ = (call);                                                                     \
        if (gv_e__ != cudaSuccess)                                                                       \
            throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(gv_e__) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__));                         \
    } while (0)
#define GV_CHECK_ABI(call)                              \
    do {                                                \
        if ((call) != 0)                                \
            throw std::runtime_error(gv_last_error()); \
    } while (0)

inline void require(bool condition, const std::string &message) {
    if (!condition)
        throw std::runtime_error(message);
}

inline double now_seconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// GV_LOG=2: phase timings of train_begin() on stderr
struct PhaseTimer {
    double last = now_seconds();
    bool on = getenv("GV_LOG") != nullptr && atoi(getenv("GV_LOG")) >= 2;
    void mark(const char *what) {
        if (on) {
            const double t = now_seconds();
            fprintf(stderr, "[gv] %-28s %8.3f s\n", what, t - last);
            last = t;
        }
    }
};

inline bool log_enabled() {
    static const bool on = getenv("GV_LOG") != nullptr && atoi(getenv("GV_LOG")) > 0;
    return on;
}

// The [|V|][dim] host matrices behind the numpy views (and their moment copies): zero-filled float arrays like
// std::vector<float>::assign(n, 0.f), but on 2-MB-aligned memory with MADV_HUGEPAGE and zeroed by a few threads --
// first-touching 1.2 GB in 4-KB pages from one thread was the largest part of build() at Youtube size, and fewer,
// larger pages also make cudaHostRegister cheaper.
class HostMatrix {
public:
    HostMatrix() {}
    HostMatrix(const HostMatrix &) = delete;
    HostMatrix &operator=(const HostMatrix &) = delete;
    ~HostMatrix() { free(base); }
    float *data() { return base; }
    const float *data() const { return base; }
    size_t size() const { return count; }
    bool empty() const { return count == 0; }
    float *begin() { return base; }
    float *end() { return base + count; }
    void clear() {
        free(base);
        base = nullptr;
        count = 0;
    }
    // n zeros (the previous contents are dropped)
    void assign(size_t n, float value) {
        if (value != 0.f)
            throw std::logic_error("HostMatrix::assign: only zero fill");
        if (n != count) {
            clear();
            if (n) {
                const size_t huge = size_t(2) << 20, bytes = (n * sizeof(float) + huge - 1) / huge * huge;
                void *memory = nullptr;
                if (posix_memalign(&memory, n * sizeof(float) >= huge ? huge : 64, bytes) != 0)
                    throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
                if (n * sizeof(float) >= huge)
                    madvise(memory, bytes, MADV_HUGEPAGE);
#endif
                base = static_cast<float *>(memory);
                count = n;
            }
        }
        zero();
    }
    void zero() {
        const size_t threads = count < (size_t(1) << 24) ? 1 : std::min<size_t>(8, std::max(1u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for (size_t t = 1; t < threads; t++)
            pool.emplace_back([this, t, threads]() {
                memset(base + count * t / threads, 0, (count * (t + 1) / threads - count * t / threads) * sizeof(float));
            });
        if (count)
            memset(base, 0, count / threads * sizeof(float));
        for (auto &thread : pool)
            thread.join();
    }

private:
    float *base = nullptr;
    size_t count = 0;
};

// RAII device allocation
struct DeviceArray {
    void *ptr = nullptr;
    size_t bytes = 0;
    DeviceArray() {}
    DeviceArray(const DeviceArray &) = delete;
    DeviceArray &operator=(const DeviceArray &) = delete;
    ~DeviceArray() { release(); }
    void release() {
        if (ptr)
            cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
    void allocate(size_t n) {
        if (n == bytes && ptr)
            return;
        release();
        if (n) {
            GV_CHECK_CUDA(cudaMalloc(&ptr, n));
            bytes = n;
        }
    }
    template<class T>
    T *as() const { return static_cast<T *>(ptr); }
    template<class T>
    void upload(const std::vector<T> &host, cudaStream_t stream = 0) {
        allocate(host.size() * sizeof(T));
        if (!host.empty()) {
            GV_CHECK_CUDA(cudaMemcpyAsync(ptr, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
            GV_CHECK_CUDA(cudaStreamSynchronize(stream));
        }
    }
    void upload(const HostMatrix &host, cudaStream_t stream = 0) {
        allocate(host.size() * sizeof(float));
        if (!host.empty()) {
            GV_CHECK_CUDA(cudaMemcpyAsync(ptr, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
            GV_CHECK_CUDA(cudaStreamSynchronize(stream));
        }
    }
};

// Host -> device copies of large PAGEABLE arrays (the graph's std::vectors).  cudaMemcpyAsync from pageable memory
// measured 1.4 GB/s on the B200 boxes (89 MB of CSR in 64 ms inside train()); here the bytes go through two page-locked
// staging buffers -- the CPU copies chunk i + 1 while the DMA engine moves chunk i.
struct StagedUploader {
    // bytes per staging buffer (GV_UPLOAD_CHUNK: a few KB makes the toy graphs of the tests take the chunked path)
    const size_t kChunk = getenv("GV_UPLOAD_CHUNK") ? std::max<size_t>(256, strtoull(getenv("GV_UPLOAD_CHUNK"), nullptr, 10))
                                                    : size_t(8) << 20;
    void *staging[2] = {nullptr, nullptr};
    cudaEvent_t moved[2] = {nullptr, nullptr};
    StagedUploader() {}
    StagedUploader(const StagedUploader &) = delete;
    StagedUploader &operator=(const StagedUploader &) = delete;
    ~StagedUploader() { release(); }
    void release() {
        for (int i = 0; i < 2; i++) {
            if (staging[i])
                cudaFreeHost(staging[i]);
            if (moved[i])
                cudaEventDestroy(moved[i]);
            staging[i] = nullptr;
            moved[i] = nullptr;
        }
    }
    // dst (device) <- src (host, any memory); returns when the bytes have left `src` and are queued on `stream`
    void copy(void *dst, const void *src, size_t bytes, cudaStream_t stream) {
        if (bytes < kChunk) {
            GV_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
            GV_CHECK_CUDA(cudaStreamSynchronize(stream));
            return;
        }
        for (int i = 0; i < 2; i++)
            if (!staging[i]) {
                GV_CHECK_CUDA(cudaMallocHost(&staging[i], kChunk));
                GV_CHECK_CUDA(cudaEventCreateWithFlags(&moved[i], cudaEventDisableTiming));
            }
        int which = 0;
        for (size_t done = 0; done < bytes; done += kChunk, which ^= 1) {
            const size_t n = std::min(kChunk, bytes - done);
            GV_CHECK_CUDA(cudaEventSynchronize(moved[which]));  // the copy that last used this buffer (none: returns at once)
            memcpy(staging[which], static_cast<const char *>(src) + done, n);
            GV_CHECK_CUDA(cudaMemcpyAsync(static_cast<char *>(dst) + done, staging[which], n, cudaMemcpyHostToDevice, stream));
            GV_CHECK_CUDA(cudaEventRecord(moved[which], stream));
        }
        GV_CHECK_CUDA(cudaStreamSynchronize(stream));
    }
    template<class T>
    void upload(DeviceArray &array, const std::vector<T> &host, cudaStream_t stream) {
        array.allocate(host.size() * sizeof(T));
        if (!host.empty())
            copy(array.ptr, host.data(), host.size() * sizeof(T), stream);
    }
};

// core/optimizer.h:42-134
struct HostOptimizer {
    gv_optimizer_t desc;
    float init_lr = 0;
    std::string type_name() const {
        static const char *names[] = {"SGD", "Momentum", "AdaGrad", "RMSprop", "Adam"};
        return desc.type < 0 ? "Default" : names[desc.type];
    }
    int num_moment() const { return desc.type <= GV_OPT_SGD ? 0 : (desc.type == GV_OPT_ADAM ? 2 : 1); }
    // LRSchedule::operator() + Optimizer::apply_schedule, core/optimizer.h:65-79,132-134
    float lr_at(int batch_id, int num_batch) const {
        float factor = 1;
        if (desc.schedule == GV_SCHEDULE_LINEAR)
            factor = std::max(1 - float(batch_id) / num_batch, 1e-4f);
        else if (desc.schedule == GV_SCHEDULE_CUSTOM && desc.schedule_fn)
            factor = desc.schedule_fn(batch_id, num_batch, desc.schedule_ctx);
        return init_lr * factor;
    }
    std::string info() const {  // Optimizer::info, core/optimizer.h:137-155
        static const char *schedules[] = {"constant", "linear", "custom"};
        std::stringstream ss;
        ss << "optimizer: " << type_name() << std::endl;
        ss << "learning rate: " << init_lr << ", lr schedule: " << schedules[desc.schedule] << std::endl;
        ss << "weight decay: " << desc.weight_decay;
        if (desc.type == GV_OPT_MOMENTUM)
            ss << std::endl << "momentum: " << desc.a;
        if (desc.type == GV_OPT_ADAGRAD)
            ss << std::endl << "epsilon: " << desc.epsilon;
        if (desc.type == GV_OPT_RMSPROP)
            ss << std::endl << "alpha: " << desc.a << ", epsilon: " << desc.epsilon;
        if (desc.type == GV_OPT_ADAM)
            ss << std::endl << "beta1: " << desc.a << ", beta2: " << desc.b << ", epsilon: " << desc.epsilon;
        return ss.str();
    }
};

// SolverMixin::partition, core/solver.h:873-887.  std::sort is unstable: the order of equal
// weights is whatever libstdc++'s introsort leaves, so the very same call is made here.
inline std::vector<std::vector<uint32_t>> partition_vertices(const std::vector<float> &weights, int num_partition) {
    std::vector<uint32_t> order(weights.size());
    for (uint32_t i = 0; i < order.size(); i++)
        order[i] = i;
    std::sort(order.begin(), order.end(), [&weights](uint32_t x, uint32_t y) { return weights[x] > weights[y]; });
    std::vector<std::vector<uint32_t>> parts(num_partition);
    const uint32_t period = num_partition * 2;
    for (uint32_t i = 0; i < order.size(); i++) {
        uint32_t slot = i % period;  // zig-zag deal: 0 1 .. P-1 P-1 .. 1 0
        if (slot >= uint32_t(num_partition))
            slot = period - 1 - slot;
        parts[slot].push_back(order[i]);
    }
    return parts;
}

// pretty::size_string, util/io.h
inline std::string size_string(uint64_t size) {
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
}

}  // namespace gv
