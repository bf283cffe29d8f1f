// =============================================================================
// tests/emu/cuda_emu.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see cuda_emu.h).
//
// 1. The fiber scheduler behind gv_emu::run(): CTAs one after another; inside a CTA one fiber per
//    CUDA thread on its own small stack, switched cooperatively at warp collectives and barriers.
//    Warps are scheduled one at a time until every lane waits at a CTA barrier or has exited, so a
//    collective costs one context switch per lane.  A rendezvous that can never complete (divergent
//    __syncthreads, a lane missing from a *_sync mask) is reported and aborts instead of hanging.
// 2. A malloc-backed subset of the CUDA runtime API (the calls libgv_b200 makes): "device" memory is
//    host memory, streams execute immediately, events carry wall-clock time stamps.
// All launches of a process are serialised by one mutex (the host runtime launches from two threads).
// =============================================================================
#include "cuda_emu.h"

#include <dlfcn.h>
#include <execinfo.h>
#include <fcntl.h>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>

#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

// the built-in variables are defined (writable) in cuda_emu_vars.cpp; the scheduler sets them through these
void gv_emu_set_thread_index(unsigned x, unsigned y, unsigned z);
void gv_emu_set_block_index(unsigned x, unsigned y, unsigned z);
void gv_emu_set_dimensions(dim3 grid, dim3 block);

// ---- context switch (x86-64 SysV: callee-saved registers + stack pointer) ---------------------------
extern "C" void gv_emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl gv_emu_switch
.type gv_emu_switch,@function
gv_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size gv_emu_switch,.-gv_emu_switch
)");

namespace gv_emu {

namespace {

constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;
constexpr int kMaxBarriers = 16;

struct Fiber {
    void *sp = nullptr;
    char *stack = nullptr;
    bool done = true;
    // waiting: runnable again once *wait_generation != wait_value
    const int *wait_generation = nullptr;
    int wait_value = 0;
    bool waits_for_block = false;
};

struct Warp {
    unsigned existing = 0;  // lanes that exist in this CTA
    unsigned arrived = 0, mask = 0;
    int generation = 0;
    uint64_t slot[2][32];
};

struct Barrier {
    int arrived = 0, generation = 0, expected = 0;
};

std::mutex g_launch_mutex;
std::vector<Fiber> g_fibers(kMaxThreads);
std::vector<Warp> g_warps(kMaxThreads / 32);
Barrier g_barriers[kMaxBarriers];
void *g_scheduler_sp = nullptr;
int g_current = -1, g_num_thread = 0, g_exited = 0;
dim3 g_block;
const std::function<void()> *g_body = nullptr;
std::vector<unsigned char> g_shared;
bool g_in_kernel = false;

[[noreturn]] void die(const std::string &message) {
    fprintf(stderr, "cuda_emu: %s (block %u,%u thread %d)\n", message.c_str(), blockIdx.x, blockIdx.y, g_current);
    fflush(stderr);
    abort();
}

void set_thread_index(int t) {
    gv_emu_set_thread_index(t % g_block.x, t / g_block.x % g_block.y, t / (g_block.x * g_block.y));
}

void yield() {
    Fiber &self = g_fibers[g_current];
    gv_emu_switch(&self.sp, g_scheduler_sp);
}

void complete_barrier_if_ready(Barrier &barrier, bool counts_exited) {
    const int expected = counts_exited ? g_num_thread - g_exited : barrier.expected;
    if (barrier.arrived > 0 && barrier.arrived >= expected) {
        barrier.arrived = 0;
        barrier.generation++;
    }
}

void fiber_entry() {
    (*g_body)();
    Fiber &self = g_fibers[g_current];
    self.done = true;
    g_exited++;
    complete_barrier_if_ready(g_barriers[0], true);  // exited threads no longer take part in __syncthreads
    gv_emu_switch(&self.sp, g_scheduler_sp);
    die("a finished fiber was resumed");
}

void prepare(Fiber &fiber) {
    if (!fiber.stack) {
        void *memory = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (memory == MAP_FAILED)
            die("cannot allocate a fiber stack");
        fiber.stack = static_cast<char *>(memory);
    }
    uintptr_t top = reinterpret_cast<uintptr_t>(fiber.stack + kStackBytes) & ~uintptr_t(15);
    void **sp = reinterpret_cast<void **>(top - 16);
    *sp = reinterpret_cast<void *>(&fiber_entry);  // "return address" of the first switch
    sp -= 6;                                       // rbp rbx r12 r13 r14 r15
    for (int i = 0; i < 6; i++)
        sp[i] = nullptr;
    fiber.sp = sp;
    fiber.done = false;
    fiber.wait_generation = nullptr;
    fiber.waits_for_block = false;
}

bool runnable(const Fiber &fiber) {
    return !fiber.done && (!fiber.wait_generation || *fiber.wait_generation != fiber.wait_value);
}

void run_block() {
    g_exited = 0;
    const int num_warp = (g_num_thread + 31) / 32;
    for (int w = 0; w < num_warp; w++) {
        Warp &warp = g_warps[w];
        const int lanes = std::min(32, g_num_thread - w * 32);
        warp.existing = lanes == 32 ? 0xFFFFFFFFu : ((1u << lanes) - 1);
        warp.arrived = warp.mask = 0;
        warp.generation = 0;
    }
    for (auto &barrier : g_barriers)
        barrier = Barrier();
    for (int t = 0; t < g_num_thread; t++)
        prepare(g_fibers[t]);
    // GV_EMU_WARP_ORDER=reverse|rotate: visit the warps of a CTA in another order (the order decides which warp runs
    // ahead of a barrier first, i.e. which write-after-read hazards between warps become visible)
    static const int order_mode = []() {
        const char *mode = getenv("GV_EMU_WARP_ORDER");
        return !mode ? 0 : (std::string(mode) == "reverse" ? 1 : (std::string(mode) == "rotate" ? 2 : 0));
    }();
    // GV_EMU_LANE_ORDER=reverse: visit the lanes of a warp from 31 down to 0.  Lanes run one after another up to their
    // next collective, so the visiting order decides whether a lane sees a shared-memory write of another lane that no
    // __syncwarp() orders: in ascending order lane 5 sees what lane 0 wrote "by luck", in descending order it does not
    static const bool reverse_lanes = []() {
        const char *mode = getenv("GV_EMU_LANE_ORDER");
        return mode && std::string(mode) == "reverse";
    }();
    int sweep = 0;
    while (g_exited < g_num_thread) {
        bool progress = false;
        sweep++;
        for (int visit = 0; visit < num_warp; visit++) {
            const int w = order_mode == 1 ? num_warp - 1 - visit : (order_mode == 2 ? (visit + sweep) % num_warp : visit);
            // run this warp until all of its lanes wait for the CTA or have exited
            for (bool warp_progress = true; warp_progress;) {
                warp_progress = false;
                const int lane_end = std::min(g_num_thread, w * 32 + 32);
                for (int lane = w * 32; lane < lane_end; lane++) {
                    const int t = reverse_lanes ? w * 32 + (lane_end - 1 - lane) : lane;
                    Fiber &fiber = g_fibers[t];
                    if (!runnable(fiber))
                        continue;
                    fiber.wait_generation = nullptr;
                    g_current = t;
                    set_thread_index(t);
                    gv_emu_switch(&g_scheduler_sp, fiber.sp);
                    warp_progress = progress = true;
                }
            }
        }
        if (!progress) {
            int warp_waiters = 0, block_waiters = 0;
            for (int t = 0; t < g_num_thread; t++)
                if (!g_fibers[t].done)
                    (g_fibers[t].waits_for_block ? block_waiters : warp_waiters)++;
            g_current = -1;
            die("deadlock: " + std::to_string(warp_waiters) + " thread(s) wait in a warp collective and " +
                std::to_string(block_waiters) + " at a CTA barrier that can never complete (" +
                std::to_string(g_exited) + " of " + std::to_string(g_num_thread) + " threads have exited)");
        }
    }
    g_current = -1;
}

}  // namespace

void run(dim3 grid, dim3 block, size_t shared_bytes, const std::function<void()> &thread_body) {
    const unsigned long long threads = (unsigned long long)block.x * block.y * block.z;
    if (threads == 0 || threads > kMaxThreads)
        die("invalid CTA size " + std::to_string(threads));
    if ((unsigned long long)grid.x * grid.y * grid.z == 0)
        die("empty grid");
    if (shared_bytes > 227 * 1024)
        die("more than 227 KB of dynamic shared memory requested");
    std::lock_guard<std::mutex> lock(g_launch_mutex);
    if (g_in_kernel)
        die("nested kernel launch");
    g_in_kernel = true;
    g_body = &thread_body;
    g_block = block;
    g_num_thread = int(threads);
    gv_emu_set_dimensions(grid, block);
    // dynamic shared memory + a canary behind it: a CTA that writes past the bytes it asked for aborts the run
    constexpr size_t kCanary = 256;
    g_shared.assign(shared_bytes + 16 + kCanary, 0xA5);
    const unsigned char *canary = static_cast<const unsigned char *>(dynamic_shared()) + shared_bytes;
    for (unsigned z = 0; z < grid.z; z++)
        for (unsigned y = 0; y < grid.y; y++)
            for (unsigned x = 0; x < grid.x; x++) {
                gv_emu_set_block_index(x, y, z);
                run_block();
                for (size_t i = 0; i < kCanary; i++)
                    if (canary[i] != 0xA5)
                        die("a CTA wrote " + std::to_string(i + 1) + "+ bytes past its " + std::to_string(shared_bytes) +
                            " bytes of dynamic shared memory");
            }
    g_body = nullptr;
    g_in_kernel = false;
}

void *resolve_kernel(void *kernel) {
    static void *library = nullptr;
    static std::string from, to;
    static bool ready = false;
    if (!ready) {
        ready = true;
        const char *path = getenv("GV_EMU_DEVICE_LIBRARY"), *rename = getenv("GV_EMU_DEVICE_NAMESPACE");
        if (path && *path) {
            library = dlopen(path, RTLD_NOW | RTLD_LOCAL);
            if (!library)
                die(std::string("cannot load the device-pass library: ") + dlerror());
            const std::string spec = rename ? rename : "";
            const size_t equal = spec.find('=');
            if (equal == std::string::npos)
                die("GV_EMU_DEVICE_NAMESPACE must be from=to");
            // Itanium mangling spells a namespace as <length><name>
            from = std::to_string(equal) + spec.substr(0, equal);
            to = std::to_string(spec.size() - equal - 1) + spec.substr(equal + 1);
        }
    }
    if (!library)
        return kernel;
    Dl_info info;
    if (!dladdr(kernel, &info) || !info.dli_sname)
        die("cannot name the kernel of a launch (is it exported from its shared object?)");
    std::string name = info.dli_sname;
    for (size_t at = name.find(from); at != std::string::npos; at = name.find(from, at + to.size()))
        name.replace(at, from.size(), to);
    void *twin = dlsym(library, name.c_str());
    if (!twin)
        die("the device-pass library has no kernel " + name);
    return twin;
}

void *dynamic_shared() {
    return reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(g_shared.data()) + 15) & ~uintptr_t(15));
}

int lane_id() {
    return g_current & 31;
}

const uint64_t *warp_gather(unsigned mask, uint64_t value) {
    Warp &warp = g_warps[g_current / 32];
    const int lane = g_current & 31;
    mask &= warp.existing;
    if (!(mask >> lane & 1))
        die("a lane outside the mask entered a warp collective");
    if (warp.arrived == 0)
        warp.mask = mask;
    else if (warp.mask != mask)
        die("lanes of one warp entered a collective with different masks");
    const int generation = warp.generation;
    warp.slot[generation & 1][lane] = value;
    warp.arrived |= 1u << lane;
    if (warp.arrived == mask) {
        warp.arrived = 0;
        warp.generation++;
    } else {
        Fiber &self = g_fibers[g_current];
        self.wait_generation = &warp.generation;
        self.wait_value = generation;
        self.waits_for_block = false;
        yield();
    }
    return warp.slot[generation & 1];
}

void block_barrier(int id, int threads) {
    if (id < 0 || id >= kMaxBarriers)
        die("barrier id out of range");
    Barrier &barrier = g_barriers[id];
    const bool whole_block = threads <= 0;
    if (!whole_block) {
        if (threads % 32 != 0 || threads > g_num_thread)
            die("bar.sync with a thread count that is not a multiple of 32 or exceeds the CTA");
        if (barrier.arrived > 0 && barrier.expected != threads)
            die("threads entered one named barrier with different counts");
        barrier.expected = threads;
    }
    const int generation = barrier.generation;
    barrier.arrived++;
    complete_barrier_if_ready(barrier, whole_block);
    if (barrier.generation == generation) {
        Fiber &self = g_fibers[g_current];
        self.wait_generation = &barrier.generation;
        self.wait_value = generation;
        self.waits_for_block = true;
        yield();
    }
}

void misaligned(const void *pointer, size_t alignment) {
    die("misaligned " + std::to_string(alignment) + "-byte access at " + std::to_string(reinterpret_cast<uintptr_t>(pointer)));
}

}  // namespace gv_emu

// GV_EMU_BACKTRACE=1: print a native backtrace on SIGSEGV (the fault may be on a fiber stack)
namespace {
void segv_handler(int) {
    void *frames[64];
    const int n = backtrace(frames, 64);
    const char message[] = "cuda_emu: SIGSEGV, native backtrace:\n";
    (void)!write(2, message, sizeof(message) - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(139);
}
struct InstallHandler {
    InstallHandler() {
        if (!getenv("GV_EMU_BACKTRACE"))
            return;
        static char stack[1 << 16];
        stack_t alt = {stack, 0, sizeof(stack)};
        sigaltstack(&alt, nullptr);
        struct sigaction action = {};
        action.sa_handler = segv_handler;
        action.sa_flags = SA_ONSTACK;
        sigaction(SIGSEGV, &action, nullptr);
    }
} g_install_handler;
}  // namespace

unsigned __activemask() {
    return 0xFFFFFFFFu;
}

void __nanosleep(unsigned) {
    sched_yield();  // a spinning "kernel" waits for another PROCESS (an emulated peer device): let it run
}

unsigned long long gv_global_timer_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// =============================================================================
// CUDA runtime subset
// =============================================================================
namespace {

thread_local cudaError_t t_last_error = cudaSuccess;

cudaError_t remember(cudaError_t error) {
    if (error != cudaSuccess)
        t_last_error = error;
    return error;
}

struct EmuEvent {
    double seconds = 0;
};

double now_seconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

cudaError_t cudaGetLastError(void) {
    const cudaError_t error = t_last_error;
    t_last_error = cudaSuccess;
    return error;
}

cudaError_t cudaPeekAtLastError(void) {
    return t_last_error;
}

const char *cudaGetErrorString(cudaError_t error) {
    switch (error) {
        case cudaSuccess: return "no error";
        case cudaErrorMemoryAllocation: return "out of memory (emulated)";
        case cudaErrorNotSupported: return "operation not supported by the CUDA emulation";
        case cudaErrorInvalidValue: return "invalid argument";
        default: return "emulated CUDA error";
    }
}

const char *cudaGetErrorName(cudaError_t error) {
    return cudaGetErrorString(error);
}

cudaError_t cudaGetDeviceCount(int *count) {
    *count = 1;
    return cudaSuccess;
}

cudaError_t cudaSetDevice(int device) {
    return remember(device == 0 ? cudaSuccess : cudaErrorInvalidDevice);
}

cudaError_t cudaGetDevice(int *device) {
    *device = 0;
    return cudaSuccess;
}

cudaError_t cudaDeviceGetAttribute(int *value, enum cudaDeviceAttr attribute, int) {
    switch (attribute) {
        case cudaDevAttrMultiProcessorCount: *value = 2; break;  // keeps persistent grids small
        case cudaDevAttrMaxSharedMemoryPerBlockOptin: *value = 227 * 1024; break;
        case cudaDevAttrWarpSize: *value = 32; break;
        default: *value = 0;
    }
    return cudaSuccess;
}

cudaError_t cudaMemGetInfo(size_t *free_bytes, size_t *total_bytes) {
    *free_bytes = size_t(160) << 30;  // pretend to be a roomy B200: the solver's automatic choices follow it
    *total_bytes = size_t(180) << 30;
    return cudaSuccess;
}

cudaError_t cudaDeviceSynchronize(void) {
    return cudaSuccess;
}

// GV_EMU_IPC=1: device allocations live in POSIX shared memory so that cudaIpcGetMemHandle /
// cudaIpcOpenMemHandle work between the processes of an emulated multi-GPU run (peer stores over "NVLink"
// are then plain stores into the other process's pool, and the peer-exchange kernels really wait for each other)
struct SharedAllocation {
    std::string name;
    size_t bytes;
};
static std::mutex g_allocation_mutex;
static std::map<void *, SharedAllocation> g_shared_allocations, g_opened_allocations;
static bool ipc_enabled() {
    static const bool on = getenv("GV_EMU_IPC") != nullptr && atoi(getenv("GV_EMU_IPC")) != 0;
    return on;
}
static void unlink_all_shared() {
    for (auto &entry : g_shared_allocations)
        shm_unlink(entry.second.name.c_str());
}

// GV_EMU_GUARD=0 switches the guard pages off (they cost a mapping per allocation)
struct GuardedAllocation {
    void *base;
    size_t bytes;
};
static std::map<void *, GuardedAllocation> g_guarded_allocations;
static bool guard_enabled() {
    static const bool on = !(getenv("GV_EMU_GUARD") != nullptr && atoi(getenv("GV_EMU_GUARD")) == 0);
    return on;
}

cudaError_t cudaMalloc(void **pointer, size_t bytes) {
    const size_t rounded = (std::max<size_t>(bytes, 1) + 255) / 256 * 256;
    void *memory = nullptr;
    if (ipc_enabled()) {
        static unsigned long long counter = 0;
        std::lock_guard<std::mutex> lock(g_allocation_mutex);
        if (counter == 0)
            atexit(unlink_all_shared);
        const std::string name = "/gv_emu_" + std::to_string(getpid()) + "_" + std::to_string(counter++);
        const int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, off_t(rounded)) != 0) {
            if (fd >= 0)
                close(fd);
            return remember(cudaErrorMemoryAllocation);
        }
        memory = mmap(nullptr, rounded, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (memory == MAP_FAILED) {
            shm_unlink(name.c_str());
            return remember(cudaErrorMemoryAllocation);
        }
        g_shared_allocations[memory] = {name, rounded};
    } else if (guard_enabled()) {
        // the allocation ends at an inaccessible page: a kernel (or copy) that runs past it faults immediately,
        // with a backtrace under GV_EMU_BACKTRACE=1 -- the emulation's stand-in for compute-sanitizer memcheck
        const size_t page = 4096, usable = (rounded + page - 1) / page * page;
        char *base = static_cast<char *>(mmap(nullptr, usable + page, PROT_READ | PROT_WRITE,
                                              MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
        if (base == MAP_FAILED)
            return remember(cudaErrorMemoryAllocation);
        mprotect(base + usable, page, PROT_NONE);
        memory = base + usable - rounded;  // 256-byte aligned: rounded is a multiple of 256
        std::lock_guard<std::mutex> lock(g_allocation_mutex);
        g_guarded_allocations[memory] = {base, usable + page};
    } else {
        memory = aligned_alloc(256, rounded);
        if (!memory)
            return remember(cudaErrorMemoryAllocation);
    }
    if (rounded <= (size_t(64) << 20))
        memset(memory, 0xA5, rounded);  // cudaMalloc does not zero: make reads of unwritten memory visible
    *pointer = memory;
    return cudaSuccess;
}

cudaError_t cudaFree(void *pointer) {
    if (!pointer)
        return cudaSuccess;
    {
        std::lock_guard<std::mutex> lock(g_allocation_mutex);
        auto found = g_shared_allocations.find(pointer);
        if (found != g_shared_allocations.end()) {
            munmap(pointer, found->second.bytes);
            shm_unlink(found->second.name.c_str());
            g_shared_allocations.erase(found);
            return cudaSuccess;
        }
        auto guarded = g_guarded_allocations.find(pointer);
        if (guarded != g_guarded_allocations.end()) {
            munmap(guarded->second.base, guarded->second.bytes);
            g_guarded_allocations.erase(guarded);
            return cudaSuccess;
        }
    }
    free(pointer);
    return cudaSuccess;
}

cudaError_t cudaMallocHost(void **pointer, size_t bytes) {
    return cudaMalloc(pointer, bytes);
}

cudaError_t cudaFreeHost(void *pointer) {
    return cudaFree(pointer);
}

cudaError_t cudaHostRegister(void *, size_t, unsigned int) {
    return cudaSuccess;
}

cudaError_t cudaHostUnregister(void *) {
    return cudaSuccess;
}

cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, enum cudaMemcpyKind) {
    if (bytes)
        memmove(dst, src, bytes);
    return cudaSuccess;
}

cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, enum cudaMemcpyKind, cudaStream_t) {
    if (bytes)
        memmove(dst, src, bytes);
    return cudaSuccess;
}

cudaError_t cudaMemset(void *pointer, int value, size_t bytes) {
    if (bytes)
        memset(pointer, value, bytes);
    return cudaSuccess;
}

cudaError_t cudaMemsetAsync(void *pointer, int value, size_t bytes, cudaStream_t) {
    if (bytes)
        memset(pointer, value, bytes);
    return cudaSuccess;
}

cudaError_t cudaStreamCreate(cudaStream_t *stream) {
    *stream = reinterpret_cast<cudaStream_t>(new int(0));
    return cudaSuccess;
}

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *stream, unsigned int) {
    return cudaStreamCreate(stream);
}

cudaError_t cudaStreamCreateWithPriority(cudaStream_t *stream, unsigned int, int) {
    return cudaStreamCreate(stream);
}

cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest) {
    *least = 0;
    *greatest = -1;
    return cudaSuccess;
}

cudaError_t cudaStreamDestroy(cudaStream_t stream) {
    delete reinterpret_cast<int *>(stream);
    return cudaSuccess;
}

cudaError_t cudaStreamSynchronize(cudaStream_t) {
    return cudaSuccess;
}

cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) {
    return cudaSuccess;
}

cudaError_t cudaEventCreate(cudaEvent_t *event) {
    *event = reinterpret_cast<cudaEvent_t>(new EmuEvent());
    return cudaSuccess;
}

cudaError_t cudaEventCreateWithFlags(cudaEvent_t *event, unsigned int) {
    return cudaEventCreate(event);
}

cudaError_t cudaEventDestroy(cudaEvent_t event) {
    delete reinterpret_cast<EmuEvent *>(event);
    return cudaSuccess;
}

cudaError_t cudaEventRecord(cudaEvent_t event, cudaStream_t) {
    reinterpret_cast<EmuEvent *>(event)->seconds = now_seconds();
    return cudaSuccess;
}

cudaError_t cudaEventSynchronize(cudaEvent_t) {
    return cudaSuccess;
}

cudaError_t cudaEventQuery(cudaEvent_t) {
    return cudaSuccess;
}

cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t begin, cudaEvent_t end) {
    *ms = float((reinterpret_cast<EmuEvent *>(end)->seconds - reinterpret_cast<EmuEvent *>(begin)->seconds) * 1e3);
    return cudaSuccess;
}

cudaError_t cudaFuncSetAttribute(const void *, enum cudaFuncAttribute, int) {
    return cudaSuccess;
}

cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *blocks, const void *, int, size_t) {
    *blocks = 1;
    return cudaSuccess;
}

cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int *blocks, const void *, int, size_t, unsigned int) {
    *blocks = 1;
    return cudaSuccess;
}

// CUDA IPC between emulated devices (= processes): see GV_EMU_IPC above; without it there is no peer
// mapping and the solver falls back to replicated sampling
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *handle, void *pointer) {
    std::lock_guard<std::mutex> lock(g_allocation_mutex);
    auto found = g_shared_allocations.find(pointer);
    if (found == g_shared_allocations.end())
        return remember(cudaErrorNotSupported);
    memset(handle, 0, sizeof(*handle));
    unsigned long long bytes = found->second.bytes;
    memcpy(handle->reserved, &bytes, sizeof(bytes));
    strncpy(handle->reserved + sizeof(bytes), found->second.name.c_str(), sizeof(handle->reserved) - sizeof(bytes) - 1);
    return cudaSuccess;
}

cudaError_t cudaIpcOpenMemHandle(void **pointer, cudaIpcMemHandle_t handle, unsigned int) {
    unsigned long long bytes = 0;
    memcpy(&bytes, handle.reserved, sizeof(bytes));
    const char *name = handle.reserved + sizeof(bytes);
    if (bytes == 0 || name[0] != '/')
        return remember(cudaErrorNotSupported);
    const int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0)
        return remember(cudaErrorInvalidValue);
    void *memory = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (memory == MAP_FAILED)
        return remember(cudaErrorMemoryAllocation);
    std::lock_guard<std::mutex> lock(g_allocation_mutex);
    g_opened_allocations[memory] = {name, size_t(bytes)};
    *pointer = memory;
    return cudaSuccess;
}

cudaError_t cudaIpcCloseMemHandle(void *pointer) {
    std::lock_guard<std::mutex> lock(g_allocation_mutex);
    auto found = g_opened_allocations.find(pointer);
    if (found == g_opened_allocations.end())
        return remember(cudaErrorInvalidValue);
    munmap(pointer, found->second.bytes);
    g_opened_allocations.erase(found);
    return cudaSuccess;
}

}  // extern "C"
