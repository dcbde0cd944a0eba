// Registry of execution contexts (context.h) and their C entry points (include/voldor_b200.h).
#include "context.h"
#include <atomic>
#include <cstdlib>
#include "../../include/py_export.h"
#include "../../include/voldor_b200.h"

namespace vb {

namespace {
// Several execution contexts with two streams each share the device's hardware work queues; with the default of 8 the
// streams of more than 6 contexts alias and serialise behind each other (measured with 32: +2 % EM-iterations/s at 8 and
// 12 windows in flight, no change at 6).  Read by the driver when the CUDA context is created, so it only takes effect
// when this library is loaded before that; a value set by the user wins.
__attribute__((constructor)) void default_work_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
std::atomic<int> g_device{-1};
std::mutex g_registry_mutex;
Context* g_contexts[kMaxContexts] = {nullptr};
thread_local int t_selected = 0;
}  // namespace

void enter_device() {
    int dev = g_device.load(std::memory_order_acquire);
    if (dev < 0) {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) return;  // no device: the CUDA calls that follow report it
        int expected = -1;
        g_device.compare_exchange_strong(expected, cur);
        dev = g_device.load();
    }
    cudaSetDevice(dev);
}

int set_library_device(int device) {
    const cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) g_device.store(device, std::memory_order_release);
    return (int)e;
}

Context* context_at(int id) {
    if (id < 0 || id >= kMaxContexts) return nullptr;
    std::lock_guard<std::mutex> lock(g_registry_mutex);
    if (!g_contexts[id]) g_contexts[id] = new Context(id);  // lives for the rest of the process, like the reference's statics
    return g_contexts[id];
}

Context& current_context() { return *context_at(t_selected); }

int select_context(int id) {
    if (!context_at(id)) return -1;
    const int prev = t_selected;
    t_selected = id;
    return prev;
}

}  // namespace vb

extern "C" {

VB_EXPORT int vb_context_select(int ctx) { return vb::select_context(ctx); }
VB_EXPORT int vb_context_current(void) { return vb::current_context().id; }
VB_EXPORT int vb_context_max(void) { return vb::kMaxContexts; }
VB_EXPORT int vb_context_srand(unsigned int seed) {
    vb::Context& cx = vb::current_context();
    std::lock_guard<std::recursive_mutex> lock(cx.mutex);
    cx.rnd->seed(seed);
    return 0;
}
VB_EXPORT int vb_debug_thread_device(void) {
    vb::enter_device();
    int dev = -1;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}
VB_EXPORT int vb_context_rand(void) {
    vb::Context& cx = vb::current_context();
    std::lock_guard<std::recursive_mutex> lock(cx.mutex);
    return cx.rnd->next();
}

}  // extern "C"
