// TEST INFRASTRUCTURE ONLY (oracle/): page-locked allocation shim for the reference build.
//
// The reference filter allocates its output with cudaMallocHost (STPSmartDeviceMemory::makeHost,
// /root/reference/SuperTerrain+/SuperAlgorithm+/Host/Private/STPSingleHistogramFilter.cpp:80,105), which the CUDA
// headers forward to cudaHostAlloc. The reference object is built once in the GPU-less container and then travels
// to the GPU box, so these symbols are resolved here: try the real CUDA runtime first (genuinely pinned memory on a
// GPU host), fall back to malloc where no driver is present.
#include <dlfcn.h>
#include <cstdlib>
#include <mutex>
#include <unordered_set>

namespace {
typedef int (*host_alloc_fn)(void**, size_t, unsigned int);
typedef int (*host_free_fn)(void*);
struct Runtime {
    host_alloc_fn alloc = nullptr;
    host_free_fn release = nullptr;
    std::mutex lock;
    std::unordered_set<void*> pinned;
    Runtime() {
        if (std::getenv("SHF_ORACLE_NO_PINNED")) return;
        const char* names[] = {"libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so.12"};
        for (const char* n : names) {
            if (void* h = dlopen(n, RTLD_NOW | RTLD_LOCAL)) {
                alloc = reinterpret_cast<host_alloc_fn>(dlsym(h, "cudaHostAlloc"));
                release = reinterpret_cast<host_free_fn>(dlsym(h, "cudaFreeHost"));
                if (alloc && release) break;
                alloc = nullptr; release = nullptr;
            }
        }
    }
};
Runtime& rt() { static Runtime r; return r; }
}

extern "C" {
int cudaHostAlloc(void** p, size_t bytes, unsigned int flags) {
    Runtime& r = rt();
    if (r.alloc) {
        void* q = nullptr;
        if (r.alloc(&q, bytes, flags) == 0 && q) {
            std::lock_guard<std::mutex> g(r.lock);
            r.pinned.insert(q);
            *p = q;
            return 0;
        }
        r.alloc = nullptr;  // no usable device/driver: stop asking
    }
    *p = std::malloc(bytes ? bytes : 1);
    return *p ? 0 : 2;
}
int cudaFreeHost(void* p) {
    Runtime& r = rt();
    {
        std::lock_guard<std::mutex> g(r.lock);
        auto it = r.pinned.find(p);
        if (it != r.pinned.end()) {
            r.pinned.erase(it);
            return r.release(p);
        }
    }
    std::free(p);
    return 0;
}
int cudaFree(void*) { return 0; }
int cudaFreeAsync(void*, void*) { return 0; }
const char* cudaGetErrorString(int) { return "cuda host shim error"; }
int shf_ref_pinned_active() { return rt().alloc != nullptr; }
}
