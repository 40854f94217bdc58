// Shared device helpers for the sm_100a kernels: mbarrier + 1-D bulk-async
// (TMA) copies, proxy fences, named barriers, warp reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/yolo_b200.h"

// Where the last failing CUDA runtime call of this thread was made ("file:line"), for
// yb_last_error_site(): a positive status alone (a cudaError_t) does not say which call it was.
namespace yb {
const char*& last_error_site();
}
#define YB_STR2(x) #x
#define YB_STR(x) YB_STR2(x)
#define YB_CUDA_TRY(expr)                                          \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) {                                   \
            yb::last_error_site() = __FILE__ ":" YB_STR(__LINE__); \
            (void)cudaGetLastError(); /* leave no stale error for the caller's own runtime calls */ \
            return (int)_e;                                        \
        }                                                          \
    } while (0)

namespace yb {

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep, do not spin
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- bulk async copies (1-D TMA; SASS: UBLKCP) ------------------------------
// global -> shared, completion counted on an mbarrier.  16-byte aligned src/dst/size.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (before a bulk store).
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- named barriers --------------------------------------------------------
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- warp reductions -------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__host__ __device__ __forceinline__ size_t align_up(size_t x, size_t a) {
    return (x + a - 1) / a * a;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel that never
// changes once raised: set it the first time a device sees the kernel instead of on every call
// (the call costs ~1 us on the launch path), and again only when a launch needs more.  Setting the
// attribute twice is harmless, so the guard needs no lock (idempotent cache, not state).
struct SmemRaised {
    int bytes[64];   // per device ordinal: the largest value set so far (zero-initialised static)
};
template <typename K>
inline cudaError_t raise_dynamic_smem_once(K kernel, int bytes, SmemRaised* state) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int* slot = &state->bytes[dev & 63];
    if (bytes <= __atomic_load_n(slot, __ATOMIC_ACQUIRE)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) __atomic_store_n(slot, bytes, __ATOMIC_RELEASE);   // racing raises are harmless
    return e;
}

}  // namespace yb
