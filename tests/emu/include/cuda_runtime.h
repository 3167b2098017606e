// TEST INFRASTRUCTURE ONLY -- a CPU stand-in for the handful of CUDA constructs the cmda_b200 kernels use, so that the
// kernel SOURCE (a transformed copy made by tests/emu/build_emu.py) can be executed on the host for small inputs where
// no GPU exists: every CUDA thread of a block is a fiber (own stack, hand-rolled switch) on ONE host thread, __syncthreads and the warp
// collectives are rendezvous points, atomics are plain read-modify-writes (nothing runs concurrently).  It checks
// kernel LOGIC (indexing, barriers, packing, carry arithmetic); it is not a product path, not an oracle and says
// nothing about performance.  Never linked into libcmda_b200.so.
#pragma once
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <vector>

#define __CUDACC__ 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
#define __shared__ static

struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct ushort2 { unsigned short x, y; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaError_t { cudaSuccess = 0 };
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered, cudaMemoryTypeHost, cudaMemoryTypeDevice, cudaMemoryTypeManaged };
struct cudaPointerAttributes { cudaMemoryType type; };
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeHost; return cudaSuccess; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = std::malloc(1); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = std::malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = std::malloc(1); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

#if !defined(__x86_64__)
#error "tests/emu switches fibers with a few lines of x86-64 assembly"
#endif
// emu_switch(&save_sp, new_sp): push the callee-saved registers, park this stack pointer, adopt the other one.
// (ucontext's swapcontext costs a signal-mask system call per switch; a block of 512 threads switches a lot.)
extern "C" void emu_switch(void** save_sp, void* new_sp);
#ifdef EMU_DEFINE_SWITCH
asm(".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_switch,.-emu_switch\n");
#endif

namespace emu {
struct Fiber {
    void* sp;
    uint3 tid;
    bool done;
    std::vector<char> stack;
};
struct WarpBox {
    unsigned long long vals[32], snap[32];
    unsigned present, snap_present;     // lanes that deposited (threads that have exited take no part: CUDA's rule
    int count, alive;                   // for *_sync is "all non-exited threads named in the mask")
    unsigned gen;
};
struct State {
    std::vector<Fiber> fibers;
    std::vector<WarpBox> warps;
    void* sched_sp = nullptr;
    Fiber* cur = nullptr;
    int nthreads = 0, alive = 0, bar_count = 0;
    unsigned bar_gen = 0;
    unsigned long progress = 0;
    std::function<void()> body;
};
inline State& st() { static State s; return s; }
inline int linear_tid() { return static_cast<int>(st().cur - st().fibers.data()); }
inline void yield() { State& s = st(); emu_switch(&s.cur->sp, s.sched_sp); }
inline void trampoline() {
    State& s = st();
    s.body();
    s.cur->done = true;
    --s.alive;
    --s.warps[linear_tid() >> 5].alive;
    emu_switch(&s.cur->sp, s.sched_sp);      // never resumed
    std::abort();
}

// one block: every thread is a fiber, resumed round-robin until all have returned
inline void run_block(dim3 block) {
    State& s = st();
    const int n = static_cast<int>(block.x * block.y * block.z);
    s.nthreads = n;
    s.alive = n;
    s.bar_count = 0;
    if (static_cast<int>(s.fibers.size()) < n) s.fibers.resize(n);
    s.warps.assign((n + 31) / 32, WarpBox{});
    for (int i = 0; i < n; ++i) ++s.warps[i >> 5].alive;
    for (int i = 0; i < n; ++i) {
        Fiber& f = s.fibers[i];
        if (f.stack.empty()) f.stack.resize(256 * 1024);
        f.done = false;
        f.tid = uint3{static_cast<unsigned>(i) % block.x, (static_cast<unsigned>(i) / block.x) % block.y,
                      static_cast<unsigned>(i) / (block.x * block.y)};
        // first switch "returns" into trampoline with the stack the ABI expects at a function entry (rsp = 16k + 8)
        uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack.data()) + f.stack.size()) & ~static_cast<uintptr_t>(15);
        void** frame = reinterpret_cast<void**>(top) - 8;       // r15 r14 r13 r12 rbx rbp | return address | pad
        for (int k = 0; k < 8; ++k) frame[k] = nullptr;
        frame[6] = reinterpret_cast<void*>(&trampoline);
        f.sp = frame;
    }
    for (;;) {
        int alive = 0;
        const unsigned long before = s.progress;
        for (int i = 0; i < n; ++i) {
            Fiber& f = s.fibers[i];
            if (f.done) continue;
            ++alive;
            s.cur = &f;
            emu_switch(&s.sched_sp, f.sp);
            if (f.done) ++s.progress;
        }
        if (!alive) break;
        if (s.progress == before) {
            std::fprintf(stderr, "cuda_emu: deadlock (a barrier or warp collective some threads never reach)\n");
            std::abort();
        }
    }
    s.cur = nullptr;
}
}  // namespace emu

inline uint3 emu_block_idx{0, 0, 0};
inline dim3 emu_block_dim, emu_grid_dim;
#define threadIdx (emu::st().cur->tid)
#define blockIdx emu_block_idx
#define blockDim emu_block_dim
#define gridDim emu_grid_dim

// grid launch: blocks one after the other (x fastest)
template <typename F>
inline void emu_launch(dim3 grid, dim3 block, F&& body) {
    emu::st().body = body;
    emu_block_dim = block;
    emu_grid_dim = grid;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                emu_block_idx = uint3{x, y, z};
                emu::run_block(block);
            }
}

// barriers and warp collectives wait for the threads that have not exited (a waiter re-checks after every switch:
// the thread it was waiting for may have returned instead of arriving)
inline void __syncthreads() {
    emu::State& s = emu::st();
    const unsigned gen = s.bar_gen;
    ++s.bar_count;
    for (;;) {
        if (s.bar_gen != gen) return;
        if (s.bar_count >= s.alive) {
            s.bar_count = 0;
            ++s.bar_gen;
            ++s.progress;
            return;
        }
        emu::yield();
    }
}

inline int __syncthreads_or(int pred) {
    emu::State& s = emu::st();
    static int acc = 0, result = 0;
    acc |= pred != 0;
    const unsigned gen = s.bar_gen;
    ++s.bar_count;
    for (;;) {
        if (s.bar_gen != gen) return result;
        if (s.bar_count >= s.alive) {
            result = acc;
            acc = 0;
            s.bar_count = 0;
            ++s.bar_gen;
            ++s.progress;
            return result;
        }
        emu::yield();
    }
}

// warp rendezvous: every lane deposits a value and sees all 32 (full masks only, as everywhere in these kernels)
inline const emu::WarpBox& emu_warp_gather(unsigned mask, unsigned long long v) {
    if (mask != 0xffffffffu) { std::fprintf(stderr, "cuda_emu: partial warp mask\n"); std::abort(); }
    emu::State& s = emu::st();
    const int t = emu::linear_tid();
    emu::WarpBox& w = s.warps[t >> 5];
    w.vals[t & 31] = v;
    w.present |= 1u << (t & 31);
    const unsigned gen = w.gen;
    ++w.count;
    for (;;) {
        if (w.gen != gen) return w;
        if (w.count >= w.alive) {
            std::memcpy(w.snap, w.vals, sizeof(w.snap));
            w.snap_present = w.present;
            w.present = 0;
            w.count = 0;
            ++w.gen;
            ++s.progress;
            return w;
        }
        emu::yield();
    }
}
template <typename T>
inline unsigned long long emu_bits(T v) {
    static_assert(sizeof(T) <= 8, "");
    unsigned long long b = 0;
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
inline T emu_unbits(unsigned long long b) {
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}
inline int emu_lane() { return emu::linear_tid() & 31; }
template <typename T>
inline T __shfl_sync(unsigned m, T v, int src) { return emu_unbits<T>(emu_warp_gather(m, emu_bits(v)).snap[src & 31]); }
template <typename T>
inline T __shfl_up_sync(unsigned m, T v, unsigned d) {
    const unsigned long long* s = emu_warp_gather(m, emu_bits(v)).snap;
    const int l = emu_lane();
    return l >= static_cast<int>(d) ? emu_unbits<T>(s[l - d]) : v;
}
template <typename T>
inline T __shfl_xor_sync(unsigned m, T v, int x) { return emu_unbits<T>(emu_warp_gather(m, emu_bits(v)).snap[(emu_lane() ^ x) & 31]); }
inline unsigned __ballot_sync(unsigned m, int pred) {
    const emu::WarpBox& w = emu_warp_gather(m, pred ? 1ull : 0ull);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((w.snap_present >> i & 1u) && w.snap[i] ? 1u : 0u) << i;
    return r;
}
template <typename T>
inline T emu_reduce(unsigned m, T v, int op) {
    const emu::WarpBox& w = emu_warp_gather(m, emu_bits(v));
    T r = v;                                   // this lane is present by construction
    bool first = true;
    for (int i = 0; i < 32; ++i) {
        if (!(w.snap_present >> i & 1u)) continue;
        const T x = emu_unbits<T>(w.snap[i]);
        r = first ? x : op == 0 ? static_cast<T>(r + x) : op == 1 ? (x < r ? x : r) : (x > r ? x : r);
        first = false;
    }
    return r;
}
inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return emu_reduce(m, v, 0); }
inline int __reduce_add_sync(unsigned m, int v) { return emu_reduce(m, v, 0); }
inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return emu_reduce(m, v, 1); }
inline int __reduce_min_sync(unsigned m, int v) { return emu_reduce(m, v, 1); }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return emu_reduce(m, v, 2); }
inline int __reduce_max_sync(unsigned m, int v) { return emu_reduce(m, v, 2); }

// atomics: nothing runs concurrently
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p = static_cast<int>(static_cast<unsigned>(o) + static_cast<unsigned>(v)); return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline unsigned atomicMax(unsigned* p, unsigned v) { const unsigned o = *p; if (v > o) *p = v; return o; }
inline int atomicMax(int* p, int v) { const int o = *p; if (v > o) *p = v; return o; }
inline unsigned atomicMin(unsigned* p, unsigned v) { const unsigned o = *p; if (v < o) *p = v; return o; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {       // shf.r.wrap
    const unsigned long long v = (static_cast<unsigned long long>(hi) << 32) | lo;
    return static_cast<unsigned>(v >> (sh & 31u));
}

template <typename T>
inline T __ldg(const T* p) { return *p; }
// shared-window addresses (cvta.to.shared) are byte offsets from emu_shared_window, which the harness points at the
// dynamic shared array of the kernels that use them
extern char* emu_shared_window;
inline size_t __cvta_generic_to_shared(const void* p) { return static_cast<size_t>(static_cast<const char*>(p) - emu_shared_window); }
inline unsigned __umulhi(unsigned a, unsigned b) { return static_cast<unsigned>((static_cast<unsigned long long>(a) * b) >> 32); }
inline int __ffs(unsigned v) { return __builtin_ffs(static_cast<int>(v)); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __uint2float_rn(unsigned v) { return static_cast<float>(v); }
inline float __int2float_rn(int v) { return static_cast<float>(v); }
inline int __float2int_rz(float v) {            // cvt.rzi.s32.f32: NaN -> 0, saturating
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT_MAX;
    if (v <= -2147483648.0f) return INT_MIN;
    return static_cast<int>(v);
}
inline int __float2int_rn(float v) {            // cvt.rni.s32.f32
    if (v != v) return 0;
    const float r = std::nearbyintf(v);
    if (r >= 2147483648.0f) return INT_MAX;
    if (r <= -2147483648.0f) return INT_MIN;
    return static_cast<int>(r);
}
inline unsigned __float2uint_rn(float v) {      // cvt.rni.u32.f32: NaN -> 0, saturating
    if (v != v) return 0u;
    const float r = std::nearbyintf(v);
    if (r >= 4294967296.0f) return 0xffffffffu;
    if (r <= 0.0f) return 0u;
    return static_cast<unsigned>(r);
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {      // prmt.b32 (default mode)
    const unsigned long long pool = (static_cast<unsigned long long>(b) << 32) | a;
    unsigned r = 0u;
    for (int i = 0; i < 4; ++i) r |= static_cast<unsigned>((pool >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
inline long long __float2ll_rn(float v) {
    if (v != v) return 0;
    const float r = std::nearbyintf(v);
    if (r >= 9223372036854775808.0f) return LLONG_MAX;
    if (r <= -9223372036854775808.0f) return LLONG_MIN;
    return static_cast<long long>(r);
}
inline long long __double2ll_rn(double v) {
    if (v != v) return 0;
    const double r = std::nearbyint(v);
    if (r >= 9223372036854775808.0) return LLONG_MAX;
    if (r <= -9223372036854775808.0) return LLONG_MIN;
    return static_cast<long long>(r);
}
inline float __double2float_rn(double v) { return static_cast<float>(v); }
inline float __ll2float_rn(long long v) { return static_cast<float>(v); }
inline float __fsqrt_rn(float v) { return std::sqrt(v); }
inline long long __double_as_longlong(double v) { return static_cast<long long>(emu_bits(v)); }
inline void __syncwarp(unsigned = 0xffffffffu) {}      // fibers run one at a time between barriers: nothing to order
inline float __int_as_float(int v) { return emu_unbits<float>(static_cast<unsigned>(v)); }
inline int __float_as_int(float v) { return static_cast<int>(emu_unbits<unsigned>(emu_bits(v))); }
inline unsigned __float_as_uint(float v) { return emu_unbits<unsigned>(emu_bits(v)); }
inline float __uint_as_float(unsigned v) { return emu_unbits<float>(v); }
using std::isnan;
using std::isinf;
template <typename A, typename B>
inline typename std::common_type<A, B>::type min(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return static_cast<T>(a) < static_cast<T>(b) ? static_cast<T>(a) : static_cast<T>(b);
}
template <typename A, typename B>
inline typename std::common_type<A, B>::type max(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return static_cast<T>(a) > static_cast<T>(b) ? static_cast<T>(a) : static_cast<T>(b);
}
