// cuda_emu.h -- TEST INFRASTRUCTURE: a minimal SIMT emulator so that theiasfm_b200/csrc/tba_engine.cu -- the engine's real host
// code AND its real kernels -- can be compiled with g++ (-x c++ -DTBA_EMULATE -include this file) and run on a machine without a GPU
// (tests/emu/Makefile -> tests/emu/libtheia_ba_b200_emu.so; `pytest -m gpu --emulate-engine`).
//   * every CUDA thread of a block is a user-level fiber (own stack, hand-rolled x86-64 switch; ucontext elsewhere); blocks run one after the other, so `__shared__` is `static`;
//   * __syncthreads / __syncwarp / __shfl_*_sync / __ballot_sync / __all_sync suspend the fiber until the whole block / warp (the
//     lanes that have not exited) has arrived, then exchange the payloads -- the semantics the kernels rely on;
//   * the TMA bulk copy + mbarrier PTX of the matvec is replaced by a memcpy + flag (tba_kernels.cuh, #ifdef TBA_EMULATE);
//   * the CUDA runtime is malloc / memcpy; kernel launches are serialised by a mutex, so atomics are plain read-modify-writes;
//   * TBA_EMU_DEVICES "devices" (default 1) share that one emulator; NCCL is tests/emu/emu_nccl.c (shared memory).
// It checks control flow, indexing, reductions and the host glue -- not performance, not memory-model subtleties.
#pragma once
#include <execinfo.h>
#include <ucontext.h>
#include <unistd.h>

#include <csignal>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <numeric>
#include <map>
#include <mutex>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
inline emu_dim3 threadIdx, blockIdx, blockDim, gridDim;  // set by the scheduler before a fiber runs

struct alignas(16) double2 { double x, y; };
struct alignas(16) double4 { double x, y, z, w; };  // CUDA 12: __align__(16), two 128-bit accesses
inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct alignas(16) float4 { float x, y, z, w; };
template <class T> inline T __ldg(const T* p) { return *p; }

using std::atan; using std::atan2; using std::cos; using std::fabs; using std::fmax; using std::fmin; using std::isfinite; using std::isinf; using std::isnan; using std::sin;
using std::sqrt; using std::tan;

// ---------------------------------------------------------------- runtime
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { const char* e = std::getenv("TBA_EMU_DEVICES"); *n = e ? std::atoi(e) : 1; return cudaSuccess; }
// 256-byte aligned like the real allocator (vector accesses of the kernels rely on it), zero-filled like calloc was
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) {
  const size_t bytes = ((n ? n : 1) + 255) & ~(size_t)255;
  *p = (T*)std::aligned_alloc(256, bytes);
  // cudaMalloc does NOT zero memory: TBA_EMU_POISON=1 fills every allocation with 0xFF bytes (NaN doubles, -1 ints) so that a kernel
  // relying on zero-initialised buffers without a cudaMemset fails here as it would (sooner or later) on the GPU
  static const bool poison = std::getenv("TBA_EMU_POISON") != nullptr;
  if (*p) std::memset((void*)*p, poison ? 0xFF : 0, bytes);
  return *p ? cudaSuccess : cudaErrorEmu;
}
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) {
  *p = (T*)std::calloc(n ? n : 1, 1);
  if (*p && std::getenv("TBA_EMU_POISON")) std::memset((void*)*p, 0xFF, n ? n : 1);
  return *p ? cudaSuccess : cudaErrorEmu;
}
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
inline int emu_last_error = cudaSuccess;  // set by emu::launch on an invalid configuration, like cudaErrorInvalidConfiguration
inline cudaError_t cudaPeekAtLastError() { return emu_last_error; }
inline cudaError_t cudaGetLastError() { const int e = emu_last_error; emu_last_error = cudaSuccess; return e; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error (emulated)" : "invalid launch configuration (emulated)"; }
enum { cudaDevAttrMultiProcessorCount = 16 };
// a small "GPU" (TBA_EMU_SMS multiprocessors, default 3): persistent kernels then give every warp a range of several slices
inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
  const char* e = getenv("TBA_EMU_SMS");
  *v = attr == cudaDevAttrMultiProcessorCount ? (e ? atoi(e) : 3) : 0;
  return cudaSuccess;
}
inline std::mutex emu_attr_mu;
inline std::map<const void*, size_t> emu_max_dyn_smem;  // kernel -> opted-in dynamic shared memory (default limit 48 KB)
template <class F> inline cudaError_t cudaFuncSetAttribute(F f, int attr, int v) {
  if (attr == cudaFuncAttributeMaxDynamicSharedMemorySize) {
    if (v > 227 * 1024) return cudaErrorEmu;
    std::lock_guard<std::mutex> lk(emu_attr_mu);
    emu_max_dyn_smem[(const void*)f] = (size_t)v;
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------- device intrinsics without synchronisation
inline double atomicAdd(double* p, double v) { const double o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }  // volatile: no contraction into an FMA
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

// ---------------------------------------------------------------- fibers
namespace emu {
enum Wait { kRun = 0, kBlockBarrier = 1, kWarpOp = 2, kSpin = 3 };
#if defined(__x86_64__)
// Minimal context switch (callee-saved registers + stack pointer): glibc's swapcontext makes a sigprocmask system call per switch,
// which dominated the emulation time.
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_switch,.-emu_switch\n");
#define EMU_FAST_SWITCH 1
#endif
struct Fiber {
  ucontext_t ctx;
  void* sp = nullptr;
  char* stack = nullptr;
  bool done = true;
  Wait wait = kRun;
  int op = 0;            // warp collective kind: 0 sync, 1 shfl (payload / src lane), 2 ballot, 3 all
  uint64_t payload = 0;  // value contributed
  int src = 0;           // shfl: lane to read from (already resolved; -1 = keep own)
  int line = 0;          // source line of the collective's call site (TBA_EMU_STRICT: all lanes of one collective must agree)
  uint64_t result = 0;
};
inline std::vector<Fiber> fibers;
inline ucontext_t sched_ctx;
inline void* sched_sp = nullptr;
inline int cur = -1;
inline const std::function<void()>* body = nullptr;
inline std::vector<char> dyn_smem_buf;
inline int sched_order_mode() {
  const char* e = std::getenv("TBA_EMU_ORDER");
  if (!e) return 0;
  return e[0] == 'r' && e[1] == 'e' ? 1 : e[0] == 'r' && e[1] == 'a' ? 2 : 0;
}
constexpr size_t kStack = 512 * 1024;

#ifdef EMU_FAST_SWITCH
inline void to_sched() { emu_switch(&fibers[cur].sp, sched_sp); }
inline void to_fiber(Fiber& f) { emu_switch(&sched_sp, f.sp); }
#else
inline void to_sched() { swapcontext(&fibers[cur].ctx, &sched_ctx); }
inline void to_fiber(Fiber& f) { swapcontext(&sched_ctx, &f.ctx); }
#endif
inline void trampoline() {
  (*body)();
  fibers[cur].done = true;
  to_sched();
  std::abort();  // a finished fiber is never resumed
}
inline void suspend(Wait w) {
  fibers[cur].wait = w;
  to_sched();
}
template <class T> inline T* dyn_smem() { return reinterpret_cast<T*>(dyn_smem_buf.data()); }

inline void run_block(unsigned nthreads) {
  if (fibers.size() < nthreads) fibers.resize(nthreads);
  for (unsigned t = 0; t < nthreads; ++t) {
    Fiber& f = fibers[t];
    if (!f.stack) f.stack = (char*)std::malloc(kStack);
#ifdef EMU_FAST_SWITCH
    {  // frame emu_switch pops: r15 r14 r13 r12 rbx rbp, then `ret` into trampoline with rsp = 8 (mod 16) as after a call
      uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
      void** p = (void**)top;
      *--p = nullptr;                      // trampoline's (never used) return address
      *--p = (void*)(void (*)())trampoline;
      for (int r = 0; r < 6; ++r) *--p = nullptr;
      f.sp = p;
    }
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = kStack; f.ctx.uc_link = &sched_ctx;
    makecontext(&f.ctx, trampoline, 0);
#endif
    f.done = false; f.wait = kRun;
  }
  // TBA_EMU_ORDER=reverse|random: the order in which the runnable fibers of a block (and the blocks of a grid) are resumed.  A kernel that
  // is correct on the GPU cannot depend on it; a missing __syncthreads() / __syncwarp() between a producer and a consumer does
  // (ascending order hides "low thread writes, high thread reads", descending order the opposite, random order both sometimes).
  static const int order_mode = sched_order_mode();
  static thread_local std::vector<unsigned> perm;
  static thread_local uint64_t rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)(std::getenv("TBA_EMU_SEED") ? std::atoll(std::getenv("TBA_EMU_SEED")) : 1);
  for (;;) {
    bool progress = false, alive = false;
    if (order_mode == 2) {
      perm.resize(nthreads);
      for (unsigned i = 0; i < nthreads; ++i) perm[i] = i;
      for (unsigned i = nthreads; i > 1; --i) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; std::swap(perm[i - 1], perm[rng % i]); }
    }
    for (unsigned i = 0; i < nthreads; ++i) {
      const unsigned t = order_mode == 0 ? i : order_mode == 1 ? nthreads - 1 - i : perm[i];
      Fiber& f = fibers[t];
      if (f.done) continue;
      alive = true;
      if (f.wait == kRun || f.wait == kSpin) {
        cur = (int)t; threadIdx.x = t;
        const Wait before = f.wait;
        f.wait = kRun;
        to_fiber(f);
        if (f.done || f.wait != kSpin || before != kSpin) progress = true;
      }
    }
    if (!alive) break;
    // block barrier: every live thread waits
    bool all_bar = true;
    for (unsigned t = 0; t < nthreads; ++t) if (!fibers[t].done && fibers[t].wait != kBlockBarrier) { all_bar = false; break; }
    if (all_bar) { for (unsigned t = 0; t < nthreads; ++t) if (!fibers[t].done) fibers[t].wait = kRun; progress = true; }
    // warp collectives: every live lane of the warp waits
    for (unsigned w0 = 0; w0 < nthreads; w0 += 32) {
      const unsigned w1 = w0 + 32 < nthreads ? w0 + 32 : nthreads;
      bool any_live = false, all_wait = true;
      for (unsigned t = w0; t < w1; ++t) if (!fibers[t].done) { any_live = true; if (fibers[t].wait != kWarpOp) all_wait = false; }
      if (!any_live || !all_wait) continue;
      static const bool strict = std::getenv("TBA_EMU_STRICT") != nullptr;
      if (strict) {  // the lanes of one warp collective must come from ONE call site with ONE operation: on the GPU lanes that reach
                     // different *_sync call sites under a full mask do not exchange data with each other (undefined / deadlock)
        int op0 = -1, line0 = -1;
        for (unsigned t = w0; t < w1; ++t) if (!fibers[t].done) {
          if (op0 < 0) { op0 = fibers[t].op; line0 = fibers[t].line; }
          else if (fibers[t].op != op0 || fibers[t].line != line0) {
            std::fprintf(stderr, "cuda_emu: divergent warp collective in block %u warp %u: op %d line %d vs op %d line %d\n", blockIdx.x, w0 / 32, op0,
                         line0, fibers[t].op, fibers[t].line);
            std::abort();
          }
        }
      }
      uint64_t ballot = 0; bool all = true;
      for (unsigned t = w0; t < w1; ++t) if (!fibers[t].done) { if (fibers[t].payload) ballot |= 1ull << (t - w0); else all = false; }
      for (unsigned t = w0; t < w1; ++t) {
        Fiber& f = fibers[t];
        if (f.done) continue;
        if (f.op == 1) { const int s = f.src; f.result = (s >= 0 && w0 + s < w1 && !fibers[w0 + s].done) ? fibers[w0 + s].payload : f.payload; }
        else if (f.op == 2) f.result = ballot;
        else if (f.op == 3) f.result = all ? 1 : 0;
      }
      for (unsigned t = w0; t < w1; ++t) if (!fibers[t].done) fibers[t].wait = kRun;
      progress = true;
    }
    if (!progress) {
      // only spinners left that made no progress: a real deadlock would hang the GPU as well
      bool only_spin = true;
      for (unsigned t = 0; t < nthreads; ++t) if (!fibers[t].done && fibers[t].wait != kSpin) only_spin = false;
      static int idle = 0;
      if (!only_spin || ++idle > 1000000) { std::fprintf(stderr, "cuda_emu: deadlock in block %u\n", blockIdx.x); std::abort(); }
    }
  }
}

inline std::mutex launch_mu;  // one emulated device: the ranks of tba_solve_multi (one host thread each) take turns
inline void segv_backtrace(int) {  // TBA_EMU_BACKTRACE=1: frames for `addr2line -e libtheia_ba_b200_emu.so`
  void* fr[64];
  const int n = backtrace(fr, 64);
  backtrace_symbols_fd(fr, n, 2);
  _exit(139);
}
template <class F>
inline void launch(const void* kernel, unsigned grid, unsigned block, size_t smem, F&& fn) {
  static const bool traced = [] { if (std::getenv("TBA_EMU_BACKTRACE")) { std::signal(SIGSEGV, segv_backtrace); std::signal(SIGABRT, segv_backtrace); } return true; }();
  (void)traced;
  std::lock_guard<std::mutex> lk(launch_mu);
  size_t smem_limit = 48 * 1024;
  { std::lock_guard<std::mutex> la(emu_attr_mu); const auto opt = emu_max_dyn_smem.find(kernel); if (opt != emu_max_dyn_smem.end()) smem_limit = opt->second; }
  if (grid == 0 || block == 0 || block > 1024 || smem > smem_limit) {  // what the driver would refuse
    std::fprintf(stderr, "cuda_emu: invalid launch configuration <<<%u, %u, %zu>>>\n", grid, block, smem);
    emu_last_error = cudaErrorEmu;
    return;
  }
  const std::function<void()> f = fn;
  body = &f;
  if (dyn_smem_buf.size() < smem + 256) dyn_smem_buf.resize(smem + 256);
  gridDim.x = grid; blockDim.x = block;
  const int order_mode = sched_order_mode();
  static const bool poison_smem = std::getenv("TBA_EMU_POISON") != nullptr;
  for (unsigned i = 0; i < grid; ++i) {
    // reverse: last block first; random: a fixed odd stride through the grid (a permutation when coprime with the grid size)
    unsigned b = order_mode == 1 ? grid - 1 - i : i;
    if (order_mode == 2) { unsigned stride = 7919u; while (std::gcd(stride, grid) != 1u) ++stride; b = (unsigned)(((uint64_t)i * stride + 3u) % grid); }
    if (poison_smem && smem) std::memset(dyn_smem_buf.data(), 0xFF, smem);  // shared memory is not zeroed at block start either
    blockIdx.x = b; run_block(block);
  }
  body = nullptr;
}

inline uint64_t warp_op(int op, uint64_t payload, int src, int line = 0) {
  Fiber& f = fibers[cur];
  f.op = op; f.payload = payload; f.src = src; f.line = line;
  suspend(kWarpOp);
  return fibers[cur].result;
}
}  // namespace emu

inline void __threadfence() {}
inline void __threadfence_system() {}
inline void __syncthreads() { emu::suspend(emu::kBlockBarrier); }
inline void __syncwarp(unsigned = 0xffffffffu, int line = __builtin_LINE()) { emu::warp_op(0, 0, -1, line); }
inline void emu_yield() { emu::suspend(emu::kSpin); }
template <class T> inline uint64_t emu_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T emu_from(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
template <class T> inline T __shfl_sync(unsigned, T v, int src_lane, int = 32, int line = __builtin_LINE()) { return emu_from<T>(emu::warp_op(1, emu_bits(v), src_lane & 31, line)); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned delta, int = 32, int line = __builtin_LINE()) {
  const int lane = threadIdx.x & 31; const int s = lane + (int)delta;
  return emu_from<T>(emu::warp_op(1, emu_bits(v), s < 32 ? s : -1, line));
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned delta, int = 32, int line = __builtin_LINE()) {
  const int lane = threadIdx.x & 31; const int s = lane - (int)delta;
  return emu_from<T>(emu::warp_op(1, emu_bits(v), s >= 0 ? s : -1, line));
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32, int line = __builtin_LINE()) { return emu_from<T>(emu::warp_op(1, emu_bits(v), (threadIdx.x & 31) ^ m, line)); }
inline unsigned __ballot_sync(unsigned, int pred, int line = __builtin_LINE()) { return (unsigned)emu::warp_op(2, pred ? 1 : 0, -1, line); }
inline int __all_sync(unsigned, int pred, int line = __builtin_LINE()) { return (int)emu::warp_op(3, pred ? 1 : 0, -1, line); }
