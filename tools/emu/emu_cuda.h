// emu_cuda.h -- SIMT emulator shim: lets the UNMODIFIED kernel sources of isaac_ros_apriltag_b200/csrc compile with g++ and run on
// the CPU, one fiber per CUDA thread, so kernel logic (barriers, warp collectives, atomics, shared-memory aliasing) can be
// checked against the oracle in a container without a GPU.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under isaac_ros_apriltag_b200/ loads the emulated library; the product path is
// libb200apriltags.so (sm_100a) and fails loudly without a GPU.  The emulated build is never timed and never shipped.
//
// Model: blocks of a launch run one after the other; inside a block every thread is a fiber with its own stack, scheduled
// round-robin and switched only at __syncthreads / warp collectives.  A collective with mask m completes when every live lane
// named in m has arrived at a collective of the same kind with the same mask (anything else is reported as a divergence bug);
// a full scheduler pass without progress is reported as a deadlock.  Atomics are plain read-modify-writes (one host thread).
// Streams: synchronous by default (every operation runs at issue); b200at_emu_async(1, seed) queues them and runs the queues in
// random dependency-respecting interleavings (emu_runtime.cpp) -- a check of the host code's event wiring.
// Force-included (g++ -include) in front of every translated source; see tools/emu/build_emu.py.
#pragma once
#define B200AT_EMU 1

// qualifiers: defined before the CUDA headers so that crt/host_defines.h keeps them
#define __host__
#define __device__
#define __global__
#define __shared__ static
#define __constant__ static
#define __grid_constant__
#define __managed__

#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <functional>
#include <type_traits>

#undef __launch_bounds__
#define __launch_bounds__(...)

// function-pointer overloads that cuda_runtime.h only declares under nvcc
template <class T>
inline cudaError_t cudaFuncSetAttribute(T *f, enum cudaFuncAttribute a, int v) {
  return ::cudaFuncSetAttribute(reinterpret_cast<const void *>(f), a, v);
}
template <class T>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, T *f, int bs, size_t smem) {
  return ::cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, reinterpret_cast<const void *>(f), bs, smem);
}

namespace emu {

struct Snapshot {
  unsigned long long v[32];
  int aux[32];
  int refs;
  Snapshot *next_free;
};

struct Fiber;
extern Fiber *g_cur;
extern uint3 g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern const char *g_kernel_name;

const uint3 &cur_tid();
int cur_lane();
void *dyn_smem();

void block_barrier();
void block_barrier_reduce(int pred, int *o_or, int *o_and, int *o_cnt);
enum Kind { K_SYNCWARP = 1, K_SHFL, K_SHFL_UP, K_SHFL_DOWN, K_SHFL_XOR, K_BALLOT, K_MATCH, K_VOTE, K_REDUX };
// all live lanes of `mask` deposit (val, aux); returns the group's snapshot (release it with done())
Snapshot *exchange(unsigned mask, int kind, unsigned long long val, int aux, unsigned *live_mask, const void *site = nullptr);
void done(Snapshot *s);

typedef void (*ThreadFn)(void *);
void run_grid(dim3 grid, dim3 block, size_t smem, ThreadFn fn, void *arg);
// a launch is an entry of its stream's queue (executed at once in the synchronous mode): `body` holds COPIES of the arguments
void enqueue_kernel(cudaStream_t s, const char *name, dim3 grid, dim3 block, size_t smem, std::function<void()> body);

template <class F>
inline void launch(const char *name, dim3 grid, dim3 block, size_t smem, cudaStream_t s, F f) {
  enqueue_kernel(s, name, grid, block, smem, std::function<void()>(f));
}
template <class F>
inline void launch(const char *name, dim3 grid, dim3 block, size_t smem, F f) {
  enqueue_kernel(nullptr, name, grid, block, smem, std::function<void()>(f));
}
template <class F>
inline void launch(const char *name, dim3 grid, dim3 block, F f) {
  enqueue_kernel(nullptr, name, grid, block, 0, std::function<void()>(f));
}

template <class T>
inline unsigned long long to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  unsigned long long b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(unsigned long long b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}

}  // namespace emu

#define threadIdx (emu::cur_tid())
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

// ---- barriers / warp collectives ----
inline void __syncthreads() { emu::block_barrier(); }
inline int __syncthreads_or(int pred) {
  int o, a, c;
  emu::block_barrier_reduce(pred, &o, &a, &c);
  return o;
}
inline int __syncthreads_and(int pred) {
  int o, a, c;
  emu::block_barrier_reduce(pred, &o, &a, &c);
  return a;
}
inline int __syncthreads_count(int pred) {
  int o, a, c;
  emu::block_barrier_reduce(pred, &o, &a, &c);
  return c;
}
__attribute__((noinline)) inline void __syncwarp(unsigned mask = 0xffffffffu) {
  unsigned live;
  emu::done(emu::exchange(mask, emu::K_SYNCWARP, 0, 0, &live, __builtin_return_address(0)));
}
inline void __threadfence() {}
inline void __threadfence_block() {}

template <class T>
__attribute__((noinline)) inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  unsigned live;
  emu::Snapshot *s = emu::exchange(mask, emu::K_SHFL, emu::to_bits(v), 0, &live, __builtin_return_address(0));
  const int lane = emu::cur_lane();
  const int base = lane & ~(width - 1);
  const int sl = base + (src & (width - 1));
  T r = (live >> sl) & 1 ? emu::from_bits<T>(s->v[sl]) : v;  // reading an inactive lane is undefined: keep own value
  emu::done(s);
  return r;
}
template <class T>
__attribute__((noinline)) inline T __shfl_xor_sync(unsigned mask, T v, int lm, int width = 32) {
  unsigned live;
  emu::Snapshot *s = emu::exchange(mask, emu::K_SHFL_XOR, emu::to_bits(v), 0, &live, __builtin_return_address(0));
  const int lane = emu::cur_lane();
  const int sl = lane ^ lm;
  T r = (sl < 32 && ((live >> sl) & 1) && (sl & ~(width - 1)) == (lane & ~(width - 1))) ? emu::from_bits<T>(s->v[sl]) : v;
  emu::done(s);
  return r;
}
template <class T>
__attribute__((noinline)) inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
  unsigned live;
  emu::Snapshot *s = emu::exchange(mask, emu::K_SHFL_UP, emu::to_bits(v), 0, &live, __builtin_return_address(0));
  const int lane = emu::cur_lane();
  const int sl = lane - (int)d;
  T r = (sl >= (lane & ~(width - 1)) && ((live >> sl) & 1)) ? emu::from_bits<T>(s->v[sl]) : v;
  emu::done(s);
  return r;
}
template <class T>
__attribute__((noinline)) inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
  unsigned live;
  emu::Snapshot *s = emu::exchange(mask, emu::K_SHFL_DOWN, emu::to_bits(v), 0, &live, __builtin_return_address(0));
  const int lane = emu::cur_lane();
  const int sl = lane + (int)d;
  T r = (sl < (lane & ~(width - 1)) + width && sl < 32 && ((live >> sl) & 1)) ? emu::from_bits<T>(s->v[sl]) : v;
  emu::done(s);
  return r;
}
__attribute__((noinline)) inline unsigned __ballot_sync(unsigned mask, int pred) {
  unsigned live;
  emu::Snapshot *s = emu::exchange(mask, emu::K_BALLOT, pred ? 1ull : 0ull, 0, &live, __builtin_return_address(0));
  unsigned r = 0;
  for (int l = 0; l < 32; l++)
    if (((live >> l) & 1) && s->v[l]) r |= 1u << l;
  emu::done(s);
  return r;
}
__attribute__((noinline)) inline int __all_sync(unsigned mask, int pred) {
  unsigned live;
  emu::Snapshot *s = emu::exchange(mask, emu::K_VOTE, pred ? 1ull : 0ull, 0, &live, __builtin_return_address(0));
  int r = 1;
  for (int l = 0; l < 32; l++)
    if (((live >> l) & 1) && !s->v[l]) r = 0;
  emu::done(s);
  return r;
}
__attribute__((noinline)) inline int __any_sync(unsigned mask, int pred) {
  unsigned live;
  emu::Snapshot *s = emu::exchange(mask, emu::K_VOTE, pred ? 1ull : 0ull, 0, &live, __builtin_return_address(0));
  int r = 0;
  for (int l = 0; l < 32; l++)
    if (((live >> l) & 1) && s->v[l]) r = 1;
  emu::done(s);
  return r;
}
template <class T>
__attribute__((noinline)) inline unsigned __match_any_sync(unsigned mask, T v) {
  unsigned live;
  const unsigned long long b = emu::to_bits(v);
  emu::Snapshot *s = emu::exchange(mask, emu::K_MATCH, b, 0, &live, __builtin_return_address(0));
  unsigned r = 0;
  for (int l = 0; l < 32; l++)
    if (((live >> l) & 1) && s->v[l] == b) r |= 1u << l;
  emu::done(s);
  return r;
}
#define EMU_REDUX(name, T, init, op)                                                  \
  __attribute__((noinline)) inline T name(unsigned mask, T v) {                                                 \
    unsigned live;                                                                    \
    emu::Snapshot *s = emu::exchange(mask, emu::K_REDUX, emu::to_bits(v), 0, &live, __builtin_return_address(0));  \
    T r = init;                                                                       \
    for (int l = 0; l < 32; l++)                                                      \
      if ((live >> l) & 1) {                                                          \
        const T x = emu::from_bits<T>(s->v[l]);                                       \
        r = op;                                                                       \
      }                                                                               \
    emu::done(s);                                                                     \
    return r;                                                                         \
  }
EMU_REDUX(__reduce_add_sync, unsigned, 0u, r + x)
EMU_REDUX(__reduce_add_sync, int, 0, r + x)
EMU_REDUX(__reduce_min_sync, unsigned, 0xffffffffu, (x < r ? x : r))
EMU_REDUX(__reduce_min_sync, int, 0x7fffffff, (x < r ? x : r))
EMU_REDUX(__reduce_max_sync, unsigned, 0u, (x > r ? x : r))
EMU_REDUX(__reduce_max_sync, int, (-0x7fffffff - 1), (x > r ? x : r))
EMU_REDUX(__reduce_or_sync, unsigned, 0u, r | x)
EMU_REDUX(__reduce_and_sync, unsigned, 0xffffffffu, r &x)
#undef EMU_REDUX
inline unsigned __activemask() { return 0xffffffffu; }

// ---- atomics (single host thread: plain read-modify-write) ----
template <class T, class U>
inline T atomicAdd(T *p, U v) {
  T o = *p;
  *p = (T)(o + (T)v);
  return o;
}
template <class T, class U>
inline T atomicSub(T *p, U v) {
  T o = *p;
  *p = (T)(o - (T)v);
  return o;
}
template <class T, class U>
inline T atomicMin(T *p, U v) {
  T o = *p;
  if ((T)v < o) *p = (T)v;
  return o;
}
template <class T, class U>
inline T atomicMax(T *p, U v) {
  T o = *p;
  if ((T)v > o) *p = (T)v;
  return o;
}
template <class T, class U>
inline T atomicOr(T *p, U v) {
  T o = *p;
  *p = (T)(o | (T)v);
  return o;
}
template <class T, class U>
inline T atomicAnd(T *p, U v) {
  T o = *p;
  *p = (T)(o & (T)v);
  return o;
}
template <class T, class U>
inline T atomicExch(T *p, U v) {
  T o = *p;
  *p = (T)v;
  return o;
}
template <class T, class U, class V>
inline T atomicCAS(T *p, U cmp, V v) {
  T o = *p;
  if (o == (T)cmp) *p = (T)v;
  return o;
}

// ---- cache-hinted loads / stores ----
template <class T>
inline T __ldg(const T *p) {
  return *p;
}
template <class T>
inline T __ldcg(const T *p) {
  return *p;
}
template <class T>
inline T __ldcs(const T *p) {
  return *p;
}
template <class T>
inline T __ldca(const T *p) {
  return *p;
}
template <class T>
inline void __stcg(T *p, T v) {
  *p = v;
}
template <class T>
inline void __stcs(T *p, T v) {
  *p = v;
}
template <class T>
inline void __stwt(T *p, T v) {
  *p = v;
}
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }

// ---- integer intrinsics ----
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  const unsigned long long xy = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    const unsigned sel = (s >> (4 * i)) & 0xf;
    unsigned b = (unsigned)(xy >> (8 * (sel & 7))) & 0xffu;
    if (sel & 8) b = (b & 0x80u) ? 0xffu : 0u;
    r |= b << (8 * i);
  }
  return r;
}
#define EMU_SIMD4(name, expr)                                \
  inline unsigned name(unsigned a, unsigned b) {             \
    unsigned r = 0;                                          \
    for (int i = 0; i < 4; i++) {                            \
      const unsigned x = (a >> (8 * i)) & 0xffu, y = (b >> (8 * i)) & 0xffu; \
      r |= ((unsigned)(expr) & 0xffu) << (8 * i);            \
    }                                                        \
    return r;                                                \
  }
EMU_SIMD4(__vminu4, (x < y ? x : y))
EMU_SIMD4(__vmaxu4, (x > y ? x : y))
EMU_SIMD4(__vcmpgtu4, (x > y ? 0xffu : 0u))
EMU_SIMD4(__vcmpgeu4, (x >= y ? 0xffu : 0u))
EMU_SIMD4(__vcmpltu4, (x < y ? 0xffu : 0u))
EMU_SIMD4(__vcmpleu4, (x <= y ? 0xffu : 0u))
EMU_SIMD4(__vcmpeq4, (x == y ? 0xffu : 0u))
EMU_SIMD4(__vcmpne4, (x != y ? 0xffu : 0u))
EMU_SIMD4(__vadd4, (x + y))
EMU_SIMD4(__vsub4, (x - y))
EMU_SIMD4(__vavgu4, ((x + y + 1) >> 1))
EMU_SIMD4(__vabsdiffu4, (x > y ? x - y : y - x))
EMU_SIMD4(__vsubus4, (x > y ? x - y : 0u))
EMU_SIMD4(__vaddus4, (x + y > 255u ? 255u : x + y))
#undef EMU_SIMD4
inline unsigned __float_as_uint(float f) { return emu::from_bits<unsigned>(emu::to_bits(f)); }
inline int __float_as_int(float f) { return emu::from_bits<int>(emu::to_bits(f)); }
inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(u); }
inline float __int_as_float(int u) { return emu::from_bits<float>((unsigned)u); }
inline long long __double_as_longlong(double d) { return emu::from_bits<long long>(emu::to_bits(d)); }
inline double __longlong_as_double(long long l) { return emu::from_bits<double>((unsigned long long)l); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline double __dsqrt_rn(double a) { return sqrt(a); }
inline double __drcp_rn(double a) { return 1.0 / a; }
inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
inline double rsqrt(double a) { return 1.0 / sqrt(a); }
inline void __nanosleep(unsigned) {}
inline void __trap() { __builtin_trap(); }

// ---- min / max overload set of the CUDA headers ----
inline int min(int a, int b) { return a < b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline float min(float a, float b) { return fminf(a, b); }
inline double min(double a, double b) { return fmin(a, b); }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
inline float max(float a, float b) { return fmaxf(a, b); }
inline double max(double a, double b) { return fmax(a, b); }
inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
