// emu_runtime.cpp -- fiber scheduler of the SIMT emulator plus the subset of the CUDA runtime API that
// isaac_ros_apriltag_b200/csrc calls, implemented on host memory (see emu_cuda.h: TEST INFRASTRUCTURE ONLY).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <deque>
#include <functional>
#include <random>
#include <vector>

namespace emu {

struct Snapshot {
  unsigned long long v[32];
  int aux[32];
  int refs;
  Snapshot *next_free;
};

enum State { S_READY = 0, S_WAIT_BAR, S_WAIT_WARP, S_DONE };

struct Fiber {
  void *sp;
  char *stack;
  int state;
  uint3 tid;
  int lane, warp;
  unsigned bar_gen;
  unsigned red_seq;
  // pending warp collective
  unsigned c_mask;
  int c_kind;
  unsigned long long c_val;
  int c_aux;
  const void *c_site;
  Snapshot *c_snap;
  unsigned c_live;
};

Fiber *g_cur = nullptr;
uint3 g_blockIdx;
dim3 g_blockDim, g_gridDim;

static void *g_sched_sp = nullptr;
static std::vector<Fiber> g_fibers;
static int g_nthreads = 0, g_live = 0;
static unsigned g_bar_gen = 0;
static int g_bar_count = 0;
static Snapshot *g_free_snaps = nullptr;
static void (*g_fn)(void *) = nullptr;
static void *g_arg = nullptr;
const char *g_kernel_name = "?";
alignas(128) static unsigned char g_dyn_smem[256 * 1024];
static constexpr size_t kStack = 256 * 1024;
static bool g_strict = getenv("B200AT_EMU_STRICT") != nullptr;
static bool g_warned_exited = false;

extern "C" void emu_switch(void **save_sp, void *new_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
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
.size emu_switch,.-emu_switch
)");

const uint3 &cur_tid() { return g_cur->tid; }
int cur_lane() { return g_cur->lane; }
void *dyn_smem() { return g_dyn_smem; }

static inline void yield() { emu_switch(&g_cur->sp, g_sched_sp); }

static void fail(const char *what) {
  fprintf(stderr, "[emu] %s in kernel %s (block %u,%u,%u; thread %u lane %d warp %d)\n", what, g_kernel_name, g_blockIdx.x, g_blockIdx.y,
          g_blockIdx.z, g_cur ? g_cur->tid.x : 0, g_cur ? g_cur->lane : -1, g_cur ? g_cur->warp : -1);
  fflush(stderr);
  abort();
}

static void release_barrier_if_complete() {
  if (g_bar_count > 0 && g_bar_count >= g_live) {
    g_bar_count = 0;
    g_bar_gen++;
  }
}

void block_barrier() {
  Fiber *f = g_cur;
  g_bar_count++;
  if (g_bar_count >= g_live) {
    g_bar_count = 0;
    g_bar_gen++;
    return;
  }
  f->bar_gen = g_bar_gen;
  f->state = S_WAIT_BAR;
  while (g_bar_gen == f->bar_gen) yield();
  f->state = S_READY;
}

// __syncthreads_or / _and / _count: three accumulator slots in rotation.  Call k accumulates into slot k % 3 and clears slot
// (k + 1) % 3: nobody can be accumulating into that one yet (that needs barrier k to complete), and nobody can still be
// reading it (it was last read after barrier k - 2, and every thread has since arrived at barrier k - 1).
static int g_red_or[3], g_red_and[3], g_red_cnt[3];
void block_barrier_reduce(int pred, int *o_or, int *o_and, int *o_cnt) {
  Fiber *f = g_cur;
  const int slot = (int)(f->red_seq % 3u), next = (int)((f->red_seq + 1u) % 3u);
  f->red_seq++;
  g_red_or[next] = 0;
  g_red_and[next] = 1;
  g_red_cnt[next] = 0;
  g_red_or[slot] |= pred ? 1 : 0;
  g_red_and[slot] &= pred ? 1 : 0;
  g_red_cnt[slot] += pred ? 1 : 0;
  block_barrier();
  *o_or = g_red_or[slot];
  *o_and = g_red_and[slot];
  *o_cnt = g_red_cnt[slot];
}

static unsigned live_lanes_of_warp(int warp) {
  unsigned m = 0;
  const int base = warp * 32;
  for (int l = 0; l < 32 && base + l < g_nthreads; l++)
    if (g_fibers[base + l].state != S_DONE) m |= 1u << l;
  return m;
}

// try to complete the collective that `f` is waiting in: all live lanes of its mask must be waiting with the same mask
static bool try_complete(Fiber *f) {
  const int base = f->warp * 32;
  const unsigned live = live_lanes_of_warp(f->warp);
  const unsigned need = f->c_mask & live;
  if (g_strict && (f->c_mask & ~live) && base + 32 <= g_nthreads) fail("collective names an exited lane (B200AT_EMU_STRICT)");
  for (int l = 0; l < 32; l++) {
    if (!((need >> l) & 1)) continue;
    Fiber &o = g_fibers[base + l];
    if (o.state != S_WAIT_WARP || o.c_snap != nullptr) return false;
    // a lane of this mask that sits in a collective with ANOTHER mask has not arrived here yet (legal: e.g. a sub-group
    // finishing a masked operation while the rest of the warp already waits at the next full-mask one); a group that can
    // never complete shows up as a deadlock
    if (o.c_mask != f->c_mask) return false;
    if (o.c_kind != f->c_kind) fail("lanes of one mask arrived at different collective operations");
    // hardware needs the lanes of one collective at ONE instruction; the same source line reached through two inlined copies
    // (a lambda called from two places, say) is two instructions.  g++ also duplicates calls on its own (both arms of a
    // short-circuit `&&` in front of a ballot, for one), so under B200AT_EMU_STRICT this is a note to look at, not an error.
    if (g_strict && o.c_site != f->c_site) {
      static int notes = 0;
      if (notes++ < 8)
        fprintf(stderr, "[emu] note: collective kind %d, mask %08x in %s: lane %d at %p, lane %d at %p (addr2line -e <emu .so>)\n", f->c_kind,
                f->c_mask, g_kernel_name, f->lane, f->c_site, o.lane, o.c_site);
    }
  }
  Snapshot *s = g_free_snaps;
  if (s)
    g_free_snaps = s->next_free;
  else
    s = new Snapshot();
  s->refs = 0;
  for (int l = 0; l < 32; l++) {
    s->v[l] = 0;
    s->aux[l] = 0;
    if (!((need >> l) & 1)) continue;
    Fiber &o = g_fibers[base + l];
    s->v[l] = o.c_val;
    s->aux[l] = o.c_aux;
    o.c_snap = s;
    o.c_live = need;
    s->refs++;
  }
  return true;
}

Snapshot *exchange(unsigned mask, int kind, unsigned long long val, int aux, unsigned *live_mask, const void *site) {
  Fiber *f = g_cur;
  f->c_site = site;
  if (!((mask >> f->lane) & 1)) fail("calling lane is not in the collective's mask");
  f->c_mask = mask;
  f->c_kind = kind;
  f->c_val = val;
  f->c_aux = aux;
  f->c_snap = nullptr;
  f->state = S_WAIT_WARP;
  if (!try_complete(f)) {
    while (f->c_snap == nullptr) yield();
  }
  f->state = S_READY;
  *live_mask = f->c_live;
  if ((mask & ~f->c_live) && !g_warned_exited && f->warp * 32 + 32 <= g_nthreads) {
    g_warned_exited = true;
    fprintf(stderr, "[emu] note: a warp collective names exited lanes (kernel %s); they are ignored\n", g_kernel_name);
  }
  return f->c_snap;
}

void done(Snapshot *s) {
  g_cur->c_snap = nullptr;
  if (--s->refs == 0) {
    s->next_free = g_free_snaps;
    g_free_snaps = s;
  }
}

static void fiber_main() {
  g_fn(g_arg);
  Fiber *f = g_cur;
  f->state = S_DONE;
  g_live--;
  release_barrier_if_complete();
  // a lane that exits may complete a collective its warp mates are waiting in
  const int base = f->warp * 32;
  for (int l = 0; l < 32 && base + l < g_nthreads; l++) {
    Fiber &o = g_fibers[base + l];
    if (o.state == S_WAIT_WARP && o.c_snap == nullptr) {
      Fiber *save = g_cur;
      g_cur = &o;
      try_complete(&o);
      g_cur = save;
    }
  }
  for (;;) yield();
}

static std::vector<char *> g_stacks;

static void init_fiber(Fiber &f, int t, dim3 block) {
  if ((int)g_stacks.size() <= t) g_stacks.resize(t + 1, nullptr);
  if (!g_stacks[t]) {
    void *p = nullptr;
    if (posix_memalign(&p, 64, kStack) != 0) abort();
    g_stacks[t] = (char *)p;
  }
  f.stack = g_stacks[t];
  f.state = S_READY;
  f.tid.x = t % block.x;
  f.tid.y = (t / block.x) % block.y;
  f.tid.z = t / (block.x * block.y);
  f.lane = t & 31;
  f.warp = t >> 5;
  f.c_snap = nullptr;
  f.c_mask = 0;
  f.red_seq = 0;
  uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
  void **sp = (void **)top;
  *--sp = nullptr;                 // fake return address of fiber_main
  *--sp = (void *)&fiber_main;     // popped by emu_switch's ret
  for (int i = 0; i < 6; i++) *--sp = nullptr;
  f.sp = sp;
}

void run_grid(dim3 grid, dim3 block, size_t smem, void (*fn)(void *), void *arg) {
  if (smem > sizeof(g_dyn_smem)) {
    fprintf(stderr, "[emu] dynamic shared memory %zu exceeds the emulator's buffer\n", smem);
    abort();
  }
  const int nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || nthreads > 1024) {
    fprintf(stderr, "[emu] bad block size %d\n", nthreads);
    abort();
  }
  g_fn = fn;
  g_arg = arg;
  g_blockDim = block;
  g_gridDim = grid;
  g_nthreads = nthreads;
  if ((int)g_fibers.size() < nthreads) g_fibers.resize(nthreads);
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        g_blockIdx.x = bx;
        g_blockIdx.y = by;
        g_blockIdx.z = bz;
        for (int t = 0; t < nthreads; t++) init_fiber(g_fibers[t], t, block);
        g_live = nthreads;
        g_bar_count = 0;
        g_bar_gen = 0;
        g_red_or[0] = g_red_cnt[0] = 0;
        g_red_and[0] = 1;
        while (g_live > 0) {
          bool progress = false;
          for (int t = 0; t < nthreads; t++) {
            Fiber &f = g_fibers[t];
            if (f.state == S_DONE) continue;
            if (f.state == S_WAIT_BAR && f.bar_gen == g_bar_gen) continue;
            if (f.state == S_WAIT_WARP && f.c_snap == nullptr) {
              g_cur = &f;
              if (!try_complete(&f)) continue;
            }
            g_cur = &f;
            emu_switch(&g_sched_sp, f.sp);
            progress = true;
          }
          if (!progress && g_live > 0) {
            fprintf(stderr, "[emu] DEADLOCK in kernel %s block (%u,%u,%u): ", g_kernel_name, bx, by, bz);
            for (int t = 0; t < nthreads; t++) {
              Fiber &f = g_fibers[t];
              if (f.state == S_WAIT_BAR) fprintf(stderr, "t%d:bar ", t);
              if (f.state == S_WAIT_WARP) fprintf(stderr, "t%d:warp(mask %08x kind %d) ", t, f.c_mask, f.c_kind);
            }
            fprintf(stderr, "\n");
            abort();
          }
        }
      }
  g_cur = nullptr;
}

}  // namespace emu

// ---------------------------------------------------------------------------------------------------------------------
// CUDA runtime subset on host memory.
// Default mode: streams and events are tokens, every operation executes synchronously at issue -- a valid serialisation,
// because the library issues work in dependency order.
// Asynchronous mode (b200at_emu_async(1, seed) or B200AT_EMU_ASYNC=seed): every stream is a queue; kernels, async copies /
// memsets, event records and event waits are queue entries; the queues only run at synchronisation points, and then in a
// RANDOM interleaving that respects stream order and event dependencies and nothing else.  A missing dependency between two
// streams (the class of bug the synchronous mode and a lucky GPU run both hide) turns into wrong results for some seeds.
// ---------------------------------------------------------------------------------------------------------------------
namespace emu {

struct EmuEvent {
  double t_ms = 0.0;
  unsigned long long recorded = 0;  // record operations issued
  unsigned long long done = 0;      // ... completed
};
struct Op {
  std::function<void()> fn;
  EmuEvent *wait_ev = nullptr;
  unsigned long long wait_gen = 0;
  EmuEvent *rec_ev = nullptr;
  unsigned long long rec_gen = 0;
};
struct EmuStream {
  std::deque<Op> q;
};
static EmuStream g_default_stream;
static std::vector<EmuStream *> g_all_streams{&g_default_stream};
static bool g_async = false;
static std::mt19937 g_rng(1);

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static EmuStream *stream_of(cudaStream_t s) { return s ? reinterpret_cast<EmuStream *>(s) : &g_default_stream; }
static bool runnable(const Op &op) { return op.wait_ev == nullptr || op.wait_ev->done >= op.wait_gen; }
static void execute(Op &op) {
  if (op.rec_ev) {
    op.rec_ev->t_ms = now_ms();
    if (op.rec_ev->done < op.rec_gen) op.rec_ev->done = op.rec_gen;
  }
  if (op.fn) op.fn();
}
// run ONE runnable queue head, chosen at random; false if every queue is empty
static bool step() {
  EmuStream *cand[64];
  int n = 0;
  bool pending = false;
  for (EmuStream *st : g_all_streams) {
    if (st->q.empty()) continue;
    pending = true;
    if (runnable(st->q.front()) && n < 64) cand[n++] = st;
  }
  if (!pending) return false;
  if (n == 0) {
    fprintf(stderr, "[emu] stream DEADLOCK: every pending queue head waits for an event nobody will record\n");
    abort();
  }
  EmuStream *st = cand[g_rng() % (unsigned)n];
  Op op = std::move(st->q.front());
  st->q.pop_front();
  execute(op);
  return true;
}
static void drain_all() {
  while (step()) {
  }
}
static void drain_stream(EmuStream *st) {
  while (!st->q.empty()) step();
}
static void enqueue(cudaStream_t s, Op op) {
  if (!g_async) {
    execute(op);
    return;
  }
  stream_of(s)->q.push_back(std::move(op));
}
void enqueue_kernel(cudaStream_t s, const char *name, dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
  Op op;
  op.fn = [name, grid, block, smem, body]() {
    g_kernel_name = name;
    struct Ctx {
      const std::function<void()> *b;
    } ctx{&body};
    run_grid(grid, block, smem, [](void *p) { (*static_cast<Ctx *>(p)->b)(); }, &ctx);
  };
  enqueue(s, std::move(op));
}

}  // namespace emu

// B200AT_EMU_ASYNC=<seed> in the environment turns the asynchronous mode on for the whole process
static const bool g_async_from_env = []() {
  if (const char *e = getenv("B200AT_EMU_ASYNC")) {
    emu::g_async = true;
    emu::g_rng.seed((unsigned)atoi(e));
  }
  return true;
}();

extern "C" void b200at_emu_async(int on, unsigned seed) {
  emu::drain_all();
  emu::g_async = on != 0;
  emu::g_rng.seed(seed);
}

extern "C" {

static cudaError_t g_last = cudaSuccess;
static int g_sms = getenv("B200AT_EMU_SMS") ? atoi(getenv("B200AT_EMU_SMS")) : 2;

cudaError_t cudaGetDeviceCount(int *n) {
  *n = 1;
  return cudaSuccess;
}
cudaError_t cudaGetDevice(int *d) {
  *d = 0;
  return cudaSuccess;
}
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaDeviceSynchronize(void) {
  emu::drain_all();
  return cudaSuccess;
}
cudaError_t cudaGetLastError(void) {
  cudaError_t e = g_last;
  g_last = cudaSuccess;
  return e;
}
cudaError_t cudaPeekAtLastError(void) { return g_last; }
const char *cudaGetErrorName(cudaError_t e) { return e == cudaSuccess ? "cudaSuccess" : "cudaError(emu)"; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "error (emulator)"; }
cudaError_t cudaDeviceGetAttribute(int *v, enum cudaDeviceAttr attr, int) {
  if (attr == cudaDevAttrMultiProcessorCount)
    *v = g_sms;
  else if (attr == cudaDevAttrMaxSharedMemoryPerBlockOptin)
    *v = 227 * 1024;
  else
    *v = 0;
  return cudaSuccess;
}

// device allocations are filled with a poison pattern: code that depends on uninitialised memory shows up as a mismatch
cudaError_t cudaMalloc(void **p, size_t n) {
  emu::drain_all();
  void *q = nullptr;
  if (posix_memalign(&q, 256, n ? n : 256) != 0) return cudaErrorMemoryAllocation;
  memset(q, 0xCD, n);
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
  emu::drain_all();
  free(p);
  return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t n) {
  void *q = nullptr;
  if (posix_memalign(&q, 256, n ? n : 256) != 0) return cudaErrorMemoryAllocation;
  memset(q, 0, n);
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
cudaError_t cudaFreeHost(void *p) {
  emu::drain_all();
  free(p);
  return cudaSuccess;
}
cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void **pd, void *ph, unsigned) {
  *pd = ph;
  return cudaSuccess;
}
// B200AT_EMU_HOSTMEM=pageable makes host pointers look like unregistered memory
cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes *a, const void *p) {
  memset(a, 0, sizeof(*a));
  const char *m = getenv("B200AT_EMU_HOSTMEM");
  if (m && !strcmp(m, "pageable")) {
    a->type = cudaMemoryTypeUnregistered;
  } else {
    a->type = cudaMemoryTypeHost;
    a->devicePointer = const_cast<void *>(p);
    a->hostPointer = const_cast<void *>(p);
  }
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, enum cudaMemcpyKind) {
  emu::drain_all();
  memmove(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, enum cudaMemcpyKind, cudaStream_t st) {
  emu::Op op;
  op.fn = [d, s, n]() { memmove(d, s, n); };
  emu::enqueue(st, std::move(op));
  return cudaSuccess;
}
static void copy2d(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h) {
  for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
}
cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, enum cudaMemcpyKind) {
  emu::drain_all();
  copy2d(d, dp, s, sp, w, h);
  return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, enum cudaMemcpyKind, cudaStream_t st) {
  emu::Op op;
  op.fn = [d, dp, s, sp, w, h]() { copy2d(d, dp, s, sp, w, h); };
  emu::enqueue(st, std::move(op));
  return cudaSuccess;
}
cudaError_t cudaMemset(void *d, int v, size_t n) {
  emu::drain_all();
  memset(d, v, n);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st) {
  emu::Op op;
  op.fn = [d, v, n]() { memset(d, v, n); };
  emu::enqueue(st, std::move(op));
  return cudaSuccess;
}
cudaError_t cudaMemset2DAsync(void *d, size_t pitch, int v, size_t w, size_t h, cudaStream_t st) {
  emu::Op op;
  op.fn = [d, pitch, v, w, h]() {
    for (size_t r = 0; r < h; r++) memset((char *)d + r * pitch, v, w);
  };
  emu::enqueue(st, std::move(op));
  return cudaSuccess;
}

cudaError_t cudaStreamCreate(cudaStream_t *s) {
  emu::EmuStream *st = new emu::EmuStream();
  emu::g_all_streams.push_back(st);
  *s = reinterpret_cast<cudaStream_t>(st);
  return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return cudaStreamCreate(s); }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { return cudaStreamCreate(s); }
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) {
  *lo = 0;
  *hi = -5;
  return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  if (!s) return cudaSuccess;
  emu::EmuStream *st = emu::stream_of(s);
  emu::drain_stream(st);
  for (size_t i = 0; i < emu::g_all_streams.size(); i++)
    if (emu::g_all_streams[i] == st) {
      emu::g_all_streams.erase(emu::g_all_streams.begin() + (long)i);
      break;
    }
  delete st;
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) {
  emu::drain_stream(emu::stream_of(s));
  return cudaSuccess;
}
// (waiting for an event that has never been recorded is a no-op, as in CUDA; the wait refers to the records issued so far)
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
  emu::EmuEvent *ev = reinterpret_cast<emu::EmuEvent *>(e);
  if (ev->recorded == 0) return cudaSuccess;
  emu::Op op;
  op.wait_ev = ev;
  op.wait_gen = ev->recorded;
  emu::enqueue(s, std::move(op));
  return cudaSuccess;
}

cudaError_t cudaEventCreate(cudaEvent_t *e) {
  *e = reinterpret_cast<cudaEvent_t>(new emu::EmuEvent());
  return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
  emu::drain_all();
  delete reinterpret_cast<emu::EmuEvent *>(e);
  return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  emu::EmuEvent *ev = reinterpret_cast<emu::EmuEvent *>(e);
  emu::Op op;
  op.rec_ev = ev;
  op.rec_gen = ++ev->recorded;
  emu::enqueue(s, std::move(op));
  return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) {
  emu::EmuEvent *ev = reinterpret_cast<emu::EmuEvent *>(e);
  while (ev->done < ev->recorded && emu::step()) {
  }
  return cudaSuccess;
}
cudaError_t cudaEventQuery(cudaEvent_t e) {
  emu::EmuEvent *ev = reinterpret_cast<emu::EmuEvent *>(e);
  return ev->done >= ev->recorded ? cudaSuccess : cudaErrorNotReady;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = (float)(reinterpret_cast<emu::EmuEvent *>(b)->t_ms - reinterpret_cast<emu::EmuEvent *>(a)->t_ms);
  return cudaSuccess;
}

cudaError_t cudaFuncSetAttribute(const void *, enum cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, const void *, int, size_t) {
  *n = 1;
  return cudaSuccess;
}
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int *n, const void *, int, size_t, unsigned) {
  *n = 1;
  return cudaSuccess;
}
// no driver: TMA descriptors cannot be encoded, the kernels take their plain-load staging path
cudaError_t cudaGetDriverEntryPoint(const char *, void **fn, unsigned long long, enum cudaDriverEntryPointQueryResult *q) {
  *fn = nullptr;
  if (q) *q = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorNotSupported;
}
cudaError_t cudaGetDriverEntryPointByVersion(const char *, void **fn, unsigned, unsigned long long, enum cudaDriverEntryPointQueryResult *q) {
  *fn = nullptr;
  if (q) *q = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorNotSupported;
}
// stream capture is not emulated: the library falls back to plain launches
cudaError_t cudaStreamBeginCapture(cudaStream_t, enum cudaStreamCaptureMode) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g) {
  *g = nullptr;
  return cudaErrorNotSupported;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }

}  // extern "C"
