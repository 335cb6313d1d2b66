// TEST INFRASTRUCTURE ONLY -- scheduler of the SIMT emulator (see cudasim.h).
#include "cudasim.h"

#include <vector>

// Context switch.  x86-64: a 12-instruction stack switch (callee-saved registers only);
// elsewhere: ucontext.  A launch of a few thousand simulated threads makes tens of millions of
// switches, and swapcontext() pays a signal-mask system call on each.
#if defined(__x86_64__)
extern "C" void cudasim_swap(void** save_sp, void* new_sp);
asm(R"(
.text
.globl cudasim_swap
.type cudasim_swap,@function
cudasim_swap:
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
.size cudasim_swap,.-cudasim_swap
)");
struct Context { void* sp = nullptr; };
static inline void ctx_switch(Context* from, Context* to) { cudasim_swap(&from->sp, to->sp); }
static inline void ctx_make(Context* c, char* stack, size_t bytes, void (*entry)()) {
  uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~uintptr_t(15);
  void** sp = reinterpret_cast<void**>(top);
  *--sp = nullptr;                               // return address slot of `entry` (it never returns)
  *--sp = reinterpret_cast<void*>(entry);        // popped by `ret` in cudasim_swap
  for (int i = 0; i < 6; ++i) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
  c->sp = sp;
}
#else
#include <ucontext.h>
struct Context { ucontext_t uc; };
static inline void ctx_switch(Context* from, Context* to) { swapcontext(&from->uc, &to->uc); }
static inline void ctx_make(Context* c, char* stack, size_t bytes, void (*entry)()) {
  getcontext(&c->uc);
  c->uc.uc_stack.ss_sp = stack;
  c->uc.uc_stack.ss_size = bytes;
  c->uc.uc_link = nullptr;
  makecontext(&c->uc, entry, 0);
}
#endif

namespace cudasim {

ThreadCtx* g_cur = nullptr;
unsigned long long g_launches = 0, g_blocks = 0, g_switches = 0;

namespace {

constexpr size_t kStackBytes = 96 * 1024;
enum State { kRunnable, kWaitBarrier, kWaitCollective, kDone };

struct Fiber {
  Context ctx;
  ThreadCtx tc;
  State state = kRunnable;
  unsigned wait_gen = 0;
  // pending warp collective
  Collective kind = kSyncWarp;
  unsigned mask = 0;
  uint64_t val = 0, result = 0;
  int aux = 0;
  bool result_ready = false;
  char* stack = nullptr;
};

struct Block {
  std::vector<Fiber> fibers;
  std::vector<unsigned> warp_arrived;   // lanes waiting in a collective, per warp
  Context sched;
  int n = 0, live = 0, bar_arrived = 0, cur = 0;
  unsigned bar_gen = 0;
  unsigned long long progress = 0;
  std::vector<unsigned char> smem;
  const std::function<void()>* body = nullptr;
};

Block g_block;
std::vector<char*> g_stacks;

[[noreturn]] void die(const char* what) {
  Block& B = g_block;
  fprintf(stderr, "[cudasim] %s (block %u,%u: %d threads, %d live, %d at __syncthreads)\n", what,
          B.fibers.empty() ? 0 : B.fibers[0].tc.blockIdx_.x, B.fibers.empty() ? 0 : B.fibers[0].tc.blockIdx_.y, B.n, B.live,
          B.bar_arrived);
  for (int i = 0; i < B.n && i < 1024; ++i) {
    const Fiber& f = B.fibers[i];
    if (f.state == kWaitCollective)
      fprintf(stderr, "  thread %d waits in a warp collective kind %d mask %08x\n", i, (int)f.kind, f.mask);
  }
  abort();
}

void release_barrier_if_complete() {
  Block& B = g_block;
  if (B.bar_arrived > 0 && B.bar_arrived == B.live) { B.bar_arrived = 0; ++B.bar_gen; }
}

void yield() {
  Block& B = g_block;
  ++g_switches;
  ctx_switch(&B.fibers[B.cur].ctx, &B.sched);
}

void fiber_main() {
  Block& B = g_block;
  (*B.body)();
  Fiber& f = B.fibers[B.cur];
  f.state = kDone;
  --B.live;
  ++B.progress;
  release_barrier_if_complete();   // exited threads no longer count towards __syncthreads
  ctx_switch(&f.ctx, &B.sched);
  abort();
}

unsigned existing_lanes(int warp) {
  const int first = warp * 32, n = g_block.n - first;
  return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
}

void complete_collective(int warp, unsigned mask) {
  Block& B = g_block;
  Fiber* lanes = &B.fibers[(size_t)warp * 32];
  const int lead = __builtin_ctz(mask);
  unsigned ballot = 0;
  for (int l = 0; l < 32; ++l)
    if (mask & (1u << l)) {
      if (lanes[l].mask != lanes[lead].mask || lanes[l].kind != lanes[lead].kind) die("lanes of one warp met in different collectives");
      if (lanes[l].kind == kBallot && lanes[l].val) ballot |= 1u << l;
    }
  for (int l = 0; l < 32; ++l) {
    if (!(mask & (1u << l))) continue;
    Fiber& f = lanes[l];
    int src = l;
    switch (f.kind) {
      case kShflIdx: src = f.aux; break;
      case kShflXor: src = l ^ f.aux; break;
      case kShflUp: src = l - f.aux; break;
      case kShflDown: src = l + f.aux; break;
      default: break;
    }
    if (src < 0 || src > 31 || !(mask & (1u << src))) src = l;   // out of range / inactive source: own value
    f.result = f.kind == kBallot ? ballot : (f.kind == kSyncWarp ? 0 : lanes[src].val);
  }
  for (int l = 0; l < 32; ++l)
    if (mask & (1u << l)) { lanes[l].result_ready = true; lanes[l].state = kRunnable; }
  B.warp_arrived[warp] &= ~mask;
}

}  // namespace

void* dyn_smem() { return g_block.smem.data(); }

void sync_threads() {
  Block& B = g_block;
  Fiber& f = B.fibers[B.cur];
  const unsigned gen = B.bar_gen;
  ++B.bar_arrived;
  ++B.progress;
  release_barrier_if_complete();
  if (B.bar_gen != gen) return;
  f.state = kWaitBarrier;
  f.wait_gen = gen;
  while (B.bar_gen == gen) yield();
  f.state = kRunnable;
}

uint64_t warp_collective(Collective kind, unsigned mask, uint64_t value, int aux) {
  Block& B = g_block;
  Fiber& f = B.fibers[B.cur];
  const int warp = B.cur >> 5, lane = B.cur & 31;
  mask &= existing_lanes(warp);       // a full mask in a partial last warp names only the lanes that exist
  if (!(mask & (1u << lane))) die("a lane executed a *_sync primitive whose mask does not name it");
  f.kind = kind; f.mask = mask; f.val = value; f.aux = aux; f.result_ready = false;
  B.warp_arrived[warp] |= 1u << lane;
  ++B.progress;
  if ((B.warp_arrived[warp] & mask) == mask) complete_collective(warp, mask);
  if (!f.result_ready) {
    f.state = kWaitCollective;
    while (!f.result_ready) yield();
  }
  f.state = kRunnable;
  return f.result;
}

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body) {
  Block& B = g_block;
  if (B.body != nullptr) { fprintf(stderr, "[cudasim] nested launch\n"); abort(); }
  const int n = (int)(block.x * block.y * block.z);
  if (n <= 0 || n > 1024) { fprintf(stderr, "[cudasim] bad block size %d\n", n); abort(); }
  ++g_launches;
  while ((int)g_stacks.size() < n) g_stacks.push_back(static_cast<char*>(malloc(kStackBytes)));
  B.body = &body;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        ++g_blocks;
        B.n = B.live = n;
        B.bar_arrived = 0; B.bar_gen = 0; B.progress = 0;
        B.fibers.assign((size_t)n, Fiber());
        B.warp_arrived.assign((size_t)(n + 31) / 32, 0u);
        B.smem.assign(dyn_smem_bytes + 16, 0xCD);    // never zeroed on a GPU either
        for (int t = 0; t < n; ++t) {
          Fiber& f = B.fibers[t];
          f.tc.threadIdx_ = {(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
          f.tc.blockIdx_ = {bx, by, bz};
          f.tc.blockDim_ = block;
          f.tc.gridDim_ = grid;
          f.stack = g_stacks[t];
          ctx_make(&f.ctx, f.stack, kStackBytes, fiber_main);
        }
        unsigned long long last_progress = ~0ull;
        while (B.live > 0) {
          if (B.progress == last_progress) die("deadlock: no simulated thread can make progress");
          last_progress = B.progress;
          for (int t = 0; t < n; ++t) {
            Fiber& f = B.fibers[t];
            if (f.state == kDone) continue;
            if (f.state == kWaitBarrier && B.bar_gen == f.wait_gen) continue;
            if (f.state == kWaitCollective && !f.result_ready) continue;
            B.cur = t;
            g_cur = &f.tc;
            ++g_switches;
            ctx_switch(&B.sched, &f.ctx);
          }
        }
      }
  B.body = nullptr;
  g_cur = nullptr;
}

}  // namespace cudasim

extern "C" {
unsigned long long cudasim_launches(void) { return cudasim::g_launches; }
unsigned long long cudasim_blocks(void) { return cudasim::g_blocks; }
unsigned long long cudasim_switches(void) { return cudasim::g_switches; }
}
