// extern "C" boundary (include/iglb200.h) and the query kernels.
#include "../../include/iglb200.h"
#include "mesh.h"
#include "cloud.h"
#include "grid.h"
#include "wn.h"
#include "icp.h"
#include "group.h"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <thread>

using plat::DevBuf;

namespace iglq {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }

#ifdef IGLB200_EMUL
void require_device() {}
struct Timer {
  std::chrono::steady_clock::time_point t0;
  void start(stream_t) { t0 = std::chrono::steady_clock::now(); }
  double stop(stream_t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};
#else
void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) throw plat::CudaError("no CUDA device available (this library has no CPU fallback)");
}
struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  void start(stream_t s) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
  double stop(stream_t s) {
    cudaEventRecord(b, s); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b); cudaEventDestroy(a); cudaEventDestroy(b);
    return ms;
  }
};
#endif

#define API_BEGIN try {
#define API_END                                                                     \
  }                                                                                 \
  catch (const std::invalid_argument &e) { return fail(IGLB200_ERR_ARG, e.what()); } \
  catch (const std::domain_error &e) { return fail(IGLB200_ERR_UNSUPPORTED, e.what()); } \
  catch (const std::bad_alloc &e) { return fail(IGLB200_ERR_OOM, e.what()); }        \
  catch (const std::exception &e) {                                                 \
    const std::string w = e.what();                                                 \
    return fail(w.find("out of memory") != std::string::npos ? IGLB200_ERR_OOM : IGLB200_ERR_CUDA, w); \
  }                                                                                 \
  catch (...) { return fail(IGLB200_ERR_INTERNAL, "unknown error"); }

// ---------------------------------------------------------------- layout shims
template <class T> struct ToCanon3 {  // strided n x 3 of T -> column-major doubles
  const T *src; int64_t rs, cs, n; double *dst;
  DEVFN void operator()(int64_t i) const { for (int j = 0; j < 3; ++j) dst[j * n + i] = (double)src[i * rs + j * cs]; }
};
template <class T> struct FromCanon3 {
  const double *src; int64_t n; T *dst; int64_t rs, cs;
  DEVFN void operator()(int64_t i) const { for (int j = 0; j < 3; ++j) dst[i * rs + j * cs] = (T)src[j * n + i]; }
};
struct ToFloat1 { const double *src; float *dst; DEVFN void operator()(int64_t i) const { dst[i] = (float)src[i]; } };

// ---------------------------------------------------------------- query kernels
struct MortonKeys {  // 30-bit Morton code of each query inside the doubled scene box
  const double *P; int64_t n;
  double lo[3], inv[3];
  uint32_t *keys, *vals;
  int shift;   // 10 - bits per axis
  static DEVFN uint32_t spread(uint32_t v) {
    v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u; v = (v | (v << 2)) & 0x09249249u;
    return v;
  }
  DEVFN void operator()(int64_t i) const {
    uint32_t c[3];
    for (int d = 0; d < 3; ++d) {
      double t = (P[d * n + i] - lo[d]) * inv[d];
      t = t < 0.0 ? 0.0 : (t > 1023.0 ? 1023.0 : t);
      if (!(t == t)) t = 0.0;
      c[d] = (uint32_t)t >> shift;
    }
    keys[i] = spread(c[0]) | (spread(c[1]) << 1) | (spread(c[2]) << 2);
    vals[i] = (uint32_t)i;
  }
};

// What the winding-number kernel of a SHARDED query needs about row t of the sorted order, left by the search kernel in the
// memory of the GPU that ran it: reading it back is local and coalesced, where S[perm[t]] and P[perm[t]] are scattered
// reads of the owner's memory over NVLink (measured on c5, 8 GPUs: the winding kernel took 19 ms on the peers, 11 on the owner)
struct ShardStash { float qx, qy, qz; uint32_t row; double sqrd; };
static_assert(sizeof(ShardStash) == 24, "stash layout");

struct ClosestArgs {  // AABB.cpp:1135-1218 semantics for a batch of queries
  const fwn::PillNode *nodes; const double *tri64; const float4 *triq;
  const double *P; int64_t n; const uint32_t *perm;
  double up, low; float absmax;
  double *sqrD; int32_t *I; double *C;
  uint32_t *counters;  // nullable: [0] node visits, [1] face tests (summed over warps)
  unsigned long long *sched = nullptr;   // work distribution state of the persistent schedule (launch_closest supplies one if null)
  // Sharded operation (several GPUs on ONE query set): chunks are not numbered from this launch's own counter but come
  // in SUPER-chunks of `super_chunks` consecutive chunks taken off a counter in the memory of the device that owns the
  // query set (a system-scope atomic over NVLink); the super-chunks this device took are listed for the kernels that follow.
  // The super-chunks are cut into `nshards` contiguous SLICES of the Morton order, one counter each (remote[2 + slice]): a
  // device works through its own slice first — one region of space, so it pulls one part of the tree through its L2
  // instead of all of it — and then helps with whichever slice has the most left.
  unsigned long long *remote = nullptr;
  uint32_t super_chunks = 0;
  uint32_t nsup = 0;            // super-chunks in the query set
  int shard = 0, nshards = 1;   // this device's slice, number of slices
  // Warps that run out of packets stay and take over parts of the walks still going on (work donation).  Off when
  // several launches share the device on purpose (the host path: the next chunk's kernel fills this one's tail).
  int donate = 1;
  // packets per chunk of the persistent schedule (a power of two, <= CP_CHUNK; launch_closest sizes it so that a small
  // launch still has ~16 chunks per SM to even out; sharded launches use CP_CHUNK: super-chunks are made of those)
  uint32_t chunk = 32;
  uint32_t *my_supers = nullptr, *my_count = nullptr;
  int64_t max_supers = 0;   // > 0: take at most this many (static partition / tests)
  const uint32_t *super_order = nullptr;   // the k-th super-chunk handed out is super_order[k]
  ShardStash *stash = nullptr;             // [n], indexed by sorted position
  uint32_t nnodes = 0;
  // the fp32 stage works in coordinates scaled by a power of two so that the mesh's max |coordinate| lies in [0.5, 1):
  // exact, and squares can neither overflow nor go denormal whatever the units of the mesh (absmax above is scaled too)
  double sigma = 1.0;
};

// Work donation (the tail of a launch): a warp that walks an unusually heavy packet while other warps have run out of
// work hands them the older half of its stack as a task (packet, sub-trees, the lanes' current upper bounds); the
// results of the pieces of a packet are merged under a lock (smaller squared distance, lower face index on a tie).
// Every warp of the launch owns a mailbox (a state word and a task record): no shared queue, hence no shared lock.
// A packet walked in pieces owns a SLOT in the launching device's memory where the pieces merge (squared distance, face)
// per query; the piece that finishes last writes the packet's results — so a sharded launch does its merging locally and
// touches the owner's arrays once, as for any other packet.
static constexpr int CP_TASK_NODES = 13, CP_MAILBOXES = 4096, CP_SLOTS = 8192;
struct CpTask { uint32_t packet, nn, slot; uint32_t node[CP_TASK_NODES]; float ub[32]; };   // 192 bytes
struct CpSlot { double d2[32]; int32_t face[32]; uint32_t pending, lock; uint32_t pad[30]; };   // 512 bytes
static_assert(sizeof(CpTask) == 192 && sizeof(CpSlot) == 512, "record sizes");
static constexpr uint32_t BOX_CLOSED = 0u, BOX_WAITING = 1u, BOX_CLAIMED = 2u, BOX_FILLED = 3u;
// state words of a launch: next chunk; per SM id: state, refill lock; sharded: lock, block counter, current slice;
// donation: idle warps, warps holding work, slots handed out, spare; mailbox states (32 bits each); task records; slots
static constexpr int64_t SCHED_DONATE = 1 + 2 * 1024 + 3, SCHED_BOX_STATE = SCHED_DONATE + 4;
static constexpr int64_t SCHED_TASKS = SCHED_BOX_STATE + CP_MAILBOXES / 2;
static constexpr int64_t SCHED_INIT_WORDS = SCHED_TASKS;
static constexpr int64_t SCHED_SLOTS = SCHED_TASKS + CP_MAILBOXES * (int64_t)(sizeof(CpTask) / 8);
static constexpr int64_t SCHED_WORDS = SCHED_SLOTS + CP_SLOTS * (int64_t)(sizeof(CpSlot) / 8);
struct InitSched {   // every SM starts with an exhausted chunk (and, sharded, the device with an exhausted super-chunk)
  unsigned long long *w; int sharded;
  DEVFN void operator()(int64_t i) const {
    if (i == 0) w[i] = sharded ? 0x7F7F7F7Full : 0ull;
    else if (i >= 1 + 2 * 1024) w[i] = 0ull;
    else w[i] = (i & 1) ? 0x7F7F7F7Full : 0ull;
  }
};

DEVFN void closest_store(const ClosestArgs &a, int64_t i, const fwn::CpResult &r) {
  a.sqrD[i] = r.sqr_d; a.I[i] = r.face;
  a.C[i] = r.c.x; a.C[a.n + i] = r.c.y; a.C[2 * a.n + i] = r.c.z;
}

#ifdef IGLB200_EMUL
// Host emulation of one warp: identical control flow, lanes looped.
void closest_packet_emul(const ClosestArgs &a, int64_t base, int32_t &seed_face) {
  using namespace fwn;
  static thread_local uint32_t stack[CP_STACK];
  CpLane L[32]; CpExact X[32]; CpQueryRef qr[32]; bool valid[32]; int64_t qi[32];
  for (int l = 0; l < 32; ++l) {
    const int64_t t = base + l;
    valid[l] = t < a.n;
    qi[l] = valid[l] ? (a.perm ? (int64_t)a.perm[t] : t) : 0;
    qr[l] = CpQueryRef{a.P, a.n, qi[l], a.sigma};
    lane_init(L[l], X[l], valid[l] ? load_query(qr[l]) : D3{0.0, 0.0, 0.0}, a.up, a.absmax, a.sigma);
    if (!valid[l] || a.low > a.up) { L[l].best_ub = -INFINITY; L[l].thr2 = -1.f; }
    else if (seed_face >= 0) {
      const float4 *tq = a.triq + TRIQ_STRIDE * (size_t)seed_face;
      float err; const float d32 = tri_dist_f32(tq[0], tq[1], tq[2], tq[3], L[l].px, L[l].py, L[l].pz, L[l].u, err);
      const float ub = d32 + err;
      if (ub < L[l].best_ub) lane_set_ub(L[l], ub);
    }
  }
  uint32_t nvis = 0, ntri = 0;
  int sp = 0; uint32_t cur = 0; bool have = true;
  while (true) {
    if (!have) {
      if (sp == 0) break;
      --sp;
      cur = stack[sp];
    }
    have = false; ++nvis;
    const PillNode n = a.nodes[cur];
    const uint32_t ch[4] = {n.child.x, n.child.y, n.child.z, n.child.w};
    const uint32_t imask = n.pad.x, tmask = n.pad.y;
    float d2[32][4];
    for (int l = 0; l < 32; ++l) for (int i = 0; i < 4; ++i) d2[l][i] = lane_pill_d2(n, i, L[l], n.pad.z);
    // faces first: they tighten the bounds the internal children are judged by
    bool want[4] = {false, false, false, false};   // voted before any face of the node is tested, as the kernel does
    for (int i = 0; i < 4; ++i)
      if ((tmask >> i) & 1u) for (int l = 0; l < 32; ++l) want[i] = want[i] || (d2[l][i] <= L[l].thr2);
    for (int i = 0; i < 4; ++i) {
      if (!want[i]) continue;
      ++ntri;
      const float4 *tq = a.triq + TRIQ_STRIDE * (size_t)ch[i];
      for (int l = 0; l < 32; ++l) {
        float err; const float d32 = tri_dist_f32(tq[0], tq[1], tq[2], tq[3], L[l].px, L[l].py, L[l].pz, L[l].u, err);
        lane_offer(L[l], X[l], a.tri64, qr[l], ch[i], d32, err);
      }
    }
    // internal children some lane still needs; the one with the smallest lower bound is walked next,
    // the others are stacked in lane order
    uint32_t wi = 0, best = 0xFFFFFFFFu;
    for (int l = 0; l < 32; ++l) for (int i = 0; i < 4; ++i) {
      if (!((imask >> i) & 1u) || !(d2[l][i] <= L[l].thr2)) continue;
      wi |= 1u << i;
      best = std::min(best, (plat::f2u(d2[l][i]) & ~3u) | (uint32_t)i);
    }
    if (wi) {
      const int nb = (int)(best & 3u);
      const uint32_t push = wi & ~(1u << nb);
      const int npush = __builtin_popcount(push);
      if (sp + npush > CP_STACK) throw std::runtime_error("closest-point stack overflow");
      for (int j = 0; j < 4; ++j) {
        if (!((push >> j) & 1u)) continue;
        stack[sp + __builtin_popcount(push & ((1u << j) - 1u))] = ch[j] & ~INTERNAL_BIT;
      }
      sp += npush;
      cur = ch[nb] & ~INTERNAL_BIT; have = true;
    }
  }
  for (int l = 0; l < 32; ++l) {
    if (!valid[l]) continue;
    CpResult r;
    if (a.low > a.up) { r.sqr_d = a.low; r.face = -1; r.c = {0.0, 0.0, 0.0}; }  // AABB.cpp:1145-1148
    else lane_finish(L[l], X[l], a.tri64, qr[l], a.up, r);
    closest_store(a, qi[l], r);
    if (l == 0) seed_face = r.face;
    if (a.stash) {
      const D3 p = load_query(qr[l]);
      a.stash[base + l] = ShardStash{(float)p.x, (float)p.y, (float)p.z, (uint32_t)qi[l], r.sqr_d};
    }
  }
  if (a.counters) { a.counters[0] += nvis; a.counters[1] += ntri; }
}
static constexpr int CP_CHUNK = 32;
static constexpr int CP_GROUP = 2;
void launch_closest(const ClosestArgs &a, stream_t) {
  int32_t seed = -1;   // groups of CP_GROUP packets: the first of a group starts unseeded
  if (!a.remote) { for (int64_t base = 0; base < a.n; base += 32) { if ((base / 32) % CP_GROUP == 0) seed = -1; closest_packet_emul(a, base, seed); } return; }
  // sharded: super-chunks off the owner's counter, listed for the kernels that follow
  const int64_t super_rows = (int64_t)a.super_chunks * CP_CHUNK * 32;
  // its own slice of the super-chunks first, then whichever slice has the most left (as next_chunk_sharded does)
  auto slice_lo = [&](int k) { return (uint32_t)((unsigned long long)a.nsup * (unsigned)k / (unsigned)a.nshards); };
  int cur = a.shard;
  while (a.max_supers <= 0 || (int64_t)*a.my_count < a.max_supers) {
    const unsigned long long k = __atomic_fetch_add(a.remote + 2 + cur, 1ull, __ATOMIC_RELAXED);
    if (slice_lo(cur) + k >= slice_lo(cur + 1)) {
      unsigned long long most = 0;
      for (int j = 0; j < a.nshards; ++j) {
        const unsigned long long size = slice_lo(j + 1) - slice_lo(j), taken = __atomic_load_n(a.remote + 2 + j, __ATOMIC_RELAXED);
        if (size > taken && size - taken > most) { most = size - taken; cur = j; }
      }
      if (most == 0) break;
      continue;
    }
    unsigned long long sc = slice_lo(cur) + k;
    if (a.super_order) sc = a.super_order[sc];
    a.my_supers[(*a.my_count)++] = (uint32_t)sc;
    for (int64_t base = (int64_t)sc * super_rows; base < std::min(a.n, ((int64_t)sc + 1) * super_rows); base += 32) {
      if ((base / 32) % CP_GROUP == 0) seed = -1;
      closest_packet_emul(a, base, seed);
    }
  }
}
uint32_t *overflow_flag() { static thread_local uint32_t f = 0; return &f; }
void reset_overflow(stream_t) { *overflow_flag() = 0; }
void throw_if_overflow(stream_t) {
  if (*overflow_flag()) { *overflow_flag() = 0; throw std::runtime_error("traversal stack overflow (tree too deep)"); }
}
#else
// 256-bit global load (sm_100: LDG.E.256): half the load instructions for the 128-byte node and the 64-byte face record
DEVFN void ld256(const void *p, float4 &a, float4 &b) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

#ifndef CP_MIN_BLOCKS
#define CP_MIN_BLOCKS 20   // one-warp blocks at 96 registers: 20 warps per SM (measured optimum, profiles/README_r1.md)
#endif
// One warp = one packet of 32 Morton-adjacent queries sharing a stack (shared memory).
// Per node: 4 box tests per lane, two warp reductions (which children does anybody still
// need; which of them is nearest), faces tested by all lanes at once.
// A warp takes CP_GROUP consecutive packets at a time and carries lane 0's winning face from one to the next: the distance to
// ANY face is an upper bound, and the winner of the Morton-adjacent packet is a close one — a search otherwise spends about
// half of its node visits before its bounds are tight (emulation, c2: -18 % node visits, -14 % face tests when every packet
// is seeded by its predecessor).  The last chunks of a launch are handed out packet by packet (short tail).
// Measured (search kernel, ms; c2 / c3 / c5): no groups 40.9 / 21.8 / 688; chunk 16, groups of 2: 39.4 / 22.0 / 631;
// **chunk 32, groups of 2: 38.6 / 21.9 / 624**; chunk 32 / 4: 38.9 / 24.5 / 615; chunk 64 / 4: 38.5 / 24.6 / 613;
// chunk 64 / 8: 38.7 / 27.1 / 628 (random queries — c3 — dislike long groups).
#ifndef CP_GROUP
#define CP_GROUP 2
#endif
#ifndef CP_ENDZONE_PACKETS
#define CP_ENDZONE_PACKETS (148 * 256)   // at most; and at most half of the launch.  Too short an end zone and the launch ends on whole groups
#endif
#ifndef CP_SKIP_ABSENT
#define CP_SKIP_ABSENT 1   // 41.0 -> 40.8 ms on c2 (with boxes it was +2 %: a cylinder test is dearer than the branch)
#endif
#ifndef CP_TOP
#define CP_TOP 0           // > 0: stage that many top nodes in shared memory (needs CP_WARPS > 1 to pay for the copy; measured: see profiles/)
#endif
#ifndef CP_WARPS
#define CP_WARPS 1         // a block is one packet: a finished packet frees its slot at once (4 warps/block: 94.6 ms)
#endif
#ifndef CP_CHUNK
#define CP_CHUNK 32   // round 1 (boxes, no groups): 1 -> 85.9 ms, 4 -> 83.1, 8 -> 82.3, 16 -> 81.5, 32 -> 82.5, 128 -> 86.2; with groups see CP_GROUP
#endif
// Sharded schedule, one level above the per-SM chunks: sched[0] = (super-chunk << 32 | chunks of it handed out).  When a
// super-chunk is used up, one thread of the device takes the next one off the owner's counter (atomicAdd over NVLink when
// the owner is a peer) and notes it in my_supers.  Returns a chunk index; chunks past the last packet end the caller's warp.
// The chunk index carries CP_SINGLE when the device is close to the end of its work (packets are then handed out one by
// one instead of in groups, so that the launch does not end on a few warps with a group each).
struct ShardSched {
  unsigned long long *sched, *remote; uint32_t super_chunks; uint32_t *my_supers, *my_count; int64_t max_supers; const uint32_t *order;
  uint32_t nsup; int shard, nshards; uint32_t endzone;   // endzone: packets per device handed out singly at the end
  DEVFN uint32_t slice_lo(int k) const { return (uint32_t)((unsigned long long)nsup * (unsigned)k / (unsigned)nshards); }
};
static constexpr unsigned long long CP_SINGLE = 0x80000000ull, CP_PAST_END = 0x0FFFFFFFull;
// super-chunks not yet handed out: in slice `best` (the one with the most left), and in all of them
static __device__ unsigned long long shard_left(const ShardSched &a, int &best) {
  unsigned long long total = 0, most = 0;
  for (int j = 0; j < a.nshards; ++j) {
    const unsigned long long size = a.slice_lo(j + 1) - a.slice_lo(j), taken = *(volatile unsigned long long *)(a.remote + 2 + j);
    const unsigned long long left = size > taken ? size - taken : 0ull;
    total += left;
    if (left > most) { most = left; best = j; }
  }
  return total;
}
static __device__ __noinline__ unsigned long long next_chunk_sharded(const ShardSched a) {
  unsigned long long *G = a.sched, *glock = a.sched + 1 + 2 * 1024, *curw = glock + 2;
  while (true) {
    const unsigned long long old = atomicAdd(G, 1ull);
    if ((uint32_t)old < a.super_chunks) {
      const unsigned long long hi = old >> 32, sc = hi & 0x7FFFFFFFull;
      const unsigned long long chunk = sc * a.super_chunks + (uint32_t)old;
      return (sc == 0x7FFFFFFFull || chunk > CP_PAST_END) ? CP_PAST_END : (chunk | (hi & CP_SINGLE));   // sentinel: its packets lie past the end
    }
    if (atomicCAS(glock, 0ull, 1ull) == 0ull) {
      if ((uint32_t)(*(volatile unsigned long long *)G) >= a.super_chunks) {
        unsigned long long word = 0x7FFFFFFFull;   // "past the end"
        const uint32_t have = *(volatile uint32_t *)a.my_count;
        if (a.max_supers <= 0 || (int64_t)have < a.max_supers) {
          const unsigned long long super_packets = (unsigned long long)a.super_chunks * CP_CHUNK;
          int cur = *curw ? (int)*curw - 1 : a.shard;
          while (true) {
            const uint32_t lo = a.slice_lo(cur), hi = a.slice_lo(cur + 1);
            const unsigned long long k = atomicAdd_system(a.remote + 2 + cur, 1ull);
            if (lo + k < hi) {
              uint32_t sc = lo + (uint32_t)k;
              bool single = false;
              if (cur != a.shard || (hi - sc - 1) * super_packets <= a.endzone) {   // not deep inside its own slice
                int best = cur;
                single = shard_left(a, best) * super_packets <= (unsigned long long)a.endzone * (unsigned)a.nshards;
              }
              if (a.order) sc = a.order[sc];
              a.my_supers[have] = sc; *(volatile uint32_t *)a.my_count = have + 1;
              word = sc | (single ? CP_SINGLE : 0ull);
              break;
            }
            int best = -1;   // this slice is used up: help with the one that has the most left
            if (shard_left(a, best) == 0ull) break;
            cur = best;
          }
          *curw = (unsigned long long)cur + 1;
        }
        __threadfence();
        atomicExch(G, word << 32);
      }
      __threadfence();
      atomicExch(glock, 0ull);
    }
  }
}
#ifndef CP_DONATE_EVERY
#define CP_DONATE_EVERY 32   // pops between two looks at the idle counter
#endif
// donor: hand the older half of the stack (at most CP_TASK_NODES sub-trees) to a warp that waits at its mailbox (four
// mailboxes are tried at random: in the tail of a launch most warps wait); returns the new height of the stack
static __device__ __noinline__ int cp_donate(unsigned long long *sched, uint32_t *stack, int sp, uint32_t packet, float best_ub, uint32_t salt,
                                             int32_t slot) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  uint32_t *state = (uint32_t *)(sched + SCHED_BOX_STATE);
  const uint32_t h = ((blockIdx.x * 2654435761u + salt * 40503u + (uint32_t)lane * 0x9E3779B9u) >> 7) % gridDim.x;
  bool claimed = false;
  if (lane < 4 && h != blockIdx.x) claimed = atomicCAS(&state[h], BOX_WAITING, BOX_CLAIMED) == BOX_WAITING;
  const unsigned won = __ballot_sync(FULL, claimed);
  if (!won) return sp;
  const int w = __ffs(won) - 1;
  if (claimed && lane != w) atomicExch(&state[h], BOX_WAITING);   // one is enough
  const uint32_t box = __shfl_sync(FULL, h, w);
  const int nd = min(sp / 2, CP_TASK_NODES);
  CpTask *tk = (CpTask *)(sched + SCHED_TASKS) + box;
  tk->ub[lane] = best_ub;
  if (lane < nd) tk->node[lane] = stack[lane];
  if (lane == 0) {
    tk->packet = packet; tk->nn = (uint32_t)nd; tk->slot = (uint32_t)slot;
    atomicAdd(&((CpSlot *)(sched + SCHED_SLOTS) + slot)->pending, 1u);   // one more piece, before anybody can finish it
  }
  __syncwarp();
  for (int k0 = 0; k0 < sp - nd; k0 += 32) {   // close the gap
    const int k = k0 + lane;
    uint32_t v = 0;
    if (k < sp - nd) v = stack[k + nd];
    __syncwarp();
    if (k < sp - nd) stack[k] = v;
    __syncwarp();
  }
  __threadfence();
  __syncwarp();
  if (lane == 0) atomicExch(&state[box], BOX_FILLED);
  return sp - nd;
}
// a warp out of packets waits at its mailbox for a task: false when nobody is left who could send one
static __device__ __noinline__ bool cp_take_task(unsigned long long *sched, uint32_t *stack, uint32_t &packet, int &sp, float &ub, bool &counted,
                                                 int32_t &slot) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  uint32_t *idle = (uint32_t *)(sched + SCHED_DONATE), *active = (uint32_t *)(sched + SCHED_DONATE + 1);
  uint32_t *state = (uint32_t *)(sched + SCHED_BOX_STATE) + blockIdx.x;
  int got = 0;
  if (lane == 0) {
    if (counted) atomicSub(active, 1u);
    atomicAdd(idle, 1u);
    atomicExch(state, BOX_WAITING);
    unsigned ns = 250, polls = 0;
    while (true) {   // (capping the waiting warps at a quarter of the launch was measured: slices 5.66 instead of 5.41 ms)
      const uint32_t st = *(volatile uint32_t *)state;
      if (st == BOX_FILLED) { got = 1; break; }
      // nobody holds work any more: withdraw — unless a donor has just claimed this mailbox, then its task is on the way
      if (st == BOX_WAITING && (++polls & 3u) == 0u && *(volatile uint32_t *)active == 0u &&
          atomicCAS(state, BOX_WAITING, BOX_CLOSED) == BOX_WAITING) break;
      __nanosleep(ns);
      if (ns < 4000u) ns *= 2u;
    }
    atomicSub(idle, 1u);
    if (got) atomicAdd(active, 1u);
  }
  got = __shfl_sync(FULL, got, 0);
  if (!got) { counted = false; return false; }
  counted = true;
  const CpTask *tk = (const CpTask *)(sched + SCHED_TASKS) + blockIdx.x;
  ub = ((const volatile float *)tk->ub)[lane];
  const uint32_t nn = *(const volatile uint32_t *)&tk->nn;
  packet = *(const volatile uint32_t *)&tk->packet;
  slot = (int32_t)*(const volatile uint32_t *)&tk->slot;
  if ((uint32_t)lane < nn) stack[lane] = ((const volatile uint32_t *)tk->node)[lane];
  __syncwarp();
  if (lane == 0) atomicExch(state, BOX_CLOSED);
  sp = (int)nn;
  return true;
}
// a packet about to be walked in pieces takes a slot ("nothing found yet", one piece: the caller); -1 when none is left
static __device__ __noinline__ int32_t cp_begin_shared(unsigned long long *sched, double up) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  uint32_t k = 0;
  if (lane == 0) k = atomicAdd((uint32_t *)(sched + SCHED_DONATE + 2), 1u);
  k = __shfl_sync(FULL, k, 0);
  if (k >= (uint32_t)CP_SLOTS) return -1;
  CpSlot *sl = (CpSlot *)(sched + SCHED_SLOTS) + k;
  sl->d2[lane] = up; sl->face[lane] = -1;
  if (lane == 0) { sl->pending = 1u; sl->lock = 0u; }
  __threadfence();
  __syncwarp();
  return (int32_t)k;
}
// a piece of a packet is done: merge its result into the slot (smaller squared distance; lower face index on a tie).  The
// piece that finishes last writes the packet's results (and its row of the stash) like any other packet's.
static __device__ __noinline__ void cp_finish_shared(const ClosestArgs a, int32_t slot, int64_t qi, int64_t t, bool valid, double sqr_d, int32_t face) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  CpSlot *sl = (CpSlot *)(a.sched + SCHED_SLOTS) + slot;
  if (lane == 0) while (atomicCAS(&sl->lock, 0u, 1u) != 0u) __nanosleep(100);
  __syncwarp();
  double d2 = *(volatile double *)&sl->d2[lane];
  int32_t f = *(volatile int32_t *)&sl->face[lane];
  if (valid && face >= 0 && (f < 0 || sqr_d < d2 || (sqr_d == d2 && face < f))) {
    d2 = sqr_d; f = face;
    *(volatile double *)&sl->d2[lane] = d2; *(volatile int32_t *)&sl->face[lane] = f;
  }
  __threadfence();
  __syncwarp();
  uint32_t left = 0;
  if (lane == 0) { atomicExch(&sl->lock, 0u); left = atomicSub(&sl->pending, 1u) - 1u; }
  left = __shfl_sync(FULL, left, 0);
  if (left != 0u || !valid) return;
  // every piece has merged (each one before it counted itself out): read the final pair
  __threadfence();
  d2 = *(volatile double *)&sl->d2[lane];
  f = *(volatile int32_t *)&sl->face[lane];
  fwn::CpResult r; r.sqr_d = d2; r.face = f; r.c = {0.0, 0.0, 0.0};
  const fwn::CpQueryRef qr{a.P, a.n, qi, a.sigma};
  const fwn::D3 p = fwn::load_query(qr);
  if (!(r.sqr_d < a.up) || r.face < 0) { r.sqr_d = a.up; r.face = -1; }
  else fwn::face_sqr_distance(a.tri64, p, (uint32_t)r.face, r.c);   // closest point of the winner
  closest_store(a, qi, r);
  if (a.stash) a.stash[t] = ShardStash{(float)p.x, (float)p.y, (float)p.z, (uint32_t)qi, r.sqr_d};
}
#ifdef CP_TIMELINE
__device__ unsigned long long *g_timeline = nullptr;   // [8 * blocks]: first request, last request that got work, exit (globaltimer, ns); visits and duration of the longest walk, donations, tasks
#endif
template <bool COUNT>
__global__ void __launch_bounds__(32 * CP_WARPS, CP_MIN_BLOCKS) closest_packet_kernel(ClosestArgs a, uint32_t *overflow) {
  using namespace fwn;
  constexpr unsigned FULL = 0xFFFFFFFFu;
  __shared__ uint32_t stacks[CP_WARPS][CP_STACK];
  __shared__ CpExact exact[32 * CP_WARPS];
  CpExact &X = exact[threadIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t *stack = stacks[warp];
#if CP_TOP > 0
  // experiment (north_star subsystem 2): the top CP_TOP nodes of the tree (breadth-first: levels are contiguous) staged in
  // shared memory, one copy per block of CP_WARPS packets
  __shared__ __align__(32) PillNode top[CP_TOP];
  const uint32_t ntop = min((uint32_t)CP_TOP, a.nnodes);
  for (uint32_t k = threadIdx.x; k < ntop * (sizeof(PillNode) / 16); k += blockDim.x) ((uint4 *)top)[k] = ((const uint4 *)a.nodes)[k];
  __syncthreads();
#endif
  // Persistent schedule: every SM hands out the packets of one chunk of CP_CHUNK Morton-consecutive packets to its
  // resident warps, so the packets that share an L1 are neighbours; chunks are taken from a global counter, so the
  // chunks in flight still form one window of the query set (L2).  sched[0] = next chunk, sched[1 + 2*smid] =
  // (chunk << 32 | packets handed out), sched[2 + 2*smid] = refill lock.
  unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
  smid &= 1023u;   // the state array has 1024 slots
  const uint32_t npackets = (uint32_t)((a.n + 31) / 32);
  const uint32_t chunk = a.chunk;
  const uint32_t nchunks = (npackets + chunk - 1) / chunk;
  const uint32_t endzone_packets = min((uint32_t)CP_ENDZONE_PACKETS, npackets / (2u * (uint32_t)a.nshards));
  const uint32_t endzone = (endzone_packets + chunk - 1) / chunk;   // chunks handed out packet by packet
  int32_t seed_face = -1;
  bool drained = false, counted = false;   // no packets left for this warp; it is counted among the warps holding work
  while (true) {
    uint32_t first = 0xFFFFFFFFu, cnt = 1;
    int task_sp = 0; float task_ub = 0.f; int32_t slot = -1;   // slot >= 0: the packet is walked in pieces
    if (!drained) {
    if (lane == 0) {
      unsigned long long *st = a.sched + 1 + 2 * smid, *lock = st + 1;
      while (true) {
        // groups of CP_GROUP packets, single packets in the end zone (decided on a peek: a stale one only mis-sizes one group)
        const uint32_t peek = (uint32_t)(*(volatile unsigned long long *)st >> 32);
        const uint32_t g = (peek & (uint32_t)CP_SINGLE) ? 1u : (uint32_t)CP_GROUP;
        const unsigned long long old = atomicAdd(st, (unsigned long long)g);
        if ((uint32_t)old < chunk) {
          first = ((uint32_t)(old >> 32) & ~(uint32_t)CP_SINGLE) * chunk + (uint32_t)old;
          cnt = min(g, chunk - (uint32_t)old);
          break;
        }
        if (atomicCAS(lock, 0ull, 1ull) == 0ull) {   // this warp refills the SM's chunk
          if ((uint32_t)(*(volatile unsigned long long *)st) >= chunk) {
            unsigned long long nc;
            if (a.remote) nc = next_chunk_sharded(ShardSched{a.sched, a.remote, a.super_chunks, a.my_supers, a.my_count, a.max_supers, a.super_order,
                                                             a.nsup, a.shard, a.nshards, endzone_packets});
            else { nc = atomicAdd(a.sched, 1ull); if (nc + endzone >= nchunks) nc |= CP_SINGLE; }
            atomicExch(st, nc << 32);
          }
          __threadfence();
          atomicExch(lock, 0ull);
        }
      }
    }
    first = __shfl_sync(FULL, first, 0);
    cnt = __shfl_sync(FULL, cnt, 0);
    if (first >= npackets) drained = true;   // the SM's chunk lies past the last packet
    else if (!counted) { counted = true; if (lane == 0) atomicAdd((uint32_t *)(a.sched + SCHED_DONATE + 1), 1u); }
    }
#ifdef CP_TIMELINE
    if (lane == 0 && g_timeline) {   // tuning aid: when each warp took its first / last piece of work and when it left
      unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (g_timeline[8 * blockIdx.x] == 0ull) g_timeline[8 * blockIdx.x] = now;
      if (first < npackets) g_timeline[8 * blockIdx.x + 1] = now; else g_timeline[8 * blockIdx.x + 2] = now;
    }
#endif
    // out of packets: help with the heavy packets still being walked, until none is left
    const bool is_task = drained;
    if (drained) {
      if (!a.donate) break;
      if (!cp_take_task(a.sched, stack, first, task_sp, task_ub, counted, slot)) break;
      cnt = 1;
    }
    if (cnt == 1) seed_face = -1;   // a single packet has no predecessor of its own
   for (uint32_t gi = 0; gi < cnt; ++gi) {
    const uint32_t packet = first + gi;
    if (packet >= npackets) break;
    const int64_t base = (int64_t)packet * 32;
    const int64_t t = base + lane;
    const bool valid = t < a.n;
    const int64_t qi = valid ? (a.perm ? (int64_t)a.perm[t] : t) : 0;
    CpLane L;
    const CpQueryRef qr{a.P, a.n, qi, a.sigma};
    lane_init(L, X, valid ? load_query(qr) : D3{0.0, 0.0, 0.0}, a.up, a.absmax, a.sigma);
    if (!valid || a.low > a.up) { L.best_ub = -INFINITY; L.thr2 = -1.f; }
    else if (is_task) { if (task_ub < L.best_ub) lane_set_ub(L, task_ub); }
    else if (seed_face >= 0) {   // warp-uniform
      const float4 *tq = a.triq + TRIQ_STRIDE * (size_t)seed_face;
      float4 T0, T1, T2, T3; ld256(tq, T0, T1); ld256(tq + 2, T2, T3);
      float err; const float d32 = tri_dist_f32(T0, T1, T2, T3, L.px, L.py, L.pz, L.u, err);
      const float ub = d32 + err;
      if (ub < L.best_ub) lane_set_ub(L, ub);
    }
    uint32_t nvis = 0, ntri = 0, npop = 0;
    int sp = 0; uint32_t cur = 0; bool have = true;
#ifdef CP_TIMELINE
    unsigned long long tl_walk0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl_walk0));
#endif
    if (is_task) { sp = task_sp - 1; cur = stack[sp]; }
    while (true) {
      if (!have) {
        if (sp == 0) break;
        --sp;
        // no re-test against the tightened bounds here: for a packet (smallest bound any lane had at the push against the
        // largest threshold any lane still has) it rejects next to nothing; the node's own tests do it
        cur = stack[sp];
        // a long walk (counted in pops: off the descending path): is anybody out of work?
        if ((++npop & (uint32_t)(CP_DONATE_EVERY - 1)) == 0u && sp >= 2 && a.donate) {
          uint32_t idle = lane == 0 ? *(volatile uint32_t *)(a.sched + SCHED_DONATE) : 0u;
          idle = __shfl_sync(FULL, idle, 0);
          if (idle) {
            if (slot < 0) slot = cp_begin_shared(a.sched, a.up);
            if (slot >= 0) sp = cp_donate(a.sched, stack, sp, packet, L.best_ub, npop, slot);
          }
        }
      }
      have = false;
      if (COUNT) ++nvis;
      const PillNode *np = a.nodes + cur;
      PillNode n;
#if CP_TOP > 0
      if (cur < ntop) n = top[cur];
      else
#endif
      { float4 c4, p4;
        ld256(&np->cx, n.cx, n.cy); ld256(&np->cz, n.cz, n.nx); ld256(&np->ny, n.ny, n.nz); ld256(&np->R, n.R, n.t); ld256(&np->child, c4, p4);
        n.child = make_uint4(__float_as_uint(c4.x), __float_as_uint(c4.y), __float_as_uint(c4.z), __float_as_uint(c4.w));
        n.pad = make_uint4(__float_as_uint(p4.x), __float_as_uint(p4.y), __float_as_uint(p4.z), 0u); }
      const uint32_t ch[4] = {n.child.x, n.child.y, n.child.z, n.child.w};
      const uint32_t imask = n.pad.x, tmask = n.pad.y;
      float d2[4];
      if (n.pad.z == 0u) {   // warp-uniform: the common node has cylinders only
#if CP_SKIP_ABSENT
        // children fill the lanes from 0 up: a node with two or three children skips the tests of its absent lanes
        const uint32_t present = imask | tmask;
        d2[0] = lane_pill_d2(n, 0, L, 0u); d2[1] = lane_pill_d2(n, 1, L, 0u);
        d2[2] = INFINITY; d2[3] = INFINITY;
        if (present & 4u) {
          d2[2] = lane_pill_d2(n, 2, L, 0u);
          if (present & 8u) d2[3] = lane_pill_d2(n, 3, L, 0u);
        }
#else
#pragma unroll
        for (int i = 0; i < 4; ++i) d2[i] = lane_pill_d2(n, i, L, 0u);
#endif
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) d2[i] = lane_pill_d2(n, i, L, n.pad.z);
      }
      // faces first: they tighten the bounds the internal children are judged by
      if (tmask) {   // warp-uniform
        // two face records in flight where two neighbouring lanes are wanted: one load round trip instead of two
        bool w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = ((tmask >> i) & 1u) && __any_sync(FULL, d2[i] <= L.thr2);
#define CP_LOADQ(p, A, B, C, D) { ld256((p), A, B); ld256((p) + 2, C, D); }
#define CP_FACE1(i) { const float4 *tq = a.triq + TRIQ_STRIDE * (size_t)ch[i]; \
          float4 T0, T1, T2, T3; CP_LOADQ(tq, T0, T1, T2, T3); if (COUNT) ++ntri; \
          float err; const float d32 = tri_dist_f32(T0, T1, T2, T3, L.px, L.py, L.pz, L.u, err); \
          if (!(d32 - err > L.best_ub)) lane_offer(L, X, a.tri64, qr, ch[i], d32, err); }
#define CP_FACE2(i, j) { const float4 *tq = a.triq + TRIQ_STRIDE * (size_t)ch[i], *tr = a.triq + TRIQ_STRIDE * (size_t)ch[j]; \
          float4 T0, T1, T2, T3, U0, U1, U2, U3; CP_LOADQ(tq, T0, T1, T2, T3); CP_LOADQ(tr, U0, U1, U2, U3); if (COUNT) ntri += 2; \
          float err, erq; const float d32 = tri_dist_f32(T0, T1, T2, T3, L.px, L.py, L.pz, L.u, err); \
          const float e32 = tri_dist_f32(U0, U1, U2, U3, L.px, L.py, L.pz, L.u, erq); \
          if (!(d32 - err > L.best_ub)) lane_offer(L, X, a.tri64, qr, ch[i], d32, err); \
          if (!(e32 - erq > L.best_ub)) lane_offer(L, X, a.tri64, qr, ch[j], e32, erq); }
        if (w[0] && w[1]) CP_FACE2(0, 1) else if (w[0]) CP_FACE1(0) else if (w[1]) CP_FACE1(1)
        if (w[2] && w[3]) CP_FACE2(2, 3) else if (w[2]) CP_FACE1(2) else if (w[3]) CP_FACE1(3)
#undef CP_FACE1
#undef CP_FACE2
      }
      if (imask) {   // warp-uniform
        uint32_t mine = 0, mykey = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool w = ((imask >> i) & 1u) && (d2[i] <= L.thr2);
          mine |= w ? (1u << i) : 0u;
          mykey = w ? min(mykey, (__float_as_uint(d2[i]) & ~3u) | (uint32_t)i) : mykey;
        }
        const uint32_t wi = __reduce_or_sync(FULL, mine);
        if (wi) {
          const uint32_t best = __reduce_min_sync(FULL, mykey);
          const int nb = (int)(best & 3u);
          const uint32_t push = wi & ~(1u << nb);
          const int npush = __popc(push);
          if (sp + npush > CP_STACK) { if (lane == 0) atomicExch(overflow, 1u); }
          else if (push) {
            if (lane < 4 && ((push >> lane) & 1u)) {
              const uint32_t c = lane == 0 ? ch[0] : lane == 1 ? ch[1] : lane == 2 ? ch[2] : ch[3];
              stack[sp + __popc(push & ((1u << lane) - 1u))] = c & ~INTERNAL_BIT;
            }
            sp += npush;
          }
          cur = (nb == 0 ? ch[0] : nb == 1 ? ch[1] : nb == 2 ? ch[2] : ch[3]) & ~INTERNAL_BIT;
          have = true;
        }
      }
      __syncwarp();
    }
#ifdef CP_TIMELINE
    if (lane == 0 && g_timeline) {
      unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      unsigned long long *tl = g_timeline + 8 * blockIdx.x;
      if (now - tl_walk0 > tl[4]) { tl[3] = npop; tl[4] = now - tl_walk0; tl[7] = packet; }
      if (is_task) tl[6]++;
    }
#endif
    int32_t won = -1;
    CpResult r; r.sqr_d = 0.0; r.face = -1; r.c = {0.0, 0.0, 0.0};
    if (valid) {
      if (a.low > a.up) r.sqr_d = a.low;  // AABB.cpp:1145-1148
      else lane_finish(L, X, a.tri64, qr, a.up, r);
      won = r.face;
      if (slot < 0) {
        closest_store(a, qi, r);
        if (a.stash) {
          const D3 p = load_query(qr);
          a.stash[t] = ShardStash{(float)p.x, (float)p.y, (float)p.z, (uint32_t)qi, r.sqr_d};
        }
      }
    }
    if (slot >= 0) { cp_finish_shared(a, slot, qi, t, valid, r.sqr_d, r.face); slot = -1; }   // warp-uniform
    seed_face = __shfl_sync(FULL, won, 0);   // lane 0's winner seeds the next packet of the group
    if (COUNT && lane == 0) { atomicAdd(&a.counters[0], nvis); atomicAdd(&a.counters[1], ntri); }
    __syncwarp();
   }
  }
}
// one flag per (host thread, device); zeroed before any kernel of any stream can see it
struct OverflowFlag {
  uint32_t *d = nullptr;
  OverflowFlag() { d = (uint32_t *)plat::dev_alloc(4); PLAT_CHECK(cudaMemset(d, 0, 4)); PLAT_CHECK(cudaDeviceSynchronize()); }
  ~OverflowFlag() { plat::dev_free(d); }
};
uint32_t *overflow_flag() { return plat::per_device<OverflowFlag>().d; }
void reset_overflow(stream_t s) { plat::dev_memset(overflow_flag(), 0, 4, s); plat::sync(s); }
// stack overflow in a walk (tree deeper than the per-warp stack can follow) must never pass silently
void throw_if_overflow(stream_t s) {
  uint32_t h = 0;
  plat::d2h(&h, overflow_flag(), 4, s);
  plat::sync(s);
  if (h) { reset_overflow(s); throw std::runtime_error("traversal stack overflow (tree too deep)"); }
}
void launch_closest(const ClosestArgs &a, stream_t s) {
  if (a.n <= 0) return;
  const int64_t warps = (a.n + 31) / 32;
  // one packet per warp, blocks in Morton order: the packets in flight form one sliding window of the
  // query set, so the tree region they touch stays L2-resident (a grid-stride loop spreads it over ~14 regions)
  int64_t blocks = std::min<int64_t>((warps + CP_WARPS - 1) / CP_WARPS, (int64_t)0x7FFFFFFF);
#if CP_TOP == 0
  static_assert(CP_WARPS == 1, "the persistent schedule hands out packets to one-warp blocks");
#endif
  if (warps > (int64_t)0xFFF00000) throw std::invalid_argument("too many queries in one call (packet index is 32 bits)");
  DevBuf<unsigned long long> own;
  ClosestArgs b = a;
  b.chunk = CP_CHUNK;
  if (!b.remote)   // 32 packets per chunk from 2.4 M queries up; a 256 Ki-row chunk of the host path gets 2
    while (b.chunk > 2u && (int64_t)b.chunk * 16 * plat::sm_count() > warps) b.chunk >>= 1;
  if (!b.sched) { own.alloc(SCHED_WORDS); b.sched = own.p; }   // no caller scratch: a private state for this launch
  plat::for_n(SCHED_INIT_WORDS, InitSched{b.sched, b.remote ? 1 : 0}, s);
  blocks = std::min<int64_t>(std::min<int64_t>(blocks, (int64_t)plat::sm_count() * CP_MIN_BLOCKS), CP_MAILBOXES);
  if (b.counters) closest_packet_kernel<true><<<(unsigned)blocks, 32 * CP_WARPS, 0, s>>>(b, overflow_flag());
  else closest_packet_kernel<false><<<(unsigned)blocks, 32 * CP_WARPS, 0, s>>>(b, overflow_flag());
  if (own.p) plat::sync(s);
  plat::launch_counter()++;
  PLAT_CHECK(cudaGetLastError());
}
#endif

// winding number and, optionally, the fused epilogue of signed_distance.cpp:99-135
struct WindingQuery {
  const fwn::FwnNode *nodes; const float4 *tri32;
  const double *P; int64_t n; const uint32_t *perm;
  float acc2; int order;
  int mode;  // 0: W only (fast_winding_number.cpp:435); 1: signed distance FWN; 2: unsigned
  double up, low; int32_t nF;
  double *W; double *S; int32_t *I; double *C;  // S holds sqrD on entry when mode!=0
  uint32_t *overflow;
  const ShardStash *stash = nullptr;   // sharded: row, query and squared distance of sorted position t come from here
  // Row of sorted position t, its query and (mode != 0) its squared distance.  Rows that need no winding number are
  // finished here; returns whether the walk must run.
  DEVFN bool begin(int64_t t, int64_t &i, float &qx, float &qy, float &qz, double &sqrd) const {
    sqrd = 0.0;
    if (stash) { const ShardStash st = stash[t]; i = st.row; qx = st.qx; qy = st.qy; qz = st.qz; sqrd = st.sqrd; }
    else {
      i = perm ? (int64_t)perm[t] : t;
      if (mode != 0) sqrd = S[i];
    }
    if (mode != 0) {
      if (sqrd >= up || sqrd < low) {
        S[i] = NAN; I[i] = nF + 1; C[i] = 0.0; C[n + i] = 0.0; C[2 * n + i] = 0.0;
        if (W) W[i] = NAN;
        return false;
      }
      if (mode == 2) { S[i] = sqrt(sqrd); if (W) W[i] = NAN; return false; }
    }
    if (!stash) { qx = (float)P[i]; qy = (float)P[n + i]; qz = (float)P[2 * n + i]; }
    return true;
  }
  DEVFN void epilogue(int64_t i, float om, double sqrd) const {
    const double FOUR_PI = 4.0 * 3.1415926535897932384626433832795;
    const double w = (double)om / FOUR_PI;
    if (mode == 0) { W[i] = w; return; }
    const double wf = (double)(float)w;  // narrowed to float by fast_winding_number.cpp:452
    if (W) W[i] = wf;
    S[i] = (1.0 - 2.0 * fabs(wf)) * sqrt(sqrd);
  }
  DEVFN void operator()(int64_t t) const {   // one thread, one query (host emulation; the device walks packets)
    int64_t i; float qx, qy, qz; double sqrd;
    if (!begin(t, i, qx, qy, qz, sqrd)) return;
    uint32_t stack[fwn::FWN_STACK];
    epilogue(i, fwn::winding_solid_angle(nodes, tri32, qx, qy, qz, acc2, order, stack, nullptr, overflow), sqrd);
  }
};

// work counters of the winding traversal (the closest-triangle walk counts per warp, see ClosestArgs)
struct CountWinding {
  const fwn::FwnNode *fnodes; const float4 *tri32;
  const double *P; int64_t n; int order;
  uint32_t *acc;  // [2] fwn nodes, fwn tris (sums; n must keep them < 2^32)
  uint32_t *overflow;
  DEVFN void operator()(int64_t i) const {
    uint32_t sn[fwn::FWN_STACK];
    uint32_t w[2] = {0, 0};
    fwn::winding_solid_angle(fnodes, tri32, (float)P[i], (float)P[n + i], (float)P[2 * n + i], 4.0f, order, sn, w, overflow);
    plat::atomic_add(&acc[0], w[0]); plat::atomic_add(&acc[1], w[1]);
  }
};

// exact winding number (wn.h) and, optionally, the epilogue of signed_distance.cpp:107-113, 611-629:
// s = 1 - 2 w, S = s * sqrt(sqrd); out-of-bounds rows as :99-105
struct ExactWindingQuery {
  const fwn::CpNode *nodes; const double *tri64; wn::View v;
  const double *P; int64_t n; const uint32_t *perm;
  int mode;   // 0: W only (winding_number.cpp:44-50); 1: signed distance
  double up, low; int32_t nF;
  double *W; double *S; int32_t *I; double *C;   // S holds sqrD on entry when mode 1
  uint32_t *overflow;
  // rows that need no winding number are finished here; returns whether the walk must run for row i
  DEVFN bool prelude(int64_t i) const {
    if (mode == 1) {
      const double sqrd = S[i];
      if (sqrd >= up || sqrd < low) {
        S[i] = NAN; I[i] = nF + 1; C[i] = 0.0; C[n + i] = 0.0; C[2 * n + i] = 0.0;
        if (W) W[i] = NAN;
        return false;
      }
    }
    return true;
  }
  DEVFN void epilogue(int64_t i, double w) const {
    if (W) W[i] = w;
    if (mode == 1) S[i] = fwn::dmul(fwn::dsub(1.0, fwn::dmul(2.0, w)), sqrt(S[i]));
  }
  DEVFN void operator()(int64_t t) const {   // one thread, one query (host emulation; the device walks packets)
    const int64_t i = perm ? (int64_t)perm[t] : t;
    if (!prelude(i)) return;
    uint32_t stack[wn::STACK];
    bool ovf = false;
    const double w = wn::winding_number(nodes, tri64, v, P[i], P[n + i], P[2 * n + i], stack, ovf);
    if (ovf) plat::atomic_max(overflow, 1u);
    epilogue(i, w);
  }
};

// SIGNED_DISTANCE_TYPE_PSEUDONORMAL epilogue: igl::pseudonormal_test (pseudonormal_test.cpp:24-150) on the
// winning face, S = s*sqrt(d2), N = the normal used (signed_distance.cpp:125-135); out-of-bounds rows as :99-105
struct PseudonormalQuery {
  const double *tri64, *pn;
  const double *P; int64_t n;
  double up, low; int32_t nF;
  double *S; int32_t *I; double *C; double *N;  // S holds sqrD on entry
  DEVFN static double dist3(const double *p, const double *q) {
    const double x = p[0] - q[0], y = p[1] - q[1], z = p[2] - q[2];
    return sqrt(x * x + y * y + z * z);
  }
  DEVFN void operator()(int64_t i) const {
    const double sqrd = S[i];
    if (sqrd >= up || sqrd < low) {
      S[i] = NAN; I[i] = nF + 1; C[i] = 0.0; C[n + i] = 0.0; C[2 * n + i] = 0.0;
      if (N) { N[i] = 0.0; N[n + i] = 0.0; N[2 * n + i] = 0.0; }
      return;
    }
    const int32_t f = I[i];
    const double q[3] = {P[i], P[n + i], P[2 * n + i]}, c[3] = {C[i], C[n + i], C[2 * n + i]};
    const double *t = tri64 + 9 * (size_t)f, *A = t, *B = t + 3, *Cv = t + 6;
    const double *np = pn + 21 * (size_t)f;   // face normal
    // doublearea(A,B,C): Kahan's Heron on the sorted edge lengths (doublearea.cpp:103-112,166-177)
    double l0 = dist3(B, Cv), l1 = dist3(Cv, A), l2 = dist3(A, B), tmp;
    if (l0 < l1) { tmp = l0; l0 = l1; l1 = tmp; }
    if (l1 < l2) { tmp = l1; l1 = l2; l2 = tmp; }
    if (l0 < l1) { tmp = l0; l0 = l1; l1 = tmp; }
    const double area = 2.0 * 0.25 * sqrt((l0 + (l1 + l2)) * (l2 - (l0 - l1)) * (l2 + (l0 - l1)) * (l0 + (l1 - l2)));
    const double MIN_DOUBLE_AREA = 1e-4, epsilon = 1e-12;
    if (area > MIN_DOUBLE_AREA) {
      double v0[3], v1[3], v2[3];
      for (int d = 0; d < 3; ++d) { v0[d] = B[d] - A[d]; v1[d] = Cv[d] - A[d]; v2[d] = c[d] - A[d]; }
      const double d00 = v0[0] * v0[0] + v0[1] * v0[1] + v0[2] * v0[2], d01 = v0[0] * v1[0] + v0[1] * v1[1] + v0[2] * v1[2];
      const double d11 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2], d20 = v2[0] * v0[0] + v2[1] * v0[1] + v2[2] * v0[2];
      const double d21 = v2[0] * v1[0] + v2[1] * v1[1] + v2[2] * v1[2];
      const double denom = d00 * d11 - d01 * d01;
      double b[3];
      b[1] = (d11 * d20 - d01 * d21) / denom;
      b[2] = (d00 * d21 - d01 * d20) / denom;
      b[0] = 1.0f - (b[1] + b[2]);
      const int type = (b[0] <= epsilon) + (b[1] <= epsilon) + (b[2] <= epsilon);
      if (type == 2) { for (int x = 0; x < 3; ++x) if (b[x] > epsilon) { np = pn + 21 * (size_t)f + 3 + 3 * x; break; } }
      else if (type == 1) { for (int x = 0; x < 3; ++x) if (b[x] <= epsilon) { np = pn + 21 * (size_t)f + 12 + 3 * x; break; } }
    } else {
      bool found = false;
      for (int v = 0; v < 3 && !found; ++v)
        if (dist3(c, t + 3 * v) < epsilon) { found = true; np = pn + 21 * (size_t)f + 3 + 3 * v; }
      for (int e = 0; e < 3 && !found; ++e) {
        const double *Sv = t + 3 * ((e + 1) % 3), *D = t + 3 * ((e + 2) % 3);
        double dms[3], smp[3];
        for (int d = 0; d < 3; ++d) { dms[d] = D[d] - Sv[d]; smp[d] = Sv[d] - c[d]; }
        const double vs = dms[0] * dms[0] + dms[1] * dms[1] + dms[2] * dms[2];
        const double tt = -(dms[0] * smp[0] + dms[1] * smp[1] + dms[2] * smp[2]) / vs;
        double sq;
        if (tt < 0) sq = smp[0] * smp[0] + smp[1] * smp[1] + smp[2] * smp[2];
        else if (tt > 1) { const double dd = dist3(c, D); sq = dd * dd; }
        else { double pr[3]; for (int d = 0; d < 3; ++d) pr[d] = (1 - tt) * Sv[d] + tt * D[d]; const double dd = dist3(c, pr); sq = dd * dd; }
        if (sqrt(sq) < epsilon) { found = true; np = pn + 21 * (size_t)f + 12 + 3 * e; }
      }
    }
    const double nx = np[0], ny = np[1], nz = np[2];
    const double dotp = (q[0] - c[0]) * nx + (q[1] - c[1]) * ny + (q[2] - c[2]) * nz;
    S[i] = (dotp >= 0 ? 1. : -1.) * sqrt(sqrd);
    if (N) { N[i] = nx; N[n + i] = ny; N[2 * n + i] = nz; }
  }
};

struct FillConst { double *S; int32_t *I; double *C; double *W; int64_t n; double sv; int32_t iv; double cv;
  DEVFN void operator()(int64_t i) const { if (S) S[i] = sv; if (I) I[i] = iv; if (C) { C[i] = cv; C[n + i] = cv; C[2 * n + i] = cv; } if (W) W[i] = sv; } };

// The two hot kernels get their own entry points (names that profilers can filter on)
// and their own launch geometry; in the host-emulation build they are plain loops.
#ifdef IGLB200_EMUL
template <class F> void launch_query(const char *, int64_t n, F f, stream_t s) { plat::for_n(n, f, s); }
#else
// packet walk: one-warp blocks (no block barrier in the dynamic fetch), 96 registers, 20 warps per SM.  c2: 32 threads x 20 blocks
// 11.5 ms; 64 x 10: 12.0; 64 x 12 (80 regs): 12.6; 64 x 8: 12.8; 128 x 5: 12.9
#ifndef WN_THREADS
#define WN_THREADS 32
#endif
#ifndef WN_MIN_BLOCKS
#define WN_MIN_BLOCKS (640 / WN_THREADS)
#endif
// one warp = one packet of 32 consecutive (Morton-adjacent) rows [t0, t0 + 32)
template <class F> DEVFN void winding_packet(const F &f, int64_t t0, uint2 *stack) {
  const int64_t t = t0 + (threadIdx.x & 31);
  int64_t i = 0;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  double sqrd = 0.0;
  const bool alive = t < f.n && f.begin(t, i, qx, qy, qz, sqrd);
  const float om = fwn::winding_solid_angle_packet(f.nodes, f.tri32, alive, qx, qy, qz, f.acc2, f.order, stack, f.overflow);
  if (alive) f.epilogue(i, om, sqrd);
}
__global__ void __launch_bounds__(WN_THREADS, WN_MIN_BLOCKS) winding_number_kernel(int64_t n, WindingQuery f) {
  __shared__ uint2 stacks[WN_THREADS / 32][fwn::FWN_PSTACK];
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x; b < n; b += (int64_t)gridDim.x * blockDim.x)
    winding_packet(f, b + (threadIdx.x & ~31), stacks[threadIdx.x >> 5]);
}
inline unsigned query_grid(int64_t n, int threads) {
  const int64_t blocks = (n + threads - 1) / threads;
  return (unsigned)std::min<int64_t>(blocks, 0x7FFFFFFF);
}
void launch_query(const char *, int64_t n, WindingQuery f, stream_t s) {
  winding_number_kernel<<<query_grid(n, WN_THREADS), WN_THREADS, 0, s>>>(n, f);
  plat::launch_counter()++;
  PLAT_CHECK(cudaGetLastError());
}
// one warp = one packet of 32 consecutive (Morton-adjacent) rows
__global__ void __launch_bounds__(64) exact_winding_kernel(int64_t n, ExactWindingQuery f) {
  __shared__ uint2 stacks[2][wn::PSTACK];
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x; b < n; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = b + threadIdx.x;
    const bool valid = t < n;
    const int64_t i = valid ? (f.perm ? (int64_t)f.perm[t] : t) : 0;
    const bool alive = valid && f.prelude(i);
    double px = 0.0, py = 0.0, pz = 0.0;
    if (alive) { px = f.P[i]; py = f.P[n + i]; pz = f.P[2 * n + i]; }
    const double w = wn::winding_number_packet(f.nodes, f.tri64, f.v, alive, px, py, pz, stacks[threadIdx.x >> 5], f.overflow);
    if (alive) f.epilogue(i, w);
  }
}
void launch_query(const char *, int64_t n, ExactWindingQuery f, stream_t s) {
  exact_winding_kernel<<<query_grid(n, 64), 64, 0, s>>>(n, f);
  plat::launch_counter()++;
  PLAT_CHECK(cudaGetLastError());
}
#endif

// scratch layout: [keys_in | keys_out | vals_in | vals_out(perm) | cub temp | schedule state of the search kernel]
struct Scratch {
  uint32_t *kin, *kout, *vin, *perm; void *temp; size_t temp_bytes; unsigned long long *sched;
};
Scratch carve(void *scratch, int64_t n) {
  Scratch s;
  const size_t a = ((size_t)n * 4 + 255) & ~size_t(255);
  char *p = (char *)scratch;
  s.kin = (uint32_t *)p; s.kout = (uint32_t *)(p + a); s.vin = (uint32_t *)(p + 2 * a); s.perm = (uint32_t *)(p + 3 * a);
  s.temp = p + 4 * a; s.temp_bytes = plat::sort_pairs_temp_bytes(n);
  s.sched = (unsigned long long *)(p + 4 * a + ((s.temp_bytes + 255) & ~size_t(255)));
  return s;
}

const uint32_t *morton_perm(const iglb200_mesh *m, const double *P, int64_t n, void *scratch, int sort_queries, stream_t s) {
  if (!sort_queries || !scratch || n < 2) return nullptr;
  Scratch sc = carve(scratch, n);
  MortonKeys mk;
  mk.P = P; mk.n = n; mk.keys = sc.kin; mk.vals = sc.vin;
  for (int d = 0; d < 3; ++d) {
    const double c = 0.5 * (m->h.bbox_lo[d] + m->h.bbox_hi[d]);
    double ext = (m->h.bbox_hi[d] - m->h.bbox_lo[d]);
    if (!(ext > 0.0)) ext = 1.0;
    mk.lo[d] = c - ext;           // box doubled about its centre
    mk.inv[d] = 1024.0 / (2.0 * ext);
  }
  // about one query per Morton cell: bits per axis = ceil(log2(n) / 3) in [6, 10].  The sort is stable, so the queries of
  // a cell keep their order; fewer key bits are fewer radix passes (16.7 M queries: 24-bit keys, 3 passes instead of 4)
  int bits = 6;
  while (bits < 10 && (int64_t(1) << (3 * bits)) < n) ++bits;
  mk.shift = 10 - bits;
  plat::Span sp("morton_sort", s);
  plat::for_n(n, mk, s);
  plat::sort_pairs_u32(sc.temp, sc.temp_bytes, sc.kin, sc.kout, sc.vin, sc.perm, n, 3 * bits, s);
  return sc.perm;
}

// set while a host-path worker drives its chunks: other workers' kernels run beside this thread's on the same device
bool &launches_share_device() { static thread_local bool f = false; return f; }
struct ShareDeviceScope {
  bool old;
  ShareDeviceScope() : old(launches_share_device()) { launches_share_device() = true; }
  ~ShareDeviceScope() { launches_share_device() = old; }
};
ClosestArgs closest_args(const iglb200_mesh *m, const double *P, int64_t n, const uint32_t *perm, double up, double low,
                         double *sqrD, int32_t *I, double *C, uint32_t *counters, unsigned long long *sched) {
  ClosestArgs a{m->pnodes(), m->tri64(), m->triq(), P, n, perm, up, low, (float)((double)m->h.absmax * m->h.sigma), sqrD, I, C, counters, sched};
  a.sigma = m->h.sigma;
  a.nnodes = (uint32_t)m->h.nnodes;
  a.donate = launches_share_device() ? 0 : 1;
  return a;
}

void check_mesh(const iglb200_mesh *m) {
  if (!m || m->h.magic != IGLB200_ARENA_MAGIC) throw std::invalid_argument("invalid mesh handle");
}
// *_dev entry points take pointers of the calling thread's current device: the copy of the mesh on that device answers
const iglb200_mesh *mesh_here(const iglb200_mesh *m) {
  check_mesh(m);
  const iglb200_mesh *r = m->on(plat::current_device());
  if (!r) throw std::invalid_argument("the mesh has no copy on the calling thread's current device (iglb200_set_device / iglb200_init)");
  return r;
}
void check_count(int64_t n) {
  if (n > (int64_t)0x7FFFFFFF) throw std::invalid_argument("more than 2^31-1 queries in one device-resident call (host-pointer calls chunk internally)");
}

void bounds_to_sqr(double lower, double upper, double &low, double &up) {  // signed_distance.cpp:81-84
  const double max_abs = std::max(std::fabs(lower), std::fabs(upper));
  up = std::pow(max_abs, 2.0);
  low = std::pow(std::max(max_abs - (upper - lower), 0.0), 2.0);
}

wn::View wn_view(const iglb200_mesh *m) {
  const char *a = m->arena;
  return wn::View{(const uint32_t *)(a + m->h.off_wn_cnt), (const uint32_t *)(a + m->h.off_wn_start),
                  (const uint32_t *)(a + m->h.off_wn_bstart), (const uint32_t *)(a + m->h.off_wn_leafface),
                  (const uint64_t *)(a + m->h.off_wn_bent)};
}

void run_signed_distance(const iglb200_mesh *m, const double *P, int64_t n, int sign_type, double lower, double upper,
                         double *S, int32_t *I, double *C, double *W, double *N, void *scratch, int sort_queries, stream_t s) {
  if (sign_type < 0 || sign_type > IGLB200_SIGN_FAST_WINDING_NUMBER) throw std::domain_error("unknown signed distance type");
  const bool exact_wn = sign_type == IGLB200_SIGN_WINDING_NUMBER || sign_type == IGLB200_SIGN_DEFAULT;
  if (exact_wn && m->h.nnodes > 0 && !m->h.off_wn_cnt)
    throw std::invalid_argument("mesh was created without IGLB200_MESH_WINDING_NUMBER");
  if (sign_type == IGLB200_SIGN_PSEUDONORMAL && m->h.nF > 0 && !m->pn())
    throw std::invalid_argument("mesh was created without IGLB200_MESH_PSEUDONORMALS");
  if (n == 0) return;
  if (m->h.nF == 0) {  // signed_distance.cpp:398-405
    plat::for_n(n, FillConst{S, I, C, W, n, NAN, -1, NAN}, s);
    if (N) plat::for_n(n, FillConst{nullptr, nullptr, N, nullptr, n, NAN, -1, NAN}, s);
    return;
  }
  if (m->h.nnodes == 0) {  // every face excluded: nothing to be close to
    plat::for_n(n, FillConst{S, I, C, W, n, NAN, (int32_t)m->h.nF + 1, 0.0}, s);
    if (N) plat::for_n(n, FillConst{nullptr, nullptr, N, nullptr, n, NAN, -1, 0.0}, s);
    return;
  }
  double low, up;
  bounds_to_sqr(lower, upper, low, up);
  const uint32_t *perm = morton_perm(m, P, n, scratch, sort_queries, s);
  { plat::Span sp("closest", s);
    launch_closest(closest_args(m, P, n, perm, up, low, S, I, C, nullptr, scratch ? carve(scratch, n).sched : nullptr), s); }
  if (sign_type == IGLB200_SIGN_PSEUDONORMAL) {
    plat::Span sp("pseudonormal", s);
    plat::for_n(n, PseudonormalQuery{m->tri64(), m->pn(), P, n, up, low, (int32_t)m->h.nF, S, I, C, N}, s);
    if (W) plat::for_n(n, FillConst{W, nullptr, nullptr, nullptr, n, NAN, 0, 0.0}, s);
    return;
  }
  if (exact_wn) {
    plat::Span sp("exact_winding", s);
    launch_query("exact_winding", n, ExactWindingQuery{m->cnodes(), m->tri64(), wn_view(m), P, n, perm, 1, up, low, (int32_t)m->h.nF,
                                                       W, S, I, C, overflow_flag()}, s);
    return;
  }
  { plat::Span sp("winding", s);
    launch_query("winding", n, WindingQuery{m->fnodes(), m->tri32(), P, n, perm, 4.0f, m->h.order,
                                            sign_type == IGLB200_SIGN_UNSIGNED ? 2 : 1, up, low, (int32_t)m->h.nF, W, S, I, C, overflow_flag()}, s); }
}

// igl::winding_number(V,F,O,W) for triangle meshes (winding_number.cpp:38-52)
void run_exact_winding(const iglb200_mesh *m, const double *O, int64_t n, double *W, void *scratch, int sort_queries, stream_t s) {
  if (n == 0) return;
  if (m->h.nnodes == 0) { plat::for_n(n, FillConst{W, nullptr, nullptr, nullptr, n, 0.0, 0, 0.0}, s); return; }
  if (!m->h.off_wn_cnt) throw std::invalid_argument("mesh was created without IGLB200_MESH_WINDING_NUMBER");
  const uint32_t *perm = morton_perm(m, O, n, scratch, sort_queries, s);
  plat::Span sp("exact_winding", s);
  launch_query("exact_winding", n, ExactWindingQuery{m->cnodes(), m->tri64(), wn_view(m), O, n, perm, 0, 0.0, 0.0, (int32_t)m->h.nF,
                                                     W, nullptr, nullptr, nullptr, overflow_flag()}, s);
}

// low_sqr_d / up_sqr_d: AABB.cpp:1135-1148 — only primitives closer than up_sqr_d count (none: the result is up_sqr_d,
// I = -1); low_sqr_d > up_sqr_d returns low_sqr_d
void run_squared_distance(const iglb200_mesh *m, const double *P, int64_t n, double *sqrD, int32_t *I, double *C,
                          void *scratch, int sort_queries, stream_t s, double low_sqr_d = 0.0, double up_sqr_d = INFINITY) {
  if (n == 0) return;
  if (m->h.nnodes == 0) {  // empty tree: the reference returns the initial bound and leaves I/C unset
    plat::for_n(n, FillConst{sqrD, I, C, nullptr, n, low_sqr_d > up_sqr_d ? low_sqr_d : up_sqr_d, -1, 0.0}, s);
    return;
  }
  const uint32_t *perm = morton_perm(m, P, n, scratch, sort_queries, s);
  plat::Span sp("closest", s);
  launch_closest(closest_args(m, P, n, perm, up_sqr_d, low_sqr_d, sqrD, I, C, nullptr, scratch ? carve(scratch, n).sched : nullptr), s);
}

void run_fwn(const iglb200_mesh *m, const double *Q, int64_t n, float accuracy_scale, double *W, void *scratch,
             int sort_queries, stream_t s) {
  if (n == 0) return;
  if (m->h.nnodes == 0) { plat::for_n(n, FillConst{W, nullptr, nullptr, nullptr, n, 0.0, 0, 0.0}, s); return; }
  const uint32_t *perm = morton_perm(m, Q, n, scratch, sort_queries, s);
  plat::Span sp("winding", s);
  launch_query("winding", n, WindingQuery{m->fnodes(), m->tri32(), Q, n, perm, accuracy_scale * accuracy_scale, m->h.order, 0, 0.0, 0.0,
                              (int32_t)m->h.nF, W, nullptr, nullptr, nullptr, overflow_flag()}, s);
}

// ---------------------------------------------------------------------------------------------------
// Sharded device-resident queries: SEVERAL GPUs answer ONE query set that lives on one of them (the owner).
//   begin  (owner)      Morton-sort the queries once (keys + permutation in the owner's scratch), zero the work counter
//   run    (every GPU)  the SAME kernels as the single-GPU call — closest-triangle search, then the winding number with
//                       the fused epilogue of signed_distance.cpp:99-135 — over the super-chunks of sorted queries this
//                       GPU takes off the owner's counter.  Queries and the permutation are read from, and S/I/C/W are
//                       written to, the owner's memory directly (peer loads/stores over NVLink): there is no separate
//                       scatter, gather or un-sort pass — the transfer is the kernels' own prologue and epilogue.
//   finish (owner)      nothing but the sentinel rows of a degenerate mesh.
// Between the phases the caller orders the devices' streams (events in one process, a tiny NCCL collective between
// processes).  The owner's scratch has one layout, computed from nP alone, so that a peer only needs its base pointer.
// ---------------------------------------------------------------------------------------------------
struct ShardLayout {
  size_t oCounter, oOrder, oSort, oPsf, oWs, total;
  static constexpr int64_t MIN_SUPER = int64_t(8) << 10;
  static constexpr int MAX_SHARDS = 64;
  static constexpr size_t COUNTER_BYTES = 1024;   // [0] pieces (winding number only), [1] overlap experiment, [2 + k] slice k
  explicit ShardLayout(int64_t n) {
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t nsup_max = (size_t)(n / MIN_SUPER + 2);
    oCounter = 0; oOrder = COUNTER_BYTES; oSort = oOrder + al(4 * 4 * nsup_max);   // keys in/out, ids in/out (the order)
    oPsf = oSort + al(iglb200_query_scratch_bytes(n));                    // sorted queries as floats, x[n] y[n] z[n]
    oWs = oPsf + al(12 * (size_t)n);                                      // solid angle of sorted row t
    total = oWs + al(4 * (size_t)n);
  }
  uint32_t *order(char *scratch, int64_t n) const { return (uint32_t *)(scratch + oOrder) + 3 * (size_t)(n / MIN_SUPER + 2); }
};
// hand-out key of a super-chunk: its index (Morton order) or, as an experiment, its distance from the mesh (see shard_begin)
struct SuperKeys {
  const double *P; int64_t n; const uint32_t *perm; int64_t super_rows; double c[3]; uint32_t *keys, *ids; int far_first;
  DEVFN void operator()(int64_t sc) const {
    int64_t t = sc * super_rows + super_rows / 2;
    if (t >= n) t = n - 1;
    const int64_t i = perm ? (int64_t)perm[t] : t;
    const double x = P[i] - c[0], y = P[n + i] - c[1], z = P[2 * n + i] - c[2];
    const float d2 = (float)(x * x + y * y + z * z);
    keys[sc] = far_first ? ~plat::f2ord(d2 == d2 ? d2 : 0.f) : (uint32_t)sc;   // descending distance, or Morton order
    ids[sc] = (uint32_t)sc;
  }
};

// Overlapped sharded signed distance (experiment, off by default — see shard_overlap): the winding number does not depend
// on the search, so it can run BESIDE it (second stream, its own work counter) over the sorted queries as floats, leaving
// the solid angle of sorted row t in the owner's scratch; the owner combines afterwards.
struct GatherSortedF { const double *P; int64_t n; const uint32_t *perm; float *Psf;
  DEVFN void operator()(int64_t t) const {
    const int64_t i = perm ? (int64_t)perm[t] : t;
    Psf[t] = (float)P[i]; Psf[n + t] = (float)P[n + i]; Psf[2 * n + t] = (float)P[2 * n + i];
  } };
struct WindingSorted {
  const fwn::FwnNode *nodes; const float4 *tri32; const float *Psf; int64_t n; float acc2; int order; float *Ws; uint32_t *overflow;
  DEVFN bool begin(int64_t t, int64_t &i, float &qx, float &qy, float &qz, double &) const {
    i = t; qx = Psf[t]; qy = Psf[n + t]; qz = Psf[2 * n + t];
    return true;
  }
  DEVFN void epilogue(int64_t t, float om, double) const { Ws[t] = om; }
  DEVFN void operator()(int64_t t) const {
    uint32_t stack[fwn::FWN_STACK];
    Ws[t] = fwn::winding_solid_angle(nodes, tri32, Psf[t], Psf[n + t], Psf[2 * n + t], acc2, order, stack, nullptr, overflow);
  }
};
// owner: S, W of every row from its squared distance (left in S by the search) and its solid angle; out-of-bounds rows
struct ShardCombine {
  int sign_type; double up, low; int32_t nF; int64_t n; const uint32_t *perm; const float *Ws;
  double *S; int32_t *I; double *C; double *W;
  DEVFN void operator()(int64_t t) const {
    const int64_t i = perm ? (int64_t)perm[t] : t;
    const double sqrd = S[i];
    if (sqrd >= up || sqrd < low) {
      S[i] = NAN; I[i] = nF + 1; C[i] = 0.0; C[n + i] = 0.0; C[2 * n + i] = 0.0;
      if (W) W[i] = NAN;
      return;
    }
    if (sign_type == IGLB200_SIGN_UNSIGNED) { S[i] = sqrt(sqrd); if (W) W[i] = NAN; return; }
    const double w = (double)Ws[t] / (4.0 * 3.1415926535897932384626433832795);
    const double wf = (double)(float)w;  // narrowed to float by fast_winding_number.cpp:452
    if (W) W[i] = wf;
    S[i] = (1.0 - 2.0 * fabs(wf)) * sqrt(sqrd);
  }
};

#ifndef IGLB200_EMUL
// winding number (+ fused epilogue) of the rows of the super-chunks listed by the search kernel; blocks take tiles of
// WN_THREADS rows off a local counter
__global__ void __launch_bounds__(WN_THREADS, WN_MIN_BLOCKS) winding_shard_kernel(WindingQuery f, const uint32_t *my_supers, const uint32_t *my_count,
                                                                                  int64_t super_rows, unsigned long long *next) {
  __shared__ unsigned long long item_s;
  __shared__ uint2 stacks[WN_THREADS / 32][fwn::FWN_PSTACK];
  const int64_t per = super_rows / WN_THREADS;           // blocks' worth of rows in a super-chunk
  const unsigned long long total = (unsigned long long)*my_count * (unsigned long long)per;
  while (true) {
    if (threadIdx.x == 0) item_s = atomicAdd(next, 1ull);
    __syncthreads();
    const unsigned long long item = item_s;
    __syncthreads();
    if (item >= total) return;
    const int64_t t0 = (int64_t)my_supers[item / per] * super_rows + (int64_t)(item % per) * WN_THREADS + (threadIdx.x & ~31);
    winding_packet(f, t0, stacks[threadIdx.x >> 5]);
  }
}
// winding number only: blocks claim pieces of `piece_rows` rows off the owner's counter (system-scope atomic) until
// the rows run out; `budget` >= 0 caps the pieces this device takes (my_count counts them)
template <class F>
__global__ void __launch_bounds__(WN_THREADS, WN_MIN_BLOCKS) winding_claim_kernel(F f, unsigned long long *remote, int64_t piece_rows,
                                                                                  int64_t budget, uint32_t *my_count) {
  __shared__ unsigned long long piece_s;
  __shared__ uint2 stacks[WN_THREADS / 32][fwn::FWN_PSTACK];
  while (true) {
    if (threadIdx.x == 0) {
      unsigned long long pc = ~0ull;
      if (budget < 0 || (int64_t)atomicAdd(my_count, 1u) < budget) pc = atomicAdd_system(remote, 1ull);
      piece_s = pc;
    }
    __syncthreads();
    const unsigned long long pc = piece_s;
    __syncthreads();
    if (pc == ~0ull || (int64_t)pc * piece_rows >= f.n) return;
    const int64_t lo = (int64_t)pc * piece_rows;
    for (int64_t t0 = lo + (threadIdx.x & ~31); t0 < lo + piece_rows; t0 += WN_THREADS) winding_packet(f, t0, stacks[threadIdx.x >> 5]);
  }
}
#endif

static constexpr int64_t FWN_PIECE_ROWS = 2048;   // winding-number-only shards: rows per claim
struct ShardWs {   // per (thread, device) state of shard_run: schedule words, the list of super-chunks taken
  DevBuf<unsigned long long> sched; DevBuf<uint32_t> supers; size_t cap = 0;
  DevBuf<ShardStash> stash;
  stream_t aux = nullptr; plat::event_t e0 = nullptr, e1 = nullptr; bool have_aux = false;
  DevBuf<uint32_t> aux_count;
  void ensure_aux() {
    if (have_aux) return;
    aux = plat::stream_create(); e0 = plat::event_create(); e1 = plat::event_create(); aux_count.alloc(1); have_aux = true;
  }
  void ensure(size_t nsup, size_t nstash, stream_t s) {
    if (!sched.p) sched.alloc(SCHED_WORDS);
    if (nsup + 1 > cap) { plat::sync(s); supers.alloc(nsup + 1); cap = nsup + 1; }
    if (nstash > stash.n) { plat::sync(s); stash.alloc(nstash); }
  }
};

int64_t shard_super_rows(int64_t n, int n_shards) {
  // >= 64 super-chunks per device so that the last ones even out the load, each large enough to keep a device's L2
  // on one region of the tree: [8 Ki, 256 Ki] rows, a multiple of a chunk (CP_CHUNK packets of 32)
  const int64_t unit = (int64_t)CP_CHUNK * 32;
  int64_t r = n / ((int64_t)std::max(n_shards, 1) * 64);
  r = std::max<int64_t>(int64_t(8) << 10, std::min<int64_t>(int64_t(256) << 10, r));
  return (r + unit - 1) / unit * unit;
}

// Default: the winding kernel runs AFTER the search on each GPU, over the rows that GPU took, from its local stash.
// IGLB200_SHARD_OVERLAP=1 (experiment, measured SLOWER: c2 57.7 vs 53.7 ms on one GPU, 14.0 vs 9.7 ms on eight) runs it
// BESIDE the search on a second stream with its own work counter, hoping to fill the search's tail: both kernels are
// persistent, so whichever gets its blocks resident first keeps the register file and the other one waits or thrashes.
bool shard_overlap() { static const bool on = [] { const char *e = std::getenv("IGLB200_SHARD_OVERLAP"); return e && e[0] == '1'; }(); return on; }

void shard_begin(const iglb200_mesh *m, const double *P, int64_t n, char *scratch, int sort_queries, int64_t super_rows, stream_t s) {
  const ShardLayout L(n);
  plat::dev_memset(scratch + L.oCounter, 0, ShardLayout::COUNTER_BYTES, s);
  if (n == 0) return;
  if (super_rows < ShardLayout::MIN_SUPER) throw std::invalid_argument("super_rows must be at least 8192");
  const uint32_t *perm = morton_perm(m, P, n, scratch + L.oSort, sort_queries, s);
  // Hand-out order of the super-chunks (the cub temp area of the query sort is free again: same stream).  Morton order by
  // default: consecutive super-chunks share tree nodes in L2.  IGLB200_SHARD_ORDER=far hands out the farthest first — the
  // idea being that a launch should not END on its costliest packets — but with cylinder bounds the far field is the cheap
  // part and the jumps cost L2 locality: measured slower (c2, 8 GPUs: search 7.0 vs 6.4 ms per rank).
  static const int far_first = [] { const char *e = std::getenv("IGLB200_SHARD_ORDER"); return (e && std::string(e) == "far") ? 1 : 0; }();
  const int64_t nsup = (n + super_rows - 1) / super_rows;
  const size_t cap = (size_t)(n / ShardLayout::MIN_SUPER + 2);
  uint32_t *kin = (uint32_t *)(scratch + L.oOrder), *kout = kin + cap, *vin = kout + cap, *vout = vin + cap;
  SuperKeys sk{P, n, perm, super_rows, {0.5 * (m->h.bbox_lo[0] + m->h.bbox_hi[0]), 0.5 * (m->h.bbox_lo[1] + m->h.bbox_hi[1]),
                                          0.5 * (m->h.bbox_lo[2] + m->h.bbox_hi[2])}, kin, vin, far_first};
  plat::for_n(nsup, sk, s);
  Scratch sc = carve(scratch + L.oSort, n);
  plat::sort_pairs_u32(sc.temp, sc.temp_bytes, kin, kout, vin, vout, nsup, 32, s);
  if (shard_overlap()) plat::for_n(n, GatherSortedF{P, n, perm, (float *)(scratch + L.oPsf)}, s);
}

bool shard_degenerate(const iglb200_mesh *m, int mode) { return (mode != 2 && m->h.nF == 0) || m->h.nnodes == 0; }

// P, S, I, C, W: the owner's arrays as seen from the calling device
void shard_run(const iglb200_mesh *m, const double *P, int64_t n, char *scratch, int mode, int sign_type, double lower, double upper,
               float accuracy_scale, int sort_queries, double *S, int32_t *I, double *C, double *W, int64_t super_rows, int64_t max_supers,
               int shard, int nshards, stream_t s) {
  if (nshards < 1 || nshards > ShardLayout::MAX_SHARDS || shard < 0 || shard >= nshards)
    throw std::invalid_argument("shard_index must lie in [0, n_shards), n_shards in [1, 64]");
  if (n == 0 || shard_degenerate(m, mode)) return;   // finish fills the sentinels
  const int64_t unit = (int64_t)CP_CHUNK * 32;
  if (super_rows <= 0 || super_rows % unit) throw std::invalid_argument("super_rows must be a positive multiple of 1024");
  const ShardLayout L(n);
  unsigned long long *remote = (unsigned long long *)(scratch + L.oCounter);
  const uint32_t *perm = (sort_queries && n >= 2) ? carve(scratch + L.oSort, n).perm : nullptr;
  const int64_t nsup = (n + super_rows - 1) / super_rows;
  ShardWs &ws = plat::per_device<ShardWs>();
  ws.ensure((size_t)nsup, (mode == 0 && !shard_overlap()) ? (size_t)n : 0, s);
  plat::dev_memset(ws.supers.p, 0, 4, s);              // [0] = count, [1..] = list
  uint32_t *my_count = ws.supers.p, *my_supers = ws.supers.p + 1;
  double low = 0.0, up = INFINITY;
  if (mode == 0) bounds_to_sqr(lower, upper, low, up);
  if (mode == 0 && sign_type != IGLB200_SIGN_FAST_WINDING_NUMBER && sign_type != IGLB200_SIGN_UNSIGNED)
    throw std::domain_error("sharded queries support the FAST_WINDING_NUMBER and UNSIGNED sign types");
  const bool overlap = mode == 0 && shard_overlap();
  if (overlap && sign_type == IGLB200_SIGN_FAST_WINDING_NUMBER) {
    ws.ensure_aux();
    plat::event_record(ws.e0, s);
    plat::stream_wait(ws.aux, ws.e0);
  }
  if (mode != 2) {
    plat::Span sp("closest", s);
    ClosestArgs a = closest_args(m, P, n, perm, up, low, S, I, C, nullptr, ws.sched.p);
    a.remote = remote; a.super_chunks = (uint32_t)(super_rows / unit); a.my_supers = my_supers; a.my_count = my_count; a.max_supers = max_supers;
    a.super_order = L.order(scratch, n);
    a.nsup = (uint32_t)nsup; a.shard = shard; a.nshards = nshards;
    a.stash = (mode == 0 && !overlap) ? ws.stash.p : nullptr;
    launch_closest(a, s);
  }
  if (overlap && sign_type == IGLB200_SIGN_FAST_WINDING_NUMBER) {
    // the winding number beside the search: second stream, second counter (remote + 1), pieces of sorted rows
    // (launched after the search so that the search's blocks are dispatched first)
    plat::Span sp("winding", ws.aux);
    const WindingSorted f{m->fnodes(), m->tri32(), (const float *)(scratch + L.oPsf), n, 4.0f, m->h.order, (float *)(scratch + L.oWs), overflow_flag()};
#ifdef IGLB200_EMUL
    for (int64_t took = 0; max_supers <= 0 || took < max_supers * (super_rows / FWN_PIECE_ROWS); ++took) {
      const unsigned long long pc = __atomic_fetch_add(remote + 1, 1ull, __ATOMIC_RELAXED);
      if ((int64_t)pc * FWN_PIECE_ROWS >= n) break;
      for (int64_t t = (int64_t)pc * FWN_PIECE_ROWS; t < std::min(n, ((int64_t)pc + 1) * FWN_PIECE_ROWS); ++t) f(t);
    }
#else
    plat::dev_memset(ws.aux_count.p, 0, 4, ws.aux);
    winding_claim_kernel<<<plat::sm_count() * WN_MIN_BLOCKS, WN_THREADS, 0, ws.aux>>>(
        f, remote + 1, FWN_PIECE_ROWS, max_supers > 0 ? max_supers * (super_rows / FWN_PIECE_ROWS) : -1, ws.aux_count.p);
    plat::launch_counter()++;
    PLAT_CHECK(cudaGetLastError());
#endif
  }
  if (overlap) {
    if (sign_type == IGLB200_SIGN_FAST_WINDING_NUMBER) { plat::event_record(ws.e1, ws.aux); plat::stream_wait(s, ws.e1); }
    return;   // the owner combines in finish
  }
  if (mode == 1) return;
  plat::Span sp("winding", s);
  const float acc = mode == 2 ? accuracy_scale : 2.0f;
  const WindingQuery f{m->fnodes(), m->tri32(), P, n, perm, acc * acc, m->h.order,
                       mode == 2 ? 0 : (sign_type == IGLB200_SIGN_UNSIGNED ? 2 : 1), up, low, (int32_t)m->h.nF, W, S, I, C, overflow_flag(),
                       mode == 0 ? ws.stash.p : nullptr};
#ifdef IGLB200_EMUL
  if (mode == 2) {   // no search ran: pieces straight off the owner's counter
    for (int64_t took = 0; max_supers <= 0 || took < max_supers; ++took) {
      const unsigned long long pc = __atomic_fetch_add(remote, 1ull, __ATOMIC_RELAXED);
      if ((int64_t)pc * FWN_PIECE_ROWS >= n) break;
      for (int64_t t = (int64_t)pc * FWN_PIECE_ROWS; t < std::min(n, ((int64_t)pc + 1) * FWN_PIECE_ROWS); ++t) f(t);
    }
  } else {
    for (uint32_t k = 0; k < *my_count; ++k)
      for (int64_t t = (int64_t)my_supers[k] * super_rows; t < std::min(n, ((int64_t)my_supers[k] + 1) * super_rows); ++t) f(t);
  }
#else
  if (mode == 2) {
    // winding number only (uniform cost per row): every block claims pieces of the rows straight off the owner's counter
    winding_claim_kernel<<<plat::sm_count() * WN_MIN_BLOCKS, WN_THREADS, 0, s>>>(f, remote, FWN_PIECE_ROWS, max_supers > 0 ? max_supers : -1, my_count);
  } else {
    unsigned long long *next = ws.sched.p + 1 + 2 * 1024 + 1;
    plat::dev_memset(next, 0, 8, s);
    winding_shard_kernel<<<plat::sm_count() * WN_MIN_BLOCKS, WN_THREADS, 0, s>>>(f, my_supers, my_count, super_rows, next);
  }
  plat::launch_counter()++;
  PLAT_CHECK(cudaGetLastError());
#endif
}

void shard_finish(const iglb200_mesh *m, int64_t n, char *scratch, int mode, int sign_type, double lower, double upper, int sort_queries,
                  double *S, int32_t *I, double *C, double *W, stream_t s) {
  if (n == 0) return;
  if (!shard_degenerate(m, mode)) {
    if (mode == 0 && shard_overlap()) {
      const ShardLayout L(n);
      double low = 0.0, up = INFINITY;
      bounds_to_sqr(lower, upper, low, up);
      const uint32_t *perm = (sort_queries && n >= 2) ? carve(scratch + L.oSort, n).perm : nullptr;
      plat::Span sp("combine", s);
      plat::for_n(n, ShardCombine{sign_type, up, low, (int32_t)m->h.nF, n, perm, (const float *)(scratch + L.oWs), S, I, C, W}, s);
    }
    return;
  }
  if (mode != 2 && m->h.nF == 0) { plat::for_n(n, FillConst{S, I, C, mode == 0 ? W : nullptr, n, NAN, -1, NAN}, s); return; }
  if (mode == 2) plat::for_n(n, FillConst{W, nullptr, nullptr, nullptr, n, 0.0, 0, 0.0}, s);
  else if (mode == 1) plat::for_n(n, FillConst{S, I, C, nullptr, n, INFINITY, -1, 0.0}, s);
  else plat::for_n(n, FillConst{S, I, C, W, n, NAN, (int32_t)m->h.nF + 1, 0.0}, s);
}

// upload a strided host matrix n x 3 as canonical column-major doubles
void upload_points(const void *P, int is_f64, int64_t n, int64_t rs, int64_t cs, DevBuf<double> &out, stream_t s) {
  out.alloc((size_t)3 * n);
  if (n == 0) return;
  if (rs <= 0 || cs <= 0) throw std::invalid_argument("strides must be positive");
  if (is_f64 && rs == 1 && cs == n) { plat::h2d(out.p, P, (size_t)24 * n, s); return; }
  const size_t span = (size_t)((n - 1) * rs + 2 * cs + 1);
  const size_t esz = is_f64 ? 8 : 4;
  DevBuf<char> raw(span * esz);
  plat::h2d(raw.p, P, span * esz, s);
  if (is_f64) plat::for_n(n, ToCanon3<double>{(const double *)raw.p, rs, cs, n, out.p}, s);
  else plat::for_n(n, ToCanon3<float>{(const float *)raw.p, rs, cs, n, out.p}, s);
  plat::sync(s);  // raw is released on return
}
void download_points(const DevBuf<double> &src, int64_t n, void *C, int is_f64, int64_t rs, int64_t cs, stream_t s) {
  if (n == 0 || !C) return;
  if (is_f64 && rs == 1 && cs == n) { plat::d2h(C, src.p, (size_t)24 * n, s); plat::sync(s); return; }
  const size_t span = (size_t)((n - 1) * rs + 2 * cs + 1);
  const size_t esz = is_f64 ? 8 : 4;
  // strided destination: only elements of the view may be written, so stage on the host
  std::vector<double> h((size_t)3 * n);
  plat::d2h(h.data(), src.p, (size_t)24 * n, s);
  plat::sync(s);
  (void)span; (void)esz;
  if (is_f64) { double *d = (double *)C; for (int64_t i = 0; i < n; ++i) for (int j = 0; j < 3; ++j) d[i * rs + j * cs] = h[j * n + i]; }
  else { float *d = (float *)C; for (int64_t i = 0; i < n; ++i) for (int j = 0; j < 3; ++j) d[i * rs + j * cs] = (float)h[j * n + i]; }
}
void download_scalar(const DevBuf<double> &src, int64_t n, void *dst, int is_f64, stream_t s) {
  if (n == 0 || !dst) return;
  if (is_f64) { plat::d2h(dst, src.p, (size_t)8 * n, s); plat::sync(s); return; }
  DevBuf<float> f((size_t)n);
  plat::for_n(n, ToFloat1{src.p, f.p}, s);
  plat::d2h(dst, f.p, (size_t)4 * n, s);
  plat::sync(s);
}


// ---------------------------------------------------------------------------------------------------
// Host-pointer calls.  The batch is cut into chunks that are dealt — one atomic counter — to worker
// threads: IGLB200_HOST_WORKERS (default 3) per device, for every device that holds a copy of the mesh
// (group.h).  A worker owns one stream, one grow-only device workspace and one grow-only pinned staging
// buffer (no cudaMalloc on the steady-state call path) and takes a chunk through
//     [pageable input: memcpy into pinned staging] -> H2D -> layout shim -> sort + kernels -> layout shim
//     -> D2H [-> memcpy out of pinned staging]
// so that one worker's copies overlap the other workers' kernels, on one GPU as on eight.  Each device
// DMA-writes its chunks straight into the caller's S/I/C when those are page-locked.  The chunk counter
// may also live in memory shared between PROCESSES (the *_coop entry points: one process per GPU under
// torchrun, all of them mapping the same host arrays).
// This is the loop of signed_distance.cpp:90-138 / AABB.cpp:1241-1265 / fast_winding_number.cpp:429-436.
// ---------------------------------------------------------------------------------------------------
struct WorkerCtx {
  int device = 0;
  stream_t s = nullptr;
  char *dbuf = nullptr; size_t dcap = 0;
  char *hbuf = nullptr; size_t hcap = 0;
  ~WorkerCtx() { plat::dev_free(dbuf); plat::pinned_free(hbuf); plat::stream_destroy(s); }
  void ensure(size_t dbytes, size_t hbytes) {
    if (!s) s = plat::stream_create();
    if (dbytes > dcap) { plat::sync(s); plat::dev_free(dbuf); dbuf = nullptr; dcap = 0; dbuf = (char *)plat::dev_alloc(dbytes); dcap = dbytes; }
    if (hbytes > hcap) { plat::sync(s); plat::pinned_free(hbuf); hbuf = nullptr; hcap = 0; hbuf = (char *)plat::pinned_alloc(hbytes); hcap = hbytes; }
  }
};

// an n x 3 host matrix as the chunk pipeline sees it
enum MatKind { MAT_NONE = 0, MAT_COLS, MAT_ROWS, MAT_OTHER };
struct HostMat {
  char *p = nullptr; int f64 = 1; int64_t rs = 0, cs = 0; MatKind kind = MAT_NONE; bool pinned = false;
  size_t esz() const { return f64 ? 8 : 4; }
};
HostMat host_mat(const void *p, int f64, int64_t n, int64_t rs, int64_t cs) {
  HostMat m; m.p = (char *)p; m.f64 = f64; m.rs = rs; m.cs = cs;
  if (!p) { m.kind = MAT_NONE; return m; }
  if (rs == 1 && (cs >= n || n <= 1)) m.kind = MAT_COLS;        // column-major (Eigen's default), any leading dimension
  else if (cs == 1 && rs == 3) m.kind = MAT_ROWS;               // dense row-major (numpy's default)
  else m.kind = MAT_OTHER;
  m.pinned = plat::host_is_pinned(p);
  return m;
}
struct HostVec { char *p = nullptr; int f64 = 1; bool pinned = false; size_t esz() const { return f64 ? 8 : 4; } };
HostVec host_vec(const void *p, int f64) { HostVec v; v.p = (char *)p; v.f64 = f64; v.pinned = p ? plat::host_is_pinned(p) : false; return v; }

struct ToFloat3Cols { const double *src; float *dst; DEVFN void operator()(int64_t i) const { dst[i] = (float)src[i]; } };
struct FromFloatCols { const float *src; double *dst; DEVFN void operator()(int64_t i) const { dst[i] = (double)src[i]; } };

struct HostQuery {
  const iglb200_mesh *m = nullptr;
  int mode = 0;   // 0: signed distance (S,I,C[,W][,N]); 1: squared distance (S=sqrD,I,C); 2: fast winding number (W); 3: exact winding number (W)
  HostMat P, C, N;
  HostVec S, W;
  int32_t *I = nullptr; bool i_pinned = false;
  int64_t n = 0;
  int sign_type = 0; double lower = 0, upper = 0; float acc = 2.0f;
  int64_t chunk = 0;
  int64_t edge = 0;                 // that many quarter-size chunks open and close the batch (0: uniform chunks)
  int64_t *counter = nullptr;       // chunks handed out so far (process-local or shared between processes)
  std::atomic<int> failed{0};
  std::atomic<long long> launches{0};
  std::mutex err_mu; std::string err; int err_code = 0;
};

// device workspace of one chunk
struct ChunkBufs { size_t oP, oS, oC, oW, oI, oN, oRaw, oX, oFlag, total; };
ChunkBufs chunk_bufs(const HostQuery &q, int64_t c) {
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  ChunkBufs b;
  b.oP = 0; b.oS = b.oP + al(24 * c); b.oC = b.oS + al(8 * c); b.oW = b.oC + al(24 * c); b.oI = b.oW + al(8 * c);
  b.oN = b.oI + al(4 * c); b.oRaw = b.oN + al(q.N.p ? 24 * c : 0);
  b.oX = b.oRaw + al(24 * c);                                   // raw: one strided/float matrix or vector at a time
  b.oFlag = b.oX + al(iglb200_query_scratch_bytes(c));
  b.total = b.oFlag + 256;
  return b;
}
// pinned staging of one chunk: [raw input | S | I | C | W | N | flag]
struct StageBufs { size_t oIn, oS, oI, oC, oW, oN, oFlag, total; };
StageBufs stage_bufs(int64_t c) {
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  StageBufs b;
  b.oIn = 0; b.oS = al(24 * c); b.oI = b.oS + al(8 * c); b.oC = b.oI + al(4 * c); b.oW = b.oC + al(24 * c); b.oN = b.oW + al(8 * c);
  b.oFlag = b.oN + al(24 * c); b.total = b.oFlag + 256;
  return b;
}

// upload rows [lo, lo+c) of a host matrix into canonical column-major doubles dP[3][c]
void chunk_upload(WorkerCtx &w, const HostMat &M, int64_t lo, int64_t c, double *dP, char *draw, char *hin) {
  const size_t e = M.esz();
  if (M.kind == MAT_COLS) {
    for (int d = 0; d < 3; ++d) {
      const char *src = M.p + (size_t)(d * M.cs + lo) * e;
      char *dst = M.f64 ? (char *)(dP + d * c) : draw + (size_t)d * c * e;
      if (!M.pinned) { std::memcpy(hin + (size_t)d * c * e, src, (size_t)c * e); src = hin + (size_t)d * c * e; }
      plat::h2d(dst, src, (size_t)c * e, w.s);
    }
    if (!M.f64) plat::for_n(3 * c, FromFloatCols{(const float *)draw, dP}, w.s);
  } else {   // dense row-major rows are one contiguous byte range
    const char *src = M.p + (size_t)lo * 3 * e;
    if (!M.pinned) { std::memcpy(hin, src, (size_t)c * 3 * e); src = hin; }
    plat::h2d(draw, src, (size_t)c * 3 * e, w.s);
    if (M.f64) plat::for_n(c, ToCanon3<double>{(const double *)draw, 3, 1, c, dP}, w.s);
    else plat::for_n(c, ToCanon3<float>{(const float *)draw, 3, 1, c, dP}, w.s);
  }
}
// enqueue the download of canonical dC[3][c] into rows [lo, lo+c) of a host matrix (through `hst` when it is pageable);
// `draw` is device scratch for the layout change (consumed by the D2H copy that follows on the same stream)
void chunk_download3(WorkerCtx &w, const HostMat &M, int64_t lo, int64_t c, const double *dC, char *draw, char *hst) {
  const size_t e = M.esz();
  if (M.kind == MAT_COLS) {
    if (!M.f64) plat::for_n(3 * c, ToFloat3Cols{dC, (float *)draw}, w.s);
    for (int d = 0; d < 3; ++d) {
      const char *src = M.f64 ? (const char *)(dC + d * c) : draw + (size_t)d * c * e;
      char *dst = M.pinned ? M.p + (size_t)(d * M.cs + lo) * e : hst + (size_t)d * c * e;
      plat::d2h(dst, src, (size_t)c * e, w.s);
    }
  } else {
    if (M.f64) plat::for_n(c, FromCanon3<double>{dC, c, (double *)draw, 3, 1}, w.s);
    else plat::for_n(c, FromCanon3<float>{dC, c, (float *)draw, 3, 1}, w.s);
    char *dst = M.pinned ? M.p + (size_t)lo * 3 * e : hst;
    plat::d2h(dst, draw, (size_t)c * 3 * e, w.s);
  }
}
void chunk_unstage3(const HostMat &M, int64_t lo, int64_t c, const char *hst) {
  if (M.pinned) return;
  const size_t e = M.esz();
  if (M.kind == MAT_COLS) for (int d = 0; d < 3; ++d) std::memcpy(M.p + (size_t)(d * M.cs + lo) * e, hst + (size_t)d * c * e, (size_t)c * e);
  else std::memcpy(M.p + (size_t)lo * 3 * e, hst, (size_t)c * 3 * e);
}

int64_t take_chunk(int64_t *counter) { return __atomic_fetch_add(counter, (int64_t)1, __ATOMIC_RELAXED); }

// Chunk k of a batch of n rows.  The first upload and the last kernels + download of a batch overlap with nothing, so
// the batch opens and closes with `edge` chunks of a quarter of the size (one per worker): [small x edge | big ... | small x edge]
struct ChunkPlan {
  int64_t n, big, small, edge, mid, total;
  ChunkPlan(int64_t n_, int64_t chunk, int64_t edge_) : n(n_), big(chunk), small(std::max<int64_t>(((chunk / 4) + 31) & ~int64_t(31), 32)), edge(edge_) {
    if (edge <= 0 || 8 * 2 * edge * small > n) { edge = 0; mid = (n + big - 1) / big; }   // shape only when the edges are < 1/8 of the batch
    else mid = (n - 2 * edge * small + big - 1) / big;
    total = 2 * edge + mid;
  }
  bool range(int64_t k, int64_t &lo, int64_t &hi) const {
    if (k >= total) return false;
    if (k < edge) { lo = k * small; hi = lo + small; }
    else if (k < edge + mid) { lo = edge * small + (k - edge) * big; hi = std::min(lo + big, n - edge * small); }
    else { lo = n - (total - k) * small; hi = lo + small; }
    return true;
  }
};

// One worker's share of a host query: chunks until the counter runs past the end.
void host_worker(WorkerCtx &w, HostQuery &q) {
  const iglb200_mesh *m = q.m->on(w.device);
  if (!m) return;
  const ShareDeviceScope share;
  const ChunkPlan plan(q.n, q.chunk, q.edge);
  const ChunkBufs B = chunk_bufs(q, q.chunk);
  const StageBufs H = stage_bufs(q.chunk);
  plat::launch_counter() = 0;
  bool sized = false;
  // three device scratch areas for layout shims are needed at most (C, N raw at once are sequenced on one stream, so one suffices)
  while (!q.failed.load(std::memory_order_relaxed)) {
    const int64_t k = take_chunk(q.counter);
    int64_t lo = 0, hi = 0;
    if (!plan.range(k, lo, hi)) break;
    if (!sized) { w.ensure(B.total, H.total); sized = true; }
    const int64_t c = hi - lo;
    char *base = w.dbuf, *hb = w.hbuf;
    double *dP = (double *)(base + B.oP), *dS = (double *)(base + B.oS), *dC = (double *)(base + B.oC), *dW = (double *)(base + B.oW);
    double *dN = q.N.p ? (double *)(base + B.oN) : nullptr;
    int32_t *dI = (int32_t *)(base + B.oI);
    char *draw = base + B.oRaw;
    chunk_upload(w, q.P, lo, c, dP, draw, hb + H.oIn);
    if (q.mode == 0) run_signed_distance(m, dP, c, q.sign_type, q.lower, q.upper, dS, dI, dC, q.W.p ? dW : nullptr, dN, base + B.oX, 1, w.s);
    else if (q.mode == 1) run_squared_distance(m, dP, c, dS, dI, dC, base + B.oX, 1, w.s, q.lower, q.upper);
    else if (q.mode == 2) run_fwn(m, dP, c, q.acc, dW, base + B.oX, 1, w.s);
    else run_exact_winding(m, dP, c, dW, base + B.oX, 1, w.s);
    // downloads (each layout shim reuses `draw`: stream order keeps them apart)
    if (q.S.p) {
      const size_t e = q.S.esz();
      const char *src = (const char *)dS;
      if (!q.S.f64) { plat::for_n(c, ToFloat1{dS, (float *)draw}, w.s); src = draw; }
      plat::d2h(q.S.pinned ? q.S.p + (size_t)lo * e : hb + H.oS, src, (size_t)c * e, w.s);
    }
    if (q.I) plat::d2h(q.i_pinned ? (char *)(q.I + lo) : hb + H.oI, dI, (size_t)4 * c, w.s);
    if (q.C.p) chunk_download3(w, q.C, lo, c, dC, draw, hb + H.oC);
    if (q.W.p) {
      const size_t e = q.W.esz();
      const char *src = (const char *)dW;
      if (!q.W.f64) { plat::for_n(c, ToFloat1{dW, (float *)draw}, w.s); src = draw; }
      plat::d2h(q.W.pinned ? q.W.p + (size_t)lo * e : hb + H.oW, src, (size_t)c * e, w.s);
    }
    if (q.N.p && dN) chunk_download3(w, q.N, lo, c, dN, draw, hb + H.oN);
    plat::d2h(hb + H.oFlag, overflow_flag(), 4, w.s);
    plat::sync(w.s);
    if (*(const uint32_t *)(hb + H.oFlag)) { reset_overflow(w.s); throw std::runtime_error("traversal stack overflow (tree too deep)"); }
    if (q.S.p && !q.S.pinned) std::memcpy(q.S.p + (size_t)lo * q.S.esz(), hb + H.oS, (size_t)c * q.S.esz());
    if (q.I && !q.i_pinned) std::memcpy(q.I + lo, hb + H.oI, (size_t)4 * c);
    if (q.C.p) chunk_unstage3(q.C, lo, c, hb + H.oC);
    if (q.W.p && !q.W.pinned) std::memcpy(q.W.p + (size_t)lo * q.W.esz(), hb + H.oW, (size_t)c * q.W.esz());
    if (q.N.p && dN) chunk_unstage3(q.N, lo, c, hb + H.oN);
  }
  q.launches += plat::launch_counter();
}

// Worker threads, IGLB200_HOST_WORKERS per device, created on first use and never destroyed (the process's exit
// must not wait for, or free CUDA resources under, a thread).  One fork-join at a time.
struct Pool {
  struct W { std::thread t; WorkerCtx ctx; uint64_t seen = 0; };
  std::mutex mu, call_mu;
  std::condition_variable cv, cv_done;
  std::vector<W *> ws;
  uint64_t gen = 0;
  int pending = 0;
  HostQuery *job = nullptr;
  int per_dev = 0;   // 0: by the size of the group (measured on c2: pageable arrays want 6 copy threads on one GPU)
  Pool() {
    const char *e = std::getenv("IGLB200_HOST_WORKERS");
    const int v = e ? std::atoi(e) : 0;
    if (v > 0 && v <= 16) per_dev = v;
  }
  int workers_for(int ndev) const { return per_dev ? per_dev : (ndev <= 1 ? 6 : (ndev <= 2 ? 4 : 3)); }
  void loop(W *w) {
    try { plat::set_device(w->ctx.device); plat::pin_thread_near_device(w->ctx.device); } catch (...) {}
    std::unique_lock<std::mutex> lk(mu);
    while (true) {
      cv.wait(lk, [&] { return gen != w->seen; });
      w->seen = gen;
      HostQuery *q = job;
      if (!q || !q->m->on(w->ctx.device)) continue;   // not taking part (pending does not count it)
      lk.unlock();
      try { host_worker(w->ctx, *q); }
      catch (const std::exception &ex) {
        std::lock_guard<std::mutex> g(q->err_mu);
        if (!q->failed.exchange(1)) {
          q->err = ex.what();
          q->err_code = dynamic_cast<const std::invalid_argument *>(&ex) ? IGLB200_ERR_ARG
                        : dynamic_cast<const std::bad_alloc *>(&ex) ? IGLB200_ERR_OOM : IGLB200_ERR_CUDA;
        }
      } catch (...) {
        std::lock_guard<std::mutex> g(q->err_mu);
        if (!q->failed.exchange(1)) { q->err = "unknown error in a worker"; q->err_code = IGLB200_ERR_INTERNAL; }
      }
      lk.lock();
      if (--pending == 0) cv_done.notify_all();
    }
  }
  void run(HostQuery &q) {
    std::lock_guard<std::mutex> call(call_mu);
    std::unique_lock<std::mutex> lk(mu);
    // workers for every device that holds a copy
    const int maxdev = std::max<int>(q.m->device + 1, (int)q.m->reps.size());
    for (int d = 0; d < maxdev; ++d) {
      if (!q.m->on(d)) continue;
      int have = 0;
      for (W *w : ws) have += w->ctx.device == d;
      const int want = workers_for(q.m->n_devices());
      for (; have < want; ++have) {
        W *w = new W; w->ctx.device = d; w->seen = gen;
        ws.push_back(w);
        w->t = std::thread([this, w] { loop(w); });
        w->t.detach();
      }
    }
    int takers = 0;
    for (W *w : ws) takers += q.m->on(w->ctx.device) ? 1 : 0;
    job = &q; pending = takers; ++gen;
    cv.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
    job = nullptr;
  }
};
Pool &pool() { static Pool *p = new Pool; return *p; }

static int64_t env_i64(const char *name, int64_t dflt) { const char *e = std::getenv(name); const long long v = e ? std::atoll(e) : 0; return v > 0 ? (int64_t)v : dflt; }

// rows per chunk of a host query of n rows on ndev devices
int64_t host_chunk_rows(int64_t n, int ndev) {
  static const int64_t cfg = env_i64("IGLB200_STREAM_CHUNK", 0);
  if (cfg) return std::min(cfg, std::max<int64_t>(n, 1));
  // large enough for the persistent search kernel to run full (>= 256 Ki rows), small enough that a device sees at
  // least four chunks so that uploads, kernels and downloads overlap
  const int64_t want = (n + (int64_t)ndev * 4 - 1) / ((int64_t)ndev * 4);
  int64_t c = std::min<int64_t>(int64_t(1) << 20, std::max<int64_t>(int64_t(1) << 18, want));
  c = (c + 31) & ~int64_t(31);
  return std::min(c, std::max<int64_t>(n, 1));
}

void run_host_query(HostQuery &q, int64_t *shared_counter, int64_t shared_chunk) {
  if (q.n <= 0) return;
  if (q.P.kind == MAT_OTHER || q.C.kind == MAT_OTHER || q.N.kind == MAT_OTHER)
    throw std::invalid_argument("host matrices must be column-major (row stride 1) or dense row-major (strides 3, 1)");
  if (q.n > (int64_t)0x7FFFFFFF * 1024) throw std::invalid_argument("too many queries");
  int64_t local_counter = 0;
  const int ndev = q.m->n_devices();
  if (shared_counter) {
    if (shared_chunk <= 0) throw std::invalid_argument("chunk_rows must be positive");
    q.counter = shared_counter; q.chunk = std::min<int64_t>(shared_chunk, int64_t(1) << 24);
  } else {
    q.counter = &local_counter; q.chunk = host_chunk_rows(q.n, ndev);
    static const int64_t edge_off = env_i64("IGLB200_STREAM_EDGE_OFF", 0), edge_cfg = env_i64("IGLB200_STREAM_EDGE", 0);
    q.edge = edge_off ? 0 : (edge_cfg ? edge_cfg : (int64_t)ndev * pool().workers_for(ndev));
  }
  static const int64_t inline_rows = env_i64("IGLB200_INLINE_ROWS", 1 << 15);
  if (!shared_counter && ndev == 1 && q.n <= inline_rows) {
    // small call: on the calling thread, with its own per-device context (no fork-join latency)
    plat::DeviceGuard dg(q.m->device);
    WorkerCtx &w = plat::per_device<WorkerCtx>();
    w.device = q.m->device;
    host_worker(w, q);
  } else {
    pool().run(q);
    if (q.failed.load()) {
      if (q.err_code == IGLB200_ERR_ARG) throw std::invalid_argument(q.err);
      if (q.err_code == IGLB200_ERR_OOM) throw std::bad_alloc();
      throw std::runtime_error(q.err);
    }
  }
  plat::launch_counter() = (int)q.launches.load();
}

}  // namespace iglq
using namespace iglq;

extern "C" {

const char *iglb200_last_error(void) { return g_err.c_str(); }
const char *iglb200_version(void) {
#ifdef IGLB200_EMUL
  return "iglb200 0.1 (HOST EMULATION BUILD - tests only)";
#else
  return "iglb200 0.1 (sm_100a)";
#endif
}

int iglb200_device_count(void) {
#ifdef IGLB200_EMUL
  return 0;
#else
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) { cudaGetLastError(); return 0; }
  if (e != cudaSuccess) return fail(IGLB200_ERR_CUDA, cudaGetErrorString(e));
  return n;
#endif
}
int iglb200_set_device(int device) {
#ifdef IGLB200_EMUL
  (void)device; return IGLB200_OK;
#else
  API_BEGIN
  require_device();
  PLAT_CHECK(cudaSetDevice(device));
  return IGLB200_OK;
  API_END
#endif
}
// One process drives n_gpus devices: devices 0..n_gpus-1 become the process's group (group.h).  Meshes created
// afterwards are replicated to every device of the group (ncclBroadcast / P2P) and host-pointer queries on them are
// dealt to all of those devices.  n_gpus == 1 returns to single-device operation.
int iglb200_init(int n_gpus) {
  API_BEGIN
  if (n_gpus < 1) throw std::invalid_argument("n_gpus must be >= 1");
  require_device();
  grp::init(n_gpus);
  return IGLB200_OK;
  API_END
}
int iglb200_group_info(int64_t info[4], char *transport, uint64_t cap) {
  API_BEGIN
  grp::State &g = grp::state();
  if (info) { info[0] = g.n; info[1] = g.nccl ? 1 : 0; info[2] = g.peer ? 1 : 0; info[3] = (int64_t)(g.last_replicate_ms * 1000.0); }
  if (transport && cap) { std::strncpy(transport, g.transport.c_str(), (size_t)cap - 1); transport[cap - 1] = 0; }
  return IGLB200_OK;
  API_END
}
// waits for outstanding work of the calling thread's device; handles stay valid (they are destroyed one by one)
void iglb200_shutdown(void) {
#ifndef IGLB200_EMUL
  int n = 0;
  if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) cudaDeviceSynchronize();
  cudaGetLastError();
#endif
}
int iglb200_host_alloc(void **out, uint64_t bytes) {
  if (!out) return fail(IGLB200_ERR_ARG, "null out");
#ifdef IGLB200_EMUL
  *out = std::malloc(bytes ? bytes : 1); return *out ? IGLB200_OK : fail(IGLB200_ERR_OOM, "malloc");
#else
  API_BEGIN
  require_device();
  PLAT_CHECK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return IGLB200_OK;
  API_END
#endif
}
int iglb200_host_free(void *p) {
#ifdef IGLB200_EMUL
  std::free(p); return IGLB200_OK;
#else
  API_BEGIN
  if (p) PLAT_CHECK(cudaFreeHost(p));
  return IGLB200_OK;
  API_END
#endif
}

// page-lock memory the caller already owns (e.g. a shared-memory mapping) so that copies to/from it are true DMA
int iglb200_host_register(void *p, uint64_t bytes) {
#ifdef IGLB200_EMUL
  (void)p; (void)bytes; return IGLB200_OK;
#else
  API_BEGIN
  require_device();
  if (!p) throw std::invalid_argument("null pointer");
  PLAT_CHECK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return IGLB200_OK;
  API_END
#endif
}
int iglb200_host_unregister(void *p) {
#ifdef IGLB200_EMUL
  (void)p; return IGLB200_OK;
#else
  API_BEGIN
  if (p) PLAT_CHECK(cudaHostUnregister(p));
  return IGLB200_OK;
  API_END
#endif
}

int iglb200_mesh_create(const void *V, int v_is_f64, int64_t nV, int64_t v_rs, int64_t v_cs, const int32_t *F, int64_t nF,
                        int64_t f_rs, int64_t f_cs, int fwn_order, iglb200_mesh **out) {
  return iglb200_mesh_create_ex(V, v_is_f64, nV, v_rs, v_cs, F, nF, f_rs, f_cs, fwn_order, 0, out);
}

int iglb200_mesh_create_ex(const void *V, int v_is_f64, int64_t nV, int64_t v_rs, int64_t v_cs, const int32_t *F, int64_t nF,
                           int64_t f_rs, int64_t f_cs, int fwn_order, int flags, iglb200_mesh **out) {
  API_BEGIN
  if (!out) throw std::invalid_argument("null out");
  *out = nullptr;
  if (nV < 0 || nF < 0 || (nV > 0 && !V) || (nF > 0 && !F)) throw std::invalid_argument("null or negative-sized input");
  if (nF > 0 && nV == 0) throw std::invalid_argument("faces without vertices");
  if ((nV > 0 && (v_rs <= 0 || v_cs <= 0)) || (nF > 0 && (f_rs <= 0 || f_cs <= 0))) throw std::invalid_argument("strides must be positive");
  if (fwn_order < 0 || fwn_order > 2) throw std::invalid_argument("fwn_order must be 0, 1 or 2");
  if (nF >= (int64_t)0x7FFFFFF0) throw std::invalid_argument("too many faces");
  require_device();
  plat::launch_counter() = 0;
  Timer t;
  t.start(nullptr);
  iglb200_mesh *m = fwn::create_mesh(V, v_is_f64, nV, v_rs, v_cs, F, nF, f_rs, f_cs, fwn_order, flags, nullptr);
  m->build_ms = t.stop(nullptr);
  m->device = plat::current_device();
  *out = m;
  if (grp::size() > 1 && m->device < grp::size()) {
    try { grp::replicate(m); }
    catch (...) { iglb200_mesh_destroy(m); *out = nullptr; throw; }
  }
  return IGLB200_OK;
  API_END
}

int iglb200_mesh_replicate(iglb200_mesh *m) {
  API_BEGIN
  check_mesh(m);
  require_device();
  grp::replicate(m);
  return IGLB200_OK;
  API_END
}

int iglb200_mesh_destroy(iglb200_mesh *m) {
  API_BEGIN
  if (!m) return IGLB200_OK;
  check_mesh(m);
  for (iglb200_mesh *r : m->reps) {
    if (!r) continue;
    if (r->owns && r->arena) plat::dev_free(r->arena);
    delete r;
  }
  if (m->owns && m->arena) plat::dev_free(m->arena);
  m->h.magic = 0;
  delete m;
  return IGLB200_OK;
  API_END
}

int iglb200_mesh_info(const iglb200_mesh *m, int64_t info[8]) {
  API_BEGIN
  check_mesh(m);
  info[0] = m->h.nF; info[1] = m->h.nnodes; info[2] = m->h.nlevels; info[3] = (int64_t)m->h.bytes; info[4] = m->h.nexcluded;
  info[5] = m->h.order; info[6] = m->h.off_pn ? 1 : 0; info[7] = m->n_devices();
  return IGLB200_OK;
  API_END
}
double iglb200_mesh_build_ms(const iglb200_mesh *m) { return m ? m->build_ms : 0.0; }

int iglb200_mesh_arena(const iglb200_mesh *m, void **device_ptr, uint64_t *bytes) {
  API_BEGIN
  check_mesh(m);
  if (device_ptr) *device_ptr = m->arena;
  if (bytes) *bytes = m->h.bytes;
  return IGLB200_OK;
  API_END
}

int iglb200_mesh_arena_download(const iglb200_mesh *m, void *host_dst, uint64_t bytes) {
  API_BEGIN
  check_mesh(m);
  if (!host_dst || bytes != m->h.bytes) throw std::invalid_argument("destination size must equal the arena size");
  plat::d2h(host_dst, m->arena, bytes, nullptr);
  plat::sync(nullptr);
  return IGLB200_OK;
  API_END
}

int iglb200_mesh_from_arena(const void *device_ptr, uint64_t bytes, int copy, iglb200_mesh **out) {
  API_BEGIN
  if (!out || !device_ptr || bytes < sizeof(ArenaHeader)) throw std::invalid_argument("bad arena");
  require_device();
  iglb200_mesh *m = new iglb200_mesh;
  plat::d2h(&m->h, device_ptr, sizeof(ArenaHeader), nullptr);
  plat::sync(nullptr);
  if (m->h.magic != IGLB200_ARENA_MAGIC || m->h.bytes != bytes) { delete m; throw std::invalid_argument("arena header mismatch"); }
  if (copy) {
    m->arena = (char *)plat::dev_alloc(bytes);
    plat::d2d(m->arena, device_ptr, bytes, nullptr);
    plat::sync(nullptr);
    m->owns = true;
  } else {
    m->arena = (char *)device_ptr;
    m->owns = false;
  }
  m->device = plat::current_device();
  *out = m;
  return IGLB200_OK;
  API_END
}

uint64_t iglb200_query_scratch_bytes(int64_t nP) {
  if (nP <= 0) return 256;
  const size_t a = ((size_t)nP * 4 + 255) & ~size_t(255);
  return 4 * a + ((plat::sort_pairs_temp_bytes(nP) + 255) & ~size_t(255)) + 8 * (size_t)SCHED_WORDS + 256;
}


// ---- sharded device-resident queries + inter-process memory
uint64_t iglb200_shard_scratch_bytes(int64_t nP) { return ShardLayout(nP < 0 ? 0 : nP).total; }
int64_t iglb200_shard_super_rows(int64_t nP, int n_shards) { return shard_super_rows(nP, n_shards); }
int iglb200_shard_begin_dev(const iglb200_mesh *m, const double *P, int64_t nP, void *shard_scratch, int sort_queries, int64_t super_rows,
                            void *stream) {
  API_BEGIN
  m = mesh_here(m);
  if (nP < 0 || !shard_scratch || (nP > 0 && !P)) throw std::invalid_argument("null pointer");
  require_device();
  check_count(nP);
  plat::launch_counter() = 0;
  shard_begin(m, P, nP, (char *)shard_scratch, sort_queries, super_rows, (stream_t)stream);
  return IGLB200_OK;
  API_END
}
int iglb200_shard_run_dev(const iglb200_mesh *m, const double *P, int64_t nP, void *shard_scratch, int mode, int sign_type, double lower,
                          double upper, float accuracy_scale, int sort_queries, double *S, int32_t *I, double *C, double *W,
                          int64_t super_rows, int64_t max_supers, int shard_index, int n_shards, void *stream) {
  API_BEGIN
  m = mesh_here(m);
  if (nP < 0 || !shard_scratch || (nP > 0 && !P)) throw std::invalid_argument("null pointer");
  if (mode < 0 || mode > 2) throw std::invalid_argument("mode must be 0 (signed distance), 1 (squared distance) or 2 (fast winding number)");
  if (nP > 0 && ((mode != 2 && (!S || !I || !C)) || (mode == 2 && !W))) throw std::invalid_argument("null output");
  require_device();
  check_count(nP);
  plat::launch_counter() = 0;
  shard_run(m, P, nP, (char *)shard_scratch, mode, sign_type, lower, upper, accuracy_scale, sort_queries, S, I, C, W, super_rows, max_supers,
            shard_index, n_shards, (stream_t)stream);
  return IGLB200_OK;
  API_END
}
int iglb200_shard_finish_dev(const iglb200_mesh *m, int64_t nP, void *shard_scratch, int mode, int sign_type, double lower, double upper,
                             int sort_queries, double *S, int32_t *I, double *C, double *W, void *stream) {
  API_BEGIN
  m = mesh_here(m);
  if (nP < 0 || !shard_scratch) throw std::invalid_argument("null pointer or negative size");
  if (mode < 0 || mode > 2) throw std::invalid_argument("mode must be 0, 1 or 2");
  if (nP > 0 && ((mode != 2 && (!S || !I || !C)) || (mode == 2 && !W))) throw std::invalid_argument("null output");
  require_device();
  plat::launch_counter() = 0;
  shard_finish(m, nP, (char *)shard_scratch, mode, sign_type, lower, upper, sort_queries, S, I, C, W, (stream_t)stream);
  return IGLB200_OK;
  API_END
}
// device memory another process of the same node can map (cudaIpc*): the owner's shard scratch under torchrun
int iglb200_ipc_alloc(uint64_t bytes, void **device_ptr, unsigned char handle[64]) {
  API_BEGIN
  if (!device_ptr || !handle) throw std::invalid_argument("null pointer");
  require_device();
  *device_ptr = plat::dev_alloc((size_t)bytes);
#ifdef IGLB200_EMUL
  std::memset(handle, 0, 64); std::memcpy(handle, device_ptr, sizeof(void *));
#else
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, *device_ptr);
  if (e != cudaSuccess) { plat::dev_free(*device_ptr); *device_ptr = nullptr; plat::check(e, "cudaIpcGetMemHandle"); }
  std::memcpy(handle, &h, 64);
#endif
  return IGLB200_OK;
  API_END
}
int iglb200_ipc_open(const unsigned char handle[64], void **device_ptr) {
  API_BEGIN
  if (!device_ptr || !handle) throw std::invalid_argument("null pointer");
  require_device();
#ifdef IGLB200_EMUL
  std::memcpy(device_ptr, handle, sizeof(void *));
#else
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  PLAT_CHECK(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
#endif
  return IGLB200_OK;
  API_END
}
int iglb200_ipc_close(void *device_ptr) {
  API_BEGIN
#ifndef IGLB200_EMUL
  if (device_ptr) PLAT_CHECK(cudaIpcCloseMemHandle(device_ptr));
#else
  (void)device_ptr;
#endif
  return IGLB200_OK;
  API_END
}
int iglb200_ipc_free(void *device_ptr) {
  API_BEGIN
  plat::dev_free(device_ptr);
  return IGLB200_OK;
  API_END
}

int iglb200_last_launch_count(void) { return plat::launch_counter(); }
#ifndef IGLB200_EMUL
// FP32 FMA micro-benchmark: 8 independent chains per thread, 4096 FMAs each
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, float a, float b) {
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 1e-3f + k;
  for (int it = 0; it < 4096; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = fmaf(x[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  if (s == 123.456f) out[0] = s;   // never true; keeps the loop alive
}
#endif
double iglb200_debug_fp32_peak_tflops(void) {
#ifdef IGLB200_EMUL
  return 0.0;
#else
  try {
    require_device();
    float *d = (float *)plat::dev_alloc(4);
    const int blocks = plat::sm_count() * 16;
    fp32_peak_kernel<<<blocks, 256>>>(d, 1.0000001f, 1e-7f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int r = 0; r < 10; ++r) fp32_peak_kernel<<<blocks, 256>>>(d, 1.0000001f, 1e-7f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    plat::dev_free(d);
    const double flops = 10.0 * (double)blocks * 256.0 * 8.0 * 4096.0 * 2.0;
    return flops / (ms * 1e-3) / 1e12;
  } catch (...) { return -1.0; }
#endif
}
#ifdef CP_TIMELINE
int iglb200_debug_timeline(unsigned long long *buf) {
  return (int)cudaMemcpyToSymbol(g_timeline, &buf, sizeof(buf));
}
#endif
int iglb200_poll_device_error(void) {
  API_BEGIN
  require_device();
  throw_if_overflow(nullptr);
  return IGLB200_OK;
  API_END
}

int iglb200_debug_visit_counts(const iglb200_mesh *m, const double *P_host_colmajor, int64_t n, double out[4]) {
  API_BEGIN
  m = mesh_here(m);
  if (n <= 0 || !P_host_colmajor || !out || m->h.nnodes == 0) throw std::invalid_argument("bad arguments");
  require_device();
  DevBuf<double> dP((size_t)3 * n), dD((size_t)n), dC((size_t)3 * n);
  DevBuf<int32_t> dI((size_t)n);
  DevBuf<uint32_t> acc(4);
  DevBuf<char> scratch(iglb200_query_scratch_bytes(n));
  plat::h2d(dP.p, P_host_colmajor, (size_t)24 * n, nullptr);
  plat::dev_memset(acc.p, 0, 16, nullptr);
  const uint32_t *perm = morton_perm(m, dP.p, n, scratch.p, 1, nullptr);
  launch_closest(closest_args(m, dP.p, n, perm, INFINITY, 0.0, dD.p, dI.p, dC.p, acc.p, carve(scratch.p, n).sched), nullptr);
  plat::for_n(n, CountWinding{m->fnodes(), m->tri32(), dP.p, n, m->h.order, acc.p + 2, overflow_flag()}, nullptr);
  uint32_t h[4];
  plat::d2h(h, acc.p, 16, nullptr);
  plat::sync(nullptr);
  // closest: per WARP (32 queries share a walk); winding: per query
  out[0] = (double)h[0] / (double)((n + 31) / 32); out[1] = (double)h[1] / (double)((n + 31) / 32);
  out[2] = (double)h[2] / (double)n; out[3] = (double)h[3] / (double)n;
  return IGLB200_OK;
  API_END
}
void iglb200_profile_enable(int on) { plat::prof_enable(on != 0); }
int iglb200_profile_collect(char *out, uint64_t cap) { return plat::prof_collect(out, (size_t)cap); }

int iglb200_signed_distance_dev(const iglb200_mesh *m, const double *P, int64_t nP, int sign_type, double lower, double upper,
                                double *S, int32_t *I, double *C, double *W, double *N, void *scratch, int sort_queries, void *stream) {
  API_BEGIN
  m = mesh_here(m);
  if (nP < 0 || (nP > 0 && (!P || !S || !I || !C))) throw std::invalid_argument("null pointer");
  require_device();
  check_count(nP);
  plat::launch_counter() = 0;
  run_signed_distance(m, P, nP, sign_type, lower, upper, S, I, C, W, N, scratch, sort_queries, (stream_t)stream);
  return IGLB200_OK;
  API_END
}
int iglb200_squared_distance_dev(const iglb200_mesh *m, const double *P, int64_t nP, double *sqrD, int32_t *I, double *C,
                                 void *scratch, int sort_queries, void *stream) {
  API_BEGIN
  m = mesh_here(m);
  if (nP < 0 || (nP > 0 && (!P || !sqrD || !I || !C))) throw std::invalid_argument("null pointer");
  require_device();
  check_count(nP);
  plat::launch_counter() = 0;
  run_squared_distance(m, P, nP, sqrD, I, C, scratch, sort_queries, (stream_t)stream);
  return IGLB200_OK;
  API_END
}
int iglb200_fwn_eval_dev(const iglb200_mesh *m, const double *Q, int64_t nQ, float accuracy_scale, double *W, void *scratch,
                         int sort_queries, void *stream) {
  API_BEGIN
  m = mesh_here(m);
  if (nQ < 0 || (nQ > 0 && (!Q || !W))) throw std::invalid_argument("null pointer");
  require_device();
  check_count(nQ);
  plat::launch_counter() = 0;
  run_fwn(m, Q, nQ, accuracy_scale, W, scratch, sort_queries, (stream_t)stream);
  return IGLB200_OK;
  API_END
}

// whole-batch path for layouts the chunk pipeline does not take (arbitrary strides): synchronous, per-call allocations
static void slow_signed_distance(const iglb200_mesh *m, const void *P, int p_is_f64, int64_t nP, int64_t p_rs, int64_t p_cs,
                                 int sign_type, double lower, double upper, void *S, int s_is_f64, int32_t *I, void *C, int c_is_f64,
                                 int64_t c_rs, int64_t c_cs, void *W, int w_is_f64, void *N, int n_is_f64, int64_t n_rs, int64_t n_cs) {
  check_count(nP);
  plat::DeviceGuard dg(m->device);
  stream_t s = nullptr;
  DevBuf<double> dP, dS((size_t)nP), dC((size_t)3 * nP), dW(W ? (size_t)nP : 0), dN(N ? (size_t)3 * nP : 0);
  DevBuf<int32_t> dI((size_t)nP);
  DevBuf<char> scratch(iglb200_query_scratch_bytes(nP));
  upload_points(P, p_is_f64, nP, p_rs, p_cs, dP, s);
  run_signed_distance(m, dP.p, nP, sign_type, lower, upper, dS.p, dI.p, dC.p, W ? dW.p : nullptr, N ? dN.p : nullptr, scratch.p, 1, s);
  if (N) download_points(dN, nP, N, n_is_f64, n_rs, n_cs, s);
  download_scalar(dS, nP, S, s_is_f64, s);
  plat::d2h(I, dI.p, (size_t)4 * nP, s);
  download_points(dC, nP, C, c_is_f64, c_rs, c_cs, s);
  if (W) download_scalar(dW, nP, W, w_is_f64, s);
  throw_if_overflow(s);
}

static int signed_distance_host(const iglb200_mesh *m, const void *P, int p_is_f64, int64_t nP, int64_t p_rs, int64_t p_cs,
                                int sign_type, double lower, double upper, void *S, int s_is_f64, int32_t *I, void *C, int c_is_f64,
                                int64_t c_rs, int64_t c_cs, void *W, int w_is_f64, void *N, int n_is_f64, int64_t n_rs, int64_t n_cs,
                                int64_t *shared_counter, int64_t shared_chunk) {
  API_BEGIN
  check_mesh(m);
  if (nP < 0 || (nP > 0 && (!P || !S || !I || !C))) throw std::invalid_argument("null pointer");
  if (nP > 0 && (c_rs <= 0 || c_cs <= 0 || p_rs <= 0 || p_cs <= 0)) throw std::invalid_argument("strides must be positive");
  if (N && nP > 0 && (n_rs <= 0 || n_cs <= 0)) throw std::invalid_argument("strides must be positive");
  if (sign_type < 0 || sign_type > IGLB200_SIGN_FAST_WINDING_NUMBER) throw std::domain_error("unknown signed distance type");
  require_device();
  plat::launch_counter() = 0;
  // N is an output of PSEUDONORMAL only (signed_distance.cpp:125-131); with #F == 0 it is NaN for every type (:403)
  if (sign_type != IGLB200_SIGN_PSEUDONORMAL && m->h.nF != 0) N = nullptr;
  HostQuery q;
  q.m = m; q.mode = 0; q.n = nP; q.sign_type = sign_type; q.lower = lower; q.upper = upper;
  q.P = host_mat(P, p_is_f64, nP, p_rs, p_cs); q.C = host_mat(C, c_is_f64, nP, c_rs, c_cs); q.N = host_mat(N, n_is_f64, nP, n_rs, n_cs);
  q.S = host_vec(S, s_is_f64); q.W = host_vec(W, w_is_f64); q.I = I; q.i_pinned = nP > 0 && plat::host_is_pinned(I);
  if (nP > 0 && (q.P.kind == MAT_OTHER || q.C.kind == MAT_OTHER || q.N.kind == MAT_OTHER)) {
    if (shared_counter) throw std::invalid_argument("cooperative calls need column-major or dense row-major matrices");
    slow_signed_distance(m, P, p_is_f64, nP, p_rs, p_cs, sign_type, lower, upper, S, s_is_f64, I, C, c_is_f64, c_rs, c_cs, W, w_is_f64,
                         N, n_is_f64, n_rs, n_cs);
    return IGLB200_OK;
  }
  run_host_query(q, shared_counter, shared_chunk);
  return IGLB200_OK;
  API_END
}

int iglb200_signed_distance(const iglb200_mesh *m, const void *P, int p_is_f64, int64_t nP, int64_t p_rs, int64_t p_cs,
                            int sign_type, double lower, double upper, void *S, int s_is_f64, int32_t *I, void *C, int c_is_f64,
                            int64_t c_rs, int64_t c_cs, void *W, int w_is_f64, void *N, int n_is_f64, int64_t n_rs, int64_t n_cs) {
  return signed_distance_host(m, P, p_is_f64, nP, p_rs, p_cs, sign_type, lower, upper, S, s_is_f64, I, C, c_is_f64, c_rs, c_cs, W, w_is_f64,
                              N, n_is_f64, n_rs, n_cs, nullptr, 0);
}
int iglb200_signed_distance_coop(const iglb200_mesh *m, const void *P, int p_is_f64, int64_t nP, int64_t p_rs, int64_t p_cs,
                                 int sign_type, double lower, double upper, void *S, int s_is_f64, int32_t *I, void *C, int c_is_f64,
                                 int64_t c_rs, int64_t c_cs, void *W, int w_is_f64, void *N, int n_is_f64, int64_t n_rs, int64_t n_cs,
                                 int64_t *shared_chunk_counter, int64_t chunk_rows) {
  if (!shared_chunk_counter) return fail(IGLB200_ERR_ARG, "null chunk counter");
  return signed_distance_host(m, P, p_is_f64, nP, p_rs, p_cs, sign_type, lower, upper, S, s_is_f64, I, C, c_is_f64, c_rs, c_cs, W, w_is_f64,
                              N, n_is_f64, n_rs, n_cs, shared_chunk_counter, chunk_rows);
}
int64_t iglb200_host_chunk_rows(int64_t nP, int n_devices) { return host_chunk_rows(nP, n_devices < 1 ? 1 : n_devices); }

int iglb200_squared_distance(const iglb200_mesh *m, const void *P, int p_is_f64, int64_t nP, int64_t p_rs, int64_t p_cs,
                             void *sqrD, int d_is_f64, int32_t *I, void *C, int c_is_f64, int64_t c_rs, int64_t c_cs) {
  return iglb200_squared_distance_bounded(m, P, p_is_f64, nP, p_rs, p_cs, 0.0, INFINITY, sqrD, d_is_f64, I, C, c_is_f64, c_rs, c_cs);
}
int iglb200_squared_distance_bounded(const iglb200_mesh *m, const void *P, int p_is_f64, int64_t nP, int64_t p_rs, int64_t p_cs,
                                     double low_sqr_d, double up_sqr_d, void *sqrD, int d_is_f64, int32_t *I, void *C, int c_is_f64,
                                     int64_t c_rs, int64_t c_cs) {
  API_BEGIN
  check_mesh(m);
  if (nP < 0 || (nP > 0 && (!P || !sqrD || !I || !C))) throw std::invalid_argument("null pointer");
  if (nP > 0 && (c_rs <= 0 || c_cs <= 0 || p_rs <= 0 || p_cs <= 0)) throw std::invalid_argument("strides must be positive");
  require_device();
  plat::launch_counter() = 0;
  if (low_sqr_d != low_sqr_d || up_sqr_d != up_sqr_d) throw std::invalid_argument("NaN bound");
  HostQuery q;
  q.m = m; q.mode = 1; q.n = nP; q.lower = low_sqr_d; q.upper = up_sqr_d;   // squared bounds in mode 1
  q.P = host_mat(P, p_is_f64, nP, p_rs, p_cs); q.C = host_mat(C, c_is_f64, nP, c_rs, c_cs);
  q.S = host_vec(sqrD, d_is_f64); q.I = I; q.i_pinned = nP > 0 && plat::host_is_pinned(I);
  if (nP > 0 && (q.P.kind == MAT_OTHER || q.C.kind == MAT_OTHER)) {
    check_count(nP);
    plat::DeviceGuard dg(m->device);
    stream_t s = nullptr;
    DevBuf<double> dP, dS((size_t)nP), dC((size_t)3 * nP);
    DevBuf<int32_t> dI((size_t)nP);
    DevBuf<char> scratch(iglb200_query_scratch_bytes(nP));
    upload_points(P, p_is_f64, nP, p_rs, p_cs, dP, s);
    run_squared_distance(m, dP.p, nP, dS.p, dI.p, dC.p, scratch.p, 1, s, low_sqr_d, up_sqr_d);
    download_scalar(dS, nP, sqrD, d_is_f64, s);
    plat::d2h(I, dI.p, (size_t)4 * nP, s);
    download_points(dC, nP, C, c_is_f64, c_rs, c_cs, s);
    throw_if_overflow(s);
    return IGLB200_OK;
  }
  run_host_query(q, nullptr, 0);
  return IGLB200_OK;
  API_END
}

static int winding_host(const iglb200_mesh *m, int mode, const void *Q, int q_is_f64, int64_t nQ, int64_t q_rs, int64_t q_cs,
                        float accuracy_scale, void *W, int w_is_f64) {
  API_BEGIN
  check_mesh(m);
  if (nQ < 0 || (nQ > 0 && (!Q || !W))) throw std::invalid_argument("null pointer");
  if (nQ > 0 && (q_rs <= 0 || q_cs <= 0)) throw std::invalid_argument("strides must be positive");
  require_device();
  plat::launch_counter() = 0;
  HostQuery q;
  q.m = m; q.mode = mode; q.n = nQ; q.acc = accuracy_scale;
  q.P = host_mat(Q, q_is_f64, nQ, q_rs, q_cs); q.W = host_vec(W, w_is_f64);
  if (nQ > 0 && q.P.kind == MAT_OTHER) {
    check_count(nQ);
    plat::DeviceGuard dg(m->device);
    stream_t s = nullptr;
    DevBuf<double> dQ, dW((size_t)nQ);
    DevBuf<char> scratch(iglb200_query_scratch_bytes(nQ));
    upload_points(Q, q_is_f64, nQ, q_rs, q_cs, dQ, s);
    if (mode == 2) run_fwn(m, dQ.p, nQ, accuracy_scale, dW.p, scratch.p, 1, s);
    else run_exact_winding(m, dQ.p, nQ, dW.p, scratch.p, 1, s);
    download_scalar(dW, nQ, W, w_is_f64, s);
    plat::sync(s);
    throw_if_overflow(s);
    return IGLB200_OK;
  }
  run_host_query(q, nullptr, 0);
  return IGLB200_OK;
  API_END
}
int iglb200_fwn_eval(const iglb200_mesh *m, const void *Q, int q_is_f64, int64_t nQ, int64_t q_rs, int64_t q_cs,
                     float accuracy_scale, void *W, int w_is_f64) {
  return winding_host(m, 2, Q, q_is_f64, nQ, q_rs, q_cs, accuracy_scale, W, w_is_f64);
}

// igl::marching_cubes(S, GV, GI, isovalue, V, F[, E2V]) (marching_cubes.cpp:114-176): cubes given by their 8 corner ids
int iglb200_marching_cubes_sparse(const double *S, const double *GV, int64_t n, int64_t gv_rs, int64_t gv_cs,
                                  const int32_t *GI, int64_t n_cubes, int64_t gi_rs, int64_t gi_cs, double isovalue,
                                  iglb200_surface **out) {
  API_BEGIN
  if (!out || n < 0 || n_cubes < 0 || (n > 0 && (!S || !GV)) || (n_cubes > 0 && !GI)) throw std::invalid_argument("null pointer or negative size");
  if (n > 0 && (gv_rs <= 0 || gv_cs <= 0)) throw std::invalid_argument("strides must be positive");
  if (n_cubes > 0 && (gi_rs <= 0 || gi_cs <= 0)) throw std::invalid_argument("strides must be positive");
  if (n > 0x7FFFFFFF) throw std::invalid_argument("more than 2^31-1 grid vertices");
  require_device();
  DevBuf<double> dS((size_t)n + 1), dGV((size_t)3 * n + 1);
  DevBuf<int32_t> dGI;
  size_t gspan = 0;
  if (n > 0) {
    plat::h2d(dS.p, S, (size_t)8 * n, nullptr);
    const size_t span = (size_t)((n - 1) * gv_rs + 2 * gv_cs + 1);
    DevBuf<double> raw(span);
    plat::h2d(raw.p, GV, span * 8, nullptr);
    plat::for_n(n, ToCanon3<double>{raw.p, gv_rs, gv_cs, n, dGV.p}, nullptr);
    plat::sync(nullptr);
  }
  if (n_cubes > 0) {
    gspan = (size_t)((n_cubes - 1) * gi_rs + 7 * gi_cs + 1);
    dGI.alloc(gspan);
    plat::h2d(dGI.p, GI, gspan * 4, nullptr);
  }
  *out = grid::marching_cubes_sparse(dS.p, dGV.p, n, dGI.p, gi_rs, gi_cs, n_cubes, isovalue, nullptr);
  return IGLB200_OK;
  API_END
}
// igl::winding_number(V,F,O,W) (winding_number.cpp:17-52): exact, hierarchical, double precision
int iglb200_winding_number(const iglb200_mesh *m, const void *O, int o_is_f64, int64_t nO, int64_t o_rs, int64_t o_cs,
                           void *W, int w_is_f64) {
  return winding_host(m, 3, O, o_is_f64, nO, o_rs, o_cs, 0.f, W, w_is_f64);
}
int iglb200_winding_number_dev(const iglb200_mesh *m, const double *O, int64_t nO, double *W, void *scratch, int sort_queries,
                               void *stream) {
  API_BEGIN
  m = mesh_here(m);
  if (nO < 0 || (nO > 0 && (!O || !W))) throw std::invalid_argument("null pointer");
  require_device();
  check_count(nO);
  plat::launch_counter() = 0;
  run_exact_winding(m, O, nO, W, scratch, sort_queries, (stream_t)stream);
  return IGLB200_OK;
  API_END
}

// igl::iterative_closest_point(VX,FX,VY,FY,num_samples,max_iters,R,t) (iterative_closest_point.cpp:40-130); Y is the mesh handle
int iglb200_iterative_closest_point(const iglb200_mesh *meshY, const void *VX, int vx_is_f64, int64_t nVX, int64_t vx_rs, int64_t vx_cs,
                                    const int32_t *FX, int64_t nFX, int64_t fx_rs, int64_t fx_cs, int num_samples, int max_iters,
                                    uint64_t seed, double R[9], double t[3]) {
  API_BEGIN
  check_mesh(meshY);
  if (!VX || !FX || !R || !t || nVX <= 0 || nFX <= 0) throw std::invalid_argument("null pointer or empty mesh X");
  if (vx_rs <= 0 || vx_cs <= 0 || fx_rs <= 0 || fx_cs <= 0) throw std::invalid_argument("strides must be positive");
  if (num_samples < 1 || max_iters < 0) throw std::invalid_argument("num_samples must be >= 1 and max_iters >= 0");
  if (meshY->h.nnodes == 0) throw std::invalid_argument("mesh Y has no (finite) faces");
  require_device();
  plat::DeviceGuard dg(meshY->device);
  plat::launch_counter() = 0;
  stream_t s = nullptr;
  const int64_t ns = num_samples;
  DevBuf<double> dVX, dVXRT((size_t)3 * nVX), dblA((size_t)nFX), cum((size_t)nFX), X((size_t)3 * ns), Pc((size_t)3 * ns), N((size_t)3 * ns),
      sqrD((size_t)ns), sys(icp::NSYS);
  DevBuf<int32_t> dF((size_t)3 * nFX), I((size_t)ns);
  DevBuf<char> scratch(iglb200_query_scratch_bytes(ns));
  upload_points(VX, vx_is_f64, nVX, vx_rs, vx_cs, dVX, s);
  {
    std::vector<int32_t> hf((size_t)3 * nFX);
    for (int64_t f = 0; f < nFX; ++f) for (int k = 0; k < 3; ++k) {
      const int32_t v = FX[f * fx_rs + k * fx_cs];
      if (v < 0 || v >= nVX) throw std::invalid_argument("FX contains a vertex index outside [0, #VX)");
      hf[(size_t)3 * f + k] = v;
    }
    plat::h2d(dF.p, hf.data(), hf.size() * 4, s);
    plat::sync(s);
  }
  double Rr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tt[3] = {0, 0, 0};
  for (int iter = 0; iter < max_iters; ++iter) {
    icp::Transform tf{dVX.p, nVX, {Rr[0], Rr[1], Rr[2], Rr[3], Rr[4], Rr[5], Rr[6], Rr[7], Rr[8]}, {tt[0], tt[1], tt[2]}, dVXRT.p};
    plat::for_n(nVX, tf, s);
    plat::for_n(nFX, icp::FaceAreas{dVXRT.p, nVX, dF.p, nFX, dblA.p}, s);
    plat::inclusive_scan_f64(dblA.p, cum.p, nFX, s);
    plat::for_n(ns, icp::Sample{dVXRT.p, nVX, dF.p, nFX, cum.p, seed, (uint64_t)iter, ns, X.p}, s);
    run_squared_distance(meshY, X.p, ns, sqrD.p, I.p, Pc.p, scratch.p, 1, s);
    plat::for_n(ns, icp::Normals{meshY->tri64(), I.p, ns, N.p}, s);
    double Rup[9], tup[3];
    icp::rigid_alignment(X.p, Pc.p, N.p, ns, sys.p, Rup, tup, s);
    const double tn[3] = {tt[0] * Rup[0] + tt[1] * Rup[3] + tt[2] * Rup[6] + tup[0], tt[0] * Rup[1] + tt[1] * Rup[4] + tt[2] * Rup[7] + tup[1],
                          tt[0] * Rup[2] + tt[1] * Rup[5] + tt[2] * Rup[8] + tup[2]};
    icp::mat3_mul(Rr, Rup, Rr);
    tt[0] = tn[0]; tt[1] = tn[1]; tt[2] = tn[2];
  }
  plat::sync(s);
  throw_if_overflow(s);
  for (int k = 0; k < 9; ++k) R[k] = Rr[k];
  for (int k = 0; k < 3; ++k) t[k] = tt[k];
  return IGLB200_OK;
  API_END
}

// ---- point cloud ----------------------------------------------------------
// ---------------------------------------------------------------- voxel grids (grid.h)
namespace {
void check_grid(const iglb200_grid *g) { if (!g) throw std::invalid_argument("null grid handle"); }
void check_surface(const iglb200_surface *s) { if (!s) throw std::invalid_argument("null surface handle"); }
struct ColsToStrided { const double *src; int64_t n; double *dst; int64_t rs, cs; DEVFN void operator()(int64_t i) const { for (int j = 0; j < 3; ++j) dst[i * rs + j * cs] = src[j * n + i]; } };
}  // namespace

int iglb200_grid_create(const double box_min[3], const double box_max[3], int s, int pad_count, iglb200_grid **out) {
  API_BEGIN
  if (!box_min || !box_max || !out) throw std::invalid_argument("null pointer");
  require_device();
  *out = grid::create(box_min, box_max, s, pad_count, nullptr);
  return IGLB200_OK;
  API_END
}
static iglb200_grid *grid_from_host(const double *GV, int64_t gv_rs, int64_t gv_cs, int nx, int ny, int nz) {
  if (nx < 1 || ny < 1 || nz < 1 || gv_rs <= 0 || gv_cs <= 0) throw std::invalid_argument("bad grid dimensions or strides");
  const int64_t n = (int64_t)nx * ny * nz;
  if (n > 0x7FFFFFFF) throw std::invalid_argument("more than 2^31-1 grid vertices");
  require_device();
  std::unique_ptr<iglb200_grid> g(new iglb200_grid);
  g->side[0] = nx; g->side[1] = ny; g->side[2] = nz; g->n = n;
  g->GV.alloc((size_t)3 * n); g->S.alloc((size_t)n);
  const size_t span = (size_t)((n - 1) * gv_rs + 2 * gv_cs + 1);
  DevBuf<double> raw(span);
  plat::h2d(raw.p, GV, span * 8, nullptr);
  plat::for_n(n, ToCanon3<double>{raw.p, gv_rs, gv_cs, n, g->GV.p}, nullptr);
  plat::sync(nullptr);
  return g.release();
}
int iglb200_grid_create_from(const double *GV, int64_t gv_rs, int64_t gv_cs, int nx, int ny, int nz, iglb200_grid **out) {
  API_BEGIN
  if (!GV || !out) throw std::invalid_argument("null pointer");
  *out = grid_from_host(GV, gv_rs, gv_cs, nx, ny, nz);
  return IGLB200_OK;
  API_END
}
int iglb200_grid_destroy(iglb200_grid *g) {
  API_BEGIN
  delete g;
  return IGLB200_OK;
  API_END
}
int iglb200_grid_info(const iglb200_grid *g, int side[3], int64_t *n, double corners[6]) {
  API_BEGIN
  check_grid(g);
  if (side) for (int c = 0; c < 3; ++c) side[c] = g->side[c];
  if (n) *n = g->n;
  if (corners) {   // first and last grid vertex
    for (int c = 0; c < 3; ++c) {
      plat::d2h(&corners[c], g->GV.p + (size_t)c * g->n, 8, nullptr);
      plat::d2h(&corners[3 + c], g->GV.p + (size_t)c * g->n + (g->n - 1), 8, nullptr);
    }
    plat::sync(nullptr);
  }
  return IGLB200_OK;
  API_END
}
int iglb200_grid_device(iglb200_grid *g, const double **GV_colmajor, double **S) {
  API_BEGIN
  check_grid(g);
  if (GV_colmajor) *GV_colmajor = g->GV.p;
  if (S) *S = g->S.p;
  return IGLB200_OK;
  API_END
}
int iglb200_grid_get_gv(const iglb200_grid *g, double *GV, int64_t gv_rs, int64_t gv_cs) {
  API_BEGIN
  check_grid(g);
  if (!GV || gv_rs <= 0 || gv_cs <= 0) throw std::invalid_argument("bad output");
  const int64_t n = g->n;
  if (gv_rs == 1 && gv_cs == n) { plat::d2h(GV, g->GV.p, (size_t)24 * n, nullptr); plat::sync(nullptr); return IGLB200_OK; }
  const size_t span = (size_t)((n - 1) * gv_rs + 2 * gv_cs + 1);
  if (span != (size_t)3 * n) throw std::invalid_argument("GV output must be dense (row- or column-major)");
  DevBuf<double> tmp(span);
  plat::for_n(n, ColsToStrided{g->GV.p, n, tmp.p, gv_rs, gv_cs}, nullptr);
  plat::d2h(GV, tmp.p, span * 8, nullptr);
  plat::sync(nullptr);
  return IGLB200_OK;
  API_END
}
int iglb200_grid_set_s(iglb200_grid *g, const double *S) {
  API_BEGIN
  check_grid(g);
  if (!S) throw std::invalid_argument("null pointer");
  plat::h2d(g->S.p, S, (size_t)8 * g->n, nullptr);
  plat::sync(nullptr);
  g->has_S = true;
  return IGLB200_OK;
  API_END
}
int iglb200_grid_get_s(const iglb200_grid *g, double *S) {
  API_BEGIN
  check_grid(g);
  if (!S) throw std::invalid_argument("null pointer");
  if (!g->has_S) throw std::invalid_argument("the grid holds no values yet");
  plat::d2h(S, g->S.p, (size_t)8 * g->n, nullptr);
  plat::sync(nullptr);
  return IGLB200_OK;
  API_END
}
int iglb200_grid_signed_distance(iglb200_grid *g, const iglb200_mesh *m, int sign_type, double lower, double upper) {
  API_BEGIN
  check_grid(g); m = mesh_here(m);
  require_device();
  const int64_t n = g->n;
  DevBuf<int32_t> I((size_t)n);
  DevBuf<double> C((size_t)3 * n);
  DevBuf<char> scratch(iglb200_query_scratch_bytes(n));
  plat::launch_counter() = 0;
  run_signed_distance(m, g->GV.p, n, sign_type, lower, upper, g->S.p, I.p, C.p, nullptr, nullptr, scratch.p, 1, nullptr);
  plat::sync(nullptr);
  throw_if_overflow(nullptr);
  g->has_S = true;
  return IGLB200_OK;
  API_END
}
int iglb200_grid_flood_fill(iglb200_grid *g) {
  API_BEGIN
  check_grid(g);
  if (!g->has_S) throw std::invalid_argument("the grid holds no values yet");
  require_device();
  grid::flood_fill(g->side, g->S.p, nullptr);
  return IGLB200_OK;
  API_END
}
int iglb200_grid_marching_cubes(const iglb200_grid *g, double shift, double isovalue, iglb200_surface **out) {
  API_BEGIN
  check_grid(g);
  if (!out) throw std::invalid_argument("null pointer");
  if (!g->has_S) throw std::invalid_argument("the grid holds no values yet");
  require_device();
  *out = grid::marching_cubes(g->S.p, g->GV.p, g->side[0], g->side[1], g->side[2], shift, isovalue, nullptr);
  return IGLB200_OK;
  API_END
}
int iglb200_surface_info(const iglb200_surface *s, int64_t *nV, int64_t *nF) {
  API_BEGIN
  check_surface(s);
  if (nV) *nV = s->nV;
  if (nF) *nF = s->nF;
  return IGLB200_OK;
  API_END
}
int iglb200_surface_get(const iglb200_surface *s, double *V, int32_t *F, int32_t *E) {
  API_BEGIN
  check_surface(s);
  if (V && s->nV) plat::d2h(V, s->V.p, (size_t)24 * s->nV, nullptr);
  if (F && s->nF) plat::d2h(F, s->F.p, (size_t)12 * s->nF, nullptr);
  if (E && s->nV) plat::d2h(E, s->E.p, (size_t)8 * s->nV, nullptr);
  plat::sync(nullptr);
  return IGLB200_OK;
  API_END
}
int iglb200_surface_device(const iglb200_surface *s, const double **V, const int32_t **F, const int32_t **E) {
  API_BEGIN
  check_surface(s);
  if (V) *V = s->V.p;
  if (F) *F = s->F.p;
  if (E) *E = s->E.p;
  return IGLB200_OK;
  API_END
}
int iglb200_surface_destroy(iglb200_surface *s) {
  API_BEGIN
  delete s;
  return IGLB200_OK;
  API_END
}
// host-pointer one-shots with the reference's argument meaning
int iglb200_flood_fill(const int res[3], double *S) {
  API_BEGIN
  if (!res || !S) throw std::invalid_argument("null pointer");
  if (res[0] < 1 || res[1] < 1 || res[2] < 1) throw std::invalid_argument("empty grid");
  const int64_t n = (int64_t)res[0] * res[1] * res[2];
  require_device();
  DevBuf<double> d((size_t)n);
  plat::h2d(d.p, S, (size_t)8 * n, nullptr);
  grid::flood_fill(res, d.p, nullptr);
  plat::d2h(S, d.p, (size_t)8 * n, nullptr);
  plat::sync(nullptr);
  return IGLB200_OK;
  API_END
}
int iglb200_marching_cubes(const double *S, const double *GV, int64_t gv_rs, int64_t gv_cs, int nx, int ny, int nz,
                           double isovalue, iglb200_surface **out) {
  API_BEGIN
  if (!S || !GV || !out) throw std::invalid_argument("null pointer");
  std::unique_ptr<iglb200_grid> hold(grid_from_host(GV, gv_rs, gv_cs, nx, ny, nz));
  iglb200_grid *g = hold.get();
  plat::h2d(g->S.p, S, (size_t)8 * g->n, nullptr);
  plat::sync(nullptr);
  *out = grid::marching_cubes(g->S.p, g->GV.p, nx, ny, nz, 0.0, isovalue, nullptr);
  return IGLB200_OK;
  API_END
}

int iglb200_cloud_create(const void *P, const void *N, const void *A, int is_f64, int64_t nP, int expansion_order,
                         iglb200_cloud **out) {
  API_BEGIN
  if (!out) throw std::invalid_argument("null out");
  *out = nullptr;
  if (nP < 0 || (nP > 0 && (!P || !N || !A))) throw std::invalid_argument("null or negative-sized input");
  if (expansion_order < 0 || expansion_order > 2) throw std::invalid_argument("expansion_order must be 0, 1 or 2");
  require_device();
  plat::launch_counter() = 0;
  *out = cloud::create(P, N, A, is_f64, nP, expansion_order, nullptr);
  return IGLB200_OK;
  API_END
}
int iglb200_cloud_destroy(iglb200_cloud *c) {
  API_BEGIN
  cloud::destroy(c);
  return IGLB200_OK;
  API_END
}
int iglb200_cloud_info(const iglb200_cloud *c, int64_t info[8]) {
  API_BEGIN
  if (!c) throw std::invalid_argument("invalid cloud handle");
  cloud::info(c, info);
  return IGLB200_OK;
  API_END
}
int iglb200_cloud_eval_dev(const iglb200_cloud *c, const double *Q, int64_t nQ, double beta, double *WN, void *stream) {
  API_BEGIN
  if (!c) throw std::invalid_argument("invalid cloud handle");
  if (nQ < 0 || (nQ > 0 && (!Q || !WN))) throw std::invalid_argument("null pointer");
  require_device();
  plat::launch_counter() = 0;
  cloud::eval(c, Q, nQ, beta, WN, (stream_t)stream);
  return IGLB200_OK;
  API_END
}
int iglb200_cloud_eval(const iglb200_cloud *c, const void *Q, int q_is_f64, int64_t nQ, int64_t q_rs, int64_t q_cs, double beta,
                       void *WN, int w_is_f64) {
  API_BEGIN
  if (!c) throw std::invalid_argument("invalid cloud handle");
  if (nQ < 0 || (nQ > 0 && (!Q || !WN))) throw std::invalid_argument("null pointer");
  require_device();
  plat::launch_counter() = 0;
  stream_t s = nullptr;
  DevBuf<double> dQ, dW((size_t)nQ);
  upload_points(Q, q_is_f64, nQ, q_rs, q_cs, dQ, s);
  cloud::eval(c, dQ.p, nQ, beta, dW.p, s);
  download_scalar(dW, nQ, WN, w_is_f64, s);
  plat::sync(s);
  return IGLB200_OK;
  API_END
}

}  // extern "C"
