// Host model of the work-donation protocol of closest_packet_kernel (libigl_b200/csrc/iglb200.cu: cp_donate,
// cp_take_task, cp_begin_shared, cp_finish_shared), one std::thread per warp and std::atomic for the device atomics.
// It checks what the protocol must guarantee whatever the interleaving:
//   * every warp terminates (no helper waits for work that cannot come, no donor waits for a helper),
//   * a task put into a mailbox is run exactly once,
//   * the pieces of a packet merge to the packet's true minimum and the result is stored exactly once, by the last piece.
// A "walk" here is a random tree of nodes with a value at every leaf; the result of a packet is the smallest leaf value
// (ties: lower leaf id) — the same order-independent rule the search uses (smaller distance, lower face index).
// Test infrastructure: states, counters and the order of operations mirror the kernel's; the arithmetic does not matter.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

namespace {
constexpr uint32_t BOX_CLOSED = 0, BOX_WAITING = 1, BOX_CLAIMED = 2, BOX_FILLED = 3;
constexpr int TASK_NODES = 13, DONATE_EVERY = 4;

struct Node { int first_child, nchild; uint64_t value; };   // leaf: nchild == 0
struct Packet { std::vector<Node> nodes; uint64_t truth; };
struct Task { uint32_t packet, nn, slot; uint32_t node[TASK_NODES]; };
struct Slot { std::atomic<uint64_t> best{~0ull}; std::atomic<uint32_t> pending{0}, lock{0}; };

struct Launch {
  std::vector<Packet> packets;
  std::atomic<uint32_t> next_packet{0}, idle{0}, active{0}, slots_out{0};
  std::vector<std::atomic<uint32_t>> box;
  std::vector<Task> task;
  std::vector<Slot> slot;
  std::vector<std::atomic<uint32_t>> stored;      // how often the result of packet p was written
  std::vector<uint64_t> result;
  std::atomic<uint64_t> tasks_sent{0}, tasks_run{0};
  Launch(int warps, size_t npackets, int nslots) : box(warps), task(warps), slot(nslots), stored(npackets), result(npackets, 0) {
    for (auto &b : box) b = BOX_CLOSED;
    for (auto &s : stored) s = 0;
  }
};

uint64_t walk_serial(const Packet &p) {
  uint64_t best = ~0ull;
  std::vector<uint32_t> st{0};
  while (!st.empty()) {
    const Node &n = p.nodes[st.back()]; st.pop_back();
    if (n.nchild == 0) { if (n.value < best) best = n.value; continue; }
    for (int c = 0; c < n.nchild; ++c) st.push_back((uint32_t)(n.first_child + c));
  }
  return best;
}

Packet make_packet(std::mt19937_64 &rng, int budget) {
  Packet p;
  p.nodes.push_back({0, 0, 0});
  std::vector<uint32_t> open{0};
  while (!open.empty()) {
    const uint32_t i = open.back(); open.pop_back();
    const bool leaf = (int)p.nodes.size() >= budget || rng() % 5 == 0;
    if (leaf) { p.nodes[i] = {0, 0, ((rng() % 1000) << 32) | i}; continue; }   // value | leaf id: ties resolved by id
    const int nc = 1 + (int)(rng() % 4);
    p.nodes[i] = {(int)p.nodes.size(), nc, 0};
    for (int c = 0; c < nc; ++c) { open.push_back((uint32_t)p.nodes.size()); p.nodes.push_back({0, 0, 0}); }
  }
  p.truth = walk_serial(p);
  return p;
}

void finish_shared(Launch &L, uint32_t slot, uint32_t packet, uint64_t mine) {
  Slot &s = L.slot[slot];
  uint32_t z = 0;
  while (!s.lock.compare_exchange_weak(z, 1u)) { z = 0; std::this_thread::yield(); }
  if (mine < s.best.load()) s.best.store(mine);
  s.lock.store(0u);
  if (s.pending.fetch_sub(1u) - 1u != 0u) return;
  L.result[packet] = s.best.load();   // every piece merged before it counted itself out
  L.stored[packet].fetch_add(1u);
}

void warp(Launch &L, int me, int warps, uint64_t seed) {
  std::mt19937_64 rng(seed);
  bool drained = false, counted = false;
  std::vector<uint32_t> stack;
  while (true) {
    uint32_t packet = 0; int32_t slot = -1; bool is_task = false;
    stack.clear();
    if (!drained) {
      packet = L.next_packet.fetch_add(1u);
      if (packet >= L.packets.size()) drained = true;
      else { if (!counted) { counted = true; L.active.fetch_add(1u); } stack.push_back(0); }
    }
    if (drained) {   // cp_take_task
      if (counted) L.active.fetch_sub(1u);
      L.idle.fetch_add(1u);
      L.box[me].store(BOX_WAITING);
      bool got = false; unsigned polls = 0;
      while (true) {
        const uint32_t st = L.box[me].load();
        if (st == BOX_FILLED) { got = true; break; }
        uint32_t w = BOX_WAITING;
        if (st == BOX_WAITING && (++polls & 3u) == 0u && L.active.load() == 0u && L.box[me].compare_exchange_strong(w, BOX_CLOSED)) break;
        std::this_thread::yield();
      }
      L.idle.fetch_sub(1u);
      if (got) L.active.fetch_add(1u);
      if (!got) { counted = false; return; }
      counted = true;
      const Task &t = L.task[me];
      packet = t.packet; slot = (int32_t)t.slot;
      for (uint32_t k = 0; k < t.nn; ++k) stack.push_back(t.node[k]);
      L.box[me].store(BOX_CLOSED);
      L.tasks_run.fetch_add(1u);
      is_task = true;
    }
    (void)is_task;
    const Packet &P = L.packets[packet];
    uint64_t best = ~0ull;
    uint32_t npop = 0;
    while (!stack.empty()) {
      const uint32_t cur = stack.back(); stack.pop_back();
      if ((++npop % DONATE_EVERY) == 0 && stack.size() >= 2 && L.idle.load() > 0) {   // cp_donate
        if (slot < 0) {   // cp_begin_shared
          const uint32_t k = L.slots_out.fetch_add(1u);
          if (k < L.slot.size()) { L.slot[k].best.store(~0ull); L.slot[k].pending.store(1u); L.slot[k].lock.store(0u); slot = (int32_t)k; }
        }
        if (slot >= 0) {
          int box = -1;
          for (int tries = 0; tries < 4 && box < 0; ++tries) {
            const int h = (int)(rng() % (uint64_t)warps);
            uint32_t w = BOX_WAITING;
            if (h != me && L.box[h].compare_exchange_strong(w, BOX_CLAIMED)) box = h;
          }
          if (box >= 0) {
            const size_t nd = std::min<size_t>(stack.size() / 2, TASK_NODES);
            Task &t = L.task[box];
            t.packet = packet; t.nn = (uint32_t)nd; t.slot = (uint32_t)slot;
            for (size_t k = 0; k < nd; ++k) t.node[k] = stack[k];
            L.slot[slot].pending.fetch_add(1u);   // one more piece, before anybody can finish it
            stack.erase(stack.begin(), stack.begin() + (long)nd);
            L.tasks_sent.fetch_add(1u);
            L.box[box].store(BOX_FILLED);
          }
        }
      }
      const Node &n = P.nodes[cur];
      for (volatile int spin = 0; spin < 300; ++spin) {}   // a node visit takes time: lets the tail of a launch exist
      if (n.nchild == 0) { if (n.value < best) best = n.value; continue; }
      for (int c = 0; c < n.nchild; ++c) stack.push_back((uint32_t)(n.first_child + c));
    }
    if (slot >= 0) finish_shared(L, (uint32_t)slot, packet, best);
    else { L.result[packet] = best; L.stored[packet].fetch_add(1u); }
  }
}
}  // namespace

int main(int argc, char **argv) {
  const int rounds = argc > 1 ? std::atoi(argv[1]) : 20;
  uint64_t sent = 0;
  for (int r = 0; r < rounds; ++r) {
    std::mt19937_64 rng(1234 + r);
    const int warps = 2 + (int)(rng() % 15);
    const size_t npackets = 1 + rng() % 60;
    // a few slots only in some rounds: running out of slots must only stop the donations
    Launch L(warps, npackets, r % 3 == 0 ? 2 : 4096);
    for (size_t p = 0; p < npackets; ++p) L.packets.push_back(make_packet(rng, rng() % 10 == 0 ? 20000 : 200));   // heavy tail
    std::vector<std::thread> th;
    for (int w = 0; w < warps; ++w) th.emplace_back(warp, std::ref(L), w, warps, rng());
    for (auto &t : th) t.join();
    for (size_t p = 0; p < npackets; ++p) {
      if (L.stored[p].load() != 1u) { std::printf("round %d: packet %zu stored %u times\n", r, p, L.stored[p].load()); return 1; }
      if (L.result[p] != L.packets[p].truth) { std::printf("round %d: packet %zu wrong result\n", r, p); return 1; }
    }
    if (L.tasks_sent.load() != L.tasks_run.load()) { std::printf("round %d: %llu tasks sent, %llu run\n", r, (unsigned long long)L.tasks_sent.load(), (unsigned long long)L.tasks_run.load()); return 1; }
    if (L.idle.load() != 0u || L.active.load() != 0u) { std::printf("round %d: idle %u active %u at the end\n", r, L.idle.load(), L.active.load()); return 1; }
    sent += L.tasks_sent.load();
  }
  std::printf("OK %d rounds, %llu tasks donated\n", rounds, (unsigned long long)sent);
  return sent > 0 ? 0 : 2;
}
