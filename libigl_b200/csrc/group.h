// The device group of a process: iglb200_init(n) makes devices 0..n-1 one group.
// A mesh created afterwards is built once (on the first device) and its flat arena
// is replicated to every other device of the group — ncclBroadcast over NVLink
// (NCCL is loaded at run time: libnccl.so.2, the copy torch ships when the process
// already has it) or, when NCCL cannot be loaded, peer-to-peer DMA.  Host-pointer
// queries on such a mesh are dealt to all devices of the group (hostpath.h).
//
// Matches the loop igl::signed_distance runs over igl::parallel_for threads
// (/root/reference/include/igl/signed_distance.cpp:90-138): queries are independent,
// the structures are read-only.
#pragma once
#include "mesh.h"
#include <mutex>
#ifndef IGLB200_EMUL
#include <dlfcn.h>
#endif

namespace grp {

struct State {
  std::mutex mu;
  int n = 1;                      // devices 0..n-1
  bool nccl = false;              // replicate with ncclBroadcast
  bool peer = false;              // peer access enabled between all pairs
  std::string transport = "none"; // what the last replication used
  double last_replicate_ms = 0.0;
#ifndef IGLB200_EMUL
  void *lib = nullptr;
  void *comms[plat::MAX_DEVICES] = {nullptr};
  cudaStream_t streams[plat::MAX_DEVICES] = {nullptr};
  // the five NCCL entry points used (prototypes as in nccl.h 2.x; ncclChar == 0)
  int (*CommInitAll)(void **comms, int ndev, const int *devlist) = nullptr;
  int (*CommDestroy)(void *comm) = nullptr;
  int (*Broadcast)(const void *send, void *recv, size_t count, int dtype, int root, void *comm, cudaStream_t s) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
#endif
};
inline State &state() { static State *s = new State; return *s; }   // never destroyed: outlives every worker thread

#ifndef IGLB200_EMUL
inline bool load_nccl(State &g) {
  if (g.lib) return true;
  const char *force = std::getenv("IGLB200_REPLICATE");
  if (force && std::string(force) == "peer") return false;
  // an NCCL that is already in the process (torch's) wins: RTLD_NOLOAD finds it by soname
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) return false;
  auto sym = [&](const char *name) { return dlsym(h, name); };
  g.CommInitAll = (int (*)(void **, int, const int *))sym("ncclCommInitAll");
  g.CommDestroy = (int (*)(void *))sym("ncclCommDestroy");
  g.Broadcast = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))sym("ncclBroadcast");
  g.GroupStart = (int (*)())sym("ncclGroupStart");
  g.GroupEnd = (int (*)())sym("ncclGroupEnd");
  g.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
  if (!g.CommInitAll || !g.CommDestroy || !g.Broadcast || !g.GroupStart || !g.GroupEnd || !g.GetErrorString) return false;
  g.lib = h;
  return true;
}
inline void nccl_check(State &g, int rc, const char *what) {
  if (rc != 0) throw std::runtime_error(std::string("NCCL: ") + what + ": " + g.GetErrorString(rc));
}
#endif

// (Re)define the group.  n == 1 drops back to single-device operation.
inline void init(int n) {
  State &g = state();
  std::lock_guard<std::mutex> lk(g.mu);
  if (n < 1) throw std::invalid_argument("n_gpus must be >= 1");
#ifdef IGLB200_EMUL
  if (n > plat::device_count()) throw std::invalid_argument("fewer devices than requested");
  g.n = n; g.peer = true; g.nccl = false;
#else
  const int have = plat::device_count();
  if (have < n) throw std::invalid_argument("fewer CUDA devices than requested");
  const int cur = plat::current_device();
  for (int d = 0; d < n; ++d) { plat::set_device(d); PLAT_CHECK(cudaFree(nullptr)); }
  // tear down the communicators of a previous, different group
  if (g.nccl && g.n != n) {
    for (int d = 0; d < g.n; ++d) { if (g.comms[d]) g.CommDestroy(g.comms[d]); g.comms[d] = nullptr; }
    g.nccl = false;
  }
  bool peer = n > 1;
  for (int a = 0; a < n && n > 1; ++a) {
    plat::set_device(a);
    if (!g.streams[a]) g.streams[a] = plat::stream_create();
    for (int b = 0; b < n; ++b) {
      if (a == b) continue;
      int can = 0;
      PLAT_CHECK(cudaDeviceCanAccessPeer(&can, a, b));
      if (!can) { peer = false; continue; }
      const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) { cudaGetLastError(); peer = false; }
    }
  }
  g.peer = peer;
  if (n > 1 && !g.nccl && load_nccl(g)) {
    int devs[plat::MAX_DEVICES];
    for (int d = 0; d < n; ++d) devs[d] = d;
    nccl_check(g, g.CommInitAll(g.comms, n, devs), "ncclCommInitAll");
    g.nccl = true;
  }
  g.n = n;
  plat::set_device(cur);
#endif
}

inline int size() { return state().n; }

// Copy the arena of `m` (built on device m->device) to every other device of the group.
inline void replicate(iglb200_mesh *m) {
  State &g = state();
  std::lock_guard<std::mutex> lk(g.mu);
  for (iglb200_mesh *r : m->reps) { if (r) { if (r->owns && r->arena) { plat::DeviceGuard dg(r->device); plat::dev_free(r->arena); } delete r; } }
  m->reps.clear();
  if (g.n <= 1 || !m->arena) return;
  if (m->device >= g.n) throw std::invalid_argument("the mesh was built on a device outside the group");
  m->reps.assign((size_t)g.n, nullptr);
  const size_t bytes = (size_t)m->h.bytes;
  for (int d = 0; d < g.n; ++d) {
    if (d == m->device) continue;
    plat::DeviceGuard dg(d);
    iglb200_mesh *r = new iglb200_mesh;
    r->h = m->h; r->device = d; r->owns = true; r->build_ms = m->build_ms;
    m->reps[(size_t)d] = r;
    r->arena = (char *)plat::dev_alloc(bytes);
  }
#ifdef IGLB200_EMUL
  for (int d = 0; d < g.n; ++d) if (m->reps[(size_t)d]) std::memcpy(m->reps[(size_t)d]->arena, m->arena, bytes);
  g.transport = "memcpy (emulation)";
#else
  const int cur = plat::current_device();
  cudaEvent_t e0, e1;
  plat::set_device(m->device);
  PLAT_CHECK(cudaDeviceSynchronize());   // the build ran on the caller's stream
  PLAT_CHECK(cudaEventCreate(&e0)); PLAT_CHECK(cudaEventCreate(&e1));
  PLAT_CHECK(cudaEventRecord(e0, g.streams[m->device]));
  if (g.nccl) {
    nccl_check(g, g.GroupStart(), "ncclGroupStart");
    for (int d = 0; d < g.n; ++d) {
      plat::set_device(d);
      const void *src = m->arena;
      void *dst = d == m->device ? (void *)m->arena : (void *)m->reps[(size_t)d]->arena;
      nccl_check(g, g.Broadcast(src, dst, bytes, 0 /*ncclChar*/, m->device, g.comms[d], g.streams[d]), "ncclBroadcast");
    }
    nccl_check(g, g.GroupEnd(), "ncclGroupEnd");
    g.transport = "ncclBroadcast";
  } else {
    // every destination pulls its copy on its own stream; NVSwitch serves them at the source's egress bandwidth
    for (int d = 0; d < g.n; ++d) {
      if (d == m->device) continue;
      plat::set_device(d);
      plat::peer_copy(m->reps[(size_t)d]->arena, d, m->arena, m->device, bytes, g.streams[d]);
    }
    g.transport = g.peer ? "cudaMemcpyPeerAsync (P2P)" : "cudaMemcpyPeerAsync (staged)";
  }
  for (int d = 0; d < g.n; ++d) { plat::set_device(d); PLAT_CHECK(cudaStreamSynchronize(g.streams[d])); }
  plat::set_device(m->device);
  PLAT_CHECK(cudaEventRecord(e1, g.streams[m->device]));
  PLAT_CHECK(cudaEventSynchronize(e1));
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
  g.last_replicate_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  plat::set_device(cur);
#endif
}

}  // namespace grp
