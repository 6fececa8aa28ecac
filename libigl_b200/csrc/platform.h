// Device/host plumbing shared by every kernel file.
//
// The product is compiled by nvcc for sm_100a only.  The SAME kernel bodies can
// also be compiled by g++ with -DIGLB200_EMUL: every launch then becomes a plain
// loop on the host.  That build exists ONLY so that `pytest -m "not gpu"` can
// check the kernels' logic against the oracle in a container without a GPU; it
// is produced under tests/_build/ by tests/emul.py and is never loaded by
// libigl_b200/.  There is no CPU fallback in the product library: without a
// CUDA device every entry point returns IGLB200_ERR_CUDA.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <stdexcept>
#include <vector>
#include <algorithm>
#include <memory>

namespace plat {
// kernels (ours + cub passes) enqueued by the calling thread; reset by the C ABI per call
inline int &launch_counter() { static thread_local int c = 0; return c; }
static constexpr int MAX_DEVICES = 64;
}  // namespace plat

#ifdef IGLB200_EMUL
namespace plat { struct Span { Span(const char *, void *) {} }; inline void prof_enable(bool) {} inline int prof_collect(char *, size_t) { return 0; } }
// ------------------------------------------------------------------ host loops
#include <vector_types.h>  // float4/uint4/double3 PODs (host-safe CUDA header)
#include <vector_functions.h>
#include <numeric>
#define HD inline
#define DEVFN inline
typedef void *stream_t;
namespace plat {
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline uint32_t atomic_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
inline uint32_t atomic_min(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
inline uint32_t atomic_max(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
inline void atomic_min64(unsigned long long *p, unsigned long long v) { if (v < *p) *p = v; }
inline void atomic_max64(unsigned long long *p, unsigned long long v) { if (v > *p) *p = v; }
inline float fsqrt(float x) { return std::sqrt(x); }
inline float fdiv(float a, float b) { return a / b; }
inline float fmul(float a, float b) { return a * b; }  // built with -ffp-contract=off
inline float fadd(float a, float b) { return a + b; }
inline float ffma(float a, float b, float c) { return a * b + c; }
inline void *dev_alloc(size_t bytes) { void *p = std::malloc(bytes ? bytes : 1); if (!p) throw std::bad_alloc(); return p; }
inline void dev_free(void *p) { std::free(p); }
inline void h2d(void *d, const void *h, size_t n, stream_t) { std::memcpy(d, h, n); }
inline void d2h(void *h, const void *d, size_t n, stream_t) { std::memcpy(h, d, n); }
inline void d2d(void *d, const void *s, size_t n, stream_t) { std::memcpy(d, s, n); }
inline void dev_memset(void *d, int v, size_t n, stream_t) { std::memset(d, v, n); }
inline void sync(stream_t) {}
// "devices" of the emulation are virtual: memory is host memory, the current device is a per-thread integer
inline int &emul_dev() { static thread_local int d = 0; return d; }
inline int current_device() { return emul_dev(); }
inline void set_device(int d) { emul_dev() = d; }
inline int device_count() { return 16; }
inline void *pinned_alloc(size_t bytes) { void *p = std::malloc(bytes ? bytes : 1); if (!p) throw std::bad_alloc(); return p; }
inline void pinned_free(void *p) { std::free(p); }
inline bool host_is_pinned(const void *) { return false; }   // every host buffer is staged: the emulation exercises that path
inline stream_t stream_create() { return nullptr; }
inline void stream_destroy(stream_t) {}
inline void peer_copy(void *d, int, const void *s_, int, size_t n, stream_t) { std::memcpy(d, s_, n); }
inline void pin_thread_near_device(int) {}
typedef void *event_t;
inline event_t event_create() { return nullptr; }
inline void event_record(event_t, stream_t) {}
inline void stream_wait(stream_t, event_t) {}
// f(i) for i in [0,n)
template <class F> void for_n(int64_t n, F f, stream_t) { for (int64_t i = 0; i < n; ++i) f(i); }
// f(i) for i in [0,*count) with the count living in "device" memory
template <class F> void for_dyn(const uint32_t *count, F f, stream_t) { for (uint32_t i = 0; i < *count; ++i) f(i); }
// f.tile(t) for t in [0,*ntiles): block-cooperative on the device, sequential here
template <class F> void tiles_dyn(const uint32_t *ntiles, F f, stream_t) { for (uint32_t t = 0; t < *ntiles; ++t) f.tile(t); }
inline void exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, stream_t) {
  uint32_t s = 0;
  for (int64_t i = 0; i < n; ++i) { uint32_t v = in[i]; out[i] = s; s += v; }
}
inline void inclusive_scan_f64(const double *in, double *out, int64_t n, stream_t) {
  double acc = 0.0;
  for (int64_t i = 0; i < n; ++i) { acc += in[i]; out[i] = acc; }
}
inline size_t sort_pairs_temp_bytes(int64_t) { return 16; }
// stable sort of (key,val) by key
inline void sort_pairs_u32(void *, size_t, const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, int64_t n, int, stream_t) {
  std::vector<uint32_t> o(n);
  std::iota(o.begin(), o.end(), 0u);
  std::stable_sort(o.begin(), o.end(), [&](uint32_t a, uint32_t b) { return kin[a] < kin[b]; });
  for (int64_t i = 0; i < n; ++i) { kout[i] = kin[o[i]]; vout[i] = vin[o[i]]; }
}
inline size_t sort_pairs64_temp_bytes(int64_t) { return 16; }
inline void sort_pairs_u64(void *, size_t, const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, int64_t n, stream_t) {
  std::vector<uint32_t> o(n);
  std::iota(o.begin(), o.end(), 0u);
  std::stable_sort(o.begin(), o.end(), [&](uint32_t a, uint32_t b) { return kin[a] < kin[b]; });
  for (int64_t i = 0; i < n; ++i) { kout[i] = kin[o[i]]; vout[i] = vin[o[i]]; }
}
}  // namespace plat
#else
// ------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <sched.h>
#include <cctype>
#define HD __host__ __device__ __forceinline__
#define DEVFN __device__ __forceinline__
typedef cudaStream_t stream_t;
namespace plat {
struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw CudaError(std::string(what) + ": " + cudaGetErrorString(e));
}
#define PLAT_CHECK(x) ::plat::check((x), #x)
HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
DEVFN uint32_t atomic_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
DEVFN uint32_t atomic_min(uint32_t *p, uint32_t v) { return atomicMin(p, v); }
DEVFN uint32_t atomic_max(uint32_t *p, uint32_t v) { return atomicMax(p, v); }
DEVFN void atomic_min64(unsigned long long *p, unsigned long long v) { atomicMin(p, v); }
DEVFN void atomic_max64(unsigned long long *p, unsigned long long v) { atomicMax(p, v); }
// IEEE round-to-nearest, never contracted into FMA regardless of -fmad
DEVFN float fsqrt(float x) { return __fsqrt_rn(x); }
DEVFN float fdiv(float a, float b) { return __fdiv_rn(a, b); }
DEVFN float fmul(float a, float b) { return __fmul_rn(a, b); }
DEVFN float fadd(float a, float b) { return __fadd_rn(a, b); }
DEVFN float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
inline void *dev_alloc(size_t bytes) { void *p = nullptr; PLAT_CHECK(cudaMalloc(&p, bytes ? bytes : 1)); return p; }
inline void dev_free(void *p) { if (p) cudaFree(p); }
inline void h2d(void *d, const void *h, size_t n, stream_t s) { if (n) PLAT_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
inline void d2h(void *h, const void *d, size_t n, stream_t s) { if (n) PLAT_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
inline void d2d(void *d, const void *s_, size_t n, stream_t s) { if (n) PLAT_CHECK(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s)); }
inline void dev_memset(void *d, int v, size_t n, stream_t s) { if (n) PLAT_CHECK(cudaMemsetAsync(d, v, n, s)); }
inline void sync(stream_t s) { PLAT_CHECK(cudaStreamSynchronize(s)); }

inline int current_device() { int d = 0; PLAT_CHECK(cudaGetDevice(&d)); return d; }
inline void set_device(int d) { PLAT_CHECK(cudaSetDevice(d)); }
inline int device_count() { int n = 0; PLAT_CHECK(cudaGetDeviceCount(&n)); return n; }
inline void *pinned_alloc(size_t bytes) { void *p = nullptr; PLAT_CHECK(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable)); return p; }
inline void pinned_free(void *p) { if (p) cudaFreeHost(p); }
// page-locked (cudaHostAlloc / cudaHostRegister) host memory can be the source/target of truly asynchronous copies
inline bool host_is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
inline stream_t stream_create() { cudaStream_t s = nullptr; PLAT_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); return s; }
inline void stream_destroy(stream_t s) { if (s) cudaStreamDestroy(s); }
inline void peer_copy(void *d, int ddev, const void *s_, int sdev, size_t n, stream_t s) {
  if (n) PLAT_CHECK(cudaMemcpyPeerAsync(d, ddev, s_, sdev, n, s));
}
typedef cudaEvent_t event_t;
inline event_t event_create() { cudaEvent_t e = nullptr; PLAT_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); return e; }
inline void event_record(event_t e, stream_t s) { PLAT_CHECK(cudaEventRecord(e, s)); }
inline void stream_wait(stream_t s, event_t e) { PLAT_CHECK(cudaStreamWaitEvent(s, e, 0)); }
// run the calling thread on the CPUs next to the GPU (pinned staging buffers are then first-touched on that NUMA node)
inline void pin_thread_near_device(int dev) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) { cudaGetLastError(); return; }
  for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
  const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return;
  char line[4096] = {0};
  const bool ok = fgets(line, sizeof(line), f) != nullptr;
  fclose(f);
  if (!ok) return;
  cpu_set_t set; CPU_ZERO(&set);
  int any = 0;
  char *save = nullptr;
  for (char *tok = strtok_r(line, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {
    int a = 0, b = 0;
    const int k = sscanf(tok, "%d-%d", &a, &b);
    if (k == 1) b = a;
    if (k >= 1) for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &set); ++any; }
  }
  if (any) sched_setaffinity(0, sizeof(set), &set);
}

inline int sm_count() {   // of the current device (cached per device)
  static int cache[64] = {0};
  int d = 0, v = 148;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return v;
  if (!cache[d]) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); cache[d] = v; }
  return cache[d];
}

template <class F> __global__ void __launch_bounds__(256) k_for_n(int64_t n, F f) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) f(i);
}
template <class F> __global__ void __launch_bounds__(256) k_for_dyn(const uint32_t *count, F f) {
  const uint32_t n = *count;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) f(i);
}
template <class F> __global__ void __launch_bounds__(256) k_tiles_dyn(const uint32_t *ntiles, F f) {
  const uint32_t n = *ntiles;
  for (uint32_t t = blockIdx.x; t < n; t += gridDim.x) { f.tile(t); __syncthreads(); }
}
template <class F> __global__ void __launch_bounds__(256) k_tiles_one(F f) { f.tile(0); }
template <class F> void for_n(int64_t n, F f, stream_t s) {
  if (n <= 0) return;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  k_for_n<<<(unsigned)blocks, 256, 0, s>>>(n, f);
  launch_counter()++;
  PLAT_CHECK(cudaGetLastError());
}
template <class F> void for_dyn(const uint32_t *count, F f, stream_t s) {
  k_for_dyn<<<sm_count() * 8, 256, 0, s>>>(count, f);
  launch_counter()++;
  PLAT_CHECK(cudaGetLastError());
}
template <class F> void tiles_dyn(const uint32_t *ntiles, F f, stream_t s) {
  k_tiles_dyn<<<sm_count() * 8, 256, 0, s>>>(ntiles, f);
  launch_counter()++;
  PLAT_CHECK(cudaGetLastError());
}
// scratch for cub, grown on demand (one per thread and device; builds are serialised per handle)
struct ScanTemp { void *p = nullptr; size_t cap = 0; ~ScanTemp() { dev_free(p); } };
inline void exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, stream_t s) {
  if (n <= 0) return;
  if (n > 0x7FFFFFFF) throw std::invalid_argument("scan of more than 2^31-1 items");
  static thread_local ScanTemp tmps[MAX_DEVICES];
  ScanTemp &tmp = tmps[current_device() % MAX_DEVICES];
  size_t need = 0;
  PLAT_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n, s));
  if (need > tmp.cap) { sync(s); dev_free(tmp.p); tmp.p = dev_alloc(need); tmp.cap = need; }
  PLAT_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, need, in, out, (int)n, s));
  launch_counter()++;
}
inline void inclusive_scan_f64(const double *in, double *out, int64_t n, stream_t s) {
  if (n <= 0) return;
  if (n > 0x7FFFFFFF) throw std::invalid_argument("scan of more than 2^31-1 items");
  static thread_local ScanTemp tmps[MAX_DEVICES];
  ScanTemp &tmp = tmps[current_device() % MAX_DEVICES];
  size_t need = 0;
  PLAT_CHECK(cub::DeviceScan::InclusiveSum(nullptr, need, in, out, (int)n, s));
  if (need > tmp.cap) { sync(s); dev_free(tmp.p); tmp.p = dev_alloc(need); tmp.cap = need; }
  PLAT_CHECK(cub::DeviceScan::InclusiveSum(tmp.p, need, in, out, (int)n, s));
  launch_counter()++;
}
inline size_t sort_pairs_temp_bytes(int64_t n) {
  size_t need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, 32);
  return need + 256;
}
inline size_t sort_pairs64_temp_bytes(int64_t n) {
  size_t need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, 64);
  return need + 256;
}
inline void sort_pairs_u64(void *temp, size_t temp_bytes, const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, int64_t n, stream_t s) {
  if (n > 0x7FFFFFFF) throw std::invalid_argument("sort of more than 2^31-1 items");
  PLAT_CHECK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kin, kout, vin, vout, (int)n, 0, 64, s));
  launch_counter() += 9;
}
inline void sort_pairs_u32(void *temp, size_t temp_bytes, const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, int64_t n, int end_bit, stream_t s) {
  if (n > 0x7FFFFFFF) throw std::invalid_argument("sort of more than 2^31-1 items");
  PLAT_CHECK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kin, kout, vin, vout, (int)n, 0, end_bit, s));
  launch_counter() += 4;
}
}  // namespace plat
#endif

#ifndef IGLB200_EMUL
namespace plat {
// Optional per-kernel timing with CUDA events on the launching stream (bench.py's
// roofline leg).  Off by default: no events are created or recorded.
struct ProfRec { std::string name; cudaEvent_t a, b; };
inline bool &prof_on() { static thread_local bool on = false; return on; }
inline std::vector<ProfRec> &prof_recs() { static thread_local std::vector<ProfRec> r; return r; }
inline void prof_enable(bool on) { prof_on() = on; }
struct Span {
  bool live; cudaStream_t s; size_t idx;
  Span(const char *name, cudaStream_t st) : live(prof_on()), s(st), idx(0) {
    if (!live) return;
    ProfRec r; r.name = name;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    idx = prof_recs().size();
    prof_recs().push_back(r);
  }
  ~Span() { if (live) cudaEventRecord(prof_recs()[idx].b, s); }
};
// "name=ms;name=ms;..." summed per name over everything recorded since the last collect
inline int prof_collect(char *out, size_t cap) {
  std::vector<std::pair<std::string, double>> acc;
  for (auto &r : prof_recs()) {
    cudaEventSynchronize(r.b);
    float ms = 0.f; cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    bool found = false;
    for (auto &p : acc) if (p.first == r.name) { p.second += ms; found = true; }
    if (!found) acc.push_back({r.name, ms});
  }
  prof_recs().clear();
  std::string sres;
  for (auto &p : acc) { char b[128]; snprintf(b, sizeof(b), "%s=%.6f;", p.first.c_str(), p.second); sres += b; }
  if (out && cap) { strncpy(out, sres.c_str(), cap - 1); out[cap - 1] = 0; }
  return (int)sres.size();
}
}  // namespace plat
#endif

namespace plat {
// one T per (host thread, device): streams, workspaces and flags must never be used on a device other than the one
// they were created on
template <class T> T &per_device() {
  static thread_local std::unique_ptr<T> slot[MAX_DEVICES];
  const int d = current_device();
  if (d < 0 || d >= MAX_DEVICES) throw std::invalid_argument("device index out of range");
  if (!slot[d]) slot[d].reset(new T());
  return *slot[d];
}
// make `device` current for the lifetime of the guard
struct DeviceGuard {
  int prev; bool switched;
  explicit DeviceGuard(int device) : prev(current_device()), switched(device != prev) { if (switched) set_device(device); }
  ~DeviceGuard() { if (switched) { try { set_device(prev); } catch (...) {} } }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};
// RAII device array
template <class T> struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  ~DevBuf() { release(); }
  void alloc(size_t count) { release(); p = (T *)dev_alloc(count * sizeof(T)); n = count; }
  void release() { if (p) dev_free(p); p = nullptr; n = 0; }
  size_t bytes() const { return n * sizeof(T); }
};

// order-preserving float <-> uint32 map so box unions can use integer atomics
HD uint32_t f2ord(float f) { uint32_t u = f2u(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
HD float ord2f(uint32_t u) { return u2f((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
HD unsigned long long d2ord(double d) {
  unsigned long long u; memcpy(&u, &d, 8);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
HD double ord2d(unsigned long long u) {
  u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
  double d; memcpy(&d, &u, 8); return d;
}
}  // namespace plat
