// TEST INFRASTRUCTURE ONLY — never linked, imported or executed by the product
// path (libigl_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library built from this.
//
// Thin extern "C" shim around the GENUINE reference soup fast-winding-number
// core.  Nothing of the reference is restated or copied here: the translation
// unit #includes the reference header where it lies
//   /root/reference/include/igl/FastWindingNumberForSoups.h  (UT_SolidAngle,
//   UT_BVH<4>; Eigen-free) and /root/reference/include/igl/parallel_for.h
// and is compiled by oracle/Makefile into oracle/_ref/libref_fwn.so.
//
// What the reference does with it (call sites this shim mirrors):
//   fast_winding_number.cpp:383-416  V->float U, F->row-major int, init(order)
//   fast_winding_number.cpp:421-437  parallel_for(#Q, computeSolidAngle/(4pi), 1000)
//
// `private` is re-defined so the tests can read UT_SolidAngle::myTree/myData
// (FastWindingNumberForSoups.h:6265-6276) and compare topology and per-node
// moments lane by lane with the device build.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <thread>
#include <chrono>
#define private public
#include <igl/FastWindingNumberForSoups.h>
#undef private
#include <igl/parallel_for.h>

using namespace igl::FastWindingNumber::HDK_Sample;

namespace {
struct RefFwn {
  std::vector<UT_Vector3T<float>> U;
  std::vector<int> F;
  UT_SolidAngle<float, float> sa;
};
}  // namespace

extern "C" {

// threads the genuine igl::parallel_for will use (default_num_threads.cpp:45-59)
int ref_num_threads() { return (int)igl::default_num_threads(); }

// U: nV x 3 row-major float;  F: nF x 3 row-major int32
void *ref_fwn_create(int nF, const int *F, int nV, const float *U, int order) {
  RefFwn *h = new RefFwn;
  h->U.resize(nV);
  for (int i = 0; i < nV; ++i)
    for (int j = 0; j < 3; ++j) h->U[i][j] = U[3 * i + j];
  h->F.assign(F, F + 3 * (size_t)nF);
  h->sa.init(nF, h->F.data(), nV, h->U.data(), order);
  return h;
}

void ref_fwn_destroy(void *p) { delete (RefFwn *)p; }

// raw solid angle (NOT divided by 4 pi), float, one per query; Q n x 3 row-major
void ref_fwn_solid_angle(void *p, const float *Q, int64_t n, float accuracy_scale,
                         float *omega) {
  RefFwn *h = (RefFwn *)p;
  igl::parallel_for(
      n,
      [&](int64_t i) {
        UT_Vector3T<float> q;
        q[0] = Q[3 * i];
        q[1] = Q[3 * i + 1];
        q[2] = Q[3 * i + 2];
        omega[i] = h->sa.computeSolidAngle(q, accuracy_scale);
      },
      1000);
}

// W = omega/(4 pi) evaluated in double exactly as fast_winding_number.cpp:435
void ref_fwn_winding(void *p, const float *Q, int64_t n, float accuracy_scale,
                     double *W) {
  RefFwn *h = (RefFwn *)p;
  igl::parallel_for(
      n,
      [&](int64_t i) {
        UT_Vector3T<float> q;
        q[0] = Q[3 * i];
        q[1] = Q[3 * i + 1];
        q[2] = Q[3 * i + 2];
        W[i] = h->sa.computeSolidAngle(q, accuracy_scale) / (4.0 * M_PI);
      },
      1000);
}

int ref_fwn_num_nodes(void *p) { return (int)((RefFwn *)p)->sa.myTree.getNumNodes(); }

// child[4] per node, reference encoding (FastWindingNumberForSoups.h:4217-4231)
void ref_fwn_nodes(void *p, uint32_t *out) {
  RefFwn *h = (RefFwn *)p;
  const int n = h->sa.myTree.getNumNodes();
  std::memcpy(out, h->sa.myTree.getNodes(), sizeof(uint32_t) * 4 * (size_t)n);
}

// BoxData per node as 23 rows of 4 lanes (FastWindingNumberForSoups.h:6477-6521)
void ref_fwn_boxdata(void *p, float *out) {
  RefFwn *h = (RefFwn *)p;
  const int n = h->sa.myTree.getNumNodes();
  static_assert(sizeof(h->sa.myData[0]) == 23 * 4 * sizeof(float), "BoxData layout");
  std::memcpy(out, h->sa.myData.get(), sizeof(float) * 92 * (size_t)n);
}

// standalone topology build from explicit boxes (nb x 6: xmin xmax ymin ymax zmin zmax)
int ref_bvh4_build(const float *boxes, int nb, uint32_t *nodes_out, int cap_nodes) {
  UT_BVH<4> t;
  t.init<UT::BVH_Heuristic::BOX_AREA, float, 3>((const UT::Box<float, 3> *)boxes,
                                                  (uint32_t)nb);
  int n = (int)t.getNumNodes();
  if (n <= cap_nodes) std::memcpy(nodes_out, t.getNodes(), sizeof(uint32_t) * 4 * (size_t)n);
  return n;
}

// the marching-cubes case tables exactly as the reference compiles them (march_cube.cpp:44 includes this
// header fragment inside the function body, as done here)
void ref_mc_tables(int *edge_flags256, int *edge_conn24, int *tri4096) {
#include <igl/marching_cubes_tables.h>
  for (int i = 0; i < 256; ++i) edge_flags256[i] = aiCubeEdgeFlags[i];
  for (int e = 0; e < 12; ++e) { edge_conn24[2 * e] = a2eConnection[e][0]; edge_conn24[2 * e + 1] = a2eConnection[e][1]; }
  for (int i = 0; i < 256; ++i) for (int k = 0; k < 16; ++k) tri4096[16 * i + k] = a2fConnectionTable[i][k];
}

}  // extern "C"
