// mc_core.h -- order-faithful replay of MeshUDF's gradient-oriented marching cubes.
//
// What it mirrors (reference, read-only):
//   meshudf/_marching_cubes_lewiner_cy.pyx:1115-1773  marching_cubes_udf (raster seed scan + BFS)
//   ...pyx:1776-1844  compute_edge_vote / my_sign / non_zero_norm / avg_cube / max_cube / dot3
//   ...pyx:422-464    Cell.set_cube            ...pyx:467-587  check_triangles / add_triangles
//   ...pyx:589-675    _add_face_from_edge_index ...pyx:677-761 get_index_in_facelayer
//   ...pyx:806-850    calculate_center_vertex   ...pyx:1847-2400 the_big_switch / check_the_big_switch
//   ...pyx:2403-2569  test_face / test_internal
//
// Design (B200): the reference walks all N^3 cubes on one CPU thread and keeps 4*N^3 int32 of
// vertex slots on the host.  Here the O(N^3) threshold scan is a separate HBM-bound CUDA kernel
// (mc_classify.cu) that emits a candidate bitmask + raster-sorted candidate list; this file is the
// O(surface) part: sign propagation + Lewiner triangulation, replayed in the reference's exact
// visiting order so vertex and face numbering come out identical.  All lattice state stays in HBM.
// The same source builds for the device (mc_replay.cu) and, for logic tests only, for the host
// (tests/ build it with g++ -ffp-contract=off); the arithmetic below is written so that both give
// the reference's IEEE results: no FMA contraction (nvcc -fmad=false), float votes accumulated
// through a double exactly as the Cython build does (see vote_accumulate()).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#define MC_HD_NOINLINE __host__ __device__ __noinline__
#define MC_LUT_STORAGE static __device__
#else
#define MC_HD inline
#define MC_HD_NOINLINE
#define MC_LUT_STORAGE static
#endif

namespace surfd_mccore {

#include "mc_luts.inc"

#if defined(__CUDACC__)
// Device: the 13 KB table blob is copied into shared memory once per kernel (mc_lut_load) -- every lookup on the
// order-dependent chain is then a ~25-cycle shared load instead of a global load that competes with the lattice traffic
// for L1.  (A function-scope __shared__ array is one object per kernel, whichever inlined copy names it.)
__device__ __forceinline__ signed char* mc_lut_smem() {
  __shared__ signed char lut[sizeof(MC_LUT_BLOB)];
  return lut;
}
__device__ __forceinline__ void mc_lut_load() {   // whole thread block
  signed char* d = mc_lut_smem();
  for (int i = threadIdx.x; i < (int)sizeof(MC_LUT_BLOB); i += blockDim.x) d[i] = MC_LUT_BLOB[i];
  __syncthreads();
}
#endif
#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
#define MC_LUTV(i) (mc_lut_smem()[i])
#define MC_LUTOFF(id) MC_LUT_OFF[id]
#define MC_LUTL1(id) MC_LUT_L1[id]
#define MC_LUTL2(id) MC_LUT_L2[id]
#elif defined(__CUDACC__)
// host pass of nvcc: the tables live in device memory only; host code never dereferences them.
#define MC_LUTV(i) ((signed char)0)
#define MC_LUTOFF(id) 0
#define MC_LUTL1(id) ((short)1)
#define MC_LUTL2(id) ((short)1)
#else
#define MC_LUTV(i) MC_LUT_BLOB[i]
#define MC_LUTOFF(id) MC_LUT_OFF[id]
#define MC_LUTL1(id) MC_LUT_L1[id]
#define MC_LUTL2(id) MC_LUT_L2[id]
#endif

// named-table accessors: the directory entries are compile-time constants (no dependent loads)
#define LUT1(name, i0) ((int)MC_LUTV(LUTOFF_##name + (i0)))
#define LUT2(name, i0, i1) ((int)MC_LUTV(LUTOFF_##name + (i0) * LUTL1_##name + (i1)))
#define LUT3(name, i0, i1, i2) ((int)MC_LUTV(LUTOFF_##name + ((i0) * LUTL1_##name + (i1)) * LUTL2_##name + (i2)))

MC_HD int lut1(int id, int i0) { return MC_LUTV(MC_LUTOFF(id) + i0); }
MC_HD int lut2(int id, int i0, int i1) { return MC_LUTV(MC_LUTOFF(id) + i0 * MC_LUTL1(id) + i1); }
MC_HD int lut3(int id, int i0, int i1, int i2) {
  return MC_LUTV(MC_LUTOFF(id) + (i0 * MC_LUTL1(id) + i1) * MC_LUTL2(id) + i2);
}

// status codes (also the C-ABI status of surfd_mc_udf)
enum { MC_OK = 0, MC_EMPTY = 1, MC_CAPACITY = 2, MC_QUEUE_OVERFLOW = 3 };

// FIFO of linear cube indices (the reference uses std::deque<(z,y,x)>, pyx:1164-1167).
struct Fifo {
  int32_t* buf;
  uint32_t mask;  // capacity-1, capacity is a power of two
  uint32_t head, tail;
  MC_HD bool empty() const { return head == tail; }
  MC_HD uint32_t size() const { return tail - head; }
  MC_HD bool push(int32_t v) {
    if (tail - head > mask) return false;
    buf[tail & mask] = v;
    ++tail;
    return true;
  }
  MC_HD int32_t front() const { return buf[head & mask]; }
  MC_HD void pop() { ++head; }
};

struct Grid {
  int N;
  const float* im;            // udf [N][N][N]   (axis order z,y,x in the pyx's naming)
  const float* grads;         // [N][N][N][3]
  const uint32_t* cand_bits;  // 1 bit per lattice index: cube anchored there passes avg/max thresholds
  const int32_t* cand_list;   // raster-sorted candidate cube indices
  int64_t n_cand;
  int64_t n_cand_total;       // device-side compaction total (may exceed the list capacity)
  int8_t* sgn;                // signed_im in {-1,0,+1}         (zero-initialised)
  uint8_t* flg;               // bit0 signed_im_mask, bit1 visited (zero-initialised)
  int32_t* face_layer;        // [4*N^3] vertex slot per (cell, edge slot), initialised to -1
  float* verts;               // [cap_v][3]  (x,y,z) = (axis2, axis1, axis0) index units, as the pyx emits
  int32_t* faces;             // [cap_f3] flat vertex indices, 3 per triangle
  int64_t cap_v, cap_f3;
  Fifo q, q_unsure, q_nontrivial;
  // results
  int64_t n_v, n_f3;
  int status;
  // statistics (diagnostics only)
  int64_t n_seed, n_accept, n_unsure_push, n_nontrivial_push;
  // cycle counters, filled only by builds with -DMC_PROFILE (device): 0 total, 1 neighbourhood fetch + edge votes,
  // 2 sign propagation, 3 case/tiling selection, 4 vertices + faces + pushes, 5 number of visits, 6 window refills
  int64_t prof[8];
};

struct Cell {
  int x, y, z;
  double v[8];    // v0..v7 in Lewiner corner order
  double vv[8];   // bit-ordered copy (index = dz*4+dy*2+dx)
  int index;
  bool v12_done;
  double v12x, v12y, v12z;
};

MC_HD float my_sign(float a) { return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f); }

MC_HD float dot3(const float* a, const float* b) {
  // float products and sums, left to right, no contraction (pyx:1843-1844)
  float p0 = a[0] * b[0];
  float p1 = a[1] * b[1];
  float p2 = a[2] * b[2];
  float s = p0 + p1;
  return s + p2;
}

MC_HD bool non_zero_norm(const float* a) {
  float s = (a[0] < 0 ? -a[0] : a[0]) + (a[1] < 0 ? -a[1] : a[1]);
  s = s + (a[2] < 0 ? -a[2] : a[2]);
  return s > 0.f;
}

// pyx:1776-1806.  dir is one of the six axis steps; g[0] pairs with the z axis, g[1] y, g[2] x.
MC_HD float edge_vote(const float* g1, const float* g2, int dz, int dy, int dx) {
  float p1, p2;
  if (dz != 0) { p1 = g1[0]; p2 = g2[0]; }
  else if (dy != 0) { p1 = g1[1]; p2 = g2[1]; }
  else { p1 = g1[2]; p2 = g2[2]; }
  if (dz + dy + dx > 0) {
    if (p2 > 0.f && p1 < 0.f) return 1.0f;
  } else {
    if (p2 < 0.f && p1 > 0.f) return 1.0f;
  }
  return dot3(g1, g2);
}

// The Cython build keeps sign_vs in a python array.array('f'): `sign_vs[i] += a*b` computes the
// float product, adds in double and rounds back to float on store (generated C++ for pyx:1302).
MC_HD float vote_accumulate(float acc, float sgn, float vote) {
  float prod = sgn * vote;
  return (float)((double)acc + (double)prod);
}

// The same in one float addition: the sum of two floats rounded to double (53 bits >= 2*24 + 2) and then to float equals
// the correctly rounded float sum (double rounding is innocuous for + - * / sqrt at that width), so the warp path skips
// the two conversions and the FP64 add.
MC_HD float vote_accumulate_f32(float acc, float sgn, float vote) {
  float prod = sgn * vote;
  return acc + prod;
}

#define MC_FLT_EPS 2.220446049250313e-16  /* np.spacing(1.0), pyx:35 */

MC_HD double dabs(double a) { return a >= 0 ? a : -a; }

MC_HD void cell_set(Cell& c, int x, int y, int z, const double* v) {
  c.x = x; c.y = y; c.z = z;
  int index = 0;
  for (int i = 0; i < 8; ++i) {
    c.v[i] = v[i];
    if (v[i] > 0.0) index |= (1 << i);
  }
  c.index = index;
  c.v12_done = false;
  // prepare_for_adding_triangles (pyx:763-781): bit-ordered copy
  c.vv[0] = v[0]; c.vv[1] = v[1]; c.vv[2] = v[3]; c.vv[3] = v[2];
  c.vv[4] = v[4]; c.vv[5] = v[5]; c.vv[6] = v[7]; c.vv[7] = v[6];
}

// pyx:677-761: slot of the vertex on edge `vi` (0..11) or the centre vertex (12)
MC_HD int64_t facelayer_index(const Grid& g, const Cell& c, int vi) {
  const int64_t nx = g.N;
  int64_t i = (int64_t)nx * nx * c.z + nx * c.y + c.x;
  int j = 0, k = 0;
  if (vi < 8) {
    if (vi >= 4) { vi -= 4; k = 1; }
    if (vi == 1) { i += 1; j = 1; }
    else if (vi == 2) { i += nx; }
    else if (vi == 3) { j = 1; }
  } else if (vi < 12) {
    j = 2;
    if (vi == 9) i += 1;
    else if (vi == 10) i += nx + 1;
    else if (vi == 11) i += nx;
  } else {
    j = 3;
  }
  i += nx * nx * k;
  return 4 * i + j;
}

MC_HD void center_vertex(Cell& c) {  // pyx:806-835 (gradient part dropped: normals are discarded by the caller)
  double w[8];
  for (int i = 0; i < 8; ++i) w[i] = 1.0 / (MC_FLT_EPS + dabs(c.v[i]));
  double fx = 0.0, fy = 0.0, fz = 0.0, ff = 0.0;
  // corner offsets in Lewiner order: v0(0,0,0) v1(1,0,0) v2(1,1,0) v3(0,1,0) v4(0,0,1) v5(1,0,1) v6(1,1,1) v7(0,1,1)
  const double ox[8] = {0, 1, 1, 0, 0, 1, 1, 0};
  const double oy[8] = {0, 0, 1, 1, 0, 0, 1, 1};
  const double oz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
  for (int i = 0; i < 8; ++i) {
    fx += ox[i] * w[i]; fy += oy[i] * w[i]; fz += oz[i] * w[i]; ff += w[i];
  }
  c.v12x = c.x + 1.0 * fx / ff;
  c.v12y = c.y + 1.0 * fy / ff;
  c.v12z = c.z + 1.0 * fz / ff;
  c.v12_done = true;
}

// Which tiling row the Lewiner switch selects (shared by the "check" and the "add" pass).
struct Tiling { int lut; int i1; int nt; int off, l1, l2; };  // i1 < 0: 2-D table; off/l1/l2: resolved directory entry

MC_HD bool test_face(const Cell& c, int face) {  // pyx:2403-2432
  int af = face < 0 ? -face : face;
  double A, B, C, D;
  const double* v = c.v;
  switch (af) {
    case 1: A = v[0]; B = v[4]; C = v[5]; D = v[1]; break;
    case 2: A = v[1]; B = v[5]; C = v[6]; D = v[2]; break;
    case 3: A = v[2]; B = v[6]; C = v[7]; D = v[3]; break;
    case 4: A = v[3]; B = v[7]; C = v[4]; D = v[0]; break;
    case 5: A = v[0]; B = v[3]; C = v[2]; D = v[1]; break;
    case 6: A = v[4]; B = v[7]; C = v[6]; D = v[5]; break;
    default: A = B = C = D = 0.0; break;  // (reference leaves them uninitialised; tables never hit this)
  }
  double ac = A * C;
  double bd = B * D;
  double AC_BD = ac - bd;
  if (AC_BD > -MC_FLT_EPS && AC_BD < MC_FLT_EPS) return face >= 0;
  double t = (double)face * A;
  t = t * AC_BD;
  return t >= 0;
}

MC_HD_NOINLINE bool test_internal(const Cell& c, int kase, int config, int subconfig, int s) {  // pyx:2435-2569
  const double* v = c.v;
  double t, At = 0.0, Bt = 0.0, Ct = 0.0, Dt = 0.0;
  int edge = -1;
  if (kase == 4 || kase == 10) {
    double a = (v[4] - v[0]) * (v[6] - v[2]) - (v[7] - v[3]) * (v[5] - v[1]);
    double b = v[2] * (v[4] - v[0]) + v[0] * (v[6] - v[2]) - v[1] * (v[7] - v[3]) - v[3] * (v[5] - v[1]);
    t = -b / (2 * a + MC_FLT_EPS);
    if (t < 0 || t > 1) return s > 0;
    At = v[0] + (v[4] - v[0]) * t;
    Bt = v[3] + (v[7] - v[3]) * t;
    Ct = v[2] + (v[6] - v[2]) * t;
    Dt = v[1] + (v[5] - v[1]) * t;
  } else {
    if (kase == 6) edge = LUT2(TEST6, config, 2);
    else if (kase == 7) edge = LUT2(TEST7, config, 4);
    else if (kase == 12) edge = LUT2(TEST12, config, 3);
    else if (kase == 13) edge = LUT3(TILING13_5_1, config, subconfig, 0);
    // per edge: t = va/(va - vb + eps); Bt,Ct,Dt interpolate three parallel edges
    // rows: {a, b, B0,B1, C0,C1, D0,D1}
    const signed char T[12][8] = {
        {0, 1, 3, 2, 7, 6, 4, 5}, {1, 2, 0, 3, 4, 7, 5, 6}, {2, 3, 1, 0, 5, 4, 6, 7}, {3, 0, 2, 1, 6, 5, 7, 4},
        {4, 5, 7, 6, 3, 2, 0, 1}, {5, 6, 4, 7, 0, 3, 1, 2}, {6, 7, 5, 4, 1, 0, 2, 3}, {7, 4, 6, 5, 2, 1, 3, 0},
        {0, 4, 3, 7, 2, 6, 1, 5}, {1, 5, 0, 4, 3, 7, 2, 6}, {2, 6, 1, 5, 0, 4, 3, 7}, {3, 7, 2, 6, 1, 5, 0, 4}};
    if (edge >= 0 && edge < 12) {
      const signed char* r = T[edge];
      t = v[r[0]] / (v[r[0]] - v[r[1]] + MC_FLT_EPS);
      At = 0;
      Bt = v[r[2]] + (v[r[3]] - v[r[2]]) * t;
      Ct = v[r[4]] + (v[r[5]] - v[r[4]]) * t;
      Dt = v[r[6]] + (v[r[7]] - v[r[6]]) * t;
    }
  }
  int test = 0;
  if (At >= 0) test += 1;
  if (Bt >= 0) test += 2;
  if (Ct >= 0) test += 4;
  if (Dt >= 0) test += 8;
  switch (test) {
    case 0: case 1: case 2: case 3: case 4: case 6: case 8: case 9: case 12: return s > 0;
    case 5: { double p = At * Ct; double q = Bt * Dt; if (p - q < MC_FLT_EPS) return s > 0; return false; }
    case 10: { double p = At * Ct; double q = Bt * Dt; if (p - q >= MC_FLT_EPS) return s > 0; return false; }
    default: return s < 0;  // 7, 11, 13, 14, 15
  }
}

// pyx:1847-2121 (and its twin :2124-2400): pick the tiling for (case, config) with the ambiguity tests.
MC_HD_NOINLINE Tiling select_tiling(const Cell& c, int kase, int config) {
  Tiling r; r.lut = -1; r.i1 = -1; r.nt = 0; r.off = 0; r.l1 = 1; r.l2 = 1;
  int sub = 0;
  switch (kase) {
    case 1: r.lut = LUT_TILING1; r.nt = 1; break;
    case 2: r.lut = LUT_TILING2; r.nt = 2; break;
    case 3:
      if (test_face(c, LUT1(TEST3, config))) { r.lut = LUT_TILING3_2; r.nt = 4; }
      else { r.lut = LUT_TILING3_1; r.nt = 2; }
      break;
    case 4:
      if (test_internal(c, kase, config, sub, LUT1(TEST4, config))) { r.lut = LUT_TILING4_1; r.nt = 2; }
      else { r.lut = LUT_TILING4_2; r.nt = 6; }
      break;
    case 5: r.lut = LUT_TILING5; r.nt = 3; break;
    case 6:
      if (test_face(c, LUT2(TEST6, config, 0))) { r.lut = LUT_TILING6_2; r.nt = 5; }
      else if (test_internal(c, kase, config, sub, LUT2(TEST6, config, 1))) { r.lut = LUT_TILING6_1_1; r.nt = 3; }
      else { r.lut = LUT_TILING6_1_2; r.nt = 9; }
      break;
    case 7:
      if (test_face(c, LUT2(TEST7, config, 0))) sub += 1;
      if (test_face(c, LUT2(TEST7, config, 1))) sub += 2;
      if (test_face(c, LUT2(TEST7, config, 2))) sub += 4;
      switch (sub) {
        case 0: r.lut = LUT_TILING7_1; r.nt = 3; break;
        case 1: r.lut = LUT_TILING7_2; r.i1 = 0; r.nt = 5; break;
        case 2: r.lut = LUT_TILING7_2; r.i1 = 1; r.nt = 5; break;
        case 3: r.lut = LUT_TILING7_3; r.i1 = 0; r.nt = 9; break;
        case 4: r.lut = LUT_TILING7_2; r.i1 = 2; r.nt = 5; break;
        case 5: r.lut = LUT_TILING7_3; r.i1 = 1; r.nt = 9; break;
        case 6: r.lut = LUT_TILING7_3; r.i1 = 2; r.nt = 9; break;
        default:
          if (test_internal(c, kase, config, sub, LUT2(TEST7, config, 3))) { r.lut = LUT_TILING7_4_2; r.nt = 9; }
          else { r.lut = LUT_TILING7_4_1; r.nt = 5; }
      }
      break;
    case 8: r.lut = LUT_TILING8; r.nt = 2; break;
    case 9: r.lut = LUT_TILING9; r.nt = 4; break;
    case 10:
      if (test_face(c, LUT2(TEST10, config, 0))) {
        if (test_face(c, LUT2(TEST10, config, 1))) { r.lut = LUT_TILING10_1_1_; r.nt = 4; }
        else { r.lut = LUT_TILING10_2; r.nt = 8; }
      } else {
        if (test_face(c, LUT2(TEST10, config, 1))) { r.lut = LUT_TILING10_2_; r.nt = 8; }
        else if (test_internal(c, kase, config, sub, LUT2(TEST10, config, 2))) { r.lut = LUT_TILING10_1_1; r.nt = 4; }
        else { r.lut = LUT_TILING10_1_2; r.nt = 8; }
      }
      break;
    case 11: r.lut = LUT_TILING11; r.nt = 4; break;
    case 12:
      if (test_face(c, LUT2(TEST12, config, 0))) {
        if (test_face(c, LUT2(TEST12, config, 1))) { r.lut = LUT_TILING12_1_1_; r.nt = 4; }
        else { r.lut = LUT_TILING12_2; r.nt = 8; }
      } else {
        if (test_face(c, LUT2(TEST12, config, 1))) { r.lut = LUT_TILING12_2_; r.nt = 8; }
        else if (test_internal(c, kase, config, sub, LUT2(TEST12, config, 2))) { r.lut = LUT_TILING12_1_1; r.nt = 4; }
        else { r.lut = LUT_TILING12_1_2; r.nt = 8; }
      }
      break;
    case 13: {
      for (int b = 0; b < 6; ++b)
        if (test_face(c, LUT2(TEST13, config, b))) sub += (1 << b);
      sub = LUT1(SUBCONFIG13, sub);
      if (sub == 0) { r.lut = LUT_TILING13_1; r.nt = 4; }
      else if (sub <= 6) { r.lut = LUT_TILING13_2; r.i1 = sub - 1; r.nt = 6; }
      else if (sub <= 18) { r.lut = LUT_TILING13_3; r.i1 = sub - 7; r.nt = 10; }
      else if (sub <= 22) { r.lut = LUT_TILING13_4; r.i1 = sub - 19; r.nt = 12; }
      else if (sub <= 26) {
        int s2 = sub - 23;
        if (test_internal(c, kase, config, s2, LUT2(TEST13, config, 6))) { r.lut = LUT_TILING13_5_1; r.i1 = s2; r.nt = 6; }
        else { r.lut = LUT_TILING13_5_2; r.i1 = s2; r.nt = 10; }
      }
      else if (sub <= 38) { r.lut = LUT_TILING13_3_; r.i1 = sub - 27; r.nt = 10; }
      else if (sub <= 44) { r.lut = LUT_TILING13_2_; r.i1 = sub - 39; r.nt = 6; }
      else if (sub == 45) { r.lut = LUT_TILING13_1_; r.nt = 4; }
      // else: "Impossible case 13" in the reference: nothing emitted
      break;
    }
    case 14: r.lut = LUT_TILING14; r.nt = 4; break;
    default: break;
  }
  if (r.lut >= 0) { r.off = MC_LUTOFF(r.lut); r.l1 = MC_LUTL1(r.lut); r.l2 = MC_LUTL2(r.lut); }
  return r;
}

MC_HD int tiling_edge(const Tiling& t, int config, int k) {
  return t.i1 < 0 ? (int)MC_LUTV(t.off + config * t.l1 + k) : (int)MC_LUTV(t.off + (config * t.l1 + t.i1) * t.l2 + k);
}

// pyx:467-526 check_triangles(2): number of distinct already-existing vertices among the tiling's
// vertex slots (first occurrence of each slot value counts; "-1" never counts).
MC_HD int check_tiling(const Grid& g, const Cell& c, const Tiling& t, int config) {
  int seen[36];
  int n = 0, result = 0;
  for (int k = 0; k < t.nt * 3; ++k) {
    int vi = tiling_edge(t, config, k);
    int fl = g.face_layer[facelayer_index(g, c, vi)];
    bool found = false;
    for (int m = 0; m < n; ++m) found = found || (seen[m] == fl);
    if (!found && fl >= 0) ++result;
    seen[n++] = fl;
  }
  return result;
}

// pyx:589-675: emit one face corner; create the vertex if its slot is still empty
MC_HD void add_face_from_edge(Grid& g, Cell& c, int vi) {
  int64_t slot = facelayer_index(g, c, vi);
  int idx = g.face_layer[slot];
  if (idx < 0) {
    double px, py, pz;
    if (vi == 12) {
      if (!c.v12_done) center_vertex(c);
      px = c.v12x; py = c.v12y; pz = c.v12z;
    } else {
      int dx1 = LUT2(EDGESRELX, vi, 0), dx2 = LUT2(EDGESRELX, vi, 1);
      int dy1 = LUT2(EDGESRELY, vi, 0), dy2 = LUT2(EDGESRELY, vi, 1);
      int dz1 = LUT2(EDGESRELZ, vi, 0), dz2 = LUT2(EDGESRELZ, vi, 1);
      double w1 = 1.0 / (MC_FLT_EPS + dabs(c.vv[dz1 * 4 + dy1 * 2 + dx1]));
      double w2 = 1.0 / (MC_FLT_EPS + dabs(c.vv[dz2 * 4 + dy2 * 2 + dx2]));
      double fx = 0.0, fy = 0.0, fz = 0.0, ff = 0.0;
      fx += (double)dx1 * w1; fy += (double)dy1 * w1; fz += (double)dz1 * w1; ff += w1;
      fx += (double)dx2 * w2; fy += (double)dy2 * w2; fz += (double)dz2 * w2; ff += w2;
      px = (double)c.x + 1.0 * fx / ff;
      py = (double)c.y + 1.0 * fy / ff;
      pz = (double)c.z + 1.0 * fz / ff;
    }
    idx = (int)g.n_v;
    if (g.n_v < g.cap_v) {
      g.verts[3 * g.n_v + 0] = (float)px;
      g.verts[3 * g.n_v + 1] = (float)py;
      g.verts[3 * g.n_v + 2] = (float)pz;
    } else {
      g.status = MC_CAPACITY;  // keep counting so the caller learns the required size
    }
    ++g.n_v;
    g.face_layer[slot] = idx;
  } else if (vi == 12 && !c.v12_done) {
    center_vertex(c);
  }
  if (g.n_f3 < g.cap_f3) g.faces[g.n_f3] = idx;
  else g.status = MC_CAPACITY;
  ++g.n_f3;
}

MC_HD void add_tiling(Grid& g, Cell& c, const Tiling& t, int config) {
  for (int k = 0; k < t.nt * 3; ++k) add_face_from_edge(g, c, tiling_edge(t, config, k));
}

MC_HD int64_t lin(const Grid& g, int z, int y, int x) { return ((int64_t)z * g.N + y) * g.N + x; }

MC_HD bool is_candidate(const Grid& g, int64_t i) { return (g.cand_bits[i >> 5] >> (i & 31)) & 1u; }

MC_HD void push_neighbours(Grid& g, int z, int y, int x) {  // pyx:1407-1418 order
  const int nb = g.N - 2;  // N{x,y,z}_bound
  bool ok = true;
  if (x + 1 < nb) ok = g.q.push((int32_t)lin(g, z, y, x + 1)) && ok;
  if (y + 1 < nb) ok = g.q.push((int32_t)lin(g, z, y + 1, x)) && ok;
  if (x - 1 >= 0) ok = g.q.push((int32_t)lin(g, z, y, x - 1)) && ok;
  if (y - 1 >= 0) ok = g.q.push((int32_t)lin(g, z, y - 1, x)) && ok;
  if (z - 1 >= 0) ok = g.q.push((int32_t)lin(g, z - 1, y, x)) && ok;
  if (z + 1 < nb) ok = g.q.push((int32_t)lin(g, z + 1, y, x)) && ok;
  if (!ok) g.status = MC_QUEUE_OVERFLOW;
}

// Corner offsets in Lewiner order (vertex_index_array_{z,y,x}, pyx:1220-1222)
#define MC_CZ(i) (((i) >> 2) & 1)
#define MC_CY(i) ((((i) & 3) >> 1))
#define MC_CX(i) ((((i) & 3) == 1 || ((i) & 3) == 2) ? 1 : 0)

// One visit of cube (z,y,x).  mode: 0 = raster seed, 1 = BFS (visit_neighbours flag True),
// 2 = BFS while serving an unsure cube's neighbours (flag False).
// Returns true when the cube was accepted and produced faces (seed: start a BFS).
MC_HD_NOINLINE bool visit_cube(Grid& g, int z, int y, int x, int mode) {
  const int N = g.N;
  const int nb = N - 2;
  int64_t ci[8];
  float cim[8];
  for (int i = 0; i < 8; ++i) {
    ci[i] = lin(g, z + MC_CZ(i), y + MC_CY(i), x + MC_CX(i));
    cim[i] = g.im[ci[i]];
  }
  int visited_vs[8];
  float sign_vs[8];
  for (int vtx = 0; vtx < 8; ++vtx) {
    visited_vs[vtx] = 0;
    sign_vs[vtx] = 0.0f;
    const int64_t c0 = ci[vtx];
    if (g.flg[c0] & 1) {
      visited_vs[vtx] = 1;
      sign_vs[vtx] = (float)g.sgn[c0];
      continue;
    }
    if (cim[vtx] == 0.0f) {
      visited_vs[vtx] = 1;
      continue;
    }
    const int zi = z + MC_CZ(vtx), yi = y + MC_CY(vtx), xi = x + MC_CX(vtx);
    const float* g1 = g.grads + 3 * c0;
    for (int d = 0; d < 6; ++d) {
      const int dz = (d == 0) - (d == 1), dy = (d == 2) - (d == 3), dx = (d == 4) - (d == 5);
      int i = 0, maxd = 1;
      while (i < maxd) {
        ++i;
        const int cz = zi + i * dz, cy = yi + i * dy, cx = xi + i * dx;
        if (cz > nb || cz < 0 || cy > nb || cy < 0 || cx > nb || cx < 0) break;
        const int64_t cn = lin(g, cz, cy, cx);
        if (g.im[cn] == 0.0f) {
          if (i < maxd) continue;
          ++maxd;
          continue;
        }
        const int8_t sn = g.sgn[cn];
        if (sn == 0) continue;
        visited_vs[vtx] += 1;
        sign_vs[vtx] = vote_accumulate(sign_vs[vtx], (float)sn, edge_vote(g1, g.grads + 3 * cn, dz, dy, dx));
      }
    }
    if (mode != 0) {
      // pyx:1584: python-object arithmetic => double division, compared with double(0.707f)
      if (visited_vs[vtx] >= 1 &&
          (double)(sign_vs[vtx] < 0 ? -sign_vs[vtx] : sign_vs[vtx]) / (double)visited_vs[vtx] < (double)0.707f &&
          !g.q.empty()) {
        if (mode == 1) {
          if (!g.q_unsure.push((int32_t)lin(g, z, y, x))) g.status = MC_QUEUE_OVERFLOW;
          ++g.n_unsure_push;
        }
        return false;
      }
    }
    g.sgn[c0] = (int8_t)my_sign(sign_vs[vtx]);
  }

  bool all_voted = true;
  for (int i = 0; i < 8; ++i) all_voted = all_voted && (visited_vs[i] >= 1);
  if (!all_voted) {
    // anchor gradient: first corner (order 0,1,3,2,4,5,7,6 in Lewiner numbering == z,y,x bit order)
    // that is already fixed and has a non-zero gradient; else first with a non-zero gradient (pyx:1310-1346)
    const int order[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    float base[3] = {0.f, 0.f, 0.f};
    float anchor_sign = 1.f;
    bool found = false;
    for (int k = 0; k < 8 && !found; ++k) {
      const int64_t c0 = ci[order[k]];
      if ((g.flg[c0] & 1) && non_zero_norm(g.grads + 3 * c0)) {
        anchor_sign = my_sign((float)g.sgn[c0]);
        base[0] = g.grads[3 * c0]; base[1] = g.grads[3 * c0 + 1]; base[2] = g.grads[3 * c0 + 2];
        found = true;
      }
    }
    for (int k = 0; k < 8 && !found; ++k) {
      const int64_t c0 = ci[order[k]];
      if (non_zero_norm(g.grads + 3 * c0)) {
        base[0] = g.grads[3 * c0]; base[1] = g.grads[3 * c0 + 1]; base[2] = g.grads[3 * c0 + 2];
        found = true;
      }
    }
    // (reference prints 'all 0 vec...' and keeps the previous base_vec; with a fresh buffer that is
    //  uninitialised memory -- we use zeros, which gives sign 0 for the unvoted corners)
    base[0] = anchor_sign * base[0]; base[1] = anchor_sign * base[1]; base[2] = anchor_sign * base[2];
    const bool check_unsure = (mode == 1) && !g.q.empty();
    for (int i = 0; i < 8; ++i) {
      if (visited_vs[i] != 0) continue;
      const float s = dot3(base, g.grads + 3 * ci[i]);
      if (check_unsure) {
        sign_vs[i] = s;
        if ((s < 0 ? -s : s) < 0.707f) {
          if (!g.q_unsure.push((int32_t)lin(g, z, y, x))) g.status = MC_QUEUE_OVERFLOW;
          ++g.n_unsure_push;
          return false;
        }
      }
      g.sgn[ci[i]] = (int8_t)my_sign(s);
    }
  }

  if (mode == 2) return false;  // neighbours of an unsure cube: tentative signs only

  double v[8];
  for (int i = 0; i < 8; ++i) {
    float p = (float)g.sgn[ci[i]] * cim[i];
    v[i] = (double)p;
  }
  Cell c;
  cell_set(c, x, y, z, v);
  for (int i = 0; i < 8; ++i) g.flg[ci[i]] |= 1;

  const int kase = LUT2(CASES, c.index, 0);
  const int64_t me = lin(g, z, y, x);
  if (kase > 0) {
    if (mode == 1) {
      const bool trivial = (kase == 1 || kase == 2 || kase == 5 || kase == 8 || kase == 9);
      if (!trivial && (!g.q.empty() || !g.q_unsure.empty())) {
        if (!g.q_nontrivial.push((int32_t)me)) g.status = MC_QUEUE_OVERFLOW;
        ++g.n_nontrivial_push;
        return false;
      }
    }
    const int config = LUT2(CASES, c.index, 1);
    const Tiling t = select_tiling(c, kase, config);
    if (mode == 1) {
      if (check_tiling(g, c, t, config) < 2) return false;
    }
    g.flg[me] |= 2;
    add_tiling(g, c, t, config);
    push_neighbours(g, z, y, x);
    ++g.n_accept;
    return true;
  }
  g.flg[me] |= 2;
  return false;
}

// pyx:1194-1771: raster scan over the (raster-sorted) candidate list; each still-unvisited candidate
// seeds a breadth-first exploration with the reference's three priority queues.
MC_HD_NOINLINE void replay(Grid& g) {
  const int N = g.N;
  g.n_v = 0; g.n_f3 = 0; g.status = MC_OK;
  g.n_seed = g.n_accept = g.n_unsure_push = g.n_nontrivial_push = 0;
  for (int64_t k = 0; k < g.n_cand; ++k) {
    const int32_t cidx = g.cand_list[k];
    if (g.flg[cidx] & 2) continue;
    int x = cidx % N, y = (cidx / N) % N, z = cidx / (N * N);
    ++g.n_seed;
    if (!visit_cube(g, z, y, x, 0)) continue;
    bool visit_neighbours = true;
    while (!g.q.empty() || !g.q_unsure.empty() || !g.q_nontrivial.empty()) {
      if (g.status == MC_QUEUE_OVERFLOW) return;
      int32_t cur;
      if (g.q.empty()) {
        if (g.q_unsure.empty()) {
          cur = g.q_nontrivial.front(); g.q_nontrivial.pop();
        } else {
          cur = g.q_unsure.front();
          if (visit_neighbours) {
            if (g.flg[cur] & 2) { g.q_unsure.pop(); continue; }
            push_neighbours(g, cur / (N * N), (cur / N) % N, cur % N);
            visit_neighbours = false;
            continue;
          } else {
            g.q_unsure.pop();
            visit_neighbours = true;
          }
        }
      } else {
        cur = g.q.front(); g.q.pop();
      }
      if (g.flg[cur] & 2) continue;
      if (!is_candidate(g, cur)) continue;
      visit_cube(g, cur / (N * N), (cur / N) % N, cur % N, visit_neighbours ? 1 : 2);
    }
  }
  if (g.status == MC_OK && g.n_v == 0) g.status = MC_EMPTY;
}

// =====================================================================================================
// Warp-cooperative variant (what the replay kernel runs).  Same visiting order and arithmetic as
// visit_cube()/replay() above; the difference is how a visit touches memory.  All 32 lanes execute the
// scalar control flow redundantly on a per-lane copy of the Grid bookkeeping (identical values, same-value
// stores), and the lanes split only the memory-heavy, order-independent parts:
//   phase 1  the 4x4x4 lattice neighbourhood of the cube (udf, sign, flags, gradients) and its 13 vertex
//            slots are fetched with one load per lane into a shared-memory CubeCache (one L2 latency
//            instead of ~50 dependent ones);
//   phase 2  the 8x6 edge votes (pure functions of the gradients) are computed one per lane;
//   then the order-dependent chain (corner by corner, direction by direction, vote_accumulate) runs
//   from the cache.  Cubes that need the reference's "look one vertex further past an exact zero" rule
//   fall back to visit_cube().  On the host (logic tests) the lane loops run sequentially.
// =====================================================================================================
#if defined(__CUDA_ARCH__)
#define MC_UNROLL _Pragma("unroll")
#else
#define MC_UNROLL
#endif
#if defined(__CUDA_ARCH__)
#define MC_LANE_LOOP(l) for (int l = (int)(threadIdx.x & 31), _mc_once = 1; _mc_once; _mc_once = 0)
#define MC_WARP_SYNC() __syncwarp()
#else
#define MC_LANE_LOOP(l) for (int l = 0; l < 32; ++l)
#define MC_WARP_SYNC()
#endif

struct CubeCache {
  float im[64];
  float gr[64 * 3];
  float vote[48];
  int32_t fl[13];
  int8_t sgn[64];
  uint8_t flg[64];
  uint8_t vstat[48];   // 0: skipped by the bounds rule, 1: usable, 2: neighbour udf == 0 (needs the extension rule)
  int8_t nsgn[48];     // sign of the neighbour vertex of (corner, direction) as fetched
  // look-ahead windows: the next <= 32 entries of the BFS queue / of the raster candidate list, one per lane, with
  // bit0 = "visited flag already set" (kept coherent by note_done()), bit1 = "is a candidate cube"
  int32_t qw_cur[32];
  int32_t sw_cur[32];
  uint8_t qw_f[32];
  uint8_t sw_f[32];
  uint8_t tedge[36];   // edge ids of the selected tiling
};

#if defined(__CUDA_ARCH__) && defined(MC_PROFILE)
#define MC_PROF_T(var) const long long var = clock64()
#define MC_PROF_ADD(slot, t0, t1) g.prof[slot] += (t1) - (t0)
#define MC_PROF_INC(slot) g.prof[slot] += 1
#else
#define MC_PROF_T(var)
#define MC_PROF_ADD(slot, t0, t1)
#define MC_PROF_INC(slot)
#endif

#if defined(__CUDA_ARCH__)
#define MC_PREFETCH(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#else
#define MC_PREFETCH(p) ((void)(p))
#endif

MC_HD int64_t facelayer_index_xyz(int64_t nx, int x, int y, int z, int vi) {
  int64_t i = nx * nx * z + nx * y + x;
  int j = 0, k = 0;
  if (vi < 8) {
    if (vi >= 4) { vi -= 4; k = 1; }
    if (vi == 1) { i += 1; j = 1; }
    else if (vi == 2) { i += nx; }
    else if (vi == 3) { j = 1; }
  } else if (vi < 12) {
    j = 2;
    if (vi == 9) i += 1;
    else if (vi == 10) i += nx + 1;
    else if (vi == 11) i += nx;
  } else {
    j = 3;
  }
  i += nx * nx * k;
  return 4 * i + j;
}

#define MC_BLK(bz, by, bx) (((bz) << 4) | ((by) << 2) | (bx))
// Lewiner corner id from its (z, y, x) bits: inverse of MC_CZ / MC_CY / MC_CX
#define MC_CORNER(cz, cy, cx) (((cz) << 2) | ((cy) ? ((cx) ? 2 : 3) : ((cx) ? 1 : 0)))

// check_tiling() on the cached slots.  A vertex index lives in exactly one face_layer slot, and the 13 edge ids of a
// cube map to 13 distinct slots, so "distinct existing vertex indices" == "distinct edge ids whose slot is filled".
MC_HD int check_tiling_c(const CubeCache& cc, const Tiling& t, int config) {
  uint32_t seen = 0;
  int result = 0;
  for (int k = 0; k < t.nt * 3; ++k) {
    const int e = tiling_edge(t, config, k);
    if (!((seen >> e) & 1u) && cc.fl[e] >= 0) ++result;
    seen |= 1u << e;
  }
  return result;
}

MC_HD void add_face_from_edge_c(Grid& g, CubeCache& cc, Cell& c, int vi) {
  int idx = cc.fl[vi];
  if (idx < 0) {
    double px, py, pz;
    if (vi == 12) {
      if (!c.v12_done) center_vertex(c);
      px = c.v12x; py = c.v12y; pz = c.v12z;
    } else {
      int dx1 = LUT2(EDGESRELX, vi, 0), dx2 = LUT2(EDGESRELX, vi, 1);
      int dy1 = LUT2(EDGESRELY, vi, 0), dy2 = LUT2(EDGESRELY, vi, 1);
      int dz1 = LUT2(EDGESRELZ, vi, 0), dz2 = LUT2(EDGESRELZ, vi, 1);
      double w1 = 1.0 / (MC_FLT_EPS + dabs(c.vv[dz1 * 4 + dy1 * 2 + dx1]));
      double w2 = 1.0 / (MC_FLT_EPS + dabs(c.vv[dz2 * 4 + dy2 * 2 + dx2]));
      double fx = 0.0, fy = 0.0, fz = 0.0, ff = 0.0;
      fx += (double)dx1 * w1; fy += (double)dy1 * w1; fz += (double)dz1 * w1; ff += w1;
      fx += (double)dx2 * w2; fy += (double)dy2 * w2; fz += (double)dz2 * w2; ff += w2;
      // the two end points of an edge differ along one axis only: along the other two the numerator is 0 or is the very
      // same sum as ff, so those quotients are exactly 0 or 1 and only one FP64 division is needed
      const double qx = fx == 0.0 ? 0.0 : (fx == ff ? 1.0 : 1.0 * fx / ff);
      const double qy = fy == 0.0 ? 0.0 : (fy == ff ? 1.0 : 1.0 * fy / ff);
      const double qz = fz == 0.0 ? 0.0 : (fz == ff ? 1.0 : 1.0 * fz / ff);
      px = (double)c.x + qx;
      py = (double)c.y + qy;
      pz = (double)c.z + qz;
    }
    idx = (int)g.n_v;
    if (g.n_v < g.cap_v) {
      g.verts[3 * g.n_v + 0] = (float)px;
      g.verts[3 * g.n_v + 1] = (float)py;
      g.verts[3 * g.n_v + 2] = (float)pz;
    } else {
      g.status = MC_CAPACITY;
    }
    ++g.n_v;
    cc.fl[vi] = idx;
    g.face_layer[facelayer_index_xyz(g.N, c.x, c.y, c.z, vi)] = idx;
  } else if (vi == 12 && !c.v12_done) {
    center_vertex(c);
  }
  if (g.n_f3 < g.cap_f3) g.faces[g.n_f3] = idx;
  else g.status = MC_CAPACITY;
  ++g.n_f3;
}

// The mutable part of a Grid, passed by value to the generic path: the warp path's own `Grid` never has its address
// taken, so the compiler keeps it in registers instead of local memory.
struct GridState {
  uint32_t qh, qt, uh, ut, nh, nt;
  int status;
  int64_t n_v, n_f3, n_seed, n_accept, n_unsure_push, n_nontrivial_push;
};
MC_HD GridState grid_state_get(const Grid& g) {
  GridState s;
  s.qh = g.q.head; s.qt = g.q.tail; s.uh = g.q_unsure.head; s.ut = g.q_unsure.tail;
  s.nh = g.q_nontrivial.head; s.nt = g.q_nontrivial.tail; s.status = g.status;
  s.n_v = g.n_v; s.n_f3 = g.n_f3; s.n_seed = g.n_seed; s.n_accept = g.n_accept;
  s.n_unsure_push = g.n_unsure_push; s.n_nontrivial_push = g.n_nontrivial_push;
  return s;
}
MC_HD void grid_state_put(Grid& g, const GridState& s) {
  g.q.head = s.qh; g.q.tail = s.qt; g.q_unsure.head = s.uh; g.q_unsure.tail = s.ut;
  g.q_nontrivial.head = s.nh; g.q_nontrivial.tail = s.nt; g.status = s.status;
  g.n_v = s.n_v; g.n_f3 = s.n_f3; g.n_seed = s.n_seed; g.n_accept = s.n_accept;
  g.n_unsure_push = s.n_unsure_push; g.n_nontrivial_push = s.n_nontrivial_push;
}
// `home` holds the constant fields (sizes, pointers, queue buffers); the state travels in `gs`.
MC_HD_NOINLINE bool visit_cube_generic(const Grid* home, GridState& gs, int z, int y, int x, int mode) {
  Grid tmp = *home;
  grid_state_put(tmp, gs);
  const bool r = visit_cube(tmp, z, y, x, mode);
  gs = grid_state_get(tmp);
  return r;
}

MC_HD bool visit_cube_w(Grid& g, CubeCache& cc, const Grid* home, int z, int y, int x, int mode, bool& done_set) {
  const int N = g.N;
  done_set = false;
  const int nb = N - 2;
  MC_PROF_T(t_begin);
  MC_PROF_INC(5);
  // ---- phase 1: cooperative fetch of the 4x4x4 neighbourhood (origin z-1,y-1,x-1) and the 13 vertex slots ----
  MC_WARP_SYNC();
  MC_LANE_LOOP(l) {
    // two lattice vertices per lane (v = l, l + 32).  Indices are clamped into the lattice so that all 12 loads (+ the
    // vertex slot) can be issued back to back -- one memory latency per visit -- and out-of-lattice entries are zeroed
    // afterwards.
    int64_t li[2];
    bool inside[2];
    MC_UNROLL
    for (int j = 0; j < 2; ++j) {
      const int v = l + 32 * j;
      const int cz = z - 1 + (v >> 4), cy = y - 1 + ((v >> 2) & 3), cx = x - 1 + (v & 3);
      inside[j] = cz >= 0 && cz < N && cy >= 0 && cy < N && cx >= 0 && cx < N;
      const int qz = cz < 0 ? 0 : (cz >= N ? N - 1 : cz), qy = cy < 0 ? 0 : (cy >= N ? N - 1 : cy), qx = cx < 0 ? 0 : (cx >= N ? N - 1 : cx);
      li[j] = lin(g, qz, qy, qx);
    }
    const float im0 = g.im[li[0]], im1 = g.im[li[1]];
    const int8_t sg0 = g.sgn[li[0]], sg1 = g.sgn[li[1]];
    const uint8_t fg0 = g.flg[li[0]], fg1 = g.flg[li[1]];
    const float a0 = g.grads[3 * li[0]], a1 = g.grads[3 * li[0] + 1], a2 = g.grads[3 * li[0] + 2];
    const float b0 = g.grads[3 * li[1]], b1 = g.grads[3 * li[1] + 1], b2 = g.grads[3 * li[1] + 2];
    const int32_t slot = g.face_layer[facelayer_index_xyz(N, x, y, z, l < 13 ? l : 0)];
    cc.im[l] = inside[0] ? im0 : 0.f; cc.sgn[l] = inside[0] ? sg0 : (int8_t)0; cc.flg[l] = inside[0] ? fg0 : (uint8_t)0;
    cc.gr[3 * l] = inside[0] ? a0 : 0.f; cc.gr[3 * l + 1] = inside[0] ? a1 : 0.f; cc.gr[3 * l + 2] = inside[0] ? a2 : 0.f;
    const int l1 = l + 32;
    cc.im[l1] = inside[1] ? im1 : 0.f; cc.sgn[l1] = inside[1] ? sg1 : (int8_t)0; cc.flg[l1] = inside[1] ? fg1 : (uint8_t)0;
    cc.gr[3 * l1] = inside[1] ? b0 : 0.f; cc.gr[3 * l1 + 1] = inside[1] ? b1 : 0.f; cc.gr[3 * l1 + 2] = inside[1] ? b2 : 0.f;
    if (l < 13) cc.fl[l] = slot;
  }
  MC_WARP_SYNC();
  // ---- phase 2: one (corner, direction) edge vote per lane ----
  MC_LANE_LOOP(l) {
    for (int it = l; it < 48; it += 32) {
      const int c = it / 6, d = it - 6 * c;
      const int dz = (d == 0) - (d == 1), dy = (d == 2) - (d == 3), dx = (d == 4) - (d == 5);
      const int bz = 1 + MC_CZ(c), by = 1 + MC_CY(c), bx = 1 + MC_CX(c);
      const int cz = z - 1 + bz + dz, cy = y - 1 + by + dy, cx = x - 1 + bx + dx;
      uint8_t st = 0; float vt = 0.f;
      if (!(cz > nb || cz < 0 || cy > nb || cy < 0 || cx > nb || cx < 0)) {
        const int nv = MC_BLK(bz + dz, by + dy, bx + dx);
        if (cc.im[nv] == 0.0f) st = 2;
        else { st = 1; vt = edge_vote(cc.gr + 3 * MC_BLK(bz, by, bx), cc.gr + 3 * nv, dz, dy, dx); }
      }
      cc.vstat[it] = st; cc.vote[it] = vt;
      cc.nsgn[it] = st ? cc.sgn[MC_BLK(bz + dz, by + dy, bx + dx)] : (int8_t)0;   // as fetched; in-cube neighbours are tracked in registers
    }
  }
  MC_WARP_SYNC();
  MC_PROF_T(t_fetched);
  MC_PROF_ADD(1, t_begin, t_fetched);
  // ---- uniform part ----
  int cb[8];
  float cim[8];
  int64_t ci[8];
  MC_UNROLL
  for (int i = 0; i < 8; ++i) {
    cb[i] = MC_BLK(1 + MC_CZ(i), 1 + MC_CY(i), 1 + MC_CX(i));
    cim[i] = cc.im[cb[i]];
    ci[i] = lin(g, z + MC_CZ(i), y + MC_CY(i), x + MC_CX(i));
  }
  // the "exact zero neighbour" extension (pyx:1287-1292) reaches outside the cached block: generic path
  MC_UNROLL
  for (int i = 0; i < 8; ++i) {
    if ((cc.flg[cb[i]] & 1) || cim[i] == 0.0f) continue;
    MC_UNROLL
    for (int d = 0; d < 6; ++d)
      if (cc.vstat[i * 6 + d] == 2) {
        GridState gs = grid_state_get(g);
        const bool r = visit_cube_generic(home, gs, z, y, x, mode);
        grid_state_put(g, gs);
        done_set = (g.flg[lin(g, z, y, x)] & 2) != 0;
        return r;
      }
  }
  // The corner chain runs on registers: the signs / flags of the cube's own 8 corners live in csgn/cflg (updated as
  // corners are decided), and each corner's six (status, vote, outside-neighbour sign) triples are loaded together
  // before its accumulation loop -- one shared-memory latency per corner instead of three dependent ones per direction.
  int8_t csgn[8];
  uint8_t cflg[8];
  MC_UNROLL
  for (int i = 0; i < 8; ++i) { csgn[i] = cc.sgn[cb[i]]; cflg[i] = cc.flg[cb[i]]; }
  int visited_vs[8];
  float sign_vs[8];
  MC_UNROLL
  for (int vtx = 0; vtx < 8; ++vtx) {
    visited_vs[vtx] = 0;
    sign_vs[vtx] = 0.0f;
    if (cflg[vtx] & 1) {
      visited_vs[vtx] = 1;
      sign_vs[vtx] = (float)csgn[vtx];
      continue;
    }
    if (cim[vtx] == 0.0f) {
      visited_vs[vtx] = 1;
      continue;
    }
    uint8_t st6[6];
    float vt6[6];
    int8_t ns6[6];
    MC_UNROLL
    for (int d = 0; d < 6; ++d) { st6[d] = cc.vstat[vtx * 6 + d]; vt6[d] = cc.vote[vtx * 6 + d]; ns6[d] = cc.nsgn[vtx * 6 + d]; }
    MC_UNROLL
    for (int d = 0; d < 6; ++d) {
      if (st6[d] != 1) continue;
      // a step along an axis stays inside the cube iff it flips that axis' corner bit from 0 to 1 (or back)
      const int cz = MC_CZ(vtx), cy = MC_CY(vtx), cx = MC_CX(vtx);
      const int nz = cz + (d == 0) - (d == 1), ny = cy + (d == 2) - (d == 3), nx = cx + (d == 4) - (d == 5);
      const bool in_cube = nz >= 0 && nz <= 1 && ny >= 0 && ny <= 1 && nx >= 0 && nx <= 1;
      const int8_t sn = in_cube ? csgn[MC_CORNER(nz & 1, ny & 1, nx & 1)] : ns6[d];
      if (sn == 0) continue;
      visited_vs[vtx] += 1;
      sign_vs[vtx] = vote_accumulate_f32(sign_vs[vtx], (float)sn, vt6[d]);
    }
    if (mode != 0) {
      // |sign| / visited < 0.707f, evaluated without the FP64 division: 0.707f * visited is exact in double (24 + 3 bits),
      // and the rounded quotient can only differ from the exact one inside half an ulp53 of the threshold, which a
      // 24-bit numerator over visited <= 6 never reaches unless it equals the product (then both tests are false).
      if (visited_vs[vtx] >= 1 &&
          (double)(sign_vs[vtx] < 0 ? -sign_vs[vtx] : sign_vs[vtx]) < (double)0.707f * (double)visited_vs[vtx] &&
          !g.q.empty()) {
        if (mode == 1) {
          if (!g.q_unsure.push((int32_t)lin(g, z, y, x))) g.status = MC_QUEUE_OVERFLOW;
          ++g.n_unsure_push;
        }
        return false;
      }
    }
    const int8_t ns = (int8_t)my_sign(sign_vs[vtx]);
    csgn[vtx] = ns;
    g.sgn[ci[vtx]] = ns;
  }

  bool all_voted = true;
  MC_UNROLL
  for (int i = 0; i < 8; ++i) all_voted = all_voted && (visited_vs[i] >= 1);
  if (!all_voted) {
    const int order[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    float base[3] = {0.f, 0.f, 0.f};
    float anchor_sign = 1.f;
    bool found = false;
    MC_UNROLL
    for (int k = 0; k < 8; ++k) {
      if (found) continue;
      const int b0 = cb[order[k]];
      if ((cflg[order[k]] & 1) && non_zero_norm(cc.gr + 3 * b0)) {
        anchor_sign = my_sign((float)csgn[order[k]]);
        base[0] = cc.gr[3 * b0]; base[1] = cc.gr[3 * b0 + 1]; base[2] = cc.gr[3 * b0 + 2];
        found = true;
      }
    }
    MC_UNROLL
    for (int k = 0; k < 8; ++k) {
      if (found) continue;
      const int b0 = cb[order[k]];
      if (non_zero_norm(cc.gr + 3 * b0)) {
        base[0] = cc.gr[3 * b0]; base[1] = cc.gr[3 * b0 + 1]; base[2] = cc.gr[3 * b0 + 2];
        found = true;
      }
    }
    base[0] = anchor_sign * base[0]; base[1] = anchor_sign * base[1]; base[2] = anchor_sign * base[2];
    const bool check_unsure = (mode == 1) && !g.q.empty();
    MC_UNROLL
    for (int i = 0; i < 8; ++i) {
      if (visited_vs[i] != 0) continue;
      const float s = dot3(base, cc.gr + 3 * cb[i]);
      if (check_unsure) {
        sign_vs[i] = s;
        if ((s < 0 ? -s : s) < 0.707f) {
          if (!g.q_unsure.push((int32_t)lin(g, z, y, x))) g.status = MC_QUEUE_OVERFLOW;
          ++g.n_unsure_push;
          return false;
        }
      }
      const int8_t ns = (int8_t)my_sign(s);
      csgn[i] = ns;
      g.sgn[ci[i]] = ns;
    }
  }

  MC_PROF_T(t_signed);
  MC_PROF_ADD(2, t_fetched, t_signed);
  if (mode == 2) return false;

  double v[8];
  MC_UNROLL
  for (int i = 0; i < 8; ++i) {
    float p = (float)csgn[i] * cim[i];
    v[i] = (double)p;
  }
  Cell c;
  cell_set(c, x, y, z, v);
  MC_UNROLL
  for (int i = 0; i < 8; ++i) g.flg[ci[i]] = (uint8_t)(cflg[i] | 1);

  const int kase = LUT2(CASES, c.index, 0);
  const int64_t me = ci[0];
  if (kase > 0) {
    if (mode == 1) {
      const bool trivial = (kase == 1 || kase == 2 || kase == 5 || kase == 8 || kase == 9);
      if (!trivial && (!g.q.empty() || !g.q_unsure.empty())) {
        if (!g.q_nontrivial.push((int32_t)me)) g.status = MC_QUEUE_OVERFLOW;
        ++g.n_nontrivial_push;
        return false;
      }
    }
    const int config = LUT2(CASES, c.index, 1);
    const Tiling t = select_tiling(c, kase, config);
    // the tiling's edge list (<= 36 table entries): one cooperative load, then both passes below read shared memory
    MC_WARP_SYNC();
    MC_LANE_LOOP(l) {
      for (int k = l; k < t.nt * 3; k += 32) cc.tedge[k] = (uint8_t)tiling_edge(t, config, k);
    }
    MC_WARP_SYNC();
    if (mode == 1) {
      uint32_t seen = 0;
      int existing = 0;   // check_tiling(): distinct edge ids whose vertex slot is already filled (see check_tiling_c)
      for (int k = 0; k < t.nt * 3; ++k) {
        const int e = cc.tedge[k];
        if (!((seen >> e) & 1u) && cc.fl[e] >= 0) ++existing;
        seen |= 1u << e;
      }
      if (existing < 2) return false;
    }
    g.flg[me] = (uint8_t)(cflg[0] | 1 | 2);
    done_set = true;
    MC_PROF_T(t_tiled);
    MC_PROF_ADD(3, t_signed, t_tiled);
    for (int k = 0; k < t.nt * 3; ++k) add_face_from_edge_c(g, cc, c, cc.tedge[k]);
    push_neighbours(g, z, y, x);
    ++g.n_accept;
    MC_PROF_T(t_emitted);
    MC_PROF_ADD(4, t_tiled, t_emitted);
    return true;
  }
  g.flg[me] = (uint8_t)(cflg[0] | 1 | 2);
  done_set = true;
  return false;
}

// linear index -> (z, y, x); `sh` = log2(N) when N is a power of two (shifts instead of three integer divisions), else -1
MC_HD void decode_index(int32_t c, int N, int sh, int& z, int& y, int& x) {
  if (sh >= 0) { x = c & (N - 1); y = (c >> sh) & (N - 1); z = c >> (2 * sh); }
  else { x = c % N; y = (c / N) % N; z = c / (N * N); }
}

// Pull the lattice neighbourhood of cube `c` towards L2 ahead of its visit (hint only; no effect on results).
MC_HD void prefetch_cube(const Grid& g, int32_t c, int sh) {
  const int N = g.N;
  int x, y, z;
  decode_index(c, N, sh, z, y, x);
  const int x0 = x > 0 ? x - 1 : 0;
  for (int r = 0; r < 16; ++r) {
    int cz = z - 1 + (r >> 2), cy = y - 1 + (r & 3);
    cz = cz < 0 ? 0 : (cz >= N ? N - 1 : cz);
    cy = cy < 0 ? 0 : (cy >= N ? N - 1 : cy);
    const int64_t i = lin(g, cz, cy, x0);
    MC_PREFETCH(g.im + i);
    MC_PREFETCH(g.grads + 3 * i);
    MC_PREFETCH(g.grads + 3 * i + 8);
    MC_PREFETCH(g.sgn + i);
    MC_PREFETCH(g.flg + i);
  }
  for (int r = 0; r < 4; ++r) MC_PREFETCH(g.face_layer + 4 * lin(g, z + (r >> 1), y + (r & 1), x));
}

// A visit set the "visited" flag of cube `me`: keep the look-ahead windows coherent.
MC_HD void note_done(CubeCache& cc, int32_t me) {
  MC_LANE_LOOP(l) {
    if (cc.qw_cur[l] == me) cc.qw_f[l] |= 1;
    if (cc.sw_cur[l] == me) cc.sw_f[l] |= 1;
  }
  MC_WARP_SYNC();
}

// replay() with visit_cube_w(); `g` is the caller's private (per-lane) copy of the bookkeeping.
// Queue pops and raster seeds are served from 32-entry look-ahead windows: one cooperative load fetches the next 32
// entries, their visited flags and candidate bits (instead of two dependent global loads per pop), and note_done()
// replays this warp's own flag updates into the windows, so the decisions are exactly those of replay().
MC_HD void replay_w(Grid& g, CubeCache& cc, const Grid* home) {
  const int N = g.N;
  int sh = -1;
  if ((N & (N - 1)) == 0) { sh = 0; while ((1 << sh) < N) ++sh; }
  g.n_v = 0; g.n_f3 = 0; g.status = MC_OK;
  g.n_seed = g.n_accept = g.n_unsure_push = g.n_nontrivial_push = 0;
  for (int i = 0; i < 8; ++i) g.prof[i] = 0;
  MC_PROF_T(t_replay0);
  MC_WARP_SYNC();
  MC_LANE_LOOP(l) { cc.qw_cur[l] = -1; cc.sw_cur[l] = -1; cc.qw_f[l] = 0; cc.sw_f[l] = 0; }
  MC_WARP_SYNC();
  uint32_t qw_base = 0, qw_n = 0;
  int64_t sw_base = 0, sw_n = 0;
  bool done_set = false;
  for (int64_t k = 0; k < g.n_cand; ++k) {
    if (k - sw_base >= sw_n) {
      sw_base = k;
      sw_n = g.n_cand - k < 32 ? g.n_cand - k : 32;
      MC_WARP_SYNC();
      MC_LANE_LOOP(l) {
        int32_t c = -1; uint8_t f = 0;
        if (l < sw_n) { c = g.cand_list[sw_base + l]; f = (g.flg[c] & 2) ? 1 : 0; }
        cc.sw_cur[l] = c; cc.sw_f[l] = f;
      }
      MC_WARP_SYNC();
    }
    const int32_t cidx = cc.sw_cur[k - sw_base];
    if (cc.sw_f[k - sw_base] & 1) continue;
    int x, y, z;
    decode_index(cidx, N, sh, z, y, x);
    ++g.n_seed;
    const bool accepted = visit_cube_w(g, cc, home, z, y, x, 0, done_set);
    if (done_set) note_done(cc, cidx);
    if (!accepted) continue;
    bool visit_neighbours = true;
    while (!g.q.empty() || !g.q_unsure.empty() || !g.q_nontrivial.empty()) {
      if (g.status == MC_QUEUE_OVERFLOW) return;
      int32_t cur;
      if (g.q.empty()) {
        if (g.q_unsure.empty()) {
          cur = g.q_nontrivial.front(); g.q_nontrivial.pop();
        } else {
          cur = g.q_unsure.front();
          if (visit_neighbours) {
            if (g.flg[cur] & 2) { g.q_unsure.pop(); continue; }
            int ux, uy, uz;
            decode_index(cur, N, sh, uz, uy, ux);
            push_neighbours(g, uz, uy, ux);
            visit_neighbours = false;
            continue;
          } else {
            g.q_unsure.pop();
            visit_neighbours = true;
          }
        }
        if (g.flg[cur] & 2) continue;
        if (!is_candidate(g, cur)) continue;
      } else {
        if (g.q.head - qw_base >= qw_n) {
          qw_base = g.q.head;
          const uint32_t avail = g.q.tail - g.q.head;
          qw_n = avail < 32u ? avail : 32u;
          MC_PROF_INC(6);
          MC_WARP_SYNC();
          MC_LANE_LOOP(l) {
            int32_t c = -1; uint8_t f = 0;
            if ((uint32_t)l < qw_n) {
              c = g.q.buf[(qw_base + (uint32_t)l) & g.q.mask];
              f = (uint8_t)(((g.flg[c] & 2) ? 1 : 0) | (is_candidate(g, c) ? 2 : 0));
              if (f == 2) prefetch_cube(g, c, sh);
            }
            cc.qw_cur[l] = c; cc.qw_f[l] = f;
          }
          MC_WARP_SYNC();
        }
        const uint32_t w = g.q.head - qw_base;
        cur = cc.qw_cur[w];
        const uint8_t wf = cc.qw_f[w];
        g.q.pop();
        if (wf != 2) continue;   // already visited, or not a candidate cube
      }
      int vx, vy, vz;
      decode_index(cur, N, sh, vz, vy, vx);
      visit_cube_w(g, cc, home, vz, vy, vx, visit_neighbours ? 1 : 2, done_set);
      if (done_set) note_done(cc, cur);
    }
  }
  MC_PROF_T(t_replay1);
  MC_PROF_ADD(0, t_replay0, t_replay1);
  if (g.status == MC_OK && g.n_v == 0) g.status = MC_EMPTY;
}

}  // namespace surfd_mccore
