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
// every branch names its table: the directory entry is a compile-time constant (no dependent table load on the chain)
#define MC_TIL(name, sub, ntri) { r.lut = LUT_##name; r.i1 = (sub); r.nt = (ntri); r.off = LUTOFF_##name; r.l1 = LUTL1_##name; r.l2 = LUTL2_##name; }
MC_HD_NOINLINE Tiling select_tiling(const Cell& c, int kase, int config) {
  Tiling r; r.lut = -1; r.i1 = -1; r.nt = 0; r.off = 0; r.l1 = 1; r.l2 = 1;
  int sub = 0;
  switch (kase) {
    case 1: MC_TIL(TILING1, -1, 1); break;
    case 2: MC_TIL(TILING2, -1, 2); break;
    case 3:
      if (test_face(c, LUT1(TEST3, config))) { MC_TIL(TILING3_2, -1, 4); }
      else { MC_TIL(TILING3_1, -1, 2); }
      break;
    case 4:
      if (test_internal(c, kase, config, sub, LUT1(TEST4, config))) { MC_TIL(TILING4_1, -1, 2); }
      else { MC_TIL(TILING4_2, -1, 6); }
      break;
    case 5: MC_TIL(TILING5, -1, 3); break;
    case 6:
      if (test_face(c, LUT2(TEST6, config, 0))) { MC_TIL(TILING6_2, -1, 5); }
      else if (test_internal(c, kase, config, sub, LUT2(TEST6, config, 1))) { MC_TIL(TILING6_1_1, -1, 3); }
      else { MC_TIL(TILING6_1_2, -1, 9); }
      break;
    case 7:
      if (test_face(c, LUT2(TEST7, config, 0))) sub += 1;
      if (test_face(c, LUT2(TEST7, config, 1))) sub += 2;
      if (test_face(c, LUT2(TEST7, config, 2))) sub += 4;
      switch (sub) {
        case 0: MC_TIL(TILING7_1, -1, 3); break;
        case 1: MC_TIL(TILING7_2, 0, 5); break;
        case 2: MC_TIL(TILING7_2, 1, 5); break;
        case 3: MC_TIL(TILING7_3, 0, 9); break;
        case 4: MC_TIL(TILING7_2, 2, 5); break;
        case 5: MC_TIL(TILING7_3, 1, 9); break;
        case 6: MC_TIL(TILING7_3, 2, 9); break;
        default:
          if (test_internal(c, kase, config, sub, LUT2(TEST7, config, 3))) { MC_TIL(TILING7_4_2, -1, 9); }
          else { MC_TIL(TILING7_4_1, -1, 5); }
      }
      break;
    case 8: MC_TIL(TILING8, -1, 2); break;
    case 9: MC_TIL(TILING9, -1, 4); break;
    case 10:
      if (test_face(c, LUT2(TEST10, config, 0))) {
        if (test_face(c, LUT2(TEST10, config, 1))) { MC_TIL(TILING10_1_1_, -1, 4); }
        else { MC_TIL(TILING10_2, -1, 8); }
      } else {
        if (test_face(c, LUT2(TEST10, config, 1))) { MC_TIL(TILING10_2_, -1, 8); }
        else if (test_internal(c, kase, config, sub, LUT2(TEST10, config, 2))) { MC_TIL(TILING10_1_1, -1, 4); }
        else { MC_TIL(TILING10_1_2, -1, 8); }
      }
      break;
    case 11: MC_TIL(TILING11, -1, 4); break;
    case 12:
      if (test_face(c, LUT2(TEST12, config, 0))) {
        if (test_face(c, LUT2(TEST12, config, 1))) { MC_TIL(TILING12_1_1_, -1, 4); }
        else { MC_TIL(TILING12_2, -1, 8); }
      } else {
        if (test_face(c, LUT2(TEST12, config, 1))) { MC_TIL(TILING12_2_, -1, 8); }
        else if (test_internal(c, kase, config, sub, LUT2(TEST12, config, 2))) { MC_TIL(TILING12_1_1, -1, 4); }
        else { MC_TIL(TILING12_1_2, -1, 8); }
      }
      break;
    case 13: {
      for (int b = 0; b < 6; ++b)
        if (test_face(c, LUT2(TEST13, config, b))) sub += (1 << b);
      sub = LUT1(SUBCONFIG13, sub);
      if (sub == 0) { MC_TIL(TILING13_1, -1, 4); }
      else if (sub <= 6) { MC_TIL(TILING13_2, sub - 1, 6); }
      else if (sub <= 18) { MC_TIL(TILING13_3, sub - 7, 10); }
      else if (sub <= 22) { MC_TIL(TILING13_4, sub - 19, 12); }
      else if (sub <= 26) {
        int s2 = sub - 23;
        if (test_internal(c, kase, config, s2, LUT2(TEST13, config, 6))) { MC_TIL(TILING13_5_1, s2, 6); }
        else { MC_TIL(TILING13_5_2, s2, 10); }
      }
      else if (sub <= 38) { MC_TIL(TILING13_3_, sub - 27, 10); }
      else if (sub <= 44) { MC_TIL(TILING13_2_, sub - 39, 6); }
      else if (sub == 45) { MC_TIL(TILING13_1_, -1, 4); }
      // else: "Impossible case 13" in the reference: nothing emitted
      break;
    }
    case 14: MC_TIL(TILING14, -1, 4); break;
    default: break;
  }
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
MC_HD int64_t lin_n(int N, int z, int y, int x) { return ((int64_t)z * N + y) * N + x; }

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

// Lattice state of the scalar replay: the reference's dense arrays (signed_im, signed_im_mask / visited, face_layer).
// visit_cube_t() below is written against this small interface so that the device chain (mc_chain.h) can run the very
// same visit on its O(surface) compact state when a cube needs the "look past an exact zero" rule.
struct DenseAcc {
  Grid& g;
  int N;
  MC_HD explicit DenseAcc(Grid& g_) : g(g_), N(g_.N) {}
  MC_HD float im(int64_t i) const { return g.im[i]; }
  MC_HD const float* gr(int64_t i) const { return g.grads + 3 * i; }
  MC_HD int sgn(int64_t i) const { return g.sgn[i]; }
  MC_HD void set_sgn(int64_t i, int s) { g.sgn[i] = (int8_t)s; }
  MC_HD bool fixed(int64_t i) const { return g.flg[i] & 1; }
  MC_HD void set_fixed(int64_t i) { g.flg[i] |= 1; }
  MC_HD void set_done(int64_t cube) { g.flg[cube] |= 2; }
  MC_HD bool q_empty() const { return g.q.empty(); }
  MC_HD bool qu_empty() const { return g.q_unsure.empty(); }
  MC_HD void push_unsure(int64_t cube) {
    if (!g.q_unsure.push((int32_t)cube)) g.status = MC_QUEUE_OVERFLOW;
    ++g.n_unsure_push;
  }
  MC_HD void push_nontrivial(int64_t cube) {
    if (!g.q_nontrivial.push((int32_t)cube)) g.status = MC_QUEUE_OVERFLOW;
    ++g.n_nontrivial_push;
  }
  MC_HD int existing(const Cell& c, const Tiling& t, int config) const { return check_tiling(g, c, t, config); }
  MC_HD void accept(Cell& c, const Tiling& t, int config) {
    add_tiling(g, c, t, config);
    push_neighbours(g, c.z, c.y, c.x);
    ++g.n_accept;
  }
};

// One visit of cube (z,y,x).  mode: 0 = raster seed, 1 = BFS (visit_neighbours flag True),
// 2 = BFS while serving an unsure cube's neighbours (flag False).
// Returns true when the cube was accepted and produced faces (seed: start a BFS).
template <class S>
MC_HD_NOINLINE bool visit_cube_t(S& s, int z, int y, int x, int mode) {
  const int N = s.N;
  const int nb = N - 2;
  int64_t ci[8];
  float cim[8];
  for (int i = 0; i < 8; ++i) {
    ci[i] = lin_n(N, z + MC_CZ(i), y + MC_CY(i), x + MC_CX(i));
    cim[i] = s.im(ci[i]);
  }
  int visited_vs[8];
  float sign_vs[8];
  for (int vtx = 0; vtx < 8; ++vtx) {
    visited_vs[vtx] = 0;
    sign_vs[vtx] = 0.0f;
    const int64_t c0 = ci[vtx];
    if (s.fixed(c0)) {
      visited_vs[vtx] = 1;
      sign_vs[vtx] = (float)s.sgn(c0);
      continue;
    }
    if (cim[vtx] == 0.0f) {
      visited_vs[vtx] = 1;
      continue;
    }
    const int zi = z + MC_CZ(vtx), yi = y + MC_CY(vtx), xi = x + MC_CX(vtx);
    const float* g1 = s.gr(c0);
    for (int d = 0; d < 6; ++d) {
      const int dz = (d == 0) - (d == 1), dy = (d == 2) - (d == 3), dx = (d == 4) - (d == 5);
      int i = 0, maxd = 1;
      while (i < maxd) {
        ++i;
        const int cz = zi + i * dz, cy = yi + i * dy, cx = xi + i * dx;
        if (cz > nb || cz < 0 || cy > nb || cy < 0 || cx > nb || cx < 0) break;
        const int64_t cn = lin_n(N, cz, cy, cx);
        if (s.im(cn) == 0.0f) {
          if (i < maxd) continue;
          ++maxd;
          continue;
        }
        const int sn = s.sgn(cn);
        if (sn == 0) continue;
        visited_vs[vtx] += 1;
        sign_vs[vtx] = vote_accumulate(sign_vs[vtx], (float)sn, edge_vote(g1, s.gr(cn), dz, dy, dx));
      }
    }
    if (mode != 0) {
      // pyx:1584: python-object arithmetic => double division, compared with double(0.707f)
      if (visited_vs[vtx] >= 1 &&
          (double)(sign_vs[vtx] < 0 ? -sign_vs[vtx] : sign_vs[vtx]) / (double)visited_vs[vtx] < (double)0.707f &&
          !s.q_empty()) {
        if (mode == 1) s.push_unsure(ci[0]);
        return false;
      }
    }
    s.set_sgn(c0, (int)my_sign(sign_vs[vtx]));
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
      if (s.fixed(c0) && non_zero_norm(s.gr(c0))) {
        anchor_sign = my_sign((float)s.sgn(c0));
        base[0] = s.gr(c0)[0]; base[1] = s.gr(c0)[1]; base[2] = s.gr(c0)[2];
        found = true;
      }
    }
    for (int k = 0; k < 8 && !found; ++k) {
      const int64_t c0 = ci[order[k]];
      if (non_zero_norm(s.gr(c0))) {
        base[0] = s.gr(c0)[0]; base[1] = s.gr(c0)[1]; base[2] = s.gr(c0)[2];
        found = true;
      }
    }
    // (reference prints 'all 0 vec...' and keeps the previous base_vec; with a fresh buffer that is
    //  uninitialised memory -- we use zeros, which gives sign 0 for the unvoted corners)
    base[0] = anchor_sign * base[0]; base[1] = anchor_sign * base[1]; base[2] = anchor_sign * base[2];
    const bool check_unsure = (mode == 1) && !s.q_empty();
    for (int i = 0; i < 8; ++i) {
      if (visited_vs[i] != 0) continue;
      const float sv = dot3(base, s.gr(ci[i]));
      if (check_unsure) {
        sign_vs[i] = sv;
        if ((sv < 0 ? -sv : sv) < 0.707f) {
          s.push_unsure(ci[0]);
          return false;
        }
      }
      s.set_sgn(ci[i], (int)my_sign(sv));
    }
  }

  if (mode == 2) return false;  // neighbours of an unsure cube: tentative signs only

  double v[8];
  for (int i = 0; i < 8; ++i) {
    float p = (float)s.sgn(ci[i]) * cim[i];
    v[i] = (double)p;
  }
  Cell c;
  cell_set(c, x, y, z, v);
  for (int i = 0; i < 8; ++i) s.set_fixed(ci[i]);

  const int kase = LUT2(CASES, c.index, 0);
  const int64_t me = ci[0];
  if (kase > 0) {
    if (mode == 1) {
      const bool trivial = (kase == 1 || kase == 2 || kase == 5 || kase == 8 || kase == 9);
      if (!trivial && (!s.q_empty() || !s.qu_empty())) {
        s.push_nontrivial(me);
        return false;
      }
    }
    const int config = LUT2(CASES, c.index, 1);
    const Tiling t = select_tiling(c, kase, config);
    if (mode == 1) {
      if (s.existing(c, t, config) < 2) return false;
    }
    s.set_done(me);
    s.accept(c, t, config);
    return true;
  }
  s.set_done(me);
  return false;
}

MC_HD_NOINLINE bool visit_cube(Grid& g, int z, int y, int x, int mode) {
  DenseAcc s(g);
  return visit_cube_t(s, z, y, x, mode);
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

}  // namespace surfd_mccore
