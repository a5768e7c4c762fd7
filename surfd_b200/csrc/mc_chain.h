// mc_chain.h -- the order-dependent part of MeshUDF's marching cubes on O(surface) state.
//
// The reference (meshudf/_marching_cubes_lewiner_cy.pyx:1115-1773) is a breadth-first walk whose sign votes, acceptance
// test ("two of the tiling's vertices already exist", pyx:1676-1690) and vertex numbering depend on the visiting order, so
// one chain of dependent visits per shape is unavoidable if vertices and faces are to come out bit-identical.  This file
// makes that chain as short as it can be and moves everything else off it:
//
//   before the chain (parallel over candidate cubes, build_record()):
//     * everything a visit reads that does NOT depend on the order -- the cube's 8 corner udf values, the 48
//       (corner, direction) edge votes (pure functions of the gradients, pyx:1776-1806), which of them are usable under
//       the reference's bounds rule, whether a neighbour's udf is exactly 0 (extension rule, pyx:1287-1292), the six
//       BFS neighbours -- is evaluated once per candidate cube and stored as a 400-byte record;
//     * lattice vertices that are a corner of some candidate cube get a compact id (rank in a bit mask), and so do the
//       candidate cubes: the mutable state (sign, "sign is final" flag, visited flag, the 4 vertex slots of the
//       reference's face_layer) is indexed by those ids -- O(surface) bytes instead of the reference's 6 N^3.
//   on the chain (one warp, replay_r()): per visit one coalesced record read (immutable: prefetched), one gather of the
//     32 vertex states + 13 vertex slots, the corner-by-corner vote accumulation, the Lewiner case/tiling selection, and
//     for an accepted cube only the *numbering* of its new vertices and faces (counters + slot stores) and a 16-byte log
//     entry;
//   after the chain (parallel over accepted cubes, emit_cube()): vertex interpolation (FP64 divisions) and the face
//     index writes, from the log -- same numbers, same order as the reference's emission inside the walk.
//
// Cubes that need the extension rule run the generic visit (mc_core.h visit_cube_t) on the same compact state.
// Builds for the device (mc.cu) and for the host (tests/native/mc_host.cpp: lanes emulated sequentially).
#pragma once
#include "mc_core.h"

namespace surfd_mccore {

#if defined(__CUDA_ARCH__)
#define MC_UNROLL _Pragma("unroll")
#define MC_LANE_LOOP(l) for (int l = (int)(threadIdx.x & 31), _mc_once = 1; _mc_once; _mc_once = 0)
#define MC_WARP_SYNC() __syncwarp()
#define MC_POPC(x) __popc(x)
#define MC_PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#else
#define MC_UNROLL
#define MC_LANE_LOOP(l) for (int l = 0; l < 32; ++l)
#define MC_WARP_SYNC()
#define MC_POPC(x) __builtin_popcount(x)
#define MC_PREFETCH_L1(p) ((void)(p))
#endif

#if defined(__CUDA_ARCH__) && defined(MC_PROFILE)
#define MC_PROF_T(var) const long long var = clock64()
#define MC_PROF_ADD(slot, t0, t1) g.prof[slot] += (t1) - (t0)
#define MC_PROF_INC(slot) g.prof[slot] += 1
#else
#define MC_PROF_T(var)
#define MC_PROF_ADD(slot, t0, t1)
#define MC_PROF_INC(slot)
#endif

// Lewiner corner id from its (z, y, x) bits: inverse of MC_CZ / MC_CY / MC_CX
#define MC_CORNER(cz, cy, cx) (((cz) << 2) | ((cy) ? ((cx) ? 2 : 3) : ((cx) ? 1 : 0)))

// face_layer slot of edge id e (pyx:677-761): the lattice vertex that owns it, as a corner of the cube, and which of
// that vertex' four slots (0: x edge, 1: y edge, 2: z edge, 3: cell centre).
MC_HD int edge_corner(int e) { return (int)((0x321047540310ull >> (4 * e)) & 15); }   // {0,1,3,0, 4,5,7,4, 0,1,2,3, 0}
MC_HD int edge_j(int e) { return (int)((0x3AA4444u >> (2 * e)) & 3); }                  // {0,1,0,1, 0,1,0,1, 2,2,2,2, 3}

struct U4 { uint32_t a, b, c, d; };

// Per candidate cube, immutable.  Direction d: 0 +z, 1 -z, 2 +y, 3 -y, 4 +x, 5 -x (the reference's loop order).
struct alignas(16) Rec {
  int32_t vid[32];     // [0..7] compact vertex ids of the corners (Lewiner order); [8 + 3*corner + axis] the corner's
                       // neighbour OUTSIDE the cube along axis (0 z, 1 y, 2 x), -1 if it can never carry a sign
  float vote[48];      // [6*corner + d] edge_vote(corner, neighbour)
  float cim[8];        // udf at the corners
  uint32_t usable[2];  // bit 6*corner + d: neighbour inside the reference's bounds and its udf != 0
  uint32_t zero_nb;    // bit corner: some in-bounds neighbour has udf == 0 (extension rule -> generic visit)
  int32_t lattice;     // linear lattice index of the cube's anchor
  int32_t nbr[6];      // BFS pushes in the reference's order (x+1, y+1, x-1, y-1, z-1, z+1): candidate id, -1 = pushed but
                       // not a candidate cube (it only keeps the queue non-empty), -2 = outside the bounds rule: not pushed
  int32_t pad[2];
};
static_assert(sizeof(Rec) == 400, "record layout");

struct Accept { int32_t cid; uint32_t tiling; int32_t nv0; int32_t nf0; };   // tiling = first table entry | triangles << 16

struct Chain {
  int N;
  const float* im;
  const float* grads;
  const uint32_t* cand_bits;     // 1 bit per lattice index (classify_kernel)
  const int32_t* cand_prefix;    // set bits before each 32-bit word
  const uint32_t* vtx_bits;      // lattice vertices that are a corner of a candidate cube
  const int32_t* vtx_prefix;
  const int32_t* cand_list;      // raster-sorted candidate cubes (lattice indices); position == compact cube id
  int64_t n_cand, n_cand_total, n_vtx, cap_vtx;
  Rec* recs;
  uint8_t* vs;                   // per vertex: bits 0-1 sign (two's complement), bit 2 "final" (signed_im_mask)
  uint8_t* done;                 // per candidate cube: visited
  int32_t* slot;                 // [4 * n_vtx] vertex index per face_layer slot, -1 = empty
  Accept* acc;
  float* verts; int32_t* faces;
  int64_t cap_v, cap_f3;
  Fifo q, q_unsure, q_nontrivial;   // entries: compact cube ids (q also holds -1 place holders)
  int64_t n_v, n_f3;
  int status;
  int64_t n_seed, n_accept, n_unsure_push, n_nontrivial_push;
  // cycle counters (-DMC_PROFILE, device): 0 total, 1 record + state fetch, 2 sign votes, 3 case / tiling, 4 numbering + pushes,
  // 5 visits, 6 queue-window refills, 7 generic visits
  int64_t prof[8];
};

MC_HD int32_t rank_in(const uint32_t* bits, const int32_t* prefix, int64_t i) {
  const uint32_t w = bits[i >> 5];
  const uint32_t b = (uint32_t)(i & 31);
  if (!((w >> b) & 1u)) return -1;
  return prefix[i >> 5] + MC_POPC(w & ((1u << b) - 1u));
}

MC_HD int st_sign(uint8_t st) { return (st & 3) == 3 ? -1 : (int)(st & 3); }
MC_HD uint8_t st_make(int sgn, int fixed) { return (uint8_t)((sgn & 3) | (fixed ? 4 : 0)); }

// ---- before the chain: one record per candidate cube (any thread / any order) ----
MC_HD void build_record(const Chain& g, int64_t k) {
  const int N = g.N;
  const int nb = N - 2;
  const int32_t cidx = g.cand_list[k];
  const int x = cidx % N, y = (cidx / N) % N, z = cidx / (N * N);
  Rec r;
  r.lattice = cidx;
  r.pad[0] = r.pad[1] = 0;
  r.usable[0] = r.usable[1] = 0;
  r.zero_nb = 0;
  for (int i = 0; i < 24; ++i) r.vid[8 + i] = -1;
  for (int c = 0; c < 8; ++c) {
    const int cz = z + MC_CZ(c), cy = y + MC_CY(c), cx = x + MC_CX(c);
    const int64_t li = lin_n(N, cz, cy, cx);
    r.vid[c] = rank_in(g.vtx_bits, g.vtx_prefix, li);
    r.cim[c] = g.im[li];
    const float* g1 = g.grads + 3 * li;
    for (int d = 0; d < 6; ++d) {
      const int dz = (d == 0) - (d == 1), dy = (d == 2) - (d == 3), dx = (d == 4) - (d == 5);
      const int nz = cz + dz, ny = cy + dy, nx = cx + dx;
      float vt = 0.f;
      if (!(nz > nb || nz < 0 || ny > nb || ny < 0 || nx > nb || nx < 0)) {
        const int64_t ln = lin_n(N, nz, ny, nx);
        if (g.im[ln] == 0.0f) {
          r.zero_nb |= 1u << c;
        } else {
          const int bit = 6 * c + d;
          r.usable[bit >> 5] |= 1u << (bit & 31);
          vt = edge_vote(g1, g.grads + 3 * ln, dz, dy, dx);
          const bool in_cube = (nz - z) >= 0 && (nz - z) <= 1 && (ny - y) >= 0 && (ny - y) <= 1 && (nx - x) >= 0 && (nx - x) <= 1;
          if (!in_cube) r.vid[8 + 3 * c + (d >> 1)] = rank_in(g.vtx_bits, g.vtx_prefix, ln);
        }
      }
      r.vote[6 * c + d] = vt;
    }
  }
  // pyx:1407-1418
  const int px[6] = {x + 1, x, x - 1, x, x, x};
  const int py[6] = {y, y + 1, y, y - 1, y, y};
  const int pz[6] = {z, z, z, z, z - 1, z + 1};
  for (int i = 0; i < 6; ++i) {
    const int c = i == 0 ? px[i] : (i == 1 ? py[i] : (i == 2 ? px[i] : (i == 3 ? py[i] : pz[i])));
    const bool pushed = (i == 0 || i == 1 || i == 5) ? (c < nb) : (c >= 0);
    r.nbr[i] = pushed ? rank_in(g.cand_bits, g.cand_prefix, lin_n(N, pz[i], py[i], px[i])) : -2;
  }
  g.recs[k] = r;
}

// ---- the generic visit on compact state (extension rule; rare) ----
struct CompactAcc {
  Chain& g;
  const Rec* rec;
  int32_t cid;
  int N;
  MC_HD CompactAcc(Chain& g_, const Rec* r, int32_t c) : g(g_), rec(r), cid(c), N(g_.N) {}
  MC_HD int32_t vrank(int64_t i) const { return rank_in(g.vtx_bits, g.vtx_prefix, i); }
  MC_HD float im(int64_t i) const { return g.im[i]; }
  MC_HD const float* gr(int64_t i) const { return g.grads + 3 * i; }
  MC_HD int sgn(int64_t i) const { const int32_t v = vrank(i); return v < 0 ? 0 : st_sign(g.vs[v]); }
  MC_HD void set_sgn(int64_t i, int s) { const int32_t v = vrank(i); g.vs[v] = (uint8_t)((g.vs[v] & 4) | (s & 3)); }
  MC_HD bool fixed(int64_t i) const { const int32_t v = vrank(i); return v >= 0 && (g.vs[v] & 4); }
  MC_HD void set_fixed(int64_t i) { const int32_t v = vrank(i); g.vs[v] |= 4; }
  MC_HD void set_done(int64_t) { g.done[cid] = 1; }
  MC_HD bool q_empty() const { return g.q.empty(); }
  MC_HD bool qu_empty() const { return g.q_unsure.empty(); }
  MC_HD void push_unsure(int64_t) {
    if (!g.q_unsure.push(cid)) g.status = MC_QUEUE_OVERFLOW;
    ++g.n_unsure_push;
  }
  MC_HD void push_nontrivial(int64_t) {
    if (!g.q_nontrivial.push(cid)) g.status = MC_QUEUE_OVERFLOW;
    ++g.n_nontrivial_push;
  }
  MC_HD int64_t slot_of(int e) const { return 4 * (int64_t)rec->vid[edge_corner(e)] + edge_j(e); }
  MC_HD int existing(const Cell&, const Tiling& t, int config) const {
    uint32_t seen = 0;
    int result = 0;
    for (int k = 0; k < t.nt * 3; ++k) {
      const int e = tiling_edge(t, config, k);
      if (!((seen >> e) & 1u) && g.slot[slot_of(e)] >= 0) ++result;
      seen |= 1u << e;
    }
    return result;
  }
  MC_HD void accept(Cell&, const Tiling& t, int config);
};

MC_HD int tiling_start(const Tiling& t, int config) {
  return t.i1 < 0 ? t.off + config * t.l1 : t.off + (config * t.l1 + t.i1) * t.l2;
}

// Numbering of an accepted cube's vertices and faces + its log entry + the BFS pushes.  `filled` = bit e set when the slot
// of edge e already holds a vertex.  The slot stores are the only lattice-state writes; positions and face indices follow
// in emit_cube().
MC_HD void accept_cube(Chain& g, const Rec& r, int32_t cid, int start, int nt, uint32_t filled) {
  const int64_t nv0 = g.n_v, nf0 = g.n_f3;
  for (int k = 0; k < nt * 3; ++k) {
    const int e = (int)MC_LUTV(start + k);
    if (!((filled >> e) & 1u)) {
      filled |= 1u << e;
      g.slot[4 * (int64_t)r.vid[edge_corner(e)] + edge_j(e)] = (int32_t)g.n_v;
      ++g.n_v;
    }
  }
  g.n_f3 += 3 * nt;
  if (g.n_v > g.cap_v || g.n_f3 > g.cap_f3) g.status = MC_CAPACITY;   // keep counting so the caller learns the required size
  Accept a;
  a.cid = cid; a.tiling = (uint32_t)start | ((uint32_t)nt << 16); a.nv0 = (int32_t)nv0; a.nf0 = (int32_t)nf0;
  g.acc[g.n_accept] = a;
  ++g.n_accept;
  bool ok = true;
  MC_UNROLL
  for (int i = 0; i < 6; ++i)
    if (r.nbr[i] != -2) ok = g.q.push(r.nbr[i]) && ok;
  if (!ok) g.status = MC_QUEUE_OVERFLOW;
}

MC_HD void CompactAcc::accept(Cell&, const Tiling& t, int config) {
  uint32_t filled = 0;
  for (int e = 0; e < 13; ++e)
    if (g.slot[slot_of(e)] >= 0) filled |= 1u << e;
  accept_cube(g, *rec, cid, tiling_start(t, config), t.nt, filled);
}

// The mutable part of a Chain, passed by value to the generic visit: the chain's own `Chain` never has its address taken,
// so the compiler keeps its counters in registers instead of local memory.
struct ChainState {
  uint32_t qh, qt, uh, ut, nh, nt;
  int status;
  int64_t n_v, n_f3, n_accept, n_unsure_push, n_nontrivial_push;
};
MC_HD ChainState chain_state_get(const Chain& g) {
  ChainState s;
  s.qh = g.q.head; s.qt = g.q.tail; s.uh = g.q_unsure.head; s.ut = g.q_unsure.tail;
  s.nh = g.q_nontrivial.head; s.nt = g.q_nontrivial.tail; s.status = g.status;
  s.n_v = g.n_v; s.n_f3 = g.n_f3; s.n_accept = g.n_accept;
  s.n_unsure_push = g.n_unsure_push; s.n_nontrivial_push = g.n_nontrivial_push;
  return s;
}
MC_HD void chain_state_put(Chain& g, const ChainState& s) {
  g.q.head = s.qh; g.q.tail = s.qt; g.q_unsure.head = s.uh; g.q_unsure.tail = s.ut;
  g.q_nontrivial.head = s.nh; g.q_nontrivial.tail = s.nt; g.status = s.status;
  g.n_v = s.n_v; g.n_f3 = s.n_f3; g.n_accept = s.n_accept;
  g.n_unsure_push = s.n_unsure_push; g.n_nontrivial_push = s.n_nontrivial_push;
}
// `home` holds the constant fields (sizes, pointers, queue buffers); the state travels in `cs`.
MC_HD_NOINLINE bool visit_generic_r(const Chain* home, ChainState& cs, const Rec* rec, int32_t cid, int mode, bool& done_set) {
  Chain tmp = *home;
  chain_state_put(tmp, cs);
  int x, y, z;
  x = rec->lattice % tmp.N; y = (rec->lattice / tmp.N) % tmp.N; z = rec->lattice / (tmp.N * tmp.N);
  CompactAcc s(tmp, rec, cid);
  const bool res = visit_cube_t(s, z, y, x, mode);
  done_set = tmp.done[cid] != 0;
  cs = chain_state_get(tmp);
  return res;
}

// ---- shared memory of the chain warp ----
struct ChainCache {
  Rec wrec[32];        // records of the BFS-queue window entries that may still be visited (copied at refill)
  Rec rec;             // record of a raster seed / priority-queue cube
  int32_t qw_cur[32];  // the window: next <= 32 entries of the BFS queue, one per lane
};

MC_HD void decode_index(int32_t c, int N, int& z, int& y, int& x) { x = c % N; y = (c / N) % N; z = c / (N * N); }

#if defined(__CUDA_ARCH__)
// every lane executes the "loop" body once: a vote collects the predicate of all lanes
#define MC_VOTE(mask, l, pred) mask = __ballot_sync(0xffffffffu, (pred))
// a per-lane register (one slot on the device, an array over the emulated lanes on the host)
#define MC_LANE_SLOTS 1
#define MC_LANE_SLOT(l) 0
#else
#define MC_VOTE(mask, l, pred) mask |= ((pred) ? (1u << (l)) : 0u)
#define MC_LANE_SLOTS 32
#define MC_LANE_SLOT(l) (l)
#endif

// record copy global -> shared: asynchronous on the device (no registers; MC_COPY_WAIT() completes all copies in flight)
MC_HD void copy_record(Rec* dst, const Rec* src, int l) {   // 16-byte piece l of the 25
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(reinterpret_cast<U4*>(dst) + l)),
               "l"(reinterpret_cast<const U4*>(src) + l) : "memory");
#else
  reinterpret_cast<U4*>(dst)[l] = reinterpret_cast<const U4*>(src)[l];
#endif
}
#if defined(__CUDA_ARCH__)
#define MC_COPY_WAIT() asm volatile("cp.async.wait_all;" ::: "memory")
#else
#define MC_COPY_WAIT()
#endif

// One visit of candidate cube `cid` (compact id); `r` is its record in shared memory.  Same decisions, same arithmetic,
// same order of effects as visit_cube_t().  All 32 lanes run the scalar control flow redundantly on registers; the lanes
// split only the memory accesses: one gather of the 32 vertex states and 13 vertex slots (distributed to every lane as bit
// masks by warp votes), the state / slot / queue stores, and the tiling's table entries.
MC_HD bool visit_r(Chain& g, const Rec& r, const Chain* home, int32_t cid, int mode, bool& done_set) {
  done_set = false;
  MC_PROF_T(t_begin);
  MC_PROF_INC(5);
  MC_WARP_SYNC();   // the previous visit's state / slot stores (other lanes) are ordered before this visit's gather
  // ---- gather: lane l holds vertex l of the record (8 corners + 24 outside neighbours), lanes 0-12 also one vertex slot ----
  uint32_t nzm = 0, ngm = 0, fxm = 0, filled = 0;   // sign != 0, sign < 0, sign final; slot of edge e holds a vertex
  MC_LANE_LOOP(l) {
    const int32_t v = r.vid[l];
    const uint8_t st = v >= 0 ? g.vs[v] : (uint8_t)0;
    int32_t fl = -1;
    if (l < 13) fl = g.slot[4 * (int64_t)r.vid[edge_corner(l)] + edge_j(l)];
    MC_VOTE(nzm, l, (st & 3) != 0);
    MC_VOTE(ngm, l, (st & 3) == 3);
    MC_VOTE(fxm, l, (st & 4) != 0);
    MC_VOTE(filled, l, fl >= 0);
  }
  // the immutable part of the visit, into registers (issued while the gather is in flight)
  float vote[48];
  float cim[8];
  MC_UNROLL
  for (int i = 0; i < 48; ++i) vote[i] = r.vote[i];
  MC_UNROLL
  for (int i = 0; i < 8; ++i) cim[i] = r.cim[i];
  const uint32_t us0 = r.usable[0], us1 = r.usable[1];
  uint32_t cim_nz = 0;
  MC_UNROLL
  for (int i = 0; i < 8; ++i) cim_nz |= (cim[i] != 0.0f) ? (1u << i) : 0u;
  MC_PROF_T(t_fetched);
  MC_PROF_ADD(1, t_begin, t_fetched);
  uint32_t cfx = fxm & 0xffu;   // corner masks, updated as corners are decided
  if (~cfx & cim_nz & r.zero_nb & 0xffu) {
    // the "exact zero neighbour" extension (pyx:1287-1292) looks further than the record: generic visit, same state
    MC_PROF_INC(7);
    ChainState cs = chain_state_get(g);
    const bool res = visit_generic_r(home, cs, &r, cid, mode, done_set);
    chain_state_put(g, cs);
    return res;
  }
  uint32_t dirty = 0;           // corners whose state byte must be written back
  // write-back of the decided corners (lane = corner); `fixed_all`: the cube's corners become final (signed_im_mask)
#define MC_FLUSH(fixed_all)                                                                                         \
  MC_LANE_LOOP(l) {                                                                                                 \
    if (l < 8 && (((dirty >> l) & 1u) || (fixed_all)))                                                              \
      g.vs[r.vid[l]] = (uint8_t)((((nzm >> l) & 1u) ? (((ngm >> l) & 1u) ? 3 : 1) : 0) | ((((cfx >> l) & 1u) || (fixed_all)) ? 4 : 0)); \
  }
  uint32_t voted = 0;           // visited_vs[i] >= 1
  MC_UNROLL
  for (int vtx = 0; vtx < 8; ++vtx) {
    if ((cfx >> vtx) & 1u) { voted |= 1u << vtx; continue; }
    if (!((cim_nz >> vtx) & 1u)) { voted |= 1u << vtx; continue; }
    float acc = 0.0f;
    int cnt = 0;
    MC_UNROLL
    for (int d = 0; d < 6; ++d) {
      const int bit = 6 * vtx + d;
      // a step along an axis stays inside the cube iff it flips that axis' corner bit from 0 to 1 (or back)
      const int cz = MC_CZ(vtx), cy = MC_CY(vtx), cx = MC_CX(vtx);
      const int nz = cz + (d == 0) - (d == 1), ny = cy + (d == 2) - (d == 3), nx = cx + (d == 4) - (d == 5);
      const bool in_cube = nz >= 0 && nz <= 1 && ny >= 0 && ny <= 1 && nx >= 0 && nx <= 1;
      const int ni = in_cube ? MC_CORNER(nz & 1, ny & 1, nx & 1) : 8 + 3 * vtx + (d >> 1);
      const bool active = (((bit < 32 ? us0 >> bit : us1 >> (bit - 32)) & 1u) != 0u) && (((nzm >> ni) & 1u) != 0u);
      const float term = ((ngm >> ni) & 1u) ? -vote[bit] : vote[bit];   // sign * vote, exact
      const float sum = vote_accumulate_f32(acc, 1.0f, term);
      acc = active ? sum : acc;
      cnt += active ? 1 : 0;
    }
    if (mode != 0) {
      // |sign| / visited < 0.707f (pyx:1584, a double division in the reference): 0.707f * n is exactly representable in
      // fp32 for n <= 6 (0.707f = 0x1.69fbe8p-1 has a 21-bit significand), and the rounded double quotient can only differ
      // from the exact one inside half an ulp53 of the threshold, which a 24-bit numerator over n <= 6 never reaches unless
      // it equals the product (then both tests are false)
      if (cnt >= 1 && (acc < 0 ? -acc : acc) < 0.707f * (float)cnt && !g.q.empty()) {
        MC_FLUSH(false);
        if (mode == 1) {
          if (!g.q_unsure.push(cid)) g.status = MC_QUEUE_OVERFLOW;
          ++g.n_unsure_push;
        }
        return false;
      }
    }
    if (cnt >= 1) voted |= 1u << vtx;
    nzm = (nzm & ~(1u << vtx)) | ((acc > 0.f || acc < 0.f) ? (1u << vtx) : 0u);
    ngm = (ngm & ~(1u << vtx)) | ((acc < 0.f) ? (1u << vtx) : 0u);
    dirty |= 1u << vtx;
  }

  if (voted != 0xffu) {
    // anchor gradient (pyx:1310-1346); the corner gradients are not part of the record: this branch is taken by seeds and
    // by the rare cube none of whose free corners has a signed neighbour
    int x, y, z;
    decode_index(r.lattice, g.N, z, y, x);
    const int order[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    const float* cg[8];
    MC_UNROLL
    for (int i = 0; i < 8; ++i) cg[i] = g.grads + 3 * lin_n(g.N, z + MC_CZ(i), y + MC_CY(i), x + MC_CX(i));
    float base[3] = {0.f, 0.f, 0.f};
    float anchor_sign = 1.f;
    bool found = false;
    MC_UNROLL
    for (int k = 0; k < 8; ++k) {
      if (found) continue;
      const int c0 = order[k];
      if (((cfx >> c0) & 1u) && non_zero_norm(cg[c0])) {
        anchor_sign = ((nzm >> c0) & 1u) ? (((ngm >> c0) & 1u) ? -1.f : 1.f) : 0.f;
        base[0] = cg[c0][0]; base[1] = cg[c0][1]; base[2] = cg[c0][2];
        found = true;
      }
    }
    MC_UNROLL
    for (int k = 0; k < 8; ++k) {
      if (found) continue;
      const int c0 = order[k];
      if (non_zero_norm(cg[c0])) {
        base[0] = cg[c0][0]; base[1] = cg[c0][1]; base[2] = cg[c0][2];
        found = true;
      }
    }
    base[0] = anchor_sign * base[0]; base[1] = anchor_sign * base[1]; base[2] = anchor_sign * base[2];
    const bool check_unsure = (mode == 1) && !g.q.empty();
    MC_UNROLL
    for (int i = 0; i < 8; ++i) {
      if ((voted >> i) & 1u) continue;
      const float sv = dot3(base, cg[i]);
      if (check_unsure && (sv < 0 ? -sv : sv) < 0.707f) {
        MC_FLUSH(false);
        if (!g.q_unsure.push(cid)) g.status = MC_QUEUE_OVERFLOW;
        ++g.n_unsure_push;
        return false;
      }
      nzm = (nzm & ~(1u << i)) | ((sv > 0.f || sv < 0.f) ? (1u << i) : 0u);
      ngm = (ngm & ~(1u << i)) | ((sv < 0.f) ? (1u << i) : 0u);
      dirty |= 1u << i;
    }
  }
  MC_PROF_T(t_signed);
  MC_PROF_ADD(2, t_fetched, t_signed);
  if (mode == 2) {   // neighbours of an unsure cube: tentative signs only
    MC_FLUSH(false);
    return false;
  }
  MC_FLUSH(true);    // the corners' signs are final from here on
#undef MC_FLUSH

  // cube index: bit i = (sign_i * udf_i > 0)
  int index = 0;
  MC_UNROLL
  for (int i = 0; i < 8; ++i) {
    const float sg = ((nzm >> i) & 1u) ? (((ngm >> i) & 1u) ? -1.f : 1.f) : 0.f;
    const float p = sg * cim[i];
    if (p > 0.f) index |= 1 << i;
  }
  const int kase = LUT2(CASES, index, 0);
  if (kase > 0) {
    const bool trivial = (kase == 1 || kase == 2 || kase == 5 || kase == 8 || kase == 9);
    if (mode == 1 && !trivial && (!g.q.empty() || !g.q_unsure.empty())) {
      if (!g.q_nontrivial.push(cid)) g.status = MC_QUEUE_OVERFLOW;
      ++g.n_nontrivial_push;
      return false;
    }
    const int config = LUT2(CASES, index, 1);
    int start, nt;
    if (kase == 1) { start = LUTOFF_TILING1 + config * LUTL1_TILING1; nt = 1; }
    else if (kase == 2) { start = LUTOFF_TILING2 + config * LUTL1_TILING2; nt = 2; }
    else if (kase == 5) { start = LUTOFF_TILING5 + config * LUTL1_TILING5; nt = 3; }
    else if (kase == 8) { start = LUTOFF_TILING8 + config * LUTL1_TILING8; nt = 2; }
    else if (kase == 9) { start = LUTOFF_TILING9 + config * LUTL1_TILING9; nt = 4; }
    else if (kase == 11) { start = LUTOFF_TILING11 + config * LUTL1_TILING11; nt = 4; }
    else if (kase == 14) { start = LUTOFF_TILING14 + config * LUTL1_TILING14; nt = 4; }
    else {
      // ambiguous cases: the face / interior tests need the signed corner values (pyx:2403-2569)
      double v[8];
      MC_UNROLL
      for (int i = 0; i < 8; ++i) {
        const float sg = ((nzm >> i) & 1u) ? (((ngm >> i) & 1u) ? -1.f : 1.f) : 0.f;
        const float p = sg * cim[i];
        v[i] = (double)p;
      }
      Cell c;
      cell_set(c, 0, 0, 0, v);
      const Tiling t = select_tiling(c, kase, config);
      start = tiling_start(t, config);
      nt = t.nt;
    }
    const int n3 = 3 * nt;
    // the tiling's edge ids, one per lane (n3 <= 30 except tiling 13.4 with 36: that one takes the sequential path)
    uint32_t seen = 0;
#if defined(__CUDA_ARCH__)
    const int lane = (int)(threadIdx.x & 31);
    const int e = (lane < n3 && n3 <= 32) ? (int)MC_LUTV(start + lane) : 13;
    if (n3 <= 32) seen = __reduce_or_sync(0xffffffffu, lane < n3 ? (1u << e) : 0u);
    else
#endif
      for (int k = 0; k < n3; ++k) seen |= 1u << (int)MC_LUTV(start + k);
    if (mode == 1) {
      // check_triangles(2) (pyx:467-526): distinct existing vertices among the tiling's slots.  A vertex index lives in
      // exactly one slot and the 13 edge ids of a cube map to 13 distinct slots, so that is the number of distinct edge ids
      // whose slot is filled.
      if (MC_POPC(seen & filled & 0x1fffu) < 2) return false;
    }
    g.done[cid] = 1;
    done_set = true;
    MC_PROF_T(t_tiled);
    MC_PROF_ADD(3, t_signed, t_tiled);
#if defined(__CUDA_ARCH__)
    if (n3 <= 32) {
      // numbering (accept_cube()) with the lanes in parallel: a tiling entry creates a vertex when it is the first
      // occurrence of its edge id and the slot is empty; new vertices are numbered in entry order
      const uint32_t same = __match_any_sync(0xffffffffu, e);
      const bool isnew = lane < n3 && (__ffs(same) - 1 == lane) && !((filled >> e) & 1u);
      const uint32_t newm = __ballot_sync(0xffffffffu, isnew);
      if (isnew) g.slot[4 * (int64_t)r.vid[edge_corner(e)] + edge_j(e)] = (int32_t)g.n_v + __popc(newm & ((1u << lane) - 1u));
      if (lane == 0) {
        Accept a;
        a.cid = cid; a.tiling = (uint32_t)start | ((uint32_t)nt << 16); a.nv0 = (int32_t)g.n_v; a.nf0 = (int32_t)g.n_f3;
        g.acc[g.n_accept] = a;
      }
      g.n_v += __popc(newm);
      g.n_f3 += n3;
      if (g.n_v > g.cap_v || g.n_f3 > g.cap_f3) g.status = MC_CAPACITY;
      ++g.n_accept;
      const int32_t nb = lane < 6 ? r.nbr[lane] : -2;
      const uint32_t pm = __ballot_sync(0xffffffffu, nb != -2);
      if (g.q.tail - g.q.head + 6u > g.q.mask) g.status = MC_QUEUE_OVERFLOW;
      else {
        if (nb != -2) g.q.buf[(g.q.tail + (uint32_t)__popc(pm & ((1u << lane) - 1u))) & g.q.mask] = nb;
        g.q.tail += (uint32_t)__popc(pm);
      }
    } else
#endif
      accept_cube(g, r, cid, start, nt, filled);
    MC_PROF_T(t_emitted);
    MC_PROF_ADD(4, t_tiled, t_emitted);
    return true;
  }
  g.done[cid] = 1;
  done_set = true;
  return false;
}

MC_HD void push_neighbours_r(Chain& g, int32_t cid) {
  bool ok = true;
  for (int i = 0; i < 6; ++i) {
    const int32_t n = g.recs[cid].nbr[i];
    if (n != -2) ok = g.q.push(n) && ok;
  }
  if (!ok) g.status = MC_QUEUE_OVERFLOW;
}

// a visit outside the BFS window (raster seed, priority queues): fetch the record first
MC_HD bool visit_direct_r(Chain& g, ChainCache& cc, const Chain* home, int32_t cid, int mode, bool& done_set) {
  MC_WARP_SYNC();
  MC_LANE_LOOP(l) { if (l < 25) copy_record(&cc.rec, g.recs + cid, l); }
  MC_COPY_WAIT();
  MC_WARP_SYNC();
  return visit_r(g, cc.rec, home, cid, mode, done_set);
}

// pyx:1194-1771 on compact ids: raster scan over the candidate list, each still-unvisited candidate seeds a breadth-first
// exploration with the reference's three priority queues.  `g` is the caller's private (per-lane) copy of the counters.
// BFS pops are served from a 32-entry look-ahead window: one cooperative load fetches the next 32 queue entries and
// their visited flags, the records of those that may still be visited are copied to shared memory, and a bit mask of the
// still-unvisited entries lets the chain jump over place holders and already-visited cubes in one step (popping an entry
// that is skipped has no effect other than advancing the head).
MC_HD void replay_r(Chain& g, ChainCache& cc, const Chain* home) {
  g.n_v = 0; g.n_f3 = 0; g.status = MC_OK;
  g.n_seed = g.n_accept = g.n_unsure_push = g.n_nontrivial_push = 0;
  for (int i = 0; i < 8; ++i) g.prof[i] = 0;
  MC_PROF_T(t_replay0);
  MC_WARP_SYNC();
  MC_LANE_LOOP(l) { cc.qw_cur[l] = -1; }
  MC_WARP_SYNC();
  uint32_t qw_base = 0, qw_n = 0, qw_valid = 0;   // window [qw_base, qw_base + qw_n) of queue positions; bit = may be visited
  int32_t nx_c[MC_LANE_SLOTS];                    // per lane: the entry of the window after this one, requested early
  uint32_t nx_base = 0xffffffffu, nx_n = 0;
  for (int i = 0; i < MC_LANE_SLOTS; ++i) nx_c[i] = -1;
  bool done_set = false;
  for (int64_t sw_base = 0; sw_base < g.n_cand; sw_base += 32) {
    // raster window: 32 consecutive candidate ids; bit = not visited yet
    uint32_t sw_open = 0;
    MC_LANE_LOOP(l) { MC_VOTE(sw_open, l, sw_base + l < g.n_cand && !g.done[sw_base + l]); }
    while (sw_open) {
      const int sb = MC_POPC((sw_open & (0u - sw_open)) - 1u);   // lowest open candidate
      sw_open &= sw_open - 1u;
      const int32_t cidx = (int32_t)(sw_base + sb);
      ++g.n_seed;
      const bool accepted = visit_direct_r(g, cc, home, cidx, 0, done_set);
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
              if (g.done[cur]) { g.q_unsure.pop(); continue; }
              push_neighbours_r(g, cur);
              visit_neighbours = false;
              continue;
            } else {
              g.q_unsure.pop();
              visit_neighbours = true;
            }
          }
          if (g.done[cur]) continue;
        } else {
          if (g.q.head - qw_base >= qw_n) {
            // refill: the entries were requested one window ahead (nx_c), so one memory latency covers the visited flags,
            // the records (copied for every real entry: waiting for the flags first would cost a second latency) and the
            // request for the following window's entries
            qw_base = g.q.head;
            const uint32_t avail = g.q.tail - g.q.head;
            qw_n = avail < 32u ? avail : 32u;
            MC_PROF_INC(6);
            qw_valid = 0;
            const bool have = nx_base == qw_base;
            const uint32_t more = avail - qw_n < 32u ? avail - qw_n : 32u;
            MC_WARP_SYNC();
            MC_LANE_LOOP(l) {
              int32_t c = -1;
              if ((uint32_t)l < qw_n) c = (have && (uint32_t)l < nx_n) ? nx_c[MC_LANE_SLOT(l)] : g.q.buf[(qw_base + (uint32_t)l) & g.q.mask];
              if (c >= 0) {
                MC_UNROLL
                for (int k = 0; k < 25; ++k) copy_record(&cc.wrec[l], g.recs + c, k);
              }
              nx_c[MC_LANE_SLOT(l)] = (uint32_t)l < more ? g.q.buf[(qw_base + 32u + (uint32_t)l) & g.q.mask] : -1;
              MC_VOTE(qw_valid, l, c >= 0 && !g.done[c]);
              cc.qw_cur[l] = c;
            }
            nx_base = qw_base + 32u; nx_n = more;
            MC_COPY_WAIT();
            MC_WARP_SYNC();
          }
          const uint32_t w = g.q.head - qw_base;
          const uint32_t ahead = qw_valid >> w;   // still-unvisited entries from the head on
          if (!ahead) { g.q.head = qw_base + qw_n; continue; }   // nothing left in the window: pop it all
          const uint32_t skip = (uint32_t)MC_POPC((ahead & (0u - ahead)) - 1u);
          g.q.head += skip + 1u;
          cur = cc.qw_cur[w + skip];
          visit_r(g, cc.wrec[w + skip], home, cur, visit_neighbours ? 1 : 2, done_set);
          if (done_set) {   // keep the window coherent: other entries naming the same cube are now visited
            uint32_t same = 0;
            MC_LANE_LOOP(l) { MC_VOTE(same, l, cc.qw_cur[l] == cur); }
            qw_valid &= ~same;
            if (cur >= sw_base && cur < sw_base + 32) sw_open &= ~(1u << (cur - sw_base));
          }
          continue;
        }
        visit_direct_r(g, cc, home, cur, visit_neighbours ? 1 : 2, done_set);
        if (done_set) {
          uint32_t same = 0;
          MC_LANE_LOOP(l) { MC_VOTE(same, l, cc.qw_cur[l] == cur); }
          qw_valid &= ~same;
          if (cur >= sw_base && cur < sw_base + 32) sw_open &= ~(1u << (cur - sw_base));
        }
      }
    }
  }
  MC_PROF_T(t_replay1);
  MC_PROF_ADD(0, t_replay0, t_replay1);
  if (g.status == MC_OK && g.n_v == 0) g.status = MC_EMPTY;
}

// ---- after the chain: positions of the vertices an accepted cube created, and its face indices (any thread / any order) ----
MC_HD void emit_cube(const Chain& g, int64_t ai) {
  const Accept a = g.acc[ai];
  const Rec& r = g.recs[a.cid];
  int x, y, z;
  decode_index(r.lattice, g.N, z, y, x);
  double v[8];
  for (int i = 0; i < 8; ++i) {
    // the corners' signs became final when this cube was accepted (signed_im_mask): what the chain saw is what is stored
    float p = (float)st_sign(g.vs[r.vid[i]]) * r.cim[i];
    v[i] = (double)p;
  }
  Cell c;
  cell_set(c, x, y, z, v);
  const int start = (int)(a.tiling & 0xffffu), nt = (int)(a.tiling >> 16);
  for (int k = 0; k < nt * 3; ++k) {
    const int vi = (int)MC_LUTV(start + k);
    const int32_t idx = g.slot[4 * (int64_t)r.vid[edge_corner(vi)] + edge_j(vi)];
    if ((int64_t)a.nf0 + k < g.cap_f3) g.faces[a.nf0 + k] = idx;
    if (idx < a.nv0 || idx >= g.cap_v) continue;   // created by an earlier cube (or no room: the caller retries)
    double px, py, pz;   // pyx:589-675, 806-850
    if (vi == 12) {
      if (!c.v12_done) center_vertex(c);
      px = c.v12x; py = c.v12y; pz = c.v12z;
    } else {
      const int dx1 = LUT2(EDGESRELX, vi, 0), dx2 = LUT2(EDGESRELX, vi, 1);
      const int dy1 = LUT2(EDGESRELY, vi, 0), dy2 = LUT2(EDGESRELY, vi, 1);
      const int dz1 = LUT2(EDGESRELZ, vi, 0), dz2 = LUT2(EDGESRELZ, vi, 1);
      const double w1 = 1.0 / (MC_FLT_EPS + dabs(c.vv[dz1 * 4 + dy1 * 2 + dx1]));
      const double w2 = 1.0 / (MC_FLT_EPS + dabs(c.vv[dz2 * 4 + dy2 * 2 + dx2]));
      double fx = 0.0, fy = 0.0, fz = 0.0, ff = 0.0;
      fx += (double)dx1 * w1; fy += (double)dy1 * w1; fz += (double)dz1 * w1; ff += w1;
      fx += (double)dx2 * w2; fy += (double)dy2 * w2; fz += (double)dz2 * w2; ff += w2;
      px = (double)c.x + 1.0 * fx / ff;
      py = (double)c.y + 1.0 * fy / ff;
      pz = (double)c.z + 1.0 * fz / ff;
    }
    g.verts[3 * (int64_t)idx + 0] = (float)px;
    g.verts[3 * (int64_t)idx + 1] = (float)py;
    g.verts[3 * (int64_t)idx + 2] = (float)pz;
  }
}

}  // namespace surfd_mccore
