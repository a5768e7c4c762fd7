// TEST INFRASTRUCTURE: host build of surfd_b200/csrc/mc_core.h so the replay logic can be checked
// against the compiled reference on a machine without a GPU.  Not part of the product path.
// g++ -O2 -ffp-contract=off -shared -fPIC -I surfd_b200/csrc tests/native/mc_host.cpp -o tests/native/_mc_host.so
#include <cstdlib>
#include <cstring>
#include <vector>
#include "mc_chain.h"
using namespace surfd_mccore;

static int run_impl(const float* im, const float* grads, int N, float* verts, int64_t cap_v, int32_t* faces, int64_t cap_f3,
                    int64_t* n_v, int64_t* n_f3, int64_t* stats, int warp_variant);

extern "C" int mc_host_run(const float* im, const float* grads, int N, float* verts, int64_t cap_v,
                           int32_t* faces, int64_t cap_f3, int64_t* n_v, int64_t* n_f3, int64_t* stats) {
  return run_impl(im, grads, N, verts, cap_v, faces, cap_f3, n_v, n_f3, stats, 0);
}
// the device pipeline (records -> compact-state chain -> parallel emission, mc_chain.h) with the chain warp's lane loops run
// sequentially: what mc.cu executes, emulated
extern "C" int mc_host_run_r(const float* im, const float* grads, int N, float* verts, int64_t cap_v,
                             int32_t* faces, int64_t cap_f3, int64_t* n_v, int64_t* n_f3, int64_t* stats) {
  return run_impl(im, grads, N, verts, cap_v, faces, cap_f3, n_v, n_f3, stats, 1);
}

static int run_impl(const float* im, const float* grads, int N, float* verts, int64_t cap_v, int32_t* faces, int64_t cap_f3,
                    int64_t* n_v, int64_t* n_f3, int64_t* stats, int warp_variant) {
  const int64_t n3 = (int64_t)N * N * N;
  // candidate classification, same arithmetic as pyx:1157-1158,1215-1218,1825-1841
  const double voxel = 2.0 / (N - 1);
  const float avg_t = (float)(1.05 * voxel), max_t = (float)(1.74 * voxel);
  std::vector<uint32_t> bits((n3 + 31) / 32, 0u);
  std::vector<int32_t> list;
  for (int z = 0; z < N - 1; ++z)
    for (int y = 0; y < N - 1; ++y)
      for (int x = 0; x < N - 1; ++x) {
        const int64_t i = ((int64_t)z * N + y) * N + x;
        const float v1 = im[i], v2 = im[i + 1], v3 = im[i + N + 1], v4 = im[i + N];
        const int64_t j = i + (int64_t)N * N;
        const float v5 = im[j], v6 = im[j + 1], v7 = im[j + N + 1], v8 = im[j + N];
        float s = v1 + v2; s = s + v3; s = s + v4; s = s + v5; s = s + v6; s = s + v7; s = s + v8;
        const float avg = (float)(0.125 * (double)s);
        float m = v7 > v8 ? v7 : v8;
        m = v6 > m ? v6 : m; m = v5 > m ? v5 : m; m = v4 > m ? v4 : m; m = v3 > m ? v3 : m; m = v2 > m ? v2 : m; m = v1 > m ? v1 : m;
        if (avg < avg_t && m <= max_t) { bits[i >> 5] |= 1u << (i & 31); list.push_back((int32_t)i); }
      }
  if (warp_variant) {
    const int64_t n_words = (n3 + 31) / 32;
    std::vector<uint32_t> vbits(n_words, 0u);
    for (int32_t i : list)
      for (int c = 0; c < 8; ++c) {
        const int64_t v = i + (int64_t)MC_CZ(c) * N * N + (int64_t)MC_CY(c) * N + MC_CX(c);
        vbits[v >> 5] |= 1u << (v & 31);
      }
    std::vector<int32_t> cpre(n_words), vpre(n_words);
    int64_t nc = 0, nv = 0;
    for (int64_t w = 0; w < n_words; ++w) { cpre[w] = (int32_t)nc; vpre[w] = (int32_t)nv; nc += __builtin_popcount(bits[w]); nv += __builtin_popcount(vbits[w]); }
    Chain g;
    memset(&g, 0, sizeof(g));
    g.N = N; g.im = im; g.grads = grads; g.cand_bits = bits.data(); g.cand_prefix = cpre.data(); g.vtx_bits = vbits.data(); g.vtx_prefix = vpre.data();
    g.cand_list = list.data(); g.n_cand = g.n_cand_total = (int64_t)list.size(); g.n_vtx = g.cap_vtx = nv;
    std::vector<Rec> recs(list.size() + 1);
    std::vector<uint8_t> vs(nv + 1, 0), done(list.size() + 1, 0);
    std::vector<int32_t> slot(4 * nv + 4, -1);
    std::vector<Accept> acc(list.size() + 1);
    g.recs = recs.data(); g.vs = vs.data(); g.done = done.data(); g.slot = slot.data(); g.acc = acc.data();
    g.verts = verts; g.cap_v = cap_v; g.faces = faces; g.cap_f3 = cap_f3;
    uint32_t cap = 1024; while (cap < 8u * list.size() + 1024u) cap <<= 1;
    std::vector<int32_t> b0(cap), b1(cap), b2(cap);
    g.q.buf = b0.data(); g.q_unsure.buf = b1.data(); g.q_nontrivial.buf = b2.data();
    g.q.mask = g.q_unsure.mask = g.q_nontrivial.mask = cap - 1;
    for (int64_t k = 0; k < g.n_cand; ++k) build_record(g, k);
    ChainCache cc;
    const Chain home = g;
    replay_r(g, cc, &home);
    for (int64_t a = 0; a < g.n_accept; ++a) emit_cube(g, a);
    *n_v = g.n_v; *n_f3 = g.n_f3;
    if (stats) { stats[0] = g.n_cand; stats[1] = g.n_seed; stats[2] = g.n_accept; stats[3] = g.n_unsure_push; stats[4] = g.n_nontrivial_push; }
    return g.status;
  }
  Grid g;
  memset(&g, 0, sizeof(g));
  g.N = N; g.im = im; g.grads = grads; g.cand_bits = bits.data(); g.cand_list = list.data(); g.n_cand = (int64_t)list.size();
  std::vector<int8_t> sgn(n3, 0); std::vector<uint8_t> flg(n3, 0); std::vector<int32_t> fl(4 * n3, -1);
  g.sgn = sgn.data(); g.flg = flg.data(); g.face_layer = fl.data();
  g.verts = verts; g.cap_v = cap_v; g.faces = faces; g.cap_f3 = cap_f3;
  uint32_t cap = 1024; while (cap < 16u * list.size() + 1024u) cap <<= 1;
  std::vector<int32_t> b0(cap), b1(cap), b2(cap);
  g.q.buf = b0.data(); g.q_unsure.buf = b1.data(); g.q_nontrivial.buf = b2.data();
  g.q.mask = g.q_unsure.mask = g.q_nontrivial.mask = cap - 1;
  replay(g);
  *n_v = g.n_v; *n_f3 = g.n_f3;
  if (stats) { stats[0] = g.n_cand; stats[1] = g.n_seed; stats[2] = g.n_accept; stats[3] = g.n_unsure_push; stats[4] = g.n_nontrivial_push; }
  return g.status;
}
