// TEST-ONLY host build of the device arithmetic headers (privacy-auction_b200/csrc/*.cuh).
// The headers emulate the PTX carry flag when compiled by g++ (pa_ptx.cuh), so
// the limb algorithms can be checked against Python integers / the oracle on a
// machine without a GPU.  This file is never linked into libpa_engine.so and
// the product has no CPU path.
#include "pa_fe.cuh"
#include <string.h>

extern "C" {
void hc_fe_op(int op, const unsigned char *a, const unsigned char *b, unsigned char *out) {
  fe x, y, r;
  fe_from_be(x, a);
  fe_from_be(y, b);
  switch (op) {
  case 0: fe_mul(r, x, y); break;
  case 1: fe_sqr(r, x); break;
  case 2: fe_add(r, x, y); break;
  case 3: fe_sub(r, x, y); break;
  case 4: fe_inv(r, x); break;
  case 5: fe_neg(r, x); break;
  case 6: r = x; break;
  case 7: { fe_set_zero(r); r.v[0] = fe_is_zero(x); fe_to_be(out, r); return; }
  case 8: { fe_set_zero(r); r.v[0] = fe_eq(x, y); fe_to_be(out, r); return; }
  case 9: fe_shl<1>(r, x); break;
  case 10: fe_shl<2>(r, x); break;
  case 11: fe_shl<3>(r, x); break;
  case 12: fe_mul3(r, x); break;
  case 13: fe_sub2(r, x, y, y); break;
  case 14: fe_sub3(r, x, y, x, y); break;
  case 15: { fe z; fe_neg(z, x); fe_sub3(r, z, y, y, y); break; }
  default: fe_set_zero(r);
  }
  fe_canon(r, r);
  fe_to_be(out, r);
}
void hc_mp_mul8(const unsigned char *a, const unsigned char *b, unsigned char *out64, int sqr) {
  fe x, y;
  fe_from_be(x, a);
  fe_from_be(y, b);
  u32 t[16];
  if (sqr) mp_sqr8(t, x.v); else mp_mul8(t, x.v, y.v);
  for (int i = 0; i < 16; ++i) {
    unsigned char *q = out64 + 4 * (15 - i);
    q[0] = t[i] >> 24; q[1] = t[i] >> 16; q[2] = t[i] >> 8; q[3] = t[i];
  }
}
}

#include "pa_smul.cuh"
#include <vector>
#include <cstdio>
#include <cstdlib>

static std::vector<u32> g_tab;
// The whole signed-window table on the host: per window a running sum e_d = e_(d-1) + B_w in Jacobian
// form and one batched inversion (building every entry with comb_entry, as the GPU does, would
// take minutes here); comb_entry itself is checked against a sample of the entries.
static void ensure_tab() {
  if (!g_tab.empty()) return;
  g_tab.assign(PA_COMB_WORDS, 0);
  aff G;
  aff_set_generator(G);
  std::vector<jac> J(PA_COMB_ENTRIES);
  std::vector<fe> pre(PA_COMB_ENTRIES);
  for (int w = 0; w < PA_COMB_WINDOWS; ++w) {
    aff Bw;
    comb_base(Bw, w, G);
    jac acc;
    jac_from_aff(acc, Bw);
    J[0] = acc;
    for (u32 d = 2; d <= PA_COMB_ENTRIES; ++d) {
      if (d == 2) jac_dbl(acc, acc); else jac_madd(acc, acc, Bw);
      J[d - 1] = acc;
    }
    fe run;
    fe_set_one(run);
    for (u32 i = 0; i < PA_COMB_ENTRIES; ++i) {
      pre[i] = run;
      fe_mul(run, run, J[i].Z);
    }
    fe inv;
    fe_inv(inv, run);
    for (int i = PA_COMB_ENTRIES - 1; i >= 0; --i) {
      fe zi;
      fe_mul(zi, inv, pre[i]);
      fe_mul(inv, inv, J[i].Z);
      aff e;
      jac_to_aff_with_zinv(e, J[i], zi);
      u32 *o = g_tab.data() + ((size_t)w * PA_COMB_ENTRIES + i) * 16;
      for (int k = 0; k < 8; ++k) { o[k] = e.x.v[k]; o[8 + k] = e.y.v[k]; }
    }
    for (u32 d : {1u, 2u, 3u, (u32)PA_COMB_ENTRIES / 3, (u32)PA_COMB_ENTRIES - 1, (u32)PA_COMB_ENTRIES}) {
      aff e;
      comb_entry(e, d, Bw);
      const u32 *o = g_tab.data() + ((size_t)w * PA_COMB_ENTRIES + (d - 1)) * 16;
      for (int k = 0; k < 8; ++k)
        if (o[k] != e.x.v[k] || o[8 + k] != e.y.v[k]) { fprintf(stderr, "hostcheck: comb_entry(%d, %u) disagrees with the running sum\n", w, d); abort(); }
    }
  }
}
static void out_jac(unsigned char *out, const jac &r) {
  aff a;
  jac_to_aff(a, r);
  aff_to_be64(out, a);
}
extern "C" {
void hc_sc_op(int op, const unsigned char *a, const unsigned char *b, unsigned char *out) {
  sc x, y, r;
  sc_from_be(x, a);
  sc_from_be(y, b);
  switch (op) {
  case 0: sc_mul(r, x, y); break;
  case 1: sc_add(r, x, y); break;
  case 2: sc_sub(r, x, y); break;
  case 3: sc_neg(r, x); break;
  default: r = x;
  }
  sc_to_be(out, r);
}
void hc_fixed_base(const unsigned char *k, unsigned char *out) {
  ensure_tab();
  sc s; sc_from_be(s, k);
  jac r; fixed_base_mul(r, s, g_tab.data());
  out_jac(out, r);
}
void hc_var_base(const unsigned char *p, const unsigned char *k, unsigned char *out) {
  aff a; aff_from_be64(a, p);
  jac P; jac_from_aff(P, a);
  sc s; sc_from_be(s, k);
  jac r; var_base_mul(r, P, s);
  out_jac(out, r);
}
void hc_lincomb2(const unsigned char *p, const unsigned char *ka, const unsigned char *q, const unsigned char *kb, unsigned char *out) {
  aff a, b; aff_from_be64(a, p); aff_from_be64(b, q);
  jac P, Q; jac_from_aff(P, a); jac_from_aff(Q, b);
  sc s, t; sc_from_be(s, ka); sc_from_be(t, kb);
  jac r; strauss<2>(r, P, s, Q, t);
  out_jac(out, r);
}
void hc_point_add(const unsigned char *p, const unsigned char *q, unsigned char *out, int mixed) {
  aff a, b; aff_from_be64(a, p); aff_from_be64(b, q);
  jac P, Q, r; jac_from_aff(P, a); jac_from_aff(Q, b);
  // give P a non-trivial Z so the Jacobian paths are exercised: (X*4, Y*8, 2)
  if (!jac_is_inf(P)) { fe two; fe_set_zero(two); two.v[0] = 2; fe t; fe_add(t, P.X, P.X); fe_add(P.X, t, t); fe_add(t, P.Y, P.Y); fe_add(t, t, t); fe_add(P.Y, t, t); P.Z = two; }
  if (mixed) jac_madd(r, P, b); else {
    if (!jac_is_inf(Q)) { fe three; fe_set_zero(three); three.v[0] = 3; fe z2, z3; fe_sqr(z2, three); fe_mul(z3, z2, three); fe_mul(Q.X, Q.X, z2); fe_mul(Q.Y, Q.Y, z3); Q.Z = three; }
    jac_add(r, P, Q);
  }
  out_jac(out, r);
}
int hc_jac_eq_aff(const unsigned char *p, const unsigned char *q) {
  aff a, b; aff_from_be64(a, p); aff_from_be64(b, q);
  jac P; jac_from_aff(P, a);
  if (!jac_is_inf(P)) { fe z; fe_set_zero(z); z.v[0] = 5; fe z2, z3; fe_sqr(z2, z); fe_mul(z3, z2, z); fe_mul(P.X, P.X, z2); fe_mul(P.Y, P.Y, z3); P.Z = z; }
  return jac_eq_aff(P, b) ? 1 : 0;
}
// the GLV halves of k: out = |k1| (5 limbs LE), |k2| (5 limbs LE), neg1, neg2 as 42 bytes
void hc_glv_split(const unsigned char *k, unsigned char *out) {
  sc s; sc_from_be(s, k);
  glv_split g;
  glv_decompose(g, s);
  for (int i = 0; i < 5; ++i) for (int b = 0; b < 4; ++b) { out[4 * i + b] = (unsigned char)(g.k1[i] >> (8 * b)); out[20 + 4 * i + b] = (unsigned char)(g.k2[i] >> (8 * b)); }
  out[40] = g.neg1 ? 1 : 0;
  out[41] = g.neg2 ? 1 : 0;
}
const unsigned int *hc_comb_table() { ensure_tab(); return g_tab.data(); }
}

#include "pa_proof.cuh"

template <int KIND> static unsigned verify_all(const unsigned char *proof, const unsigned char *stmt, u64 id, int nchk) {
  ensure_tab();
  sc ch1;
  verify_derive<KIND>(ch1, proof, stmt, id);
  unsigned mask = 0;
  for (int j = 0; j < nchk; ++j)
    if (verify_check_one<KIND>(j, proof, stmt, ch1, g_tab.data())) mask |= 1u << j;
  return mask;
}
template <int KIND> static void prove_all(unsigned char *proof, const unsigned char *stmt, u64 id, const unsigned char *secrets,
                                          const unsigned char *rnd, int branch) {
  typedef proof_kind<KIND> K;
  ensure_tab();
  for (int j = 0; j < K::NEPS; ++j) {
    jac r;
    int e = prove_op_one<KIND>(r, branch, j, stmt, rnd, g_tab.data());
    out_jac(proof + 64 * e, r);
  }
  prove_respond<KIND>(proof, stmt, id, secrets, rnd, branch);
}
// the prover with witnesses (pa_proof.cuh): extended secrets; must publish the same bytes as prove_all
template <int KIND> static void prove_all_wit(unsigned char *proof, const unsigned char *stmt, u64 id, const unsigned char *secrets,
                                              const unsigned char *rnd, int branch, int veto_i, int veto_j, int cbit) {
  typedef proof_kind<KIND> K;
  ensure_tab();
  for (int j = 0; j < K::NEPS; ++j) {
    jac r;
    int e = prove_op_one_wit<KIND>(r, branch, j, stmt, rnd, secrets, veto_i, veto_j, cbit, g_tab.data());
    out_jac(proof + 64 * e, r);
  }
  prove_respond<KIND>(proof, stmt, id, secrets, rnd, branch);
}
extern "C" {
void hc_proof_prove_w(int kind, unsigned char *proof, const unsigned char *stmt, unsigned long long id,
                      const unsigned char *secrets, const unsigned char *rnd, int branch, int veto_i, int veto_j, int cbit) {
  switch (kind) {
  case PA_COM: prove_all_wit<PA_COM>(proof, stmt, id, secrets, rnd, branch, veto_i, veto_j, cbit); break;
  case PA_S1: prove_all_wit<PA_S1>(proof, stmt, id, secrets, rnd, branch, veto_i, veto_j, cbit); break;
  default: prove_all_wit<PA_S2>(proof, stmt, id, secrets, rnd, branch, veto_i, veto_j, cbit);
  }
}
unsigned hc_proof_verify(int kind, const unsigned char *proof, const unsigned char *stmt, unsigned long long id) {
  switch (kind) {
  case PA_POK: return verify_all<PA_POK>(proof, stmt, id, 1);
  case PA_COM: return verify_all<PA_COM>(proof, stmt, id, 4);
  case PA_S1: return verify_all<PA_S1>(proof, stmt, id, 8);
  default: return verify_all<PA_S2>(proof, stmt, id, 16);
  }
}
void hc_proof_prove(int kind, unsigned char *proof, const unsigned char *stmt, unsigned long long id,
                    const unsigned char *secrets, const unsigned char *rnd, int branch) {
  switch (kind) {
  case PA_POK: prove_all<PA_POK>(proof, stmt, id, secrets, rnd, branch); break;
  case PA_COM: prove_all<PA_COM>(proof, stmt, id, secrets, rnd, branch); break;
  case PA_S1: prove_all<PA_S1>(proof, stmt, id, secrets, rnd, branch); break;
  default: prove_all<PA_S2>(proof, stmt, id, secrets, rnd, branch);
  }
}
void hc_stream_draw(unsigned long long seed, unsigned long long stream, unsigned long long ctr, unsigned char *out) {
  u32 d[8];
  pa_stream_draw(d, seed, stream, ctr);
  for (int i = 0; i < 8; ++i) { out[4*i] = d[i] >> 24; out[4*i+1] = d[i] >> 16; out[4*i+2] = d[i] >> 8; out[4*i+3] = d[i]; }
}
void hc_challenge(const unsigned char *pts, int k, unsigned long long id, unsigned char *out) {
  const unsigned char *p[32];
  for (int i = 0; i < k; ++i) p[i] = pts + 64 * i;
  sc h;
  challenge_hash(h, p, k, id);
  sc_to_be(out, h);
}
}
