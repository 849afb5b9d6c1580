// The four non-interactive zero-knowledge proofs of the SEAL protocol, as data:
//   POK  NIZKPoKDLog      Schnorr proof of knowledge of a discrete log      SEAL/bidder.cpp:90-136
//   COM  NIZKPoWFCom      1-of-2 OR proof, commitment well-formedness        SEAL/bidder.cpp:149-299
//   S1   NIZKPoWFStage1   1-of-2 OR proof, cryptogram before the junction    SEAL/bidder.cpp:318-571
//   S2   NIZKPoWFStage2   1-of-3 OR proof, cryptogram after the junction     SEAL/bidder.cpp:598-1101
//
// Every point the provers publish and every check the verifiers run has the
// shape  a*P + b*Q  (P is the generator in half of them), so a proof kind is
// described by small tables — which point and scalar feed each operation —
// and ONE generic device routine (lincomb_op) does all the curve work.  A
// thread owns one operation (not one proof): a stage-2 verification is 16
// independent threads.  Tables are ordered so that, whatever branch of the OR
// a prover is in, operation j has the same shape in every lane of a warp.
//
// Wire layouts (bytes; points 64, scalars 32, in the order of SEAL/types.h):
//   POK  eps | rho                                                     =   96
//   COM  eps11 eps12 eps21 eps22 | rho1 rho2 ch2                       =  352
//   S1   eps11..14 eps21..24 | rho11 rho12 rho21 rho22 ch2             =  672
//   S2   eps11 eps12 eps13 eps11' eps12' eps13' eps21 eps22 eps23 eps21' eps22'
//        eps23' eps31 eps32 eps31' eps32' | rho11 rho12 rho13 rho21 rho22 rho23
//        rho31 rho32 ch2 ch3                                           = 1344
// Statement points are passed in the order of the reference's parameter lists:
//   POK (X)   COM (phi, A, B)   S1 (b, X, Y, R, c, A, B)
//   S2 (Bi, Xi, Ri, Bj, Xj, Rj, Ci, A, B, Yi, Yj)
#pragma once
#include "pa_sha256.cuh"
#include "pa_smul.cuh"

enum { PA_POK = 0, PA_COM = 1, PA_S1 = 2, PA_S2 = 3 };

// point source codes
#define PS_G (-1)      // the generator
#define PS_NONE (-2)   // term absent
#define PS_STMT 16     // PS_STMT + k : statement point k
#define PS_CG 48       // (commitment point) - g   ("phi/g", "c/g", "Ci/g")
// scalar source codes, verifier: 0.. = published scalar k, VS_CH1 = ch - ch2 (- ch3)
#define VS_CH1 16
// scalar source codes, prover: 0.. = drawn scalar k (draw order), SS_ZERO = 0
#define SS_ZERO 16

struct pa_op {
  signed char P, a, Q, b, E;  // E: index of the eps this operation produces / is compared with
};

template <int KIND> struct proof_kind;

template <> struct proof_kind<PA_POK> {
  static constexpr int NEPS = 1, NSC = 1, NSTMT = 1, NRND = 1, NSECRET = 1, NBRANCH = 1, NHASH = 2, CG_STMT = 0;
  static constexpr int REC = NEPS * 64 + NSC * 32;
};
template <> struct proof_kind<PA_COM> {
  static constexpr int NEPS = 4, NSC = 3, NSTMT = 3, NRND = 3, NSECRET = 1, NBRANCH = 2, NHASH = 7, CG_STMT = 0;
  static constexpr int REC = NEPS * 64 + NSC * 32;
};
template <> struct proof_kind<PA_S1> {
  static constexpr int NEPS = 8, NSC = 5, NSTMT = 7, NRND = 5, NSECRET = 2, NBRANCH = 2, NHASH = 15, CG_STMT = 4;
  static constexpr int REC = NEPS * 64 + NSC * 32;
};
template <> struct proof_kind<PA_S2> {
  static constexpr int NEPS = 16, NSC = 10, NSTMT = 11, NRND = 11, NSECRET = 3, NBRANCH = 3, NHASH = 27, CG_STMT = 6;
  static constexpr int REC = NEPS * 64 + NSC * 32;
};

// ---- Fiat-Shamir input order (after the generator): eps k, or PS_STMT + k -------
template <int KIND> PA_HD int hash_src(int i);
template <> PA_HD int hash_src<PA_POK>(int i) {  // {g, g^v, g^x}            SEAL/hash.cpp:25
  const signed char t[2] = {0, PS_STMT + 0};
  return t[i];
}
template <> PA_HD int hash_src<PA_COM>(int i) {  // {g, eps11, eps12, eps21, eps22, phi, A, B}   SEAL/hash.cpp:75
  const signed char t[7] = {0, 1, 2, 3, PS_STMT + 0, PS_STMT + 1, PS_STMT + 2};
  return t[i];
}
template <> PA_HD int hash_src<PA_S1>(int i) {  // {g, eps11..24, b, X, Y, R, c, A, B}           SEAL/hash.cpp:130-132
  const signed char t[15] = {0, 1, 2, 3, 4, 5, 6, 7, PS_STMT + 0, PS_STMT + 1, PS_STMT + 2, PS_STMT + 3,
                             PS_STMT + 4, PS_STMT + 5, PS_STMT + 6};
  return t[i];
}
template <> PA_HD int hash_src<PA_S2>(int i) {
  // {g, 16 eps, Xi, Xj, A, Bi, Bj, B, Ri, Rj, Ci, Yi, Yj}                                    SEAL/hash.cpp:191-196
  const signed char t[27] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
                             PS_STMT + 1, PS_STMT + 4, PS_STMT + 7, PS_STMT + 0, PS_STMT + 3, PS_STMT + 8,
                             PS_STMT + 2, PS_STMT + 5, PS_STMT + 6, PS_STMT + 9, PS_STMT + 10};
  return t[i];
}

// ---- verification checks:  a*P + b*Q == eps[E] ----------------------------------
template <int KIND> PA_HD pa_op verify_op(int j);
template <> PA_HD pa_op verify_op<PA_POK>(int) {  // g^rho * X^h == eps         SEAL/bidder.cpp:127-131
  return pa_op{PS_G, 0, PS_STMT + 0, VS_CH1, 0};
}
template <> PA_HD pa_op verify_op<PA_COM>(int j) {  // SEAL/bidder.cpp:255-296; scalars rho1=0 rho2=1 ch2=2
  const pa_op t[4] = {
      {PS_G, 0, PS_STMT + 1, VS_CH1, 0},         // g^rho1 * A^ch1 == eps11
      {PS_G, 1, PS_STMT + 1, 2, 2},              // g^rho2 * A^ch2 == eps21
      {PS_STMT + 2, 0, PS_STMT + 0, VS_CH1, 1},  // B^rho1 * phi^ch1 == eps12
      {PS_STMT + 2, 1, PS_CG, 2, 3},             // B^rho2 * (phi/g)^ch2 == eps22
  };
  return t[j];
}
template <> PA_HD pa_op verify_op<PA_S1>(int j) {
  // SEAL/bidder.cpp:487-568; stmt b=0 X=1 Y=2 R=3 c=4 A=5 B=6; scalars rho11=0 rho12=1 rho21=2 rho22=3 ch2=4
  const pa_op t[8] = {
      {PS_G, 0, PS_STMT + 1, VS_CH1, 0},         // g^rho11 * X^ch1 == eps11
      {PS_G, 1, PS_STMT + 5, VS_CH1, 1},         // g^rho12 * A^ch1 == eps12
      {PS_G, 2, PS_STMT + 1, 4, 4},              // g^rho21 * X^ch2 == eps21
      {PS_G, 3, PS_STMT + 5, 4, 5},              // g^rho22 * A^ch2 == eps22
      {PS_STMT + 2, 0, PS_STMT + 0, VS_CH1, 2},  // Y^rho11 * b^ch1 == eps13
      {PS_STMT + 6, 1, PS_STMT + 4, VS_CH1, 3},  // B^rho12 * c^ch1 == eps14
      {PS_STMT + 3, 2, PS_STMT + 0, 4, 6},       // R^rho21 * b^ch2 == eps23
      {PS_STMT + 6, 3, PS_CG, 4, 7},             // B^rho22 * (c/g)^ch2 == eps24
  };
  return t[j];
}
template <> PA_HD pa_op verify_op<PA_S2>(int j) {
  // SEAL/bidder.cpp:937-1098; stmt Bi=0 Xi=1 Ri=2 Bj=3 Xj=4 Rj=5 Ci=6 A=7 B=8 Yi=9 Yj=10
  // scalars rho11=0 rho12=1 rho13=2 rho21=3 rho22=4 rho23=5 rho31=6 rho32=7 ch2=8 ch3=9
  const pa_op t[16] = {
      {PS_G, 0, PS_STMT + 1, VS_CH1, 0},   // check 1   g^rho11 * Xi^ch1 == eps11
      {PS_G, 1, PS_STMT + 4, VS_CH1, 1},   // check 2   g^rho12 * Xj^ch1 == eps12
      {PS_G, 2, PS_STMT + 7, VS_CH1, 2},   // check 3   g^rho13 * A^ch1  == eps13
      {PS_G, 3, PS_STMT + 1, 8, 6},        // check 7   g^rho21 * Xi^ch2 == eps21
      {PS_G, 4, PS_STMT + 4, 8, 7},        // check 8   g^rho22 * Xj^ch2 == eps22
      {PS_G, 5, PS_STMT + 7, 8, 8},        // check 9   g^rho23 * A^ch2  == eps23
      {PS_G, 6, PS_STMT + 1, 9, 12},       // check 13  g^rho31 * Xi^ch3 == eps31
      {PS_G, 7, PS_STMT + 4, 9, 13},       // check 14  g^rho32 * Xj^ch3 == eps32
      {PS_STMT + 2, 0, PS_STMT + 0, VS_CH1, 3},  // check 4   Ri^rho11 * Bi^ch1 == eps11'
      {PS_STMT + 5, 1, PS_STMT + 3, VS_CH1, 4},  // check 5   Rj^rho12 * Bj^ch1 == eps12'
      {PS_STMT + 8, 2, PS_CG, VS_CH1, 5},        // check 6   B^rho13 * (Ci/g)^ch1 == eps13'
      {PS_STMT + 9, 3, PS_STMT + 0, 8, 9},       // check 10  Yi^rho21 * Bi^ch2 == eps21'
      {PS_STMT + 5, 4, PS_STMT + 3, 8, 10},      // check 11  Rj^rho22 * Bj^ch2 == eps22'
      {PS_STMT + 8, 5, PS_STMT + 6, 8, 11},      // check 12  B^rho23 * Ci^ch2 == eps23'
      {PS_STMT + 9, 6, PS_STMT + 0, 9, 14},      // check 15  Yi^rho31 * Bi^ch3 == eps31'
      {PS_STMT + 10, 7, PS_STMT + 3, 9, 15},     // check 16  Yj^rho32 * Bj^ch3 == eps32'
  };
  return t[j];
}

// ---- prover: eps[E] = a*P + b*Q, scalars are drawn values (draw order)
template <int KIND> PA_HD pa_op prove_op(int branch, int j);
template <> PA_HD pa_op prove_op<PA_POK>(int, int) {  // eps = g^v           SEAL/bidder.cpp:97-98
  return pa_op{PS_G, 0, PS_NONE, 0, 0};
}
template <> PA_HD pa_op prove_op<PA_COM>(int branch, int j) {
  // draws: r1=0, then bit 0: ch2=1 rho2=2 (SEAL/bidder.cpp:165-169) / bit 1: ch1=1 rho1=2 (:187-188)
  const pa_op t[2][4] = {
      {{PS_G, 0, PS_NONE, 0, 0},               // eps11 = g^r1                 :171
       {PS_STMT + 2, 0, PS_NONE, 0, 1},        // eps12 = B^r1                 :173
       {PS_G, 2, PS_STMT + 1, 1, 2},           // eps21 = g^rho2 * A^ch2       :175
       {PS_STMT + 2, 2, PS_CG, 1, 3}},         // eps22 = B^rho2 * (phi/g)^ch2 :178-185
      {{PS_G, 0, PS_NONE, 0, 2},               // eps21 = g^r1                 :200
       {PS_STMT + 2, 0, PS_NONE, 0, 3},        // eps22 = B^r1                 :202
       {PS_G, 2, PS_STMT + 1, 1, 0},           // eps11 = g^rho1 * A^ch1       :190-193
       {PS_STMT + 2, 2, PS_STMT + 0, 1, 1}},   // eps12 = B^rho1 * phi^ch1     :195-198
  };
  return t[branch][j];
}
template <> PA_HD pa_op prove_op<PA_S1>(int branch, int j) {
  // draws: r11=0 r12=1, then bit 0: rho21=2 rho22=3 ch2=4 (SEAL/bidder.cpp:347-353) / bit 1: rho11=2 rho12=3 ch1=4 (:386-388)
  const pa_op t[2][8] = {
      {{PS_G, 0, PS_NONE, 0, 0},               // eps11 = g^r11                :355
       {PS_G, 1, PS_NONE, 0, 1},               // eps12 = g^r12                :357
       {PS_STMT + 2, 0, PS_NONE, 0, 2},        // eps13 = Y^r11                :359
       {PS_STMT + 6, 1, PS_NONE, 0, 3},        // eps14 = B^r12                :361
       {PS_G, 2, PS_STMT + 1, 4, 4},           // eps21 = g^rho21 * X^ch2      :364-366
       {PS_G, 3, PS_STMT + 5, 4, 5},           // eps22 = g^rho22 * A^ch2      :369-371
       {PS_STMT + 3, 2, PS_STMT + 0, 4, 6},    // eps23 = R^rho21 * b^ch2      :374-376
       {PS_STMT + 6, 3, PS_CG, 4, 7}},         // eps24 = B^rho22 * (c/g)^ch2  :379-384
      {{PS_G, 0, PS_NONE, 0, 4},               // eps21 = g^r11                :410
       {PS_G, 1, PS_NONE, 0, 5},               // eps22 = g^r12                :412
       {PS_STMT + 3, 0, PS_NONE, 0, 6},        // eps23 = R^r11                :414
       {PS_STMT + 6, 1, PS_NONE, 0, 7},        // eps24 = B^r12                :417
       {PS_G, 2, PS_STMT + 1, 4, 0},           // eps11 = g^rho11 * X^ch1      :391-393
       {PS_G, 3, PS_STMT + 5, 4, 1},           // eps12 = g^rho12 * A^ch1      :396-398
       {PS_STMT + 2, 2, PS_STMT + 0, 4, 2},    // eps13 = Y^rho11 * b^ch1      :401-403
       {PS_STMT + 6, 3, PS_STMT + 4, 4, 3}},   // eps14 = B^rho12 * c^ch1      :406-408
  };
  return t[branch][j];
}
template <> PA_HD pa_op prove_op<PA_S2>(int branch, int j) {
  // draws 0..2 = r11 r12 r13 (SEAL/bidder.cpp:643-645), then
  //   branch 0 (bi=1):       rho21=3 rho22=4 rho23=5 rho31=6 rho32=7 rho33=8(unused) ch2=9 ch3=10   :648-655
  //   branch 1 (bi=0,bj=1):  rho11=3 rho12=4 rho13=5 rho31=6 rho32=7 rho33=8(unused) ch1=9 ch3=10   :692-699
  //   branch 2 (bj=0):       3,4,5 overwritten; rho21=6 rho22=7 rho23=8 ch1=9 ch2=10 :749-756
  //                          rho11 = rho12 = rho13 = 0
  // stmt Bi=0 Xi=1 Ri=2 Bj=3 Xj=4 Rj=5 Ci=6 A=7 B=8 Yi=9 Yj=10
  const pa_op t[3][16] = {
      {{PS_G, 0, PS_NONE, 0, 0},                 // eps11  = g^r11              :657
       {PS_G, 1, PS_NONE, 0, 1},                 // eps12  = g^r12              :658
       {PS_G, 2, PS_NONE, 0, 2},                 // eps13  = g^r13              :659
       {PS_STMT + 2, 0, PS_NONE, 0, 3},          // eps11' = Ri^r11             :660
       {PS_STMT + 5, 1, PS_NONE, 0, 4},          // eps12' = Rj^r12             :661
       {PS_STMT + 8, 2, PS_NONE, 0, 5},          // eps13' = B^r13              :662
       {PS_G, 3, PS_STMT + 1, 9, 6},             // eps21  = g^rho21 * Xi^ch2   :664
       {PS_G, 4, PS_STMT + 4, 9, 7},             // eps22  = g^rho22 * Xj^ch2   :665
       {PS_G, 5, PS_STMT + 7, 9, 8},             // eps23  = g^rho23 * A^ch2    :666
       {PS_G, 6, PS_STMT + 1, 10, 12},           // eps31  = g^rho31 * Xi^ch3   :680
       {PS_G, 7, PS_STMT + 4, 10, 13},           // eps32  = g^rho32 * Xj^ch3   :681
       {PS_STMT + 9, 3, PS_STMT + 0, 9, 9},      // eps21' = Yi^rho21 * Bi^ch2  :668-670
       {PS_STMT + 5, 4, PS_STMT + 3, 9, 10},     // eps22' = Rj^rho22 * Bj^ch2  :672-674
       {PS_STMT + 8, 5, PS_STMT + 6, 9, 11},     // eps23' = B^rho23 * Ci^ch2   :676-678
       {PS_STMT + 9, 6, PS_STMT + 0, 10, 14},    // eps31' = Yi^rho31 * Bi^ch3  :683-685
       {PS_STMT + 10, 7, PS_STMT + 3, 10, 15}},  // eps32' = Yj^rho32 * Bj^ch3  :687-689
      {{PS_G, 0, PS_NONE, 0, 6},                 // eps21  = g^r11              :701
       {PS_G, 1, PS_NONE, 0, 7},                 // eps22  = g^r12              :702
       {PS_G, 2, PS_NONE, 0, 8},                 // eps23  = g^r13              :703
       {PS_STMT + 9, 0, PS_NONE, 0, 9},          // eps21' = Yi^r11             :704
       {PS_STMT + 5, 1, PS_NONE, 0, 10},         // eps22' = Rj^r12             :705
       {PS_STMT + 8, 2, PS_NONE, 0, 11},         // eps23' = B^r13              :706
       {PS_G, 3, PS_STMT + 1, 9, 0},             // eps11  = g^rho11 * Xi^ch1   :709-711
       {PS_G, 4, PS_STMT + 4, 9, 1},             // eps12  = g^rho12 * Xj^ch1   :713-715
       {PS_G, 5, PS_STMT + 7, 9, 2},             // eps13  = g^rho13 * A^ch1    :717-719
       {PS_G, 6, PS_STMT + 1, 10, 12},           // eps31  = g^rho31 * Xi^ch3   :736
       {PS_G, 7, PS_STMT + 4, 10, 13},           // eps32  = g^rho32 * Xj^ch3   :738
       {PS_STMT + 2, 3, PS_STMT + 0, 9, 3},      // eps11' = Ri^rho11 * Bi^ch1  :721-723
       {PS_STMT + 5, 4, PS_STMT + 3, 9, 4},      // eps12' = Rj^rho12 * Bj^ch1  :725-727
       {PS_STMT + 8, 5, PS_CG, 9, 5},            // eps13' = B^rho13 * (Ci/g)^ch1 :729-734
       {PS_STMT + 9, 6, PS_STMT + 0, 10, 14},    // eps31' = Yi^rho31 * Bi^ch3  :741-743
       {PS_STMT + 10, 7, PS_STMT + 3, 10, 15}},  // eps32' = Yj^rho32 * Bj^ch3  :745-747
      {{PS_G, 0, PS_NONE, 0, 12},                // eps31  = g^r11              :811
       {PS_G, 1, PS_NONE, 0, 13},                // eps32  = g^r12              :812
       {PS_G, SS_ZERO, PS_STMT + 1, 9, 0},       // eps11  = g^0 * Xi^ch1       :759-761
       {PS_STMT + 9, 0, PS_NONE, 0, 14},         // eps31' = Yi^r11             :813
       {PS_STMT + 10, 1, PS_NONE, 0, 15},        // eps32' = Yj^r12             :814
       {PS_STMT + 2, SS_ZERO, PS_STMT + 0, 9, 3},  // eps11' = Ri^0 * Bi^ch1    :771-773
       {PS_G, SS_ZERO, PS_STMT + 4, 9, 1},       // eps12  = g^0 * Xj^ch1       :763-765
       {PS_G, SS_ZERO, PS_STMT + 7, 9, 2},       // eps13  = g^0 * A^ch1        :767-769
       {PS_G, 6, PS_STMT + 1, 10, 6},            // eps21  = g^rho21 * Xi^ch2   :787-789
       {PS_G, 7, PS_STMT + 4, 10, 7},            // eps22  = g^rho22 * Xj^ch2   :791-793
       {PS_G, 8, PS_STMT + 7, 10, 8},            // eps23  = g^rho23 * A^ch2    :795-797
       {PS_STMT + 5, SS_ZERO, PS_STMT + 3, 9, 4},  // eps12' = Rj^0 * Bj^ch1    :775-777
       {PS_STMT + 8, SS_ZERO, PS_CG, 9, 5},      // eps13' = B^0 * (Ci/g)^ch1   :779-784
       {PS_STMT + 9, 6, PS_STMT + 0, 10, 9},     // eps21' = Yi^rho21 * Bi^ch2  :799-801
       {PS_STMT + 5, 7, PS_STMT + 3, 10, 10},    // eps22' = Rj^rho22 * Bj^ch2  :803-805
       {PS_STMT + 8, 8, PS_STMT + 6, 10, 11}},   // eps23' = B^rho23 * Ci^ch2   :807-809
  };
  return t[branch][j];
}

// ---- the one curve routine: r = a*P + b*Q ---------------------------------------------
// P == NULL means the generator (comb table); Q == NULL means the term is absent.
PA_HD void lincomb_op(jac &r, const jac *P, const sc &a, const jac *Q, const sc &b, const u32 *comb) {
  if (P == nullptr) {
    fixed_base_mul(r, a, comb);
    if (Q != nullptr) {
      jac t;
      var_base_mul(t, *Q, b);
      jac_add(r, r, t);
    }
  } else if (Q == nullptr) {
    var_base_mul(r, *P, a);
  } else {
    strauss<2>(r, *P, a, *Q, b);
  }
}

// ---- operand loading ---------------------------------------------------------------------
// wire point / scalar at a 16-byte aligned address
PA_HD void load_wire_point(aff &a, const unsigned char *p) {
#if defined(__CUDA_ARCH__)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 x0 = q[0], x1 = q[1], y0 = q[2], y1 = q[3];
  a.x.v[7] = __byte_perm(x0.x, 0, 0x0123); a.x.v[6] = __byte_perm(x0.y, 0, 0x0123);
  a.x.v[5] = __byte_perm(x0.z, 0, 0x0123); a.x.v[4] = __byte_perm(x0.w, 0, 0x0123);
  a.x.v[3] = __byte_perm(x1.x, 0, 0x0123); a.x.v[2] = __byte_perm(x1.y, 0, 0x0123);
  a.x.v[1] = __byte_perm(x1.z, 0, 0x0123); a.x.v[0] = __byte_perm(x1.w, 0, 0x0123);
  a.y.v[7] = __byte_perm(y0.x, 0, 0x0123); a.y.v[6] = __byte_perm(y0.y, 0, 0x0123);
  a.y.v[5] = __byte_perm(y0.z, 0, 0x0123); a.y.v[4] = __byte_perm(y0.w, 0, 0x0123);
  a.y.v[3] = __byte_perm(y1.x, 0, 0x0123); a.y.v[2] = __byte_perm(y1.y, 0, 0x0123);
  a.y.v[1] = __byte_perm(y1.z, 0, 0x0123); a.y.v[0] = __byte_perm(y1.w, 0, 0x0123);
#else
  aff_from_be64(a, p);
#endif
}
PA_HD void load_wire_scalar(sc &k, const unsigned char *p) {
#if defined(__CUDA_ARCH__)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 hi = q[0], lo = q[1];
  k.v[7] = __byte_perm(hi.x, 0, 0x0123); k.v[6] = __byte_perm(hi.y, 0, 0x0123);
  k.v[5] = __byte_perm(hi.z, 0, 0x0123); k.v[4] = __byte_perm(hi.w, 0, 0x0123);
  k.v[3] = __byte_perm(lo.x, 0, 0x0123); k.v[2] = __byte_perm(lo.y, 0, 0x0123);
  k.v[1] = __byte_perm(lo.z, 0, 0x0123); k.v[0] = __byte_perm(lo.w, 0, 0x0123);
  sc_reduce(k);
#else
  sc_from_be(k, p);
#endif
}

// resolve a point source code; returns false for the generator / an absent term
template <int KIND>
PA_HD bool load_src_point(jac &P, int code, const unsigned char *proof, const unsigned char *stmt) {
  typedef proof_kind<KIND> K;
  if (code == PS_G || code == PS_NONE) return false;
  aff a;
  if (code == PS_CG) {
    // tmp = g; invert; tmp = c + tmp          SEAL/bidder.cpp:178-180, 286-288, 559-561, 989-991
    aff g, ng;
    load_wire_point(a, stmt + 64 * K::CG_STMT);
    aff_set_generator(g);
    aff_neg(ng, g);
    jac_from_aff(P, a);
    jac_madd(P, P, ng);
    return true;
  }
  load_wire_point(a, code >= PS_STMT ? stmt + 64 * (code - PS_STMT) : proof + 64 * code);
  jac_from_aff(P, a);
  return true;
}

// verifier, check j of one proof: a*P + b*Q == eps[E] ?   (`ch1` from verify_derive)
template <int KIND>
PA_HD bool verify_check_one(int j, const unsigned char *proof, const unsigned char *stmt, const sc &ch1,
                            const u32 *comb) {
  typedef proof_kind<KIND> K;
  pa_op op = verify_op<KIND>(j);
  const unsigned char *scs = proof + K::NEPS * 64;
  jac P, Q, r;
  sc a, b;
  bool hp = load_src_point<KIND>(P, op.P, proof, stmt);
  bool hq = load_src_point<KIND>(Q, op.Q, proof, stmt);
  if (op.a == VS_CH1) a = ch1; else load_wire_scalar(a, scs + 32 * op.a);
  if (op.b == VS_CH1) b = ch1; else load_wire_scalar(b, scs + 32 * op.b);
  lincomb_op(r, hp ? &P : nullptr, a, hq ? &Q : nullptr, b, comb);
  aff e;
  load_wire_point(e, proof + 64 * op.E);
  return jac_eq_aff(r, e);  // EC_POINT_cmp, SEAL/bidder.cpp:131
}

// prover, operation j of one proof: r = eps[E] (Jacobian); returns E
template <int KIND>
PA_HD int prove_op_one(jac &r, int branch, int j, const unsigned char *stmt, const unsigned char *rnd,
                       const u32 *comb) {
  pa_op op = prove_op<KIND>(branch, j);
  jac P, Q;
  sc a, b;
  bool hp = load_src_point<KIND>(P, op.P, nullptr, stmt);
  bool hq = load_src_point<KIND>(Q, op.Q, nullptr, stmt);
  if (op.a == SS_ZERO) sc_set_zero(a); else load_wire_scalar(a, rnd + 32 * op.a);
  if (hq) load_wire_scalar(b, rnd + 32 * op.b); else sc_set_zero(b);
  lincomb_op(r, hp ? &P : nullptr, a, hq ? &Q : nullptr, b, comb);
  return op.E;
}

// ---- prover with witnesses ---------------------------------------------------------------
// A prover knows the discrete logarithm of most points of its own statement: X = g^x, R = g^r,
// A = g^alpha, B = g^beta, c = g^(alpha beta + bit), and its cryptogram is g^(r x) when it vetoes,
// Y^x when it does not.  Only Y (the combination of the OTHER bidders' keys) is foreign.  So every
// published point  a*P + b*Q  equals  kf*g + kv*Y  with two scalars computed mod n from the draws
// and the secrets: one fixed-base multiplication and at most one variable-base multiplication
// instead of up to two variable-base ones.  The points are the same group elements, hence the same
// bytes (the reference calls EC_POINT_mul on the bases as given, SEAL/bidder.cpp:171-202, 355-417,
// 657-814).  Extended secrets:
//   COM (alpha, beta)    S1 (x, alpha, r, beta)    S2 (xi, xj, alpha, ri, rj, beta)
// flags: veto_i / veto_j say which form the cryptograms Bi / Bj have, cbit is the committed bit.
struct wit_term {
  sc u;       // the point's g-part: u*g                 (valid if has_u)
  sc v;       // the point's foreign part: v*stmt[ybase]  (valid if ybase >= 0)
  bool has_u;
  int ybase;
};
template <int KIND> struct wit_kind;
template <> struct wit_kind<PA_COM> { static constexpr int NSECRET = 2; };
template <> struct wit_kind<PA_S1> { static constexpr int NSECRET = 4; };
template <> struct wit_kind<PA_S2> { static constexpr int NSECRET = 6; };

// dlog of the commitment point c = phi: alpha * beta + bit  (- 1 for c / g)
PA_HD void wit_commit_dlog(sc &u, const unsigned char *alpha, const unsigned char *beta, int cbit, bool over_g) {
  sc a, b, one;
  load_wire_scalar(a, alpha);
  load_wire_scalar(b, beta);
  sc_mul(u, a, b);
  sc_set_zero(one);
  one.v[0] = 1;
  if (cbit) sc_add(u, u, one);
  if (over_g) sc_sub(u, u, one);
}
PA_HD void wit_set_u(wit_term &w, const unsigned char *s) {
  load_wire_scalar(w.u, s);
  w.has_u = true;
  w.ybase = -1;
}
PA_HD void wit_set_prod(wit_term &w, const unsigned char *s0, const unsigned char *s1) {
  sc a, b;
  load_wire_scalar(a, s0);
  load_wire_scalar(b, s1);
  sc_mul(w.u, a, b);
  w.has_u = true;
  w.ybase = -1;
}
PA_HD void wit_set_y(wit_term &w, int ybase, const unsigned char *s) {  // s == NULL: the foreign point itself
  if (s) {
    load_wire_scalar(w.v, s);
  } else {
    sc_set_zero(w.v);
    w.v.v[0] = 1;
  }
  w.has_u = false;
  w.ybase = ybase;
}
template <int KIND>
PA_HD void wit_resolve(wit_term &w, int code, const unsigned char *s, int veto_i, int veto_j, int cbit);
template <>
PA_HD void wit_resolve<PA_COM>(wit_term &w, int code, const unsigned char *s, int, int, int cbit) {
  // stmt phi=0 A=1 B=2; secrets alpha, beta
  if (code == PS_CG || code == PS_STMT + 0) {
    wit_commit_dlog(w.u, s, s + 32, cbit, code == PS_CG);
    w.has_u = true;
    w.ybase = -1;
  } else {
    wit_set_u(w, code == PS_STMT + 1 ? s : s + 32);
  }
}
template <>
PA_HD void wit_resolve<PA_S1>(wit_term &w, int code, const unsigned char *s, int veto_i, int, int cbit) {
  // stmt b=0 X=1 Y=2 R=3 c=4 A=5 B=6; secrets x, alpha, r, beta
  switch (code) {
    case PS_STMT + 0: if (veto_i) wit_set_prod(w, s + 64, s); else wit_set_y(w, 2, s); break;  // b = R^x | Y^x
    case PS_STMT + 1: wit_set_u(w, s); break;
    case PS_STMT + 2: wit_set_y(w, 2, nullptr); break;
    case PS_STMT + 3: wit_set_u(w, s + 64); break;
    case PS_STMT + 5: wit_set_u(w, s + 32); break;
    case PS_STMT + 6: wit_set_u(w, s + 96); break;
    default:  // c, c/g
      wit_commit_dlog(w.u, s + 32, s + 96, cbit, code == PS_CG);
      w.has_u = true;
      w.ybase = -1;
  }
}
template <>
PA_HD void wit_resolve<PA_S2>(wit_term &w, int code, const unsigned char *s, int veto_i, int veto_j, int cbit) {
  // stmt Bi=0 Xi=1 Ri=2 Bj=3 Xj=4 Rj=5 Ci=6 A=7 B=8 Yi=9 Yj=10; secrets xi, xj, alpha, ri, rj, beta
  switch (code) {
    case PS_STMT + 0: if (veto_i) wit_set_prod(w, s + 96, s); else wit_set_y(w, 9, s); break;          // Bi = Ri^xi | Yi^xi
    case PS_STMT + 1: wit_set_u(w, s); break;
    case PS_STMT + 2: wit_set_u(w, s + 96); break;
    case PS_STMT + 3: if (veto_j) wit_set_prod(w, s + 128, s + 32); else wit_set_y(w, 10, s + 32); break;  // Bj = Rj^xj | Yj^xj
    case PS_STMT + 4: wit_set_u(w, s + 32); break;
    case PS_STMT + 5: wit_set_u(w, s + 128); break;
    case PS_STMT + 7: wit_set_u(w, s + 64); break;
    case PS_STMT + 8: wit_set_u(w, s + 160); break;
    case PS_STMT + 9: wit_set_y(w, 9, nullptr); break;
    case PS_STMT + 10: wit_set_y(w, 10, nullptr); break;
    default:  // Ci, Ci/g
      wit_commit_dlog(w.u, s + 64, s + 160, cbit, code == PS_CG);
      w.has_u = true;
      w.ybase = -1;
  }
}

// prover with witnesses, operation j of one proof: r = eps[E] = kf*g + kv*stmt[ybase]; returns E.
// The two terms of an operation never involve two different foreign points (prove_op tables).
template <int KIND>
PA_HD int prove_op_one_wit(jac &r, int branch, int j, const unsigned char *stmt, const unsigned char *rnd,
                           const unsigned char *secrets, int veto_i, int veto_j, int cbit, const u32 *comb) {
  pa_op op = prove_op<KIND>(branch, j);
  sc kf, kv, a, t;
  sc_set_zero(kf);
  sc_set_zero(kv);
  int ybase = -1;
  for (int side = 0; side < 2; ++side) {
    int code = side ? op.Q : op.P, scode = side ? op.b : op.a;
    if (code == PS_NONE) continue;
    if (scode == SS_ZERO) continue;  // the term vanishes (stage-2 branch 3)
    load_wire_scalar(a, rnd + 32 * scode);
    if (code == PS_G) {
      sc_add(kf, kf, a);
      continue;
    }
    wit_term w;
    wit_resolve<KIND>(w, code, secrets, veto_i, veto_j, cbit);
    if (w.has_u) {
      sc_mul(t, a, w.u);
      sc_add(kf, kf, t);
    } else {
      sc_mul(t, a, w.v);
      sc_add(kv, kv, t);
      ybase = w.ybase;
    }
  }
  fixed_base_mul(r, kf, comb);
  if (ybase >= 0 && !sc_is_zero(kv)) {
    aff y;
    jac Y, v;
    load_wire_point(y, stmt + 64 * ybase);
    jac_from_aff(Y, y);
    var_base_mul(v, Y, kv);
    jac_add(r, r, v);
  }
  return op.E;
}

// ---- Fiat-Shamir challenge of one proof ------------------------------------------------
template <int KIND>
PA_HD void proof_challenge(sc &ch, const unsigned char *proof, const unsigned char *stmt, u64 id) {
  typedef proof_kind<KIND> K;
  const unsigned char *pts[K::NHASH];
  for (int i = 0; i < K::NHASH; ++i) {
    int s = hash_src<KIND>(i);
    pts[i] = s >= PS_STMT ? stmt + 64 * (s - PS_STMT) : proof + 64 * s;
  }
  challenge_hash(ch, pts, K::NHASH, id);
}

// Is a 64-byte wire encoding acceptable?  Either the point at infinity (64 zero bytes) or canonical
// coordinates (both < p) satisfying y^2 = x^3 + 7: what EC_POINT_set_affine_coordinates enforces when a
// point enters libcrypto in the reference.  The a = 0 group formulas never use the constant 7, so an
// off-curve point would silently be processed on another curve, and a coordinate >= p would hash
// differently from its reduced twin: both are refused here, before any check runs.
PA_HD bool wire_point_valid(const unsigned char *p) {
  aff a, c;
  load_wire_point(a, p);
  if (aff_is_inf(a)) return true;
  fe_canon(c.x, a.x);
  fe_canon(c.y, a.y);
  bool canonical = true;
#pragma unroll
  for (int k = 0; k < 8; ++k) canonical &= c.x.v[k] == a.x.v[k] && c.y.v[k] == a.y.v[k];
  return canonical && aff_on_curve(a);
}
// every point a verifier receives: the eps of the proof and the statement
template <int KIND>
PA_HD bool proof_points_valid(const unsigned char *proof, const unsigned char *stmt) {
  typedef proof_kind<KIND> K;
  bool ok = true;
  for (int i = 0; i < K::NEPS; ++i) ok &= wire_point_valid(proof + 64 * i);
  for (int i = 0; i < K::NSTMT; ++i) ok &= wire_point_valid(stmt + 64 * i);
  return ok;
}

// verifier: the challenge share that is not published.  ch1 = ch - ch2 (- ch3)
// (SEAL/bidder.cpp:253, 485, 934-935); for POK it is h itself (:125).
template <int KIND>
PA_HD void verify_derive(sc &ch1, const unsigned char *proof, const unsigned char *stmt, u64 id) {
  typedef proof_kind<KIND> K;
  sc ch;
  proof_challenge<KIND>(ch, proof, stmt, id);
  if (KIND == PA_POK) {
    ch1 = ch;
    return;
  }
  const unsigned char *scs = proof + K::NEPS * 64;
  sc t;
  sc_from_be(t, scs + 32 * (K::NSC - (KIND == PA_S2 ? 2 : 1)));  // ch2
  sc_sub(ch1, ch, t);
  if (KIND == PA_S2) {
    sc_from_be(t, scs + 32 * (K::NSC - 1));  // ch3
    sc_sub(ch1, ch1, t);
  }
}

// prover: responses written behind the eps points of the proof record.
//   secrets: POK (x)   COM (alpha)   S1 (x, alpha)   S2 (xi, xj, alpha)
template <int KIND>
PA_HD void prove_respond(unsigned char *proof, const unsigned char *stmt, u64 id, const unsigned char *secrets,
                         const unsigned char *rnd, int branch) {
  typedef proof_kind<KIND> K;
  unsigned char *out = proof + K::NEPS * 64;
  sc ch, t, u;
  proof_challenge<KIND>(ch, proof, stmt, id);
  sc r[K::NRND];
  for (int i = 0; i < K::NRND; ++i) sc_from_be(r[i], rnd + 32 * i);
  sc s[K::NSECRET];
  for (int i = 0; i < K::NSECRET; ++i) sc_from_be(s[i], secrets + 32 * i);
  if (KIND == PA_POK) {
    // rho = v - h*x                                                 SEAL/bidder.cpp:102-103
    sc_mul(t, ch, s[0]);
    sc_sub(u, r[0], t);
    sc_to_be(out, u);
  } else if (KIND == PA_COM) {
    // real challenge = ch - simulated; rho_real = r1 - alpha*ch_real     :208-217
    sc chr, rho;
    sc_sub(chr, ch, r[1]);
    sc_mul(t, chr, s[0]);
    sc_sub(rho, r[0], t);
    if (branch == 0) {  // publish rho1 = real, rho2 = drawn, ch2 = drawn
      sc_to_be(out, rho);
      sc_to_be(out + 32, r[2]);
      sc_to_be(out + 64, r[1]);
    } else {  // publish rho1 = drawn, rho2 = real, ch2 = real challenge
      sc_to_be(out, r[2]);
      sc_to_be(out + 32, rho);
      sc_to_be(out + 64, chr);
    }
  } else if (KIND == PA_S1) {
    // real challenge = ch - drawn; rho_x = r11 - x*ch_real; rho_alpha = r12 - alpha*ch_real   :424-436
    sc chr, rx, ra;
    sc_sub(chr, ch, r[4]);
    sc_mul(t, chr, s[0]);
    sc_sub(rx, r[0], t);
    sc_mul(t, chr, s[1]);
    sc_sub(ra, r[1], t);
    if (branch == 0) {  // rho11 rho12 real; rho21 rho22 ch2 drawn
      sc_to_be(out, rx);
      sc_to_be(out + 32, ra);
      sc_to_be(out + 64, r[2]);
      sc_to_be(out + 96, r[3]);
      sc_to_be(out + 128, r[4]);
    } else {  // rho11 rho12 drawn; rho21 rho22 real; ch2 = real challenge
      sc_to_be(out, r[2]);
      sc_to_be(out + 32, r[3]);
      sc_to_be(out + 64, rx);
      sc_to_be(out + 96, ra);
      sc_to_be(out + 128, chr);
    }
  } else {
    // real challenge = ch - the two drawn ones                               :824-862
    sc chr, r1, r2, r3, zero;
    sc_set_zero(zero);
    sc_sub(chr, ch, r[9]);
    sc_sub(chr, chr, r[10]);
    sc_mul(t, s[0], chr);
    sc_sub(r1, r[0], t);  // r11 - xi*ch
    sc_mul(t, s[1], chr);
    sc_sub(r2, r[1], t);  // r12 - xj*ch
    sc_mul(t, s[2], chr);
    sc_sub(r3, r[2], t);  // r13 - alpha*ch
    // out: rho11 rho12 rho13 rho21 rho22 rho23 rho31 rho32 ch2 ch3
    if (branch == 0) {
      sc_to_be(out, r1); sc_to_be(out + 32, r2); sc_to_be(out + 64, r3);
      sc_to_be(out + 96, r[3]); sc_to_be(out + 128, r[4]); sc_to_be(out + 160, r[5]);
      sc_to_be(out + 192, r[6]); sc_to_be(out + 224, r[7]);
      sc_to_be(out + 256, r[9]); sc_to_be(out + 288, r[10]);
    } else if (branch == 1) {
      sc_to_be(out, r[3]); sc_to_be(out + 32, r[4]); sc_to_be(out + 64, r[5]);
      sc_to_be(out + 96, r1); sc_to_be(out + 128, r2); sc_to_be(out + 160, r3);
      sc_to_be(out + 192, r[6]); sc_to_be(out + 224, r[7]);
      sc_to_be(out + 256, chr); sc_to_be(out + 288, r[10]);
    } else {
      sc_to_be(out, zero); sc_to_be(out + 32, zero); sc_to_be(out + 64, zero);
      sc_to_be(out + 96, r[6]); sc_to_be(out + 128, r[7]); sc_to_be(out + 160, r[8]);
      sc_to_be(out + 192, r1); sc_to_be(out + 224, r2);
      sc_to_be(out + 256, r[10]); sc_to_be(out + 288, chr);
    }
  }
}
