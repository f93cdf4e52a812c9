// Shot sampling walk and amplitude read-out (MeasureAll / GetElementOfVector of the reference,
// Simulator.cpp:148-224, 415-431) as per-thread device functions.
#ifndef JKQ_SAMPLE_CUH
#define JKQ_SAMPLE_CUH

#include "jkq_engine.cuh"

// Philox-4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).
// key = seed, counter = (shot_lo, shot_hi, block, 0): the stream of a shot depends only on the seed
// and the GLOBAL shot index, so counts are identical however the shots are split over GPUs.
JD void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// One shot: n dependent 16-byte loads from the root.  bitpos[L] = output bit of level L (>= 64:
// not measured).  Word j of Philox block b drives step 4b+j counted from the root; the branch to
// child 0 is taken when u < q0 with u = (word + 0.5) / 2^32 (the reference draws rand()/RAND_MAX,
// Simulator.cpp:188-196).
#ifdef JKQ_EMU
JD JSNode ld_snode(const JSNode *p) { return *p; }
#else
JD JSNode ld_snode(const JSNode *p) {
  uint4 v = __ldg((const uint4 *)p);
  JSNode s;
  s.c0 = v.x; s.c1 = v.y; s.q0 = __hiloint2double((int)v.w, (int)v.z);
  return s;
}
#endif

template <bool LAZY = false>
JD uint64_t sample_walk(const JSNode *sn, uint32_t root, int n, const uint8_t *bitpos, uint64_t seed, uint64_t shot) {
  uint32_t cur = root, c[4] = {0, 0, 0, 0};
  uint64_t key = 0;
  int have = -1;  // Philox block held in c[]
  for (int step = 0; step < n; step++) {
    const JSNode s = ld_snode(sn + cur);
    // A branch that is certain needs no random number: u lies strictly between 0 and 1, so u < q0 is decided by q0
    // alone when q0 >= 1 (child 0) or q0 <= 0 (child 1).  The stream is counter-based -- word j of block b belongs to
    // step 4b + j whether or not the steps before it drew theirs -- so a block is only computed when one of its four
    // steps has a real choice to make: same outcomes bit for bit, one Philox block per GHZ shot instead of n / 4.
    // LAZY is chosen per diagram (jkq_b200.cu sample_impl: fewer than half of its nodes have a real choice); a diagram
    // like QFT's, where every node has one, draws each block when it starts, with the node load in flight.
    uint32_t bit;
    if (LAZY && s.q0 >= 1.0) bit = 0;
    else if (LAZY && !(s.q0 > 0.0)) bit = 1;
    else {
      if (LAZY ? have != (step >> 2) : (step & 3) == 0) {
        have = step >> 2;
        c[0] = (uint32_t)shot; c[1] = (uint32_t)(shot >> 32); c[2] = (uint32_t)have; c[3] = 0;
        philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      }
      const uint32_t word = (step & 3) == 0 ? c[0] : (step & 3) == 1 ? c[1] : (step & 3) == 2 ? c[2] : c[3];
      const double u = JKQ_DMUL(JKQ_DADD((double)word, 0.5), 1.0 / 4294967296.0);
      bit = !(u < s.q0);
    }
    const uint32_t bp = bitpos[n - 1 - step];
    if (bit && bp < 64) key |= 1ull << bp;
    cur = bit ? s.c1 : s.c0;
  }
  return key;
}

// Amplitude of basis state `index` (bit q = qubit q; qubit q sits on level n-1-q): product of the
// edge weights along one root-to-terminal path, FP64 on the canonical values.
template <int MODE> JD void amp_walk(const Ctx &c, uint64_t index, double *re, double *im) {
  const JEng *E = c.E;
  const int n = (int)E->n;
  uint32_t cur = E->root;
  double ar = jre<MODE>(c, E->rootw), ai = jim<MODE>(c, E->rootw);
  for (int L = n - 1; L >= 0; L--) {
    JNode nd = ld_s<MODE>(E->nodes + cur);
    uint32_t bit = (uint32_t)(index >> (n - 1 - L)) & 1u;
    uint64_t w = bit ? nd_w1(nd) : nd_w0(nd);
    if (w == 0) { *re = 0.0; *im = 0.0; return; }
    if (w != JW_ONE) {
      double br = jre<MODE>(c, w), bi = jim<MODE>(c, w);
      double tr = JKQ_DSUB(JKQ_DMUL(ar, br), JKQ_DMUL(ai, bi));
      ai = JKQ_DADD(JKQ_DMUL(ar, bi), JKQ_DMUL(ai, br));
      ar = tr;
    }
    cur = bit ? nd.c1 : nd_c0(nd);
  }
  *re = ar; *im = ai;
}

#endif  // JKQ_SAMPLE_CUH
