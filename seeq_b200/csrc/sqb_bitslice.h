// sqb_bitslice.h -- the line-bit-sliced Levenshtein automaton (core of K2).
//
// One 32-bit word holds ONE BIT OF STATE FOR 32 DIFFERENT LINES.  The automaton
// of a line is the column of vertical deltas of the search-type edit-distance
// matrix (Myers 1999): row j of the pattern contributes two bit-planes, Pv[j]
// (delta +1) and Mv[j] (delta -1).  Feeding one text column (one byte of each
// of the 32 lines) walks the rows bottom-up, carrying the horizontal delta:
//
//    Xh = Eq | Mh_in          Ph_out = Mv | ~(Xh | Pv)     Mh_out = Pv & Xh
//    Xv = Eq | Mv             Pv'    = Mh_in | ~(Xv | Ph_in)   Mv' = Ph_in & Xv
//
// = 6 three-input logic ops (LOP3) per row for 32 text bytes.  Eq is the class
// mask of the row: which of the 32 lines carry a byte matching pattern row j.
// The distance D = value of the last row is kept as a bit-sliced counter and
// updated with the last row's horizontal delta; the report state machine of
// the reference (libseeq.c:278-330: streak / match flag / best distance) is
// evaluated with bitwise logic on those planes, for 32 lines at once.
//
// The R >= m rows of a kernel instance are padded AT THE BOTTOM with R-m rows
// whose Eq is all-ones and whose deltas start at 0: they stay 0 and hand a 0
// horizontal delta to the first pattern row, which is the free-start boundary
// of the search recurrence.  The last row is therefore always row R-1.
//
// This header compiles for host and device: tests/host_bitslice.cpp runs the
// same code on the CPU against the oracle.
#ifndef SQB_BITSLICE_H_
#define SQB_BITSLICE_H_

#include <stdint.h>

#ifdef __CUDACC__
#define SQB_BS_HD __host__ __device__ __forceinline__
#else
#define SQB_BS_HD static inline
#endif

namespace sqb {

// class codes of the tokenizer (K1), three bit-planes p2 p1 p0
//   0 A   1 C   2 G   3 T/U   4 N   5 STOP   6 SKIP   7 NULL
// Eq slots of one text column
constexpr int kBsMaxCustom = 6;          // bracket classes other than a single base, N and "any" a pattern may use
enum BsSlot { BS_A = 0, BS_C, BS_G, BS_T, BS_N, BS_ANY, BS_CUSTOM0, BS_CUSTOM1, BS_ONES = BS_CUSTOM0 + kBsMaxCustom, BS_SLOTS };

enum BsMode { BS_FIRST = 0, BS_BEST = 1, BS_ALL = 2 };

constexpr int kBsMaxRows = 128;         // rows of the longest pattern: 4 parts of 32 rows
constexpr int kBsBestBits = 4;          // best distance planes: tau + 1 <= 15

// Patterns of more than 32 positions are cut into G PARTS of R rows each, part p
// = rows [p*R, (p+1)*R).  Rows only depend on the rows below them (the horizontal
// delta flows upwards), so the parts form a pipeline: part p works on text column
// c - p while part p-1 works on column c - p + 1 and hands over the horizontal
// delta (ph, mh) of its top row.  In the CUDA kernel the parts of a group of 32
// lines sit in G lanes of one warp and the hand-over is one pair of shuffles.
struct BsPattern {
   int32_t  m, tau;
   int32_t  rows;                       // R: rows per part of the kernel instance
   int32_t  parts;                      // G: 1, 2 or 4;  R * G >= m
   int32_t  ncustom;                    // custom classes in use (0..kBsMaxCustom)
   uint8_t  slot[kBsMaxRows];           // Eq slot of every row (pad rows: BS_ONES)
   uint32_t slot_off[kBsMaxRows];       // the same as byte offset slot * 32 lanes * 4 (kernel smem layout)
   uint32_t custom[kBsMaxCustom][5];    // custom classes: all-ones where A,C,G,T,N belongs to the class
   uint32_t tau_plane[8];               // bit k of tau, replicated: 0 or ~0
   uint32_t m_plane[8];                 // bit k of m
   uint32_t best_plane[kBsBestBits];    // bit k of tau + 1
};

// planes of the distance counter: it runs from 0 to the number of rows
constexpr int bs_score_bits(int rows) { return rows < 16 ? 4 : (rows < 32 ? 5 : (rows < 64 ? 6 : (rows < 128 ? 7 : 8))); }

template <int R, int G = 1> struct BsState {
   static constexpr int B = bs_score_bits(R * G);
   uint32_t pv[R], mv[R];               // the rows of ONE part
   uint32_t s[B];                       // distance of the last row, bit-sliced
   uint32_t bd[kBsBestBits];            // best distance so far (BS_BEST)
   uint32_t alive;                      // lines still being scanned
   uint32_t flag;                       // the reference's `match` suppress flag
   uint32_t hit;                        // lines with at least one (reported) event
   uint32_t stopped;                    // lines that ran into a STOP byte
};

// part: which part of the pattern this state holds (0 .. G-1); `valid` = lines in
// use.  Only the LAST part tracks the distance and reports, the others keep
// alive == 0 and never produce an event.
template <int R, int G>
SQB_BS_HD void bs_reset(BsState<R, G> &st, const BsPattern &p, uint32_t valid, int part = 0)
{
   const int pad = R * G - p.m;                   // wildcard rows at the bottom (all inside part 0)
#pragma unroll
   for (int j = 0; j < R; j++) {
      st.pv[j] = (part * R + j) >= pad ? ~0u : 0u;
      st.mv[j] = 0u;
   }
#pragma unroll
   for (int k = 0; k < BsState<R, G>::B; k++) st.s[k] = p.m_plane[k];
#pragma unroll
   for (int k = 0; k < kBsBestBits; k++) st.bd[k] = p.best_plane[k];
   st.alive = part == G - 1 ? valid : 0u;
   st.flag = 0u;
   st.hit = 0u;
   st.stopped = 0u;
}

// Eq slots and the class masks of one column from the three code planes
SQB_BS_HD void bs_classes(uint32_t p0, uint32_t p1, uint32_t p2, const BsPattern &p, uint32_t *slots,
                          uint32_t &anybase, uint32_t &stop, uint32_t &skip)
{
   const uint32_t a = ~p2 & ~p1 & ~p0, c = ~p2 & ~p1 & p0, g = ~p2 & p1 & ~p0, t = ~p2 & p1 & p0;
   const uint32_t n = p2 & ~p1 & ~p0;
   slots[BS_A] = a;
   slots[BS_C] = c;
   slots[BS_G] = g;
   slots[BS_T] = t;
   slots[BS_N] = n;
   anybase = ~p2 | n;
   slots[BS_ANY] = anybase;
   for (int k = 0; k < kBsMaxCustom; k++)
      slots[BS_CUSTOM0 + k] = (a & p.custom[k][0]) | (c & p.custom[k][1]) | (g & p.custom[k][2]) | (t & p.custom[k][3]) |
                              (n & p.custom[k][4]);
   slots[BS_ONES] = ~0u;
   stop = p2 & ~p1 & p0;
   skip = p2 & p1 & ~p0;
}

// The rows of one part for one text column.  eq(j) returns the Eq mask of row j
// of this part.  ph / mh: on entry the horizontal delta handed over by the part
// below (0 for part 0), on return the delta of this part's top row.
//
// (Skipping the wildcard rows of part 0 with a computed jump into the unrolled
// rows was tried: the basic-block boundaries stop the scheduler from hoisting
// the Eq loads, k2_bitslice<12> went from 447 to 558 us on cfg2.)
//
// A NULL column (class 7: no base, no stop, no skip) leaves a state that is
// still at its reset value unchanged and hands 0 upwards: a mismatch against
// the initial column [0, 1, .., m] reproduces it.  The kernel feeds NULL columns
// to the upper parts while the pipeline fills.
template <int R, int G, bool SKIP, class EqOf>
SQB_BS_HD void bs_rows(BsState<R, G> &st, const EqOf &eq, uint32_t skipc, uint32_t &ph, uint32_t &mh)
{
   static_assert(R <= 32, "at most 32 rows per part");
   (void)skipc;
#pragma unroll
   for (int j = 0; j < R; j++) {
      const uint32_t e = eq(j);
      const uint32_t pv = st.pv[j], mv = st.mv[j];
      const uint32_t xh = e | mh;
      const uint32_t xv = e | mv;
      const uint32_t ph_out = mv | ~(xh | pv);
      const uint32_t mh_out = pv & xh;
      uint32_t pv_new = mh | ~(xv | ph);
      uint32_t mv_new = ph & xv;
      if (SKIP) {                                 // an ignored byte leaves the automaton untouched
         pv_new = (pv & skipc) | (pv_new & ~skipc);
         mv_new = (mv & skipc) | (mv_new & ~skipc);
      }
      st.pv[j] = pv_new;
      st.mv[j] = mv_new;
      ph = ph_out;
      mh = mh_out;
   }
}

// The report state machine of the LAST part for one text column.  ph / mh = the
// horizontal delta of the top row (the last pattern position).  Returns the
// event mask: bit r set <=> line r reports a match ending at this column with
// distance `streak` = the value held by st.s BEFORE the call (returned in
// streak[]).  `quiet` = lines whose events update the suppress flag but are not
// reported (the warm-up of a line segment); pass 0 otherwise.
template <int R, int G, int MODE>
SQB_BS_HD uint32_t bs_report(BsState<R, G> &st, const BsPattern &p, uint32_t ph, uint32_t mh, uint32_t anybase,
                             uint32_t stopc, uint32_t *streak, uint32_t quiet = 0u)
{
   constexpr int B = BsState<R, G>::B;
   const uint32_t base = anybase & st.alive;      // lines that feed a base to the automaton
   const uint32_t stop = stopc & st.alive;        // lines that end here

   // ---- streak = distance before this column: <= tau ?  == 0 ? ---------------
   uint32_t gt = 0u, nz = 0u;
#pragma unroll
   for (int k = 0; k < B; k++) {
      const uint32_t s = st.s[k], t = p.tau_plane[k];
      gt = (s & ~t) | (~(s ^ t) & gt);
      nz |= s;
      streak[k] = s;
   }
   const uint32_t le = ~gt, zero = ~nz;

   // ---- report state machine (libseeq.c:267-330) ------------------------------
   const uint32_t rise = (ph & base) | stop;      // the capped distance goes up (a terminal byte is tau+1)
   const uint32_t fall = mh & base;
   const uint32_t active = base | stop;
   st.flag &= rise | ~active;                     // :278  any non-rise clears the flag
   uint32_t evt = active & le & (zero | rise) & ~st.flag;      // :286-288
   st.flag |= evt;                                // also during a warm-up
   evt &= ~quiet;
   if (MODE == BS_BEST) {
      uint32_t lt = 0u;                           // streak < best distance (4 low bits decide: streak <= tau)
#pragma unroll
      for (int k = 0; k < kBsBestBits; k++) {
         const uint32_t s = k < B ? st.s[k] : 0u, d = st.bd[k];
         lt = (~s & d) | (~(s ^ d) & lt);
      }
      // libseeq.c:286-288 leaves the flag alone when `streak < best_d` fails; setting it
      // anyway is unobservable: until the flag is cleared the distance only rises, so
      // no later candidate of the run can beat best_d either (BEST = strict running
      // minimum over the SQ_ALL events)
      evt &= lt;
#pragma unroll
      for (int k = 0; k < kBsBestBits; k++) {
         const uint32_t s = k < B ? st.s[k] : 0u;
         st.bd[k] = (evt & s) | (~evt & st.bd[k]);
      }
   }
   st.hit |= evt;
   st.stopped |= stop;
   st.alive &= ~stop;
   if (MODE == BS_FIRST) st.alive &= ~evt;        // :330  the first match ends the scan of the line

   // ---- distance += rise - fall (ripple over the bit-planes) -----------------
   const uint32_t inc = ph & base;
   uint32_t carry = inc | fall;
#pragma unroll
   for (int k = 0; k < B; k++) {
      const uint32_t s = st.s[k];
      st.s[k] = s ^ carry;
      carry &= s ^ fall;
   }
   return evt;
}

// One text column of a single-part automaton (G == 1): rows + report.
template <int R, int MODE, bool SKIP, class EqOf>
SQB_BS_HD uint32_t bs_step(BsState<R, 1> &st, const BsPattern &p, const EqOf &eq, uint32_t anybase, uint32_t stopc,
                           uint32_t skipc, uint32_t *streak, uint32_t quiet = 0u)
{
   uint32_t ph = 0u, mh = 0u;
   bs_rows<R, 1, SKIP>(st, eq, skipc, ph, mh);
   return bs_report<R, 1, MODE>(st, p, ph, mh, anybase, stopc, streak, quiet);
}

// ---------------------------------------------------------------------------
// tau <= 2: the same automaton as bit-sliced NFA levels (Wu-Manber 1992) instead
// of Myers' delta encoding.  r[d][j] = "a suffix of the text read so far matches
// the first j+1 rows with at most d errors" (d = 0 .. tau), one bit per line:
//
//    r0[j]' = r0[j-1] & Eq[j]
//    rd[j]' = (rd[j-1] & Eq[j]) | r(d-1)[j-1] | r(d-1)[j] | r(d-1)[j-1]'
//              match              substitution   insertion    deletion
//
// with row -1 all ones (free start).  That is 2 tau + 1 LOP3 per row instead of
// 6, and the capped distance of the last row comes for free in unary: "D <= d" is
// r[d][R-1], so the ripple counter and the comparisons of the Myers version
// disappear as well (cfg2, tau = 1: 59 instead of 115 logic ops per column).
// D(i) = min{d : r[d][last]} or tau + 1 is exactly the capped NW value of the
// reference (libseeq.c:779-786); the report state machine is the one of
// bs_report, written for unary distances.
// ---------------------------------------------------------------------------
template <int R, int T> struct BsWmState {     // T = tau + 1 levels (2 or 3)
   uint32_t r[T][R];
   uint32_t lp[T];                             // streak: distance before this column <= d
   uint32_t bd[T];                             // best distance so far <= d (BS_BEST)
   uint32_t alive, flag, hit, stopped;
};

template <int R, int T> SQB_BS_HD void bs_wm_reset(BsWmState<R, T> &st, const BsPattern &p, uint32_t valid)
{
   const int pad = R - p.m;                     // wildcard rows at the bottom: always reachable
#pragma unroll
   for (int d = 0; d < T; d++) {
#pragma unroll
      for (int j = 0; j < R; j++) st.r[d][j] = (j < pad || j - pad + 1 <= d) ? ~0u : 0u;    // j+1-pad deletions
      st.lp[d] = 0u;                            // the distance starts at tau + 1 (libseeq.c:240-244)
      st.bd[d] = 0u;
   }
   st.alive = valid;
   st.flag = 0u;
   st.hit = 0u;
   st.stopped = 0u;
}

// One text column.  Returns the event mask; streak[d] = "distance before this column
// <= d" (unary: the distance of line r is the number of d with bit r of streak[d] clear).
template <int R, int T, int MODE, bool SKIP, class EqOf>
SQB_BS_HD uint32_t bs_wm_step(BsWmState<R, T> &st, const EqOf &eq, uint32_t anybase, uint32_t stopc, uint32_t skipc,
                              uint32_t *streak, uint32_t quiet = 0u)
{
   (void)skipc;
   uint32_t oldp[T], newp[T];                   // row j-1 before / after this column
#pragma unroll
   for (int d = 0; d < T; d++) oldp[d] = newp[d] = ~0u;
#pragma unroll
   for (int j = 0; j < R; j++) {
      const uint32_t e = eq(j);
      uint32_t old[T], nw[T];
#pragma unroll
      for (int d = 0; d < T; d++) old[d] = st.r[d][j];
      nw[0] = oldp[0] & e;
#pragma unroll
      for (int d = 1; d < T; d++) nw[d] = (oldp[d] & e) | oldp[d - 1] | old[d - 1] | newp[d - 1];
#pragma unroll
      for (int d = 0; d < T; d++) {
         if (SKIP) nw[d] = (old[d] & skipc) | (nw[d] & ~skipc);      // an ignored byte leaves the automaton untouched
         st.r[d][j] = nw[d];
         oldp[d] = old[d];
         newp[d] = nw[d];
      }
   }
   const uint32_t base = anybase & st.alive;    // lines that feed a base to the automaton
   const uint32_t stop = stopc & st.alive;      // lines that end here (their distance becomes tau + 1)
   const uint32_t active = base | stop;
   uint32_t rise = 0u;                          // streak < new distance
#pragma unroll
   for (int d = 0; d < T; d++) {
      streak[d] = st.lp[d];
      rise |= st.lp[d] & ~(newp[d] & base);
   }
   st.flag &= rise | ~active;                   // libseeq.c:278  any non-rise clears the flag
   uint32_t evt = active & st.lp[T - 1] & (st.lp[0] | rise) & ~st.flag;     // :286-288
   st.flag |= evt;                              // also during a warm-up
   evt &= ~quiet;
   if (MODE == BS_BEST) {
      uint32_t lt = 0u;                         // streak < best distance
#pragma unroll
      for (int d = 0; d < T; d++) lt |= st.lp[d] & ~st.bd[d];
      evt &= lt;
#pragma unroll
      for (int d = 0; d < T; d++) st.bd[d] |= evt & st.lp[d];
   }
   st.hit |= evt;
   st.stopped |= stop;
   st.alive &= ~stop;
   if (MODE == BS_FIRST) st.alive &= ~evt;      // :330
#pragma unroll
   for (int d = 0; d < T; d++) st.lp[d] = (base & newp[d]) | (~base & st.lp[d]);    // ignored bytes keep the streak
   return evt;
}

// distance of line r from unary planes
template <int T> SQB_BS_HD uint32_t bs_value_unary(const uint32_t *planes, int r)
{
   uint32_t v = 0;
#pragma unroll
   for (int d = 0; d < T; d++) v += ((planes[d] >> r) & 1u) ^ 1u;
   return v;
}

// distance of line r from bit-planes
template <int B> SQB_BS_HD uint32_t bs_value(const uint32_t *planes, int r)
{
   uint32_t v = 0;
#pragma unroll
   for (int k = 0; k < B; k++) v |= ((planes[k] >> r) & 1u) << k;
   return v;
}

}  // namespace sqb
#endif
