// tests/host_bitslice.cpp -- TEST HARNESS: runs the bit-sliced automaton of
// seeq_b200/csrc/sqb_bitslice.h on the CPU, lane by lane, exactly as the CUDA
// kernel drives it (32 lines per word, class nibbles, NULL padding in front of
// a line start that is not 16-byte aligned), so that its events can be compared
// with the oracle without a GPU.  Built by tests/test_bitslice_host.py (g++).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "sqb_tables.h"

using namespace sqb;

struct Ev { uint64_t line, end, dist; };

// One group of <= 32 lines through a G-part automaton, driven like the kernel:
// part p works on column t - p at iteration t (NULL columns while the pipeline
// fills) and receives the horizontal delta part p-1 produced one iteration ago.
template <int R, int G, int MODE, bool SKIP>
static void run_group(const std::vector<uint8_t> &cls, size_t n, const std::vector<size_t> &begin, size_t l0,
                      size_t l1, const BsPattern &p, std::vector<Ev> &out)
{
   BsState<R, G> st[G];
   const int nl = (int)(l1 - l0);
   for (int g = 0; g < G; g++) bs_reset(st[g], p, nl == 32 ? ~0u : ((1u << nl) - 1u), g);
   uint32_t out_ph[G] = {0}, out_mh[G] = {0};
   uint32_t streak[8];
   for (long t = 0; st[G - 1].alive; t++) {
      uint32_t in_ph[G], in_mh[G];
      for (int g = 0; g < G; g++) {
         in_ph[g] = g ? out_ph[g - 1] : 0u;
         in_mh[g] = g ? out_mh[g - 1] : 0u;
      }
      for (int g = 0; g < G; g++) {
         const long col = t - g;
         uint32_t p0 = 0, p1 = 0, p2 = 0;
         for (int r = 0; r < nl; r++) {
            const size_t b = begin[l0 + r], a = b & ~(size_t)15;
            const long pos = (long)a + col;
            uint8_t c = (col < 0 || (size_t)pos < b) ? kClsNull : ((size_t)pos >= n ? kClsStop : (cls[pos] & 7));
            p0 |= (uint32_t)(c & 1) << r;
            p1 |= (uint32_t)((c >> 1) & 1) << r;
            p2 |= (uint32_t)((c >> 2) & 1) << r;
         }
         for (int r = nl; r < 32; r++) { p0 |= 1u << r; p2 |= 1u << r; }     // STOP
         uint32_t slots[BS_SLOTS], anybase, stop, skip;
         bs_classes(p0, p1, p2, p, slots, anybase, stop, skip);
         auto eq = [&](int j) { return slots[p.slot[g * R + j]]; };
         uint32_t ph = in_ph[g], mh = in_mh[g];
         bs_rows<R, G, SKIP>(st[g], eq, skip, ph, mh, g == 0 ? R * G - p.m : 0);
         out_ph[g] = ph;
         out_mh[g] = mh;
         if (g == G - 1) {
            const uint32_t evt = bs_report<R, G, MODE>(st[g], p, ph, mh, anybase, stop, streak);
            for (int r = 0; r < nl; r++)
               if ((evt >> r) & 1u) {
                  const size_t b = begin[l0 + r];
                  out.push_back(Ev{(uint64_t)(l0 + r), (uint64_t)(col - (long)(b & 15)),
                                   bs_value<BsState<R, G>::B>(streak, r)});
               }
         }
      }
   }
}

template <int R, int G>
static void run_rows(int mode, bool skip, const std::vector<uint8_t> &cls, size_t n, const std::vector<size_t> &begin,
                     size_t l0, size_t l1, const BsPattern &p, std::vector<Ev> &out)
{
#define CASE(M)                                                                       \
   if (mode == M) {                                                                   \
      if (skip) run_group<R, G, M, true>(cls, n, begin, l0, l1, p, out);              \
      else run_group<R, G, M, false>(cls, n, begin, l0, l1, p, out);                  \
   }
   CASE(BS_FIRST) CASE(BS_BEST) CASE(BS_ALL)
#undef CASE
}

// Returns the number of (line, end, dist) events written to out (3 x u64 each,
// line 1-based, ordered by line then end), -1 if the pattern is not supported by
// the bit-sliced path, -2 if cap is too small.
extern "C" long bs_host_scan(const char *buf, size_t n, const unsigned char *keys, int m, int tau, int options,
                             uint64_t *out, long cap)
{
   BsPattern p;
   if (!build_bs_pattern(keys, m, tau, &p)) return -1;
   ClassTable ct;
   build_class_table(options, &ct);
   std::vector<uint8_t> cls(n);
   for (size_t i = 0; i < n; i++) cls[i] = ct.code[(unsigned char)buf[i]];
   std::vector<size_t> begin;
   if (n > 0) begin.push_back(0);
   for (size_t i = 0; i + 1 < n; i++) if (buf[i] == '\n') begin.push_back(i + 1);
   const int match = options & OPT_MATCH;
   const int mode = match == OPT_ALL ? BS_ALL : (match == OPT_BEST ? BS_BEST : BS_FIRST);
   const bool skip = (options & OPT_NONDNA) == OPT_IGNORE;
   std::vector<Ev> ev;
   for (size_t l0 = 0; l0 < begin.size(); l0 += 32) {
      const size_t l1 = std::min(begin.size(), l0 + 32);
#define SHAPE(R, G) if (p.rows == R && p.parts == G) run_rows<R, G>(mode, skip, cls, n, begin, l0, l1, p, ev);
      SHAPE(8, 1) SHAPE(12, 1) SHAPE(16, 1) SHAPE(24, 1) SHAPE(32, 1)
      SHAPE(20, 2) SHAPE(24, 2) SHAPE(32, 2)
      SHAPE(20, 4) SHAPE(24, 4) SHAPE(28, 4) SHAPE(32, 4)
#undef SHAPE
   }
   std::stable_sort(ev.begin(), ev.end(), [](const Ev &a, const Ev &b) { return a.line < b.line; });
   long k = 0;
   for (size_t i = 0; i < ev.size(); i++) {
      if (mode == BS_BEST && i + 1 < ev.size() && ev[i + 1].line == ev[i].line) continue;   // the last improvement wins
      if (k >= cap) return -2;
      out[3 * k + 0] = ev[i].line + 1;
      out[3 * k + 1] = ev[i].end;
      out[3 * k + 2] = ev[i].dist;
      k++;
   }
   return k;
}
