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
// g_skew generalises the pipeline: part p runs p * g_skew columns behind and takes
// what part p-1 produced g_skew iterations ago.  The kernels use 1; a skew of 2
// takes the hand-over off the critical path between consecutive columns of a lane
// (the parts of two consecutive columns could then be interleaved).
static int g_skew = 1;
extern "C" void bs_set_skew(int skew) { g_skew = skew < 1 ? 1 : skew; }

template <int R, int G, int MODE, bool SKIP>
static void run_group(const std::vector<uint8_t> &cls, size_t n, const std::vector<size_t> &begin, size_t l0,
                      size_t l1, const BsPattern &p, std::vector<Ev> &out)
{
   BsState<R, G> st[G];
   const int nl = (int)(l1 - l0);
   for (int g = 0; g < G; g++) bs_reset(st[g], p, nl == 32 ? ~0u : ((1u << nl) - 1u), g);
   const int skew = g_skew;
   std::vector<uint32_t> hist_ph((size_t)skew * G, 0u), hist_mh((size_t)skew * G, 0u);   // outputs of the last `skew` iterations
   uint32_t out_ph[G] = {0}, out_mh[G] = {0};
   uint32_t streak[8];
   for (long t = 0; st[G - 1].alive; t++) {
      uint32_t in_ph[G], in_mh[G];
      const size_t slot = (size_t)(t % skew) * G;          // holds the outputs of iteration t - skew
      for (int g = 0; g < G; g++) {
         in_ph[g] = g ? hist_ph[slot + g - 1] : 0u;
         in_mh[g] = g ? hist_mh[slot + g - 1] : 0u;
      }
      for (int g = 0; g < G; g++) {
         const long col = t - (long)g * skew;
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
         bs_rows<R, G, SKIP>(st[g], eq, skip, ph, mh);
         out_ph[g] = ph;
         out_mh[g] = mh;
         hist_ph[slot + g] = ph;                             // read again at iteration t + skew
         hist_mh[slot + g] = mh;
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
      SHAPE(8, 1) SHAPE(10, 1) SHAPE(12, 1) SHAPE(16, 1) SHAPE(20, 1) SHAPE(24, 1) SHAPE(32, 1)
      SHAPE(20, 2) SHAPE(24, 2) SHAPE(32, 2)
      SHAPE(20, 4) SHAPE(24, 4) SHAPE(26, 4) SHAPE(28, 4) SHAPE(32, 4)
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


// ---------------------------------------------------------------------------
// Long lines cut into segments (sqb_tables.h: kCutStride / kCutWindow), with the
// stride and window as parameters so that short random lines exercise many cuts.
// Mirrors the GPU pipeline: K1 (cut decision), pack (warm-up in front of a
// continuation, NULL columns behind a segment that is followed by another one),
// K2 (quiet warm-up, stopped lines), k_seg_reduce (one result per real line;
// segments behind a STOP are dead).
// ---------------------------------------------------------------------------
struct Seg {
   size_t start;        // first byte whose events this segment reports (A, or the line start)
   size_t line;         // real line index
   size_t lbeg;         // start of the real line
   bool cont, follow;
};

template <int R, int G, int MODE>
static void run_group_cut(const std::vector<uint8_t> &cls, size_t n, const std::vector<Seg> &segs, size_t l0, size_t l1,
                          uint32_t wup, const BsPattern &p, std::vector<std::vector<Ev>> &per_seg,
                          std::vector<uint8_t> &segstop)
{
   BsState<R, G> st[G];
   const int nl = (int)(l1 - l0);
   for (int g = 0; g < G; g++) bs_reset(st[g], p, nl == 32 ? ~0u : ((1u << nl) - 1u), g);
   uint32_t contmask = 0, followmask = 0;
   std::vector<size_t> s0(nl), len(nl);
   for (int r = 0; r < nl; r++) {
      const Seg &sg = segs[l0 + r];
      if (sg.cont) contmask |= 1u << r;
      if (sg.follow) followmask |= 1u << r;
      s0[r] = sg.start - (sg.cont ? wup : 0);
      const size_t next = l0 + r + 1 < segs.size() ? segs[l0 + r + 1].start : n;
      len[r] = (next - sg.start) + (sg.cont ? wup : 0) + (sg.follow ? 1 : 0);
   }
   size_t ncols = 0;
   for (int r = 0; r < nl; r++) ncols = std::max(ncols, len[r] + 1);
   uint32_t out_ph[G] = {0}, out_mh[G] = {0};
   uint32_t streak[8];
   for (long t = 0; t < (long)ncols + G - 1 && st[G - 1].alive; t++) {
      uint32_t in_ph[G], in_mh[G];
      for (int g = 0; g < G; g++) {
         in_ph[g] = g ? out_ph[g - 1] : 0u;
         in_mh[g] = g ? out_mh[g - 1] : 0u;
      }
      for (int g = 0; g < G; g++) {
         const long col = t - g;
         uint32_t p0 = 0, p1 = 0, p2 = 0;
         for (int r = 0; r < nl; r++) {
            uint8_t c;
            if (col < 0) c = kClsNull;
            else if (segs[l0 + r].follow && (size_t)col >= len[r]) c = kClsNull;      // silent end of a segment
            else {
               const size_t pos = s0[r] + (size_t)col;
               c = pos >= n ? kClsStop : (cls[pos] & 7);
            }
            p0 |= (uint32_t)(c & 1) << r;
            p1 |= (uint32_t)((c >> 1) & 1) << r;
            p2 |= (uint32_t)((c >> 2) & 1) << r;
         }
         for (int r = nl; r < 32; r++) { p0 |= 1u << r; p2 |= 1u << r; }     // STOP
         uint32_t slots[BS_SLOTS], anybase, stop, skip;
         bs_classes(p0, p1, p2, p, slots, anybase, stop, skip);
         auto eq = [&](int j) { return slots[p.slot[g * R + j]]; };
         uint32_t ph = in_ph[g], mh = in_mh[g];
         bs_rows<R, G, false>(st[g], eq, skip, ph, mh);
         out_ph[g] = ph;
         out_mh[g] = mh;
         if (g == G - 1) {
            const uint32_t quiet = (col <= (long)wup) ? contmask : 0u;
            const uint32_t evt = bs_report<R, G, MODE>(st[g], p, ph, mh, anybase, stop, streak, quiet);
            for (int r = 0; r < nl; r++)
               if ((evt >> r) & 1u) {
                  const Seg &sg = segs[l0 + r];
                  // column -> end offset inside the real line (what the finish kernels do)
                  const uint64_t end = (uint64_t)(s0[r] + (size_t)col - sg.lbeg);
                  per_seg[l0 + r].push_back(Ev{(uint64_t)sg.line, end, bs_value<BsState<R, G>::B>(streak, r)});
               }
         }
      }
   }
   for (int r = 0; r < nl; r++)
      if (((st[G - 1].stopped & followmask) >> r) & 1u) segstop[l0 + r] = 1;
}

extern "C" long bs_host_scan_cut(const char *buf, size_t n, const unsigned char *keys, int m, int tau, int options,
                                 size_t stride, size_t window, uint64_t *out, long cap, long *ncuts)
{
   BsPattern p;
   if (!build_bs_pattern(keys, m, tau, &p)) return -1;
   if ((options & OPT_NONDNA) == OPT_IGNORE) return -1;      // no cuts with SQ_IGNORE
   const uint32_t wup = bs_warmup(m, tau);
   if (wup > window) return -1;
   ClassTable ct;
   build_class_table(options, &ct);
   std::vector<uint8_t> cls(n);
   for (size_t i = 0; i < n; i++) cls[i] = ct.code[(unsigned char)buf[i]];
   // K1: line starts and cuts
   std::vector<Seg> segs;
   {
      std::vector<uint8_t> is_start(n + 1, 0);
      if (n > 0) is_start[0] = 1;
      for (size_t i = 0; i + 1 < n; i++) if (buf[i] == '\n') is_start[i + 1] = 1;
      std::vector<uint8_t> is_cut(n + 1, 0);
      *ncuts = 0;
      for (size_t a = window; a < n; a += stride) {
         bool any = a - window == 0;                            // the first line start sits in the window of a == window
         for (size_t q = a - window; q + 1 <= a && !any; q++) any = buf[q] == '\n';   // newline in [a - window, a - 1]
         if (!any) { is_cut[a] = 1; (*ncuts)++; }
      }
      size_t line = 0, lbeg = 0;
      bool have = false;
      for (size_t i = 0; i < n; i++) {
         if (is_start[i]) {
            if (have) line++;
            have = true;
            lbeg = i;
            segs.push_back(Seg{i, line, lbeg, false, false});
         } else if (is_cut[i]) {
            segs.push_back(Seg{i, line, lbeg, true, false});
         }
      }
      for (size_t k = 0; k + 1 < segs.size(); k++) segs[k].follow = segs[k + 1].cont;
   }
   const int match = options & OPT_MATCH;
   const int mode = match == OPT_ALL ? BS_ALL : (match == OPT_BEST ? BS_BEST : BS_FIRST);
   std::vector<std::vector<Ev>> per_seg(segs.size());
   std::vector<uint8_t> segstop(segs.size(), 0);
   for (size_t l0 = 0; l0 < segs.size(); l0 += 32) {
      const size_t l1 = std::min(segs.size(), l0 + 32);
#define SHAPE(R, G)                                                                                          \
   if (p.rows == R && p.parts == G) {                                                                        \
      if (mode == BS_FIRST) run_group_cut<R, G, BS_FIRST>(cls, n, segs, l0, l1, wup, p, per_seg, segstop);   \
      if (mode == BS_BEST) run_group_cut<R, G, BS_BEST>(cls, n, segs, l0, l1, wup, p, per_seg, segstop);     \
      if (mode == BS_ALL) run_group_cut<R, G, BS_ALL>(cls, n, segs, l0, l1, wup, p, per_seg, segstop);       \
   }
      SHAPE(8, 1) SHAPE(10, 1) SHAPE(12, 1) SHAPE(16, 1) SHAPE(20, 1) SHAPE(24, 1) SHAPE(32, 1)
      SHAPE(20, 2) SHAPE(24, 2) SHAPE(32, 2)
      SHAPE(20, 4) SHAPE(24, 4) SHAPE(26, 4) SHAPE(28, 4) SHAPE(32, 4)
#undef SHAPE
   }
   // k_seg_reduce: one result per real line, segments behind a STOP are dead
   long k = 0;
   for (size_t s = 0; s < segs.size();) {
      size_t e = s + 1;
      while (e < segs.size() && segs[e].cont) e++;
      bool dead = false, have = false;
      Ev best{0, 0, 0};
      for (size_t q = s; q < e; q++) {
         if (!dead) {
            for (size_t i = 0; i < per_seg[q].size(); i++) {
               const Ev &ev = per_seg[q][i];
               if (mode == BS_ALL) {
                  if (k >= cap) return -2;
                  out[3 * k] = ev.line + 1; out[3 * k + 1] = ev.end; out[3 * k + 2] = ev.dist; k++;
               } else if (mode == BS_FIRST) {
                  if (!have) { best = ev; have = true; }
               } else {
                  // inside a segment the last improvement wins; across segments a strictly smaller distance
                  const bool last_of_seg = i + 1 == per_seg[q].size();
                  if (last_of_seg && (!have || ev.dist < best.dist)) { best = ev; have = true; }
               }
            }
         }
         if (segstop[q]) dead = true;
      }
      if (mode != BS_ALL && have) {
         if (k >= cap) return -2;
         out[3 * k] = best.line + 1; out[3 * k + 1] = best.end; out[3 * k + 2] = best.dist; k++;
      }
      s = e;
   }
   return k;
}


// ---------------------------------------------------------------------------
// tau <= 2: the NFA-level automaton (bs_wm_step), plain and with segment cuts.
// cut == 0: every line is one segment.
// ---------------------------------------------------------------------------
template <int R, int T, int MODE, bool SKIP>
static void run_group_wm(const std::vector<uint8_t> &cls, size_t n, const std::vector<Seg> &segs, size_t l0, size_t l1,
                         uint32_t wup, const BsPattern &p, std::vector<std::vector<Ev>> &per_seg,
                         std::vector<uint8_t> &segstop)
{
   BsWmState<R, T> st;
   const int nl = (int)(l1 - l0);
   bs_wm_reset(st, p, nl == 32 ? ~0u : ((1u << nl) - 1u));
   uint32_t contmask = 0, followmask = 0;
   std::vector<size_t> s0(nl), len(nl);
   for (int r = 0; r < nl; r++) {
      const Seg &sg = segs[l0 + r];
      if (sg.cont) contmask |= 1u << r;
      if (sg.follow) followmask |= 1u << r;
      s0[r] = sg.start - (sg.cont ? wup : 0);
      const size_t next = l0 + r + 1 < segs.size() ? segs[l0 + r + 1].start : n;
      len[r] = (next - sg.start) + (sg.cont ? wup : 0) + (sg.follow ? 1 : 0);
   }
   uint32_t streak[4];
   for (size_t col = 0; st.alive; col++) {
      uint32_t p0 = 0, p1 = 0, p2 = 0;
      for (int r = 0; r < nl; r++) {
         uint8_t c;
         if (segs[l0 + r].follow && col >= len[r]) c = kClsNull;
         else {
            const size_t pos = s0[r] + col;
            c = pos >= n ? kClsStop : (cls[pos] & 7);
         }
         p0 |= (uint32_t)(c & 1) << r;
         p1 |= (uint32_t)((c >> 1) & 1) << r;
         p2 |= (uint32_t)((c >> 2) & 1) << r;
      }
      for (int r = nl; r < 32; r++) { p0 |= 1u << r; p2 |= 1u << r; }     // STOP
      uint32_t slots[BS_SLOTS], anybase, stop, skip;
      bs_classes(p0, p1, p2, p, slots, anybase, stop, skip);
      auto eq = [&](int j) { return slots[p.slot[j]]; };
      const uint32_t quiet = col <= wup ? contmask : 0u;
      const uint32_t evt = bs_wm_step<R, T, MODE, SKIP>(st, eq, anybase, stop, skip, streak, quiet);
      for (int r = 0; r < nl; r++)
         if ((evt >> r) & 1u) {
            const Seg &sg = segs[l0 + r];
            per_seg[l0 + r].push_back(Ev{(uint64_t)sg.line, (uint64_t)(s0[r] + col - sg.lbeg), bs_value_unary<T>(streak, r)});
         }
      // a followed segment whose columns are used up is over (the kernel stops at the tile's last column)
      bool any = false;
      for (int r = 0; r < nl; r++)
         if (((st.alive >> r) & 1u) && !(segs[l0 + r].follow && col + 1 >= len[r] + 1)) any = true;
      if (!any) break;
   }
   for (int r = 0; r < nl; r++)
      if (((st.stopped & followmask) >> r) & 1u) segstop[l0 + r] = 1;
}

extern "C" long bs_host_scan_wm(const char *buf, size_t n, const unsigned char *keys, int m, int tau, int options,
                                size_t stride, size_t window, uint64_t *out, long cap, long *ncuts)
{
   BsPattern p;
   if (!build_bs_pattern(keys, m, tau, &p) || p.parts != 1 || tau > 2) return -1;
   const bool skipmode = (options & OPT_NONDNA) == OPT_IGNORE;
   const bool cut = stride != 0;
   if (cut && skipmode) return -1;
   const uint32_t wup = bs_warmup(m, tau);
   if (cut && wup > window) return -1;
   ClassTable ct;
   build_class_table(options, &ct);
   std::vector<uint8_t> cls(n);
   for (size_t i = 0; i < n; i++) cls[i] = ct.code[(unsigned char)buf[i]];
   std::vector<Seg> segs;
   {
      std::vector<uint8_t> is_cut(n + 1, 0);
      *ncuts = 0;
      if (cut)
         for (size_t a = window; a < n; a += stride) {
            bool any = a - window == 0;
            for (size_t q = a - window; q + 1 <= a && !any; q++) any = buf[q] == '\n';
            if (!any) { is_cut[a] = 1; (*ncuts)++; }
         }
      size_t line = 0, lbeg = 0;
      bool have = false;
      for (size_t i = 0; i < n; i++) {
         if (i == 0 || buf[i - 1] == '\n') {
            if (have) line++;
            have = true;
            lbeg = i;
            segs.push_back(Seg{i, line, lbeg, false, false});
         } else if (is_cut[i]) {
            segs.push_back(Seg{i, line, lbeg, true, false});
         }
      }
      for (size_t k = 0; k + 1 < segs.size(); k++) segs[k].follow = segs[k + 1].cont;
   }
   const int match = options & OPT_MATCH;
   const int mode = match == OPT_ALL ? BS_ALL : (match == OPT_BEST ? BS_BEST : BS_FIRST);
   std::vector<std::vector<Ev>> per_seg(segs.size());
   std::vector<uint8_t> segstop(segs.size(), 0);
   for (size_t l0 = 0; l0 < segs.size(); l0 += 32) {
      const size_t l1 = std::min(segs.size(), l0 + 32);
#define WM3(R, T, M)                                                                                    \
   if (mode == M) {                                                                                     \
      if (skipmode) run_group_wm<R, T, M, true>(cls, n, segs, l0, l1, wup, p, per_seg, segstop);        \
      else run_group_wm<R, T, M, false>(cls, n, segs, l0, l1, wup, p, per_seg, segstop);                \
   }
#define WM2(R, T) if (p.rows == R && tau + 1 == T) { WM3(R, T, BS_FIRST) WM3(R, T, BS_BEST) WM3(R, T, BS_ALL) }
#define WM1(R) WM2(R, 1) WM2(R, 2) WM2(R, 3)
      WM1(8) WM1(10) WM1(12) WM1(16) WM1(20) WM1(24) WM1(32)
#undef WM1
#undef WM2
#undef WM3
   }
   long k = 0;
   for (size_t s = 0; s < segs.size();) {
      size_t e = s + 1;
      while (e < segs.size() && segs[e].cont) e++;
      bool dead = false, have = false;
      Ev best{0, 0, 0};
      for (size_t q = s; q < e; q++) {
         if (!dead) {
            for (size_t i = 0; i < per_seg[q].size(); i++) {
               const Ev &ev = per_seg[q][i];
               if (mode == BS_ALL) {
                  if (k >= cap) return -2;
                  out[3 * k] = ev.line + 1; out[3 * k + 1] = ev.end; out[3 * k + 2] = ev.dist; k++;
               } else if (mode == BS_FIRST) {
                  if (!have) { best = ev; have = true; }
               } else {
                  const bool last_of_seg = i + 1 == per_seg[q].size();
                  if (last_of_seg && (!have || ev.dist < best.dist)) { best = ev; have = true; }
               }
            }
         }
         if (segstop[q]) dead = true;
      }
      if (mode != BS_ALL && have) {
         if (k >= cap) return -2;
         out[3 * k] = best.line + 1; out[3 * k + 1] = best.end; out[3 * k + 2] = best.dist; k++;
      }
      s = e;
   }
   return k;
}

// SQB_FASTQ chunk alignment (sqb_tables.h: fastq_record_start), as the chunk pipeline calls it
extern "C" long bs_fastq_record_start(const char *text, size_t lo, size_t cut, size_t nbytes)
{
   const size_t q = sqb::fastq_record_start(text, lo, cut, nbytes);
   return q == (size_t)-1 ? -1L : (long)q;
}

// sysfs cpulist parser of the NUMA placement (sqb_tables.h: parse_cpulist)
extern "C" int bs_parse_cpulist(const char *s, unsigned char *cpus, int maxcpus)
{
   return sqb::parse_cpulist(s, cpus, maxcpus);
}

// ---------------------------------------------------------------------------------------------------------------
// The arithmetic of the fused tokenise + pack kernel (seeq_b200/csrc/sqb_k12_arith.h) on the host
// ---------------------------------------------------------------------------------------------------------------
#include "sqb_k12_arith.h"

extern "C" uint32_t k12_nl_flags(uint32_t w) { return sqb::nl_flags(w); }
extern "C" uint32_t k12_chunk_before(uint32_t v) { return sqb::chunk_before(v); }
extern "C" uint32_t k12_chunk_byte_of(uint32_t u) { return sqb::chunk_byte_of(u); }
extern "C" uint32_t k12_chunk_flags(const unsigned char *p32)
{
   uint32_t w[8];
   memcpy(w, p32, 32);
   return sqb::chunk_flags_words(w);
}
extern "C" void k12_table32(int options, uint32_t *out256)
{
   sqb::ClassTable ct;
   sqb::ClassTable32 t;
   sqb::build_class_table(options, &ct);
   sqb::build_class_table32(ct, &t);
   memcpy(out256, t.w, sizeof t.w);
}

// One group of up to 32 lines the way k12_scan_pack builds its planes: every line is read from the aligned word at or
// in front of its start (`text` must be 4-byte aligned and readable up to 35 + 32 bytes past the longest line), four
// columns per "lane", one table look-up and one multiply-add per byte into the accumulator of its column and line octet,
// PRMT assembly, lead (NULL) columns OR-ed in.  planes: [block of 32 columns][plane][32 columns].  Returns the columns.
extern "C" uint32_t k12_host_group(const unsigned char *text, const uint32_t *starts, const uint32_t *lens, int nlines, int options,
                                   uint32_t *planes, uint32_t planes_cap_words)
{
   sqb::ClassTable ct;
   sqb::ClassTable32 t;
   sqb::build_class_table(options, &ct);
   sqb::build_class_table32(ct, &t);
   uint32_t aligned[32], lead[32], ncols = 0, m1 = 0, m2 = 0, m3 = 0;
   for (int r = 0; r < 32; r++) {
      const int rr = r < nlines ? r : 0;                       // a slot without a line re-reads slot 0
      aligned[r] = starts[rr] & ~3u;
      lead[r] = r < nlines ? (starts[rr] & 3u) : 0u;
      if (r < nlines) ncols = std::max(ncols, lens[r] + lead[r]);
      if (lead[r] >= 1) m1 |= 1u << r;
      if (lead[r] >= 2) m2 |= 1u << r;
      if (lead[r] == 3) m3 |= 1u << r;
   }
   const uint32_t nblk = (ncols + 31u) >> 5;
   if (nblk * 96u > planes_cap_words) return 0xffffffffu;
   for (uint32_t cb = 0; cb < nblk; cb++)
      for (uint32_t cq = 0; cq < 8; cq++) {                    // one lane of the kernel
         uint32_t A[4][4], P[3][4];
         for (int q = 0; q < 4; q++) {
            for (int j = 0; j < 4; j++) A[j][q] = 0;
            for (int i = 0; i < 8; i++) {
               uint32_t w;
               memcpy(&w, text + aligned[8 * q + i] + cb * 32u + cq * 4u, 4);
               const uint32_t x[4] = {w & 0xffu, sqb::k12_prmt(w, 0u, 0x4441u), sqb::k12_prmt(w, 0u, 0x4442u), w >> 24};
               for (int j = 0; j < 4; j++) A[j][q] = t.w[x[j]] * (1u << i) + A[j][q];
            }
         }
         for (int j = 0; j < 4; j++) sqb::k12_planes_of_column(A[j], P[0][j], P[1][j], P[2][j]);
         if (cb == 0 && cq == 0)
            for (int p = 0; p < 3; p++) {
               P[p][0] |= m1;
               P[p][1] |= m2;
               P[p][2] |= m3;
            }
         for (int p = 0; p < 3; p++)
            for (int j = 0; j < 4; j++) planes[cb * 96u + (uint32_t)p * 32u + cq * 4u + (uint32_t)j] = P[p][j];
      }
   return ncols;
}
