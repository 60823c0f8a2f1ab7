// sqb_k2_bitslice.cuh -- K2, line-bit-sliced forward matcher (patterns <= 32
// positions over many short lines; the integer-pipe-lean path).
//
// Lines are handled in TILES of 1024 consecutive lines = one warp of the match
// kernel, in which LANE g owns the 32 lines [g*32, g*32+32) of the tile, one bit
// each (sqb_bitslice.h).  The text reaches it as bit-planes:
//
//   k15_tile_cols / k15_scan   columns per tile (longest line + 1) and their
//                              exclusive prefix = where the tile's planes live;
//                              also the device-side decision whether this scan
//                              is bit-sliced at all (enough lines, no long line)
//   k15_pack    a warp takes 64 consecutive lines (2 groups).  Lane L streams
//               the class nibbles (K1's `codes`) of lines L and 32+L with
//               16-byte loads, realigns them to the line start with funnel
//               shifts, and the warp transposes each 32x32 bit block with a
//               5-stage shuffle butterfly: lane j then holds plane (j&3) of
//               column (j>>2) for 32 lines.  Stored as planes[tile][column][g]
//               = {p0,p1,p2,-} (uint4), 32 B per store and column.
//   k2_bitslice lane g loads ONE uint4 per column (the warp: 512 contiguous
//               bytes), derives the Eq slots and runs bs_step = 6 LOP3 per
//               pattern row for 32 text bytes.  Events (rare) leave the
//               bit-sliced world through a per-lane slow path.
//
// A first version fused the pack step into the match kernel (32 lines x 32
// lanes per warp, 16-column windows through shared memory): parity was green,
// but every window touched a different 64-byte block of 1024 lines, the
// working set of the resident warps (155 MB) exceeded L2 and ncu showed 8.5 GB
// of DRAM reads for 0.75 GB of codes.  Packing line-major in its own kernel
// keeps the working set at 64 lines per warp and reads every byte once.
#pragma once

#include <type_traits>

#include "sqb_bitslice.h"
#include "sqb_kernels.cuh"

namespace sqb {

constexpr int kBsWarps   = 4;                       // warps per CTA of the match kernel
constexpr int kBsThreads = kBsWarps * 32;
constexpr int kBsTileLines = 1024;                  // lines per warp tile
constexpr int kBsBlock   = 4;                       // columns per prefetch block of the match kernel
#ifndef SQB_PACK_CTAS
#define SQB_PACK_CTAS 4                             // CTAs per SM of k15_pack (A/B knob: 5 -> 51 registers, 6 -> 42)
#endif
#ifndef SQB_G2_BLOCK
#define SQB_G2_BLOCK 6                              // prefetch block of the multi-part matcher (r1y, r1z: 2 -12 %, 6 +1.5 %, 8 -10 %)
#endif
#ifndef SQB_G2_RADDR_SMEM
#define SQB_G2_RADDR_SMEM 0                         // 1: Eq slot addresses of the multi-part matcher from shared memory
#endif                                              //    instead of R registers (A/B knob)
#ifndef SQB_WM_4CTA_ROWS
#define SQB_WM_4CTA_ROWS 60                         // NFA-level instances of up to this many state planes run 4 CTAs per SM (128 registers; m = 20, tau = 2: 0.524 -> 0.515 ms, r4v)
#endif
#ifndef SQB_G2_CTAS
#define SQB_G2_CTAS 3                               // CTAs per SM of the multi-part matcher with R <= 24 (A/B knob)
#endif

// one group of 32 lines of the fused tokenise + pack kernel (sqb_k12_fused.cuh)
struct GroupDesc {
   uint32_t tile;       // K1 tile the lines start in
   uint32_t first;      // local entry index of slot 0 (consecutive entries) / index into gent (line filter)
   uint32_t meta;       // lines in the group (1..32) | columns << 8
   uint32_t poff;       // first uint4 of the group's planes: [block of 32 columns][plane 0..2][32 columns] words
   uint32_t lead_lo, lead_hi;   // bit r: bits 0 / 1 of the NULL columns in front of line r (its start & 3)
   uint32_t pad0, pad1;
};

struct BsPrepArgs {
   const uint32_t *ls;
   uint32_t max_lines;
   uint32_t n;                    // text bytes
   unsigned long long *ctr;
   uint32_t *tile_cols;           // columns of every tile
   uint32_t *tile_off;            // exclusive prefix of tile_cols
   uint32_t max_tiles;
   unsigned long long planes_cap; // columns the plane buffer can hold
   BsGate gate;
   const uint32_t *lid;           // line of every ls entry (segment cuts), or nullptr
   uint32_t wup;                  // warm-up bytes in front of a continuation segment
   int cuts_possible;             // this scan could have cut its long lines but did not (see k15_scan)
   const uint32_t *act;           // line filter: the entries of ls to look at (nullptr: all of them)
};

// the bit-sliced kernels number the lines they look at 0 .. nact-1 ("slots": tile =
// slot / 1024, group = slot / 32); with the line filter slot q is entry act[q] of ls
__device__ __forceinline__ uint32_t bs_nslots(const unsigned long long *ctr, const uint32_t *act, uint32_t max_lines)
{
   return (uint32_t)min(ctr[act ? C_NACTIVE : C_NPSEUDO], (unsigned long long)max_lines);
}

// With segment cuts an entry l of ls is a CONTINUATION if it belongs to the same
// line as the entry before it, and is FOLLOWED if the next entry continues it.
// A continuation scans from ls[l] - wup; a followed segment scans up to and
// including the first byte of the next one and then falls silent (NULL columns).
struct SegShape {
   uint32_t begin;                // first text byte fed to the automaton
   uint32_t limit;                // columns to feed (0xffffffff: until the STOP of the line)
   bool cont, follow;
};
__device__ __forceinline__ SegShape seg_shape(const uint32_t *ls, const uint32_t *lid, uint32_t wup, uint32_t l,
                                              uint32_t nlines)
{
   SegShape g{ls[l], 0xffffffffu, false, false};
   if (lid) {            // callers pass nullptr when K1 made no cut
      const uint32_t me = lid[l];
      g.cont = l > 0u && lid[l - 1u] == me;
      g.follow = l + 1u < nlines && lid[l + 1u] == me;
      if (g.cont) g.begin -= wup;
      if (g.follow) g.limit = ls[l + 1u] + 1u - g.begin;
   }
   return g;
}

// columns of every tile = longest line (with its terminator) + 1; one warp per tile
static __global__ void __launch_bounds__(kThreads) k15_tile_cols(const BsPrepArgs a)
{
   const int lane = threadIdx.x & 31;
   const uint32_t nents = (uint32_t)min(a.ctr[C_NPSEUDO], (unsigned long long)a.max_lines);
   const uint32_t nlines = bs_nslots(a.ctr, a.act, a.max_lines);
   const uint32_t ntiles = (nlines + kBsTileLines - 1) / kBsTileLines;
   const uint32_t wid = (blockIdx.x * kThreads + threadIdx.x) >> 5, nw = (gridDim.x * kThreads) >> 5;
   const uint32_t *lid = a.ctr[C_NCUTS] != 0ull ? a.lid : nullptr;
   for (uint32_t t = wid; t < ntiles; t += nw) {
      uint32_t mx = 0;
      const uint32_t l0 = t * kBsTileLines;
#pragma unroll 4
      for (int k = 0; k < 32; k++) {
         const uint32_t q = l0 + (uint32_t)k * 32u + (uint32_t)lane;
         if (q < nlines) {
            const uint32_t l = a.act ? a.act[q] : q;
            uint32_t len = a.ls[l + 1] - a.ls[l];
            if (lid) {
               const SegShape g = seg_shape(a.ls, lid, a.wup, l, nents);
               len = g.follow ? g.limit : a.ls[l + 1] - g.begin;
            }
            mx = max(mx, len);
         }
      }
      mx = __reduce_max_sync(kFull, mx);
      if (lane == 0) a.tile_cols[t] = mx + 1u;
   }
}

// exclusive scan of tile_cols (one CTA) and the decision: 1 = bit-sliced scan,
// 0 = word-parallel kernels, 2 = bit-sliced but the plane buffer is too small
// (the host repeats the scan with the exact size, ctr[C_BS_COLS]), 3 = there are
// long lines and the scan did not cut them: the host repeats it with segment
// cuts (and keeps cutting from then on); no matcher runs
static __global__ void __launch_bounds__(1024) k15_scan(const BsPrepArgs a)
{
   __shared__ CtaScanSmem cs;
   const unsigned long long nl_dev = a.ctr[C_NPSEUDO];
   const uint32_t nlines = bs_nslots(a.ctr, a.act, a.max_lines);
   const uint32_t ntiles = (nlines + kBsTileLines - 1) / kBsTileLines;
   uint32_t mx = 0;
   const unsigned long long cols = cta_scan_u32(a.tile_cols, a.tile_off, ntiles, cs, &mx);   // total <= text bytes < 2^32
   if (threadIdx.x == 0) {
      // with segment cuts the entries of ls are segments: only the bit-sliced kernel knows them
      const bool want = nl_dev <= a.max_lines &&
                        ((a.lid != nullptr && a.ctr[C_NCUTS] != 0ull) || (nl_dev >= a.gate.min_lines && mx <= a.gate.max_line + 1u));
      a.ctr[C_BS_COLS] = cols;
      if (a.cuts_possible && nl_dev <= a.max_lines && mx > a.gate.max_line + 1u) a.ctr[C_BS_SELECTED] = 3ull;
      else a.ctr[C_BS_SELECTED] = !want ? 0ull : (cols <= a.planes_cap ? 1ull : 2ull);
   }
}

struct BsPackArgs {
   const uint4 *codes;            // 32 class nibbles per 16 bytes
   uint32_t ncode16;              // entries of codes[]
   const uint32_t *ls;
   uint32_t max_lines;
   const unsigned long long *ctr;
   const uint32_t *tile_cols, *tile_off;
   uint4 *planes;                 // [tile_off + column][32 groups] {p0, p1, p2, -}
   const uint32_t *lid;           // segment cuts (or nullptr)
   uint32_t wup;
   uint32_t *gmask, *gfollow;     // out, per group of 32 slots: continuations / followed segments
   const uint32_t *act;           // line filter (or nullptr)
};

// class nibbles at columns >= limit of a followed segment become NULL (7): the
// segment falls silent there (block = the 32 columns starting at c0)
__device__ __forceinline__ void null_fill(uint32_t (&w)[4], uint32_t c0, uint32_t limit)
{
   if (limit >= c0 + 32u) return;
   const uint32_t keep = limit > c0 ? limit - c0 : 0u;          // 0..31 nibbles stay
#pragma unroll
   for (int i = 0; i < 4; i++) {
      const uint32_t k = keep > 8u * (uint32_t)i ? min(keep - 8u * (uint32_t)i, 8u) : 0u;
      if (k < 8u) w[i] |= 0x77777777u << (4u * k);
   }
}

// 32x32 bit transpose across the warp: on return lane j holds, in bit i, bit j
// of lane i's input.  keep[s] / rot[s] are per-lane constants of stage s.  The
// first two stages exchange whole half-words and bytes: one PRMT each (sel16 /
// sel8, per-lane selectors) instead of mask + funnel shift + merge.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, const uint32_t (&keep)[5], const uint32_t (&rot)[5],
                                                     uint32_t sel16, uint32_t sel8)
{
   x = __byte_perm(x, __shfl_xor_sync(kFull, x, 16), sel16);
   x = __byte_perm(x, __shfl_xor_sync(kFull, x, 8), sel8);
#pragma unroll
   for (int s = 2; s < 5; s++) {
      const uint32_t y = __shfl_xor_sync(kFull, x, 16 >> s);
      x = (x & keep[s]) | (__funnelshift_l(y, y, rot[s]) & ~keep[s]);
   }
   return x;
}

// nibble stream of one line, realigned: every call returns the next 32 nibbles
// (128 bits) starting at the line start
struct NibbleStream {
   const uint4 *codes;
   uint32_t chunk, last;          // next 16-byte chunk to load, last valid chunk
   uint32_t ws, bs;               // word and bit shift of the line start inside its first chunk
   uint4 prev, cur;               // the two chunks the next call combines: cur is loaded one call ahead,
                                  // its latency hides behind the transposes of the block before
   bool valid;

   __device__ __forceinline__ uint4 load()
   {
      uint4 v = cur;
      if (valid) v = codes[chunk];
      chunk = min(chunk + 1u, last);
      return v;
   }
   __device__ __forceinline__ void open(const uint4 *c, uint32_t ncode16, uint32_t begin, bool ok)
   {
      codes = c;
      last = ncode16 - 1u;
      valid = ok;
      chunk = min(begin >> 5, last);
      ws = (begin & 31u) >> 3;
      bs = (begin & 7u) * 4u;
      cur = make_uint4(0x55555555u, 0x55555555u, 0x55555555u, 0x55555555u);
      prev = load();
      cur = load();
   }
   __device__ __forceinline__ void next(uint32_t (&out)[4])
   {
      const uint32_t w[8] = {prev.x, prev.y, prev.z, prev.w, cur.x, cur.y, cur.z, cur.w};
      uint32_t u[5];
#pragma unroll
      for (int i = 0; i < 5; i++) {
         const uint32_t lo = (ws & 1u) ? w[i + 1] : w[i];
         const uint32_t hi = (ws & 1u) ? w[i + 3] : w[i + 2];
         u[i] = (ws & 2u) ? hi : lo;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) out[i] = __funnelshift_r(u[i], u[i + 1], bs);
      prev = cur;
      cur = load();
   }
};

static __global__ void __launch_bounds__(kThreads, SQB_PACK_CTAS) k15_pack(const BsPackArgs a)
{
   if (a.ctr[C_BS_SELECTED] != 1ull) return;
   const int lane = threadIdx.x & 31;
   const uint32_t nents = (uint32_t)min(a.ctr[C_NPSEUDO], (unsigned long long)a.max_lines);
   const uint32_t nlines = bs_nslots(a.ctr, a.act, a.max_lines);
   const uint32_t npairs = (nlines + 63u) / 64u;
   const uint32_t *lid = a.ctr[C_NCUTS] != 0ull ? a.lid : nullptr;
   uint32_t keep[5], rot[5];
   {
      const uint32_t m[5] = {0x0000FFFFu, 0x00FF00FFu, 0x0F0F0F0Fu, 0x33333333u, 0x55555555u};
#pragma unroll
      for (int s = 0; s < 5; s++) {
         const uint32_t d = 16u >> s;
         keep[s] = (lane & d) ? ~m[s] : m[s];
         rot[s] = (lane & d) ? 32u - d : d;
      }
   }
   // stage 0: low half of x | low half of y << 16, or (lanes 16..31) y >> 16 | high half of x; stage 1 alike per byte
   const uint32_t sel16 = (lane & 16) ? 0x3276u : 0x5410u, sel8 = (lane & 8) ? 0x3715u : 0x6240u;
   const uint32_t wid = (blockIdx.x * kThreads + threadIdx.x) >> 5, nw = (gridDim.x * kThreads) >> 5;
   for (uint32_t pair = wid; pair < npairs; pair += nw) {
      const uint32_t tile = pair >> 4;                          // 16 pairs of groups per tile
      const uint32_t g0 = (pair & 15u) * 2u;                    // even group of the pair inside the tile
      const uint32_t ncols = a.tile_cols[tile];
      uint4 *out = a.planes + (size_t)a.tile_off[tile] * 32u;
      const uint32_t la = pair * 64u + (uint32_t)lane, lb = la + 32u;
      NibbleStream sa, sb;
      SegShape ga{0u, 0xffffffffu, false, false}, gb{0u, 0xffffffffu, false, false};
      if (la < nlines) ga = seg_shape(a.ls, lid, a.wup, a.act ? a.act[la] : la, nents);
      if (lb < nlines) gb = seg_shape(a.ls, lid, a.wup, a.act ? a.act[lb] : lb, nents);
      sa.open(a.codes, a.ncode16, ga.begin, la < nlines);
      sb.open(a.codes, a.ncode16, gb.begin, lb < nlines);
      if (lid) {
         const uint32_t ca = __ballot_sync(kFull, ga.cont), cb = __ballot_sync(kFull, gb.cont);
         const uint32_t fa = __ballot_sync(kFull, ga.follow), fb = __ballot_sync(kFull, gb.follow);
         if (lane == 0) {
            a.gmask[pair * 2u] = ca;
            a.gmask[pair * 2u + 1u] = cb;
            a.gfollow[pair * 2u] = fa;
            a.gfollow[pair * 2u + 1u] = fb;
         }
      }
      // what this lane stores per 8-column block: 8 bytes = two planes of one group
      //   even lanes: group g0,     planes (j&2), (j&2)+1 of column j>>2
      //   odd  lanes: group g0 + 1, the same planes
      const uint32_t slot16 = (g0 + (uint32_t)(lane & 1)) * 2u + (uint32_t)((lane >> 1) & 1);   // in 8-byte units
      for (uint32_t c0 = 0; c0 < ncols; c0 += 32) {
         uint32_t wa[4], wb[4];
         sa.next(wa);
         sb.next(wb);
         if (lid) {
            null_fill(wa, c0, ga.limit);
            null_fill(wb, c0, gb.limit);
         }
#pragma unroll
         for (int k = 0; k < 4; k++) {
            const uint32_t ta = warp_transpose32(wa[k], keep, rot, sel16, sel8);
            const uint32_t tb = warp_transpose32(wb[k], keep, rot, sel16, sel8);
            // lane j holds plane (j&3) of column (j>>2): pair up planes 0|1 and 2|3
            const uint32_t give = (lane & 1) ? ta : tb;
            const uint32_t got = __shfl_xor_sync(kFull, give, 1);
            const uint2 v = (lane & 1) ? make_uint2(got, tb) : make_uint2(ta, got);
            const uint32_t col = c0 + 8u * (uint32_t)k + (uint32_t)(lane >> 2);
            if (col < ncols) reinterpret_cast<uint2 *>(out + (size_t)col * 32u)[slot16] = v;
         }
      }
   }
}

struct K2BsArgs {
   const uint4 *planes;
   const uint32_t *tile_cols, *tile_off;
   uint32_t max_lines;
   unsigned long long *ctr;
   unsigned long long *res;       // BS_FIRST / BS_BEST: per-line (dist << 32 | end); preset to kNoMatch
   uint32_t *cnt;                 // BS_ALL: per-line event count
   Event *ev;                     // BS_ALL: unordered events
   uint32_t ev_cap;
   int count_only;                // counts only: no res / event stores
   const uint32_t *gmask, *gfollow;  // segment cuts: per group, continuations / followed segments (or nullptr)
   uint8_t *segstop;              // out: followed segments that ran into a STOP (the rest of the line is dead)
   uint32_t wup;                  // a continuation reports the events that end after its warm-up
   const uint32_t *act;           // line filter: slot -> entry of ls (nullptr: identity)
   // fused tokenise + pack (sqb_k12_fused.cuh): the planes come per GROUP of 32 lines, described by gdesc
   const GroupDesc *gdesc;        // nullptr: planes[tile][column][group] of k15_pack
   uint32_t gdesc_cap;
   const uint16_t *gent;          // line filter: local entry index of every slot (or nullptr: consecutive)
   const uint32_t *k1_tile_base;  // first ls entry of every K1 tile (k1_scan_tiles)
};

struct BsWarpSmem {
   // Eq slots of the column in flight, [buffer][slot][lane].  Two buffers, used by the
   // even and odd columns (G == 1): the slot stores of column c+1 do not have to wait
   // for the Eq loads of column c
   uint32_t slots[2][BS_SLOTS][32];
};
struct BsWarpSmemAll : BsWarpSmem {
   uint32_t cnt[32 * 32];         // events per line of the warp's groups (BS_ALL)
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
   uint32_t v;
   asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");     // ordered after the slot stores
   return v;
}

// R rows per part, G parts per group of 32 lines (sqb_bitslice.h).  A warp serves
// 32 / G groups of one tile: lane = part * (32 / G) + group, so that the lanes
// of one part read neighbouring uint4 of one plane column.  Part p runs p
// columns behind part 0 and receives the horizontal delta of the part below
// with one pair of shuffles per column; only the last part reports.
//
// WM > 0 selects the NFA-level automaton (bs_wm_step, tau = WM - 1 <= 2, G == 1)
// instead of Myers' delta encoding: fewer logic ops per column for small tau.
// CUSTOM = false: an instance for patterns without bracket classes (single-part automata only).  The code of the
// bracket classes costs such patterns 8 % when it is merely SKIPPED at run time (metric shape, r4t: 0.571 against 0.525 ms).
template <int R, int G, int MODE, bool SKIP, int WM = 0, bool FUSED = false, bool CUSTOM = true>
__global__ void __launch_bounds__(kBsThreads, WM ? (R * WM <= 24 ? 6 : (R * WM <= SQB_WM_4CTA_ROWS ? 4 : 3)) : (G > 1 ? (R <= 24 ? SQB_G2_CTAS : 2) : (R <= 16 ? 4 : 3)))
k2_bitslice(const K2BsArgs a, const __grid_constant__ BsPattern pat)
{
   static_assert(WM == 0 || G == 1, "the NFA-level automaton is single-part");
   static_assert(!FUSED || G == 1, "the fused tokenise + pack kernel serves single-part automata");
   using Smem = typename std::conditional<MODE == BS_ALL, BsWarpSmemAll, BsWarpSmem>::type;
   using State = typename std::conditional<WM != 0, BsWmState<R, (WM ? WM : 1)>, BsState<R, G>>::type;
   extern __shared__ __align__(128) uint8_t dyn[];
   __shared__ uint32_t s_red[2][kBsWarps];
   __shared__ uint32_t s_off[G > 1 ? R * G : 1];

   if (a.ctr[C_BS_SELECTED] != 1ull) return;              // the word-parallel kernel takes this scan
   constexpr int NG = 32 / G;                             // groups per warp
   constexpr int kBsBlock = G > 1 ? SQB_G2_BLOCK : sqb::kBsBlock;   // (shadows the namespace constant)
   constexpr int B = WM ? WM : BsState<R, G>::B;          // planes of the distance handed out with an event
   auto value_of = [](const uint32_t *planes, int r) -> uint32_t {
      if constexpr (WM != 0) return bs_value_unary<(WM ? WM : 1)>(planes, r);
      else return bs_value<BsState<R, G>::B>(planes, r);
   };
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int part = lane / NG, gl = lane % NG;
   Smem &sm = reinterpret_cast<Smem *>(dyn)[warp];
   const uint32_t nlines = bs_nslots(a.ctr, a.act, a.max_lines);
   const uint32_t ntiles = (nlines + kBsTileLines - 1) / kBsTileLines;
   // fused tokenise + pack (single-part automata only): the work items are 32 GROUPS each
   // (FUSED is a template parameter: as a run-time branch the two fetch paths cost the tight instances
   // -- 80 registers at 6 CTAs per SM -- a few spills)
   constexpr bool fused = FUSED;
   const uint32_t ngroups_f = fused ? (uint32_t)min(a.ctr[C_NGROUPS], (unsigned long long)a.gdesc_cap) : 0u;
   const uint32_t nitems = fused ? (ngroups_f + 31u) / 32u : ntiles * (uint32_t)G;          // (tile, quarter) pairs

   sm.slots[0][BS_ONES][lane] = ~0u;
   sm.slots[1][BS_ONES][lane] = ~0u;
   const uint32_t *slot_base = &sm.slots[0][0][lane];
   // shared-memory address of the Eq mask of every row of this lane's part: with
   // G > 1 the slot of a row differs between the lanes of a warp, so the addresses
   // live in registers (from a staged copy of the table: a per-lane index into the
   // kernel parameters would serialise in the constant bank)
   uint32_t raddr[(G > 1 && !SQB_G2_RADDR_SMEM) ? R : 1];
   const uint32_t sb = smem_addr(slot_base);
   if (G > 1) {
      for (int i = tid; i < R * G; i += kBsThreads) s_off[i] = pat.slot_off[i];
      __syncthreads();
      if (!SQB_G2_RADDR_SMEM) {
#pragma unroll
         for (int j = 0; j < R; j++) raddr[(G > 1 && !SQB_G2_RADDR_SMEM) ? j : 0] = sb + s_off[part * R + j];
      }
   }

   uint32_t my_matched = 0, my_events = 0;
   const bool has_cuts = a.gmask != nullptr && a.ctr[C_NCUTS] != 0ull;

   for (uint32_t item = blockIdx.x * kBsWarps + warp; item < nitems; item += gridDim.x * kBsWarps) {
      const uint32_t tile = item / (uint32_t)G, q = item % (uint32_t)G;
      const uint32_t group = q * (uint32_t)NG + (uint32_t)gl;           // group of the tile served by this lane
      const uint32_t line0 = tile * kBsTileLines;
      uint32_t mine, ncols, my_ncols = 0u, gent0 = 0u;
      uint32_t slot0 = line0 + group * 32u, ent0;
      uint32_t lead_lo = 0u, lead_hi = 0u;                // fused: NULL columns in front of the lines (two bits each)
      bool consecutive = true;
      const uint4 *col;
      if (fused) {
         // lane = one group of the fused kernel: its descriptor says where the lines and the planes are
         const uint32_t gi = item * 32u + (uint32_t)lane;
         GroupDesc d{0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
         if (gi < ngroups_f) d = a.gdesc[gi];
         mine = d.meta & 0xffu;
         my_ncols = d.meta >> 8;
         gent0 = d.first;
         ent0 = mine ? a.k1_tile_base[d.tile] + (a.gent ? 0u : d.first) : 0u;
         col = a.planes + d.poff;
         lead_lo = d.lead_lo;
         lead_hi = d.lead_hi;
         ncols = __reduce_max_sync(kFull, my_ncols);
      } else {
         const uint32_t left = nlines - line0;                              // > 0
         mine = left > group * 32u ? min(left - group * 32u, 32u) : 0u;
         // slot -> entry of ls.  With the line filter the 32 slots of a lane usually map to
         // consecutive entries when the filter drops nothing around them: one look-up per
         // tile then serves every event of the lane
         ent0 = slot0;
         if (a.act && mine > 0u) {
            ent0 = a.act[slot0];
            consecutive = a.act[slot0 + mine - 1u] - ent0 == mine - 1u;
         }
         ncols = a.tile_cols[tile];
         col = a.planes + (size_t)a.tile_off[tile] * 32u + group;
      }
      State st;
      if constexpr (WM != 0) bs_wm_reset(st, pat, mine == 32u ? ~0u : ((1u << mine) - 1u));
      else bs_reset(st, pat, mine == 32u ? ~0u : ((1u << mine) - 1u), part);
      if (MODE == BS_ALL) {
#pragma unroll
         for (int i = 0; i < 32; i++) static_cast<BsWarpSmemAll &>(sm).cnt[i * 32 + lane] = 0u;
         __syncwarp();            // lane g counts in cnt[g * 32 + r]: words other lanes have just cleared (racecheck, r3a)
      }
      uint32_t lane_events = 0;
      auto entry_of = [&](uint32_t r) -> uint32_t {
         if (fused) return a.gent ? ent0 + (uint32_t)a.gent[gent0 + r] : ent0 + r;
         return consecutive ? ent0 + r : a.act[slot0 + r];
      };
      auto lead_of = [&](int r) -> uint32_t { return ((lead_lo >> r) & 1u) | (((lead_hi >> r) & 1u) << 1); };
      uint32_t qmask = 0u, fmask = 0u;
      if (has_cuts) {
         qmask = a.gmask[tile * 32u + group];
         fmask = a.gfollow[tile * 32u + group];
      }
      const uint32_t niter = ncols + (uint32_t)(G - 1);

      // columns are consumed in blocks of kBsBlock; the next block is in flight while
      // this one is matched (global latency >> one column of work).  Iteration t of
      // part p is column t - p; before the line start that is a NULL column.
      auto fetch = [&](uint32_t t) {
         const int c = (int)t - part;
         uint4 v = col[(size_t)min((uint32_t)max(c, 0), ncols - 1u) * 32u];          // ncols >= 1
         if (G > 1 && c < 0) v = make_uint4(~0u, ~0u, ~0u, 0u);
         return v;
      };
      // fused layout: [block of 32 columns][plane][32 columns] words; block b of four columns of the lane's own
      // group = three uint4 {p0, p1, p2 of 4 columns}, 128 bytes apart; a lane without a group (or past its last
      // block) re-reads a valid block: its lines are dead by then
      const uint32_t my_last = ((my_ncols + 3u) >> 2) - (my_ncols ? 1u : 0u);
      auto fetchb = [&](uint32_t b, uint4 &q0, uint4 &q1, uint4 &q2) {
         const uint32_t bb = min(b, my_last);
         const uint4 *p = col + (size_t)(bb >> 3) * 24u + (bb & 7u);
         q0 = p[0];
         q1 = p[8];
         q2 = p[16];
      };
      uint4 nxt[kBsBlock];
      if (fused) {
         fetchb(0u, nxt[0], nxt[1], nxt[2]);
      } else {
#pragma unroll
         for (int k = 0; k < kBsBlock; k++) nxt[k] = fetch((uint32_t)k);
      }
      uint32_t ph_prev = 0u, mh_prev = 0u;               // what this part handed upwards one iteration ago
#pragma unroll 1
      for (uint32_t t0 = 0; t0 < niter; t0 += kBsBlock) {
         if (!__any_sync(kFull, st.alive != 0u)) break;
         uint32_t P0[kBsBlock], P1[kBsBlock], P2[kBsBlock];
         if constexpr (fused) {
            {
               const uint4 q0 = nxt[0], q1 = nxt[1], q2 = nxt[2];
               P0[0] = q0.x; P0[1] = q0.y; P0[2] = q0.z; P0[3] = q0.w;
               P1[0] = q1.x; P1[1] = q1.y; P1[2] = q1.z; P1[3] = q1.w;
               P2[0] = q2.x; P2[1] = q2.y; P2[2] = q2.z; P2[3] = q2.w;
               fetchb(t0 / (uint32_t)kBsBlock + 1u, nxt[0], nxt[1], nxt[2]);
            }
         } else {
#pragma unroll
            for (int k = 0; k < kBsBlock; k++) {
               P0[k] = nxt[k].x;
               P1[k] = nxt[k].y;
               P2[k] = nxt[k].z;
               nxt[k] = fetch(t0 + kBsBlock + (uint32_t)k);
            }
         }
#pragma unroll
         for (int k = 0; k < kBsBlock; k++) {
         const int kk = k;
         // iterations >= niter: every line is dead, nothing happens
         const uint32_t c = t0 + (uint32_t)k - (uint32_t)part;       // column of this lane (last part: >= 0 while alive)
         const uint32_t p0 = P0[k], p1 = P1[k], p2 = P2[k];
         uint32_t anybase, stop, skip;
         {
            const uint32_t na = ~p2 & ~p1 & ~p0, nc = ~p2 & ~p1 & p0, ng = ~p2 & p1 & ~p0, nt = ~p2 & p1 & p0;
            const uint32_t nn = p2 & ~p1 & ~p0;
            anybase = ~p2 | nn;
            stop = p2 & ~p1 & p0;
            skip = p2 & p1 & ~p0;
            auto &sl = sm.slots[G == 1 ? (kk & 1) : 0];
            sl[BS_A][lane] = na;
            sl[BS_C][lane] = nc;
            sl[BS_G][lane] = ng;
            sl[BS_T][lane] = nt;
            sl[BS_N][lane] = nn;
            sl[BS_ANY][lane] = anybase;
            // bracket classes beyond the single bases: the first two inline, the others in a ROLLED loop (uniform trip
            // count) -- unrolled for all six they made every instance 7 % slower whether a pattern had one or not (r4q:
            // code size), and a rolled loop from the first class on cost the patterns with one class 6 % (r4r)
            if (G > 1) {
               // (multi-part instances: every class through the rolled loop -- the inline code cost cfg4 3.5 %, r5b)
#pragma unroll 1
               for (int q = 0; q < pat.ncustom; q++)
                  sl[BS_CUSTOM0 + q][lane] = (na & pat.custom[q][0]) | (nc & pat.custom[q][1]) | (ng & pat.custom[q][2]) |
                                             (nt & pat.custom[q][3]) | (nn & pat.custom[q][4]);
            } else if (CUSTOM && pat.ncustom > 0) {
               sl[BS_CUSTOM0][lane] = (na & pat.custom[0][0]) | (nc & pat.custom[0][1]) | (ng & pat.custom[0][2]) |
                                      (nt & pat.custom[0][3]) | (nn & pat.custom[0][4]);
               if (pat.ncustom > 1) {
                  sl[BS_CUSTOM1][lane] = (na & pat.custom[1][0]) | (nc & pat.custom[1][1]) | (ng & pat.custom[1][2]) |
                                         (nt & pat.custom[1][3]) | (nn & pat.custom[1][4]);
#pragma unroll 1
                  for (int q = 2; q < pat.ncustom; q++)
                     sl[BS_CUSTOM0 + q][lane] = (na & pat.custom[q][0]) | (nc & pat.custom[q][1]) | (ng & pat.custom[q][2]) |
                                                (nt & pat.custom[q][3]) | (nn & pat.custom[q][4]);
               }
            }
         }
         uint32_t ph = 0u, mh = 0u;
         if (G > 1) {
            ph = __shfl_up_sync(kFull, ph_prev, NG);
            mh = __shfl_up_sync(kFull, mh_prev, NG);
            if (part == 0) ph = mh = 0u;
         }
         auto eq = [&](int j) -> uint32_t {
            if (G > 1) {
               if (SQB_G2_RADDR_SMEM) return lds_u32(sb + s_off[(G > 1 ? part * R + j : 0)]);
               return lds_u32(raddr[(G > 1 && !SQB_G2_RADDR_SMEM) ? j : 0]);
            }
            return *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(slot_base) +
                                                       (kk & 1) * (int)sizeof(sm.slots[0]) + pat.slot_off[j]);
         };
         uint32_t streak[B];
         uint32_t evt;
         if constexpr (WM != 0) {
            evt = bs_wm_step<R, (WM ? WM : 1), MODE, SKIP>(st, eq, anybase, stop, skip, streak, c <= a.wup ? qmask : 0u);
         } else {
            bs_rows<R, G, SKIP>(st, eq, skip, ph, mh);
            ph_prev = ph;
            mh_prev = mh;
            evt = bs_report<R, G, MODE>(st, pat, ph, mh, anybase, stop, streak, c <= a.wup ? qmask : 0u);
         }

         // ---- events leave the bit-sliced world here (rare) ---------------------
         if (MODE == BS_ALL) {
            if (__any_sync(kFull, evt != 0u)) {
               // warp-aggregated room for the events of this column
               const uint32_t ne = (uint32_t)__popc(evt);
               uint32_t inc = ne;
#pragma unroll
               for (int d = 1; d < 32; d <<= 1) {
                  const uint32_t t = __shfl_up_sync(kFull, inc, d);
                  if (lane >= d) inc += t;
               }
               unsigned long long base = 0;
               if (!a.count_only) {
                  if (lane == 31) base = atomicAdd(&a.ctr[C_EVENTS], (unsigned long long)inc);
                  base = __shfl_sync(kFull, base, 31);
               }
               unsigned long long idx = base + inc - ne;
               uint32_t e = evt;
               while (e) {
                  const int r = __ffs(e) - 1;
                  e &= e - 1;
                  const uint32_t line = entry_of((uint32_t)r);
                  const uint32_t rank = static_cast<BsWarpSmemAll &>(sm).cnt[gl * 32 + r]++;
                  if (!a.count_only) {
                     if (idx < a.ev_cap) a.ev[idx] = Event{line, rank, c - lead_of(r), value_of(streak, r)};
                     idx++;
                  }
               }
               lane_events += ne;
            }
         } else if (evt != 0u && !a.count_only) {
            uint32_t e = evt;
            while (e) {
               const int r = __ffs(e) - 1;
               e &= e - 1;
               a.res[entry_of((uint32_t)r)] = ((unsigned long long)value_of(streak, r) << 32) | (c - lead_of(r));
            }
         }
         }
      }
      my_matched += (uint32_t)__popc(st.hit);
      my_events += lane_events;
      for (uint32_t ss = st.stopped & fmask; ss; ss &= ss - 1u) a.segstop[entry_of((uint32_t)(__ffs(ss) - 1))] = 1;
      if (MODE == BS_ALL && !a.count_only) {
         __syncwarp();
         if (fused) {
            // lane L writes the count of slot L of group i: the group's whereabouts come from lane i
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
               const uint32_t e0 = __shfl_sync(kFull, ent0, i), g0 = __shfl_sync(kFull, gent0, i), mi = __shfl_sync(kFull, mine, i);
               if ((uint32_t)lane < mi)
                  a.cnt[a.gent ? e0 + (uint32_t)a.gent[g0 + (uint32_t)lane] : e0 + (uint32_t)lane] =
                     static_cast<BsWarpSmemAll &>(sm).cnt[i * 32 + lane];
            }
         } else {
#pragma unroll 4
            for (int i = 0; i < NG; i++) {
               const uint32_t slot = line0 + (q * (uint32_t)NG + (uint32_t)i) * 32u + (uint32_t)lane;
               if (slot < nlines) a.cnt[a.act ? a.act[slot] : slot] = static_cast<BsWarpSmemAll &>(sm).cnt[i * 32 + lane];
            }
         }
         __syncwarp();
      }
   }

   if (a.count_only) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
         my_matched += __shfl_xor_sync(kFull, my_matched, d);
         my_events += __shfl_xor_sync(kFull, my_events, d);
      }
      if (lane == 0) {
         s_red[0][warp] = my_matched;
         s_red[1][warp] = my_events;
      }
      __syncthreads();
      if (tid == 0) {
         unsigned long long sm_ = 0, se = 0;
         for (int w = 0; w < kBsWarps; w++) {
            sm_ += s_red[0][w];
            se += s_red[1][w];
         }
         if (sm_) atomicAdd(&a.ctr[C_NMATCHED], sm_);
         if (MODE == BS_ALL && se) atomicAdd(&a.ctr[C_NRECS], se);
      }
   }
}

}  // namespace sqb
