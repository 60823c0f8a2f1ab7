// sqb_k2_bitslice.cuh -- K2, line-bit-sliced forward matcher (patterns <= 32
// positions over many short lines; the integer-pipe-lean path).
//
// A warp owns a tile of 1024 consecutive lines; in the automaton phase LANE g
// owns the 32 lines [g*32, g*32+32) of the tile, one bit each (sqb_bitslice.h).
// The text is consumed as the class nibbles written by K1 (`codes`, 4 bits per
// text byte).  Work alternates between two phases of 16 text columns:
//
//   pack   32 rounds; in round i lane L loads the 8 code bytes (16 columns) of
//          line i*32+L at its 16-byte-aligned position, the warp transposes
//          the 32x32 bit matrix with a 5-stage shuffle butterfly (twice: two
//          words), and lane j stores "bit j of all 32 lines" = plane (j&3) of
//          column (j>>2) of group i to shared memory;
//   match  lane g reads the three class planes of its group column by column
//          (conflict-free: row stride 33), derives the Eq slots, and runs one
//          bs_step per column = 6 LOP3 per pattern row for 32 text bytes.
//
// Lines do not start 16-byte aligned: the nibbles in front of a line start are
// forced to the NULL class (no pattern row matches it), which leaves the initial
// automaton untouched, and the reported end is corrected by (begin & 15).
// Events (rare) leave the bit-sliced world through a per-lane slow path.
#pragma once

#include <type_traits>

#include "sqb_bitslice.h"
#include "sqb_kernels.cuh"

namespace sqb {

constexpr int kBsWarps   = 4;                       // warps per CTA
constexpr int kBsThreads = kBsWarps * 32;
constexpr int kBsCols    = 16;                      // text columns per phase
constexpr int kBsStride  = 33;                      // padded row of the plane buffer (words)
constexpr int kBsTileLines = 1024;                  // lines per warp tile

struct K2BsArgs {
   const uint2 *codes;            // 16 class nibbles per 16 text bytes
   uint32_t ncode8;               // entries of codes[]
   uint32_t n;                    // text bytes
   const uint32_t *ls;
   uint32_t max_lines;
   unsigned long long *ctr;
   unsigned long long *res;       // BS_FIRST / BS_BEST: per-line (dist << 32 | end); preset to kNoMatch
   uint32_t *cnt;                 // BS_ALL: per-line event count
   Event *ev;                     // BS_ALL: unordered events
   uint32_t ev_cap;
   int count_only;                // counts only: no res / event stores
   BsGate bs;
};

struct BsWarpSmem {
   uint32_t planes[2][32 * kBsStride];
   uint32_t slots[BS_SLOTS][32];
};
struct BsWarpSmemAll : BsWarpSmem {
   uint32_t cnt[32 * 32];         // events per line of the tile (BS_ALL)
};

// 32x32 bit transpose across the warp: on return lane j holds, in bit i, bit j
// of lane i's input.  keep[s] / rot[s] are per-lane constants of stage s.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, const uint32_t (&keep)[5], const uint32_t (&rot)[5])
{
#pragma unroll
   for (int s = 0; s < 5; s++) {
      const uint32_t y = __shfl_xor_sync(kFull, x, 16 >> s);
      x = (x & keep[s]) | (__funnelshift_l(y, y, rot[s]) & ~keep[s]);
   }
   return x;
}

template <int R, int MODE, bool SKIP>
__global__ void __launch_bounds__(kBsThreads, R <= 16 ? 4 : 3)
k2_bitslice(const K2BsArgs a, const __grid_constant__ BsPattern pat)
{
   using Smem = typename std::conditional<MODE == BS_ALL, BsWarpSmemAll, BsWarpSmem>::type;
   extern __shared__ __align__(16) uint8_t dyn[];
   __shared__ uint32_t s_red[2][kBsWarps];

   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   Smem &sm = reinterpret_cast<Smem *>(dyn)[warp];

   const unsigned long long nl_dev = a.ctr[C_NLINES];
   if (!bs_selected(a.bs, nl_dev, a.n)) return;                 // the word-parallel kernel takes this scan
   const uint32_t nlines = (uint32_t)min(nl_dev, (unsigned long long)a.max_lines);
   const uint32_t ntiles = (nlines + kBsTileLines - 1) / kBsTileLines;

   // per-lane constants of the transpose butterfly
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
   sm.slots[BS_ONES][lane] = ~0u;
   const uint32_t *slot_base = &sm.slots[0][lane];
   constexpr int B = BsState<R>::B;

   uint32_t my_matched = 0, my_events = 0;

   for (uint32_t tile = blockIdx.x * kBsWarps + warp; tile < ntiles; tile += gridDim.x * kBsWarps) {
      const uint32_t line0 = tile * kBsTileLines;
      const uint32_t left = nlines - line0;                       // > 0
      const uint32_t mine = left > (uint32_t)lane * 32u ? min(left - (uint32_t)lane * 32u, 32u) : 0u;
      BsState<R> st;
      bs_reset(st, pat, mine == 32u ? ~0u : ((1u << mine) - 1u));
      if (MODE == BS_ALL) {
#pragma unroll
         for (int i = 0; i < 32; i++) static_cast<BsWarpSmemAll &>(sm).cnt[i * 32 + lane] = 0u;
      }
      uint32_t lane_events = 0;

      for (uint32_t col0 = 0; __any_sync(kFull, st.alive != 0u); col0 += kBsCols) {
         // ---- pack: 32 rounds, lane = line of the round --------------------------
         __syncwarp();
#pragma unroll 4
         for (int i = 0; i < 32; i++) {
            const uint32_t line = line0 + (uint32_t)i * 32u + (uint32_t)lane;
            uint2 v = make_uint2(0x55555555u, 0x55555555u);            // STOP
            if (line < nlines) {
               const uint32_t begin = a.ls[line];
               const uint32_t at = min((begin >> 4) + (col0 >> 4), a.ncode8 - 1u);
               v = a.codes[at];
               if (col0 == 0) {                   // NULL class in front of the line start
                  const uint32_t o = begin & 15u;
                  const uint32_t lo = o >= 8u ? ~0u : ((1u << (4u * o)) - 1u);
                  const uint32_t hi = o > 8u ? ((1u << (4u * (o - 8u))) - 1u) : 0u;
                  v.x |= 0x77777777u & lo;
                  v.y |= 0x77777777u & hi;
               }
            }
            sm.planes[0][i * kBsStride + lane] = warp_transpose32(v.x, keep, rot);
            sm.planes[1][i * kBsStride + lane] = warp_transpose32(v.y, keep, rot);
         }
         __syncwarp();

         // ---- match: lane = group of 32 lines -------------------------------------
#pragma unroll 1
         for (int c = 0; c < kBsCols; c++) {
            if (c && !__any_sync(kFull, st.alive != 0u)) break;
            const uint32_t *pl = &sm.planes[c >> 3][lane * kBsStride + 4 * (c & 7)];
            const uint32_t p0 = pl[0], p1 = pl[1], p2 = pl[2];
            uint32_t anybase, stop, skip;
            {
               const uint32_t na = ~p2 & ~p1 & ~p0, nc = ~p2 & ~p1 & p0, ng = ~p2 & p1 & ~p0, nt = ~p2 & p1 & p0;
               const uint32_t nn = p2 & ~p1 & ~p0;
               anybase = ~p2 | nn;
               stop = p2 & ~p1 & p0;
               skip = p2 & p1 & ~p0;
               sm.slots[BS_A][lane] = na;
               sm.slots[BS_C][lane] = nc;
               sm.slots[BS_G][lane] = ng;
               sm.slots[BS_T][lane] = nt;
               sm.slots[BS_N][lane] = nn;
               sm.slots[BS_ANY][lane] = anybase;
               if (pat.ncustom > 0)
                  sm.slots[BS_CUSTOM0][lane] = (na & pat.custom[0][0]) | (nc & pat.custom[0][1]) |
                                               (ng & pat.custom[0][2]) | (nt & pat.custom[0][3]) |
                                               (nn & pat.custom[0][4]);
               if (pat.ncustom > 1)
                  sm.slots[BS_CUSTOM1][lane] = (na & pat.custom[1][0]) | (nc & pat.custom[1][1]) |
                                               (ng & pat.custom[1][2]) | (nt & pat.custom[1][3]) |
                                               (nn & pat.custom[1][4]);
            }
            uint32_t streak[B];
            auto eq = [&](int j) -> uint32_t {
               return *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(slot_base) + pat.slot_off[j]);
            };
            const uint32_t evt = bs_step<R, MODE, SKIP>(st, pat, eq, anybase, stop, skip, streak);

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
                     const uint32_t line = line0 + (uint32_t)lane * 32u + (uint32_t)r;
                     const uint32_t rank = static_cast<BsWarpSmemAll &>(sm).cnt[lane * 32 + r]++;
                     if (!a.count_only) {
                        const uint32_t end = col0 + (uint32_t)c - (a.ls[line] & 15u);
                        if (idx < a.ev_cap) a.ev[idx] = Event{line, rank, end, bs_value<B>(streak, r)};
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
                  const uint32_t line = line0 + (uint32_t)lane * 32u + (uint32_t)r;
                  const uint32_t end = col0 + (uint32_t)c - (a.ls[line] & 15u);
                  a.res[line] = ((unsigned long long)bs_value<B>(streak, r) << 32) | end;
               }
            }
         }
      }
      my_matched += (uint32_t)__popc(st.hit);
      my_events += lane_events;
      if (MODE == BS_ALL && !a.count_only) {
         __syncwarp();
#pragma unroll 4
         for (int i = 0; i < 32; i++) {
            const uint32_t line = line0 + (uint32_t)i * 32u + (uint32_t)lane;
            if (line < nlines) a.cnt[line] = static_cast<BsWarpSmemAll &>(sm).cnt[i * 32 + lane];
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
