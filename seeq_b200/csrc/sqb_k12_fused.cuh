// sqb_k12_fused.cuh -- K1 + bit-plane pack in ONE kernel: text in, line starts and bit-planes out.
//
// The two-kernel path (k1_scan_classify -> class nibbles in HBM -> k15_pack) moves every text byte
// through DRAM three times (text in, nibbles out and in, planes out) and pays the pack's scattered
// 16-byte loads / 32-byte store pieces on the L1 data pipe.  Here a CTA stages one 32 KiB tile of text
// (plus `ov` bytes of overlap so that the tile's last line is complete) with ONE TMA bulk copy and
//
//   1. finds the newlines with byte-SIMD arithmetic on 16-byte vectors (exact zero-byte test of
//      w ^ 0x0A0A0A0A: three logic/add ops per four text bytes, no table), counts them (popc), places
//      them with warp-aggregated prefix sums and writes the tile's ordered list of line starts: into
//      shared memory for step 2 and -- allocated with one atomic per tile -- into ls_raw (as K1);
//   2. forms GROUPS of 32 consecutive lines that start in the tile (tile-local: no CTA waits for
//      another) and builds their bit-planes ALREADY TRANSPOSED: a warp takes a group, LANE = TEXT
//      COLUMN, and walks the 32 lines of the group: one LDS.U8 fetches byte (line r, column c), one
//      table look-up turns it into {p0, p1, p2} spread over three bytes of a word, and one multiply-add
//      drops the three class bits into bit r & 7 of the line-octet's accumulator.  After 32 lines four
//      accumulators hold the three plane words of column c (bit r = line r): 9 PRMT put them together.
//      No transpose, no shuffles, no funnel shifts: 4 instructions per text byte (the three-kernel
//      version before this one: 12), two of them on the shared-memory pipe, which bounds the kernel;
//   3. stores the plane words with fully coalesced 128-byte stores (lane = column).
//
// Plane layout of a group (allocated with one atomic per tile, any order across tiles):
//   [block of 32 columns][plane 0..2][32 columns] words = 384 bytes per 32 columns = 3 bits per text byte
// (the two-kernel path stores {p0,p1,p2,-}: 4 bits).  A group descriptor {K1 tile, first local entry,
// lines | columns << 8, plane offset} tells the matcher what its lane holds; the line NUMBER of slot r
// is tile_base[tile] + first + r, known after k1_scan_tiles like every line number of the scan.
// With the line filter (FASTQ-like input with -x 0) only the lines that are not dead on arrival are
// grouped -- the others are never classified at all; `gent` then holds the local entry index of every slot.
// The columns of a line behind its terminator hold whatever follows in the text: the matcher's lines die
// at their STOP column and ignore the rest.
//
// What the kernel does not handle is detected on the device and sent back to the two-kernel path by
// the host (ctr[C_FUSED_OVF], one re-run, the engine remembers): a line longer than the staged overlap,
// more than kFMaxEntries line starts in a tile.  Not used at all with segment cuts, FASTA
// headers, SQB_FASTQ, multi-part automata and pattern sets (sqb_engine.cu: use_fused).
//
// Replaces, like K1 + pack: the getline loop of seeqFileMatch (/root/reference/src/seeq.c:361-377)
// and the translate step of seeqStringMatch (libseeq.c:250-264).
#pragma once

#include "sqb_k2_bitslice.cuh"

namespace sqb {

constexpr uint32_t kFMaxEntries = 1024;                 // line starts per tile the fused path handles (lines of 32 bytes on average)
constexpr uint32_t kFMaxGroups  = kFMaxEntries / 32;
#ifndef SQB_K12_CTAS
#define SQB_K12_CTAS 4                                  // CTAs per SM: 64 registers, 42 KB of shared memory each
#endif
constexpr uint32_t kFMaxOverlap = 4096;                 // bytes staged behind a tile at most = longest line of the fused path
constexpr uint32_t kFChunks     = kK1Tile / 32;         // 32-byte chunks of a tile (1024): four per thread

// (GroupDesc: sqb_k2_bitslice.cuh)

struct K12Args {
   const uint8_t *text;
   uint32_t n;
   uint32_t *ls_raw;              // out: line starts, tile segments in allocation order (| kDeadBit)
   uint32_t ls_cap;
   unsigned long long *ctr;
   uint32_t *tile_cnt, *tile_off; // out: as K1
   uint32_t *tile_alive;          // out (FILTER): live entries per tile (statistics: C_NACTIVE)
   uint32_t filter_k;
   uint32_t skip;                 // as K1Args::skip
   uint32_t ov;                   // overlap bytes staged behind a tile: multiple of 32, 32 .. kFMaxOverlap
   GroupDesc *gdesc;              // out
   uint32_t gdesc_cap;
   uint16_t *gent;                // out (FILTER): [group * 32 + slot] local entry index
   uint4 *planes;                 // out
   uint32_t planes_cap;           // uint4 units
};

// byte -> {p0, p1, p2, newline} in the four bytes of a word (bit 0 of each)
struct ClassTable32 {
   uint32_t w[256];
};
static inline void build_class_table32(const ClassTable &ct, ClassTable32 *out)
{
   for (int b = 0; b < 256; b++) {
      const uint32_t c = ct.code[b];
      out->w[b] = (c & 1u) | (((c >> 1) & 1u) << 8) | (((c >> 2) & 1u) << 16) | (((c >> 3) & 1u) << 24);
   }
}

// dynamic shared memory: the text stage (+ 32 columns read past the longest line, + the STOP byte behind the
// buffer), the class table, the tile's list of line starts, the live entries, the line starts of every warp's group
__host__ __device__ constexpr uint32_t k12_text_bytes(uint32_t ov) { return (kK1Tile + ov + 64u + 127u) & ~127u; }
__host__ __device__ constexpr uint32_t k12_smem_bytes(uint32_t ov)
{
   return k12_text_bytes(ov) + 1024u + (kFMaxEntries + 4u) * 4u + kFMaxEntries * 2u + (uint32_t)kWarps * 32u * 4u;
}
// plane units (uint4) of a group of `ncols` columns
__host__ __device__ constexpr uint32_t k12_group_units(uint32_t ncols) { return ((ncols + 31u) >> 5) * 24u; }

// Newline flags of four text bytes: bit 7 of every byte that is '\n' (exact: no carry crosses a byte).
__device__ __forceinline__ uint32_t nl_flags(uint32_t w)
{
   const uint32_t t = ((w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;     // bit 7: the low seven bits differ from '\n'
   return ~(t | w) & 0x80808080u;                                         // ... and bit 7 of the byte itself is clear
}

// The flags of a 32-byte chunk live in ONE word: the flags of text word k (0..7) rotated left by k, so byte b of
// word k sits at bit (8 b + 7 + k) & 31.  In "u-space" (the word rotated right by 7) that is bit u = 8 b + k.
// u-space mask of the bytes in front of byte index v = 4 k + b (v = 0..32) in TEXT order:
__device__ __forceinline__ uint32_t chunk_before(uint32_t v)
{
   if (v >= 32u) return ~0u;
   const uint32_t k = v >> 2, b = v & 3u;
   return (((1u << k) - 1u) * 0x01010101u) | ((0x01010101u << k) & ((1u << (8u * b)) - 1u));
}

__device__ __forceinline__ uint32_t chunk_flags(const uint8_t *p, uint32_t h)
{
   // the half of the chunk a lane loads first alternates every four lanes: the 16-byte loads of a
   // quarter-warp (eight lanes, 32 bytes apart) then cover all 32 banks once
   const uint4 va = *reinterpret_cast<const uint4 *>(p + 16u * h);
   const uint4 vb = *reinterpret_cast<const uint4 *>(p + 16u * (h ^ 1u));
   const uint32_t ra = 4u * h, rb = 4u * (h ^ 1u);
   uint32_t f, acc;
   f = nl_flags(va.x); acc = __funnelshift_l(f, f, ra);
   f = nl_flags(va.y); acc |= __funnelshift_l(f, f, ra + 1u);
   f = nl_flags(va.z); acc |= __funnelshift_l(f, f, ra + 2u);
   f = nl_flags(va.w); acc |= __funnelshift_l(f, f, ra + 3u);
   f = nl_flags(vb.x); acc |= __funnelshift_l(f, f, rb);
   f = nl_flags(vb.y); acc |= __funnelshift_l(f, f, rb + 1u);
   f = nl_flags(vb.z); acc |= __funnelshift_l(f, f, rb + 2u);
   f = nl_flags(vb.w); acc |= __funnelshift_l(f, f, rb + 3u);
   return __funnelshift_r(acc, acc, 7);          // u-space
}

template <bool FILTER>
__global__ void __launch_bounds__(kThreads, SQB_K12_CTAS) k12_scan_pack(const K12Args a, const __grid_constant__ ClassTable32 ct)
{
   extern __shared__ __align__(128) uint8_t dyn[];
   __shared__ uint64_t bar;
   __shared__ uint32_t s_tile[2], s_base, s_ovnl, s_nlive, s_pbase, s_gbase, s_skip;
   __shared__ uint32_t s_wsum[4][kWarps];                      // line starts per (pass, warp)
   __shared__ uint32_t s_gcols[kFMaxGroups];
   __shared__ uint32_t s_gpre[kFMaxGroups];                    // plane units of the tile's groups in front of each

   const uint32_t n = a.n, ov = a.ov;
   const uint32_t stage = kK1Tile + ov;                         // text bytes staged per tile
   const uint32_t ntiles = (n + kK1Tile - 1) / kK1Tile;
   const uint32_t n16 = (n + 15u) & ~15u;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   uint8_t *buf = dyn;                                          // the text of the tile (TMA)
   uint32_t *lut = reinterpret_cast<uint32_t *>(dyn + k12_text_bytes(ov));   // [256]
   uint32_t *lst = lut + 256;                                   // [entries + 1] offset in the stage | kDeadBit
   uint16_t *live = reinterpret_cast<uint16_t *>(lst + kFMaxEntries + 4u);   // FILTER: entries that are alive
   uint32_t *gl = reinterpret_cast<uint32_t *>(live + kFMaxEntries) + warp * 32;   // line starts of this warp's group
   const uint32_t half = ((uint32_t)lane >> 2) & 1u;
   const bool has_ov = (uint32_t)tid * 32u < ov;

   auto issue = [&](uint32_t t) {          // (tid 0) TMA of tile t into the text stage
      const uint32_t start = t * kK1Tile;
      uint32_t bytes = n16 - start;
      if (bytes > stage) bytes = stage;
      mbar_expect_tx(&bar, bytes);
      bulk_g2s(buf, a.text + start, bytes, &bar);
   };
   lut[tid] = ct.w[tid];
   if (tid == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
      const uint32_t t = (uint32_t)atomicAdd(&a.ctr[C_TICKET_K1], 1ull);
      s_tile[0] = t;
      s_ovnl = 0xffffffffu;
      if (t < ntiles) issue(t);
   }
   __syncthreads();
   uint32_t phase = 0;

   // One text stage: the TMA of the next tile is issued when every warp is done with this one (barrier F);
   // the other CTAs of the SM compute meanwhile.  The tile numbers travel through s_tile[iteration parity]:
   // the next one is drawn before barrier A of this iteration.
   for (uint32_t iter = 0;; iter++) {
      const uint32_t tile = s_tile[iter & 1u];
      if (tile >= ntiles) break;
      const uint32_t tile0 = tile * kK1Tile;
      uint32_t next_tile = 0xffffffffu;                        // tid 0
      mbar_wait(&bar, phase);
      phase ^= 1u;

      if (tid == 0) {
         next_tile = (uint32_t)atomicAdd(&a.ctr[C_TICKET_K1], 1ull);
         s_tile[(iter + 1u) & 1u] = next_tile;
         // the bytes in front of the buffer are not looked at (this thread scans them itself, below)
         if (tile == 0u)
            for (uint32_t i = 0; i < a.skip; i++) buf[i] = 'A';
      }

      // ---- line scan: chunk tid of the four 8 KiB passes; the overlap: one chunk per thread ----
      uint32_t c[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
         const uint32_t o = (uint32_t)i * (kK1Tile / 4u) + (uint32_t)tid * 32u;
         c[i] = chunk_flags(buf + o, half);
         // a newline at p opens a line at p + 1 only if p + 1 < n
         const uint32_t p = tile0 + o;
         if (p + 33u > n) c[i] &= chunk_before(p + 1u >= n ? 0u : n - 1u - p);
      }
      if (has_ov) {
         // only the first newline (= end of the tile's last line) is of interest; the last byte of the buffer counts
         const uint32_t o = kK1Tile + (uint32_t)tid * 32u;
         uint32_t x = chunk_flags(buf + o, half);
         const uint32_t p = tile0 + o;
         if (p + 32u > n) x &= chunk_before(p >= n ? 0u : n - p);
         uint32_t best = 32u;
         while (x) {
            const uint32_t u = (uint32_t)__ffs(x) - 1u;
            x &= x - 1u;
            best = min(best, 4u * (u & 7u) + (u >> 3));
         }
         if (best < 32u) atomicMin(&s_ovnl, o + best);
      }
      const uint32_t first = (tile == 0u && tid == 0 && n > a.skip) ? 1u : 0u;     // the line at the start of the buffer
      const uint32_t c0n = (uint32_t)__popc(c[0]) + first, c1n = (uint32_t)__popc(c[1]);
      const uint32_t c2n = (uint32_t)__popc(c[2]), c3n = (uint32_t)__popc(c[3]);
      // two packed inclusive warp scans (a lane has at most 33 starts per pass: 16 bits hold a warp's sum)
      uint32_t x01 = c0n | (c1n << 16), x23 = c2n | (c3n << 16);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         const uint32_t t01 = __shfl_up_sync(kFull, x01, d), t23 = __shfl_up_sync(kFull, x23, d);
         if (lane >= d) {
            x01 += t01;
            x23 += t23;
         }
      }
      if (lane == 31) {
         s_wsum[0][warp] = x01 & 0xffffu;
         s_wsum[1][warp] = x01 >> 16;
         s_wsum[2][warp] = x23 & 0xffffu;
         s_wsum[3][warp] = x23 >> 16;
      }
      __syncthreads();                       // A: s_wsum, s_ovnl, the STOP byte
      uint32_t base[4], tile_total = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) {
         uint32_t before = 0, tot = 0;
#pragma unroll
         for (int w2 = 0; w2 < kWarps; w2++) {
            const uint32_t x = s_wsum[i][w2];
            if (w2 < warp) before += x;
            tot += x;
         }
         base[i] = tile_total + before;
         tile_total += tot;
      }
      base[0] += (x01 & 0xffffu) - c0n;
      base[1] += (x01 >> 16) - c1n;
      base[2] += (x23 & 0xffffu) - c2n;
      base[3] += (x23 >> 16) - c3n;
      if (tid == 0) {
         const uint32_t ovnl = s_ovnl;
         s_ovnl = 0xffffffffu;
         // one STOP byte behind the end of the buffer closes a last line without a newline (every scanning
         // thread is past the stage; the filter and the classification read it after barrier B)
         if (n - tile0 < stage + 32u) buf[n - tile0] = 0;
         const uint32_t at = (uint32_t)atomicAdd(&a.ctr[C_LS_CURSOR], (unsigned long long)tile_total);
         a.tile_cnt[tile] = tile_total;
         a.tile_off[tile] = at;
         s_base = at;
         // where the tile's last line ends (exclusive, with its terminator): the first newline of the
         // overlap, or one STOP column behind the end of the buffer
         uint32_t end = 0xffffffffu;
         if (ovnl != 0xffffffffu) end = ovnl + 1u;
         if (tile0 + stage > n) end = min(end, n - tile0 + 1u);
         if (tile_total > kFMaxEntries || (tile_total != 0u && end == 0xffffffffu)) {
            s_skip = 1u;
            atomicMax(&a.ctr[C_FUSED_OVF], 1ull);
         } else {
            s_skip = 0u;
            lst[tile_total] = end;
         }
      }
      // ---- the tile's list of line starts, in text order ----
      if (first) {
         lst[0] = a.skip;
         base[0] += 1u;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
         const uint32_t o = (uint32_t)i * (kK1Tile / 4u) + (uint32_t)tid * 32u + 1u;
         uint32_t x = c[i];
         while (x) {
            const uint32_t u = (uint32_t)__ffs(x) - 1u;
            x &= x - 1u;
            const uint32_t v = 4u * (u & 7u) + (u >> 3);                       // byte index of the newline in its chunk
            const uint32_t idx = base[i] + (uint32_t)__popc(c[i] & chunk_before(v));
            if (idx < kFMaxEntries) lst[idx] = o + v;
         }
      }
      __syncthreads();                       // B: the list, s_base, s_skip
      const bool skip_tile = s_skip != 0u;   // (uniform) the host repeats the scan on the two-kernel path
      uint32_t ngroups = 0, nlive = 0;
      if (!skip_tile) {
         // ---- ls_raw (coalesced); FILTER: a STOP among the first filter_k class codes of a line kills it;
         //      no filter: the columns of every group = its longest line with the terminator ----
         const uint32_t gbase_ls = s_base;
         for (uint32_t j0 = (uint32_t)warp * 32u; j0 < tile_total; j0 += (uint32_t)kThreads) {
            const uint32_t j = j0 + (uint32_t)lane;
            uint32_t len = 0;
            if (j < tile_total) {
               const uint32_t o = lst[j];
               uint32_t fl = 0u;
               if (FILTER) {
                  for (uint32_t i = 0; i < a.filter_k; i++) {                  // filter_k <= 8
                     const uint32_t e = lut[buf[o + i]];
                     if ((e & 0x00010101u) == 0x00010001u) {                   // STOP = 101
                        fl = kDeadBit;
                        break;
                     }
                  }
                  if (fl) live[j] = 1;                                         // (flags for the compaction below)
                  else live[j] = 0;
               } else {
                  len = lst[j + 1u] - o;
               }
               if (gbase_ls + j < a.ls_cap) a.ls_raw[gbase_ls + j] = (tile0 + o) | fl;
            }
            if (!FILTER) {
               len = __reduce_max_sync(kFull, len);
               if (lane == 0) s_gcols[j0 >> 5] = len;
            }
         }
         nlive = tile_total;
      }
      __syncthreads();                       // C: (no filter) the group columns; (filter) the dead flags
      if (FILTER) {
         // the entries that are alive, in order (one warp; a few rounds).  The flags sit in live[] itself:
         // entry i is read in round i / 32 and the compacted list never overtakes the reader
         if (!skip_tile && warp == 0) {
            uint32_t run = 0;
            for (uint32_t i0 = 0; i0 < tile_total; i0 += 32) {
               const uint32_t i = i0 + (uint32_t)lane;
               const bool lv = i < tile_total && live[i] == 0;
               const uint32_t bal = __ballot_sync(kFull, lv);
               __syncwarp();
               if (lv) live[run + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
               __syncwarp();
               run += (uint32_t)__popc(bal);
            }
            if (lane == 0) {
               s_nlive = run;
               a.tile_alive[tile] = run;
            }
         }
         if (skip_tile && tid == 0) a.tile_alive[tile] = 0u;
         __syncthreads();                    // C2
         if (!skip_tile) {
            nlive = s_nlive;
            const uint32_t ng = (nlive + 31u) >> 5;
            for (uint32_t g = (uint32_t)warp; g < ng; g += kWarps) {
               const uint32_t j = g * 32u + (uint32_t)lane;
               uint32_t len = 0;
               if (j < nlive) {
                  const uint32_t i = (uint32_t)live[j];
                  len = lst[i + 1u] - lst[i];
               }
               len = __reduce_max_sync(kFull, len);
               if (lane == 0) s_gcols[g] = len;
            }
         }
         __syncthreads();                    // D
      }
      ngroups = (nlive + 31u) >> 5;
      // plane units (uint4) of the groups in front of every group of the tile, the tile's room in the plane
      // buffer and its group numbers: one warp
      if (!skip_tile && warp == 0) {
         uint32_t tot = 0, longest = 0;
         for (uint32_t g0 = 0; g0 < ngroups; g0 += 32) {
            const uint32_t g = g0 + (uint32_t)lane;
            const uint32_t cols = g < ngroups ? s_gcols[g] : 0u;
            const uint32_t u = k12_group_units(cols);
            uint32_t x = u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
               const uint32_t t = __shfl_up_sync(kFull, x, d);
               if (lane >= d) x += t;
            }
            if (g < ngroups) s_gpre[g] = tot + x - u;
            tot += __shfl_sync(kFull, x, 31);
            longest = max(longest, __reduce_max_sync(kFull, cols));
         }
         if (lane == 0) {
            if (longest > ov) {
               // a line longer than the overlap: its group would read beyond the staged text
               s_skip = 1u;
               atomicMax(&a.ctr[C_FUSED_OVF], 1ull);
            } else {
               const uint32_t pb = (uint32_t)min(atomicAdd(&a.ctr[C_PLANE_UNITS], (unsigned long long)tot), 0xffffffffull);
               const uint32_t gb = (uint32_t)atomicAdd(&a.ctr[C_NGROUPS], (unsigned long long)ngroups);
               s_pbase = pb;
               s_gbase = gb;
               // beyond a capacity nothing is stored; the counters go on counting and the host repeats the scan
               if ((unsigned long long)pb + tot > a.planes_cap || gb + ngroups > a.gdesc_cap) s_skip = 1u;
            }
         }
      }
      __syncthreads();                       // E
      if (!skip_tile && s_skip == 0u) {
         const uint32_t pbase = s_pbase, gbase = s_gbase;
         // ---- classify + transpose: one group per warp and round, lane = column ----
         for (uint32_t g = (uint32_t)warp; g < ngroups; g += kWarps) {
            const uint32_t ncols = s_gcols[g];
            const uint32_t j = g * 32u + (uint32_t)lane;
            const bool have = j < nlive;
            const uint32_t ent = have ? (FILTER ? (uint32_t)live[j] : j) : 0u;
            const uint32_t begin = lst[have ? ent : (FILTER ? (uint32_t)live[g * 32u] : g * 32u)];     // a lane without a line re-reads slot 0
            if (lane == 0) {
               GroupDesc d;
               d.tile = tile;
               d.first = FILTER ? (gbase + g) * 32u : g * 32u;
               d.meta = min(nlive - g * 32u, 32u) | (ncols << 8);
               d.poff = pbase + s_gpre[g];
               a.gdesc[gbase + g] = d;
            }
            if (FILTER) a.gent[(size_t)(gbase + g) * 32u + (uint32_t)lane] = (uint16_t)ent;
            __syncwarp();
            gl[lane] = begin;
            __syncwarp();
            uint32_t L[32];                                          // the 32 line starts (the same in every lane)
#pragma unroll
            for (int q = 0; q < 8; q++) {
               const uint4 v = reinterpret_cast<const uint4 *>(gl)[q];
               L[4 * q] = v.x;
               L[4 * q + 1] = v.y;
               L[4 * q + 2] = v.z;
               L[4 * q + 3] = v.w;
            }
            uint32_t *out = reinterpret_cast<uint32_t *>(a.planes + (size_t)pbase + s_gpre[g]) + lane;
            const uint8_t *col = buf + lane;
            for (uint32_t c0 = 0; c0 < ncols; c0 += 32) {
               uint32_t acc[4] = {0u, 0u, 0u, 0u};                   // [- | p2 | p1 | p0] bytes of the lines 8q .. 8q+7
#pragma unroll
               for (int r = 0; r < 32; r++) {
                  const uint32_t e = lut[col[L[r]]];
                  acc[r >> 3] = mad_u32(e, 1u << (r & 7), acc[r >> 3]);
               }
               const uint32_t t0 = __byte_perm(acc[0], acc[1], 0x5140u), t1 = __byte_perm(acc[2], acc[3], 0x5140u);
               const uint32_t t2 = __byte_perm(acc[0], acc[1], 0x7362u), t3 = __byte_perm(acc[2], acc[3], 0x7362u);
               out[0] = __byte_perm(t0, t1, 0x5410u);
               out[32] = __byte_perm(t0, t1, 0x7632u);
               out[64] = __byte_perm(t2, t3, 0x5410u);
               out += 96;
               col += 32;
            }
         }
      }
      fence_proxy_async();                   // this thread's reads of the text stage come before its refill by TMA
      __syncthreads();                       // F: the stage, the lists and the group tables are free
      if (tid == 0 && next_tile < ntiles) issue(next_tile);
   }
}

}  // namespace sqb
