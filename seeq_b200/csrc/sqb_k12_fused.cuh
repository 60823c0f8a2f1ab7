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
//      another) and builds their bit-planes ALREADY TRANSPOSED.  A block of 32 lines x 32 columns is
//      the work of eight lanes: a lane takes four neighbouring columns and walks the 32 lines; per line ONE
//      aligned 32-bit load fetches its four text bytes, every byte goes through the 256-entry class table
//      in shared memory, whose entry holds {p0, p1, p2} in three bytes of a word, and one multiply-add drops
//      the three class bits of (line r, column c) into bit r & 7 of the column's accumulator of that line
//      octet.  After 32 lines seven PRMT per column put the plane words together (bit r = line r).  No
//      transpose across lanes, no shuffles, no funnel shifts.  Per four text bytes: 5.5 accesses to shared
//      memory, 9 multiply-adds (FMA pipe), 5 logic ops -- the three pipes are loaded about evenly.  (A lane per
//      column with one LDS.U8 per byte: twice the shared-memory accesses, bound by that pipe at 0.8 of its
//      peak.  Classifying four bytes at a time in registers -- the low three bits of a byte are a perfect hash
//      of the alphabet, PRMT is a 4-way look-up in an 8-byte register table -- needs no table in memory but
//      15 logic ops per four bytes on the half-rate ALU pipe: measured slower.)
//   3. stores the plane words as 16-byte pieces, 128 contiguous bytes per plane and block.
//
// Lines start at any byte, 32-bit loads do not: a line is read from the aligned word at or in front of its
// start, and its planes begin with lead = start & 3 NULL columns (class 7: all three planes set, by OR-ing the
// group's lead masks into the first columns).  A NULL column leaves an automaton that is still in its reset
// state untouched (sqb_bitslice.h), so the matcher only has to subtract the lead from the end columns of
// the events it reports (the two bits of every line's lead travel in the group descriptor).
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

#include "sqb_k12_arith.h"
#include "sqb_k2_bitslice.cuh"

namespace sqb {

constexpr uint32_t kFMaxEntries = 1024;                 // line starts per tile the fused path handles (lines of 32 bytes on average)
constexpr uint32_t kFMaxGroups  = kFMaxEntries / 32;
#ifndef SQB_K12_THREADS
#define SQB_K12_THREADS 256                             // threads per CTA: 256 (four 8 KiB scan passes) or 512 (two 16 KiB passes)
#endif
#ifndef SQB_K12_STAGES
#define SQB_K12_STAGES 1                                // text stages: 1, or 2 (the next tile's copy runs under this tile)
#endif
// CTAs per SM.  The kernel is a chain of short phases between CTA barriers, so what keeps the SM busy is the NUMBER of
// independent CTAs, not the copy / compute overlap inside one (measured, cfg2, r4i): two stages, 3 CTAs (80 registers,
// 72 KB) 0.70 ms; one stage, 4 CTAs (64 registers) 0.67 ms; one stage, 5 CTAs (48 registers, 39 KB) 0.64 ms; 512
// threads, two stages, 2 CTAs 0.90 ms.
#ifndef SQB_K12_CTAS
#define SQB_K12_CTAS (SQB_K12_STAGES == 1 ? 5 : (SQB_K12_THREADS == 256 ? 3 : 2))
#endif
constexpr uint32_t kFStages = SQB_K12_STAGES;
constexpr int kFThreads = SQB_K12_THREADS, kFWarps = kFThreads / 32;
constexpr int kFPasses = (int)(kK1Tile / 32u) / kFThreads;          // 32-byte chunks per thread and tile
static_assert(kFPasses * kFWarps == 32 && (kFPasses == 2 || kFPasses == 4), "one (pass, warp) sum per lane");
constexpr uint32_t kFMaxOverlap = 4096;                 // bytes staged behind a tile at most = longest line of the fused path

// (GroupDesc: sqb_k2_bitslice.cuh)

// (ClassTable32, nl_flags, chunk_before, the plane assembly: sqb_k12_arith.h -- host-compilable, pinned on the CPU)

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
   uint32_t four;                 // == 4, as a run-time value: table address = byte * four + base stays an IMAD (FMA
                                  // pipe); with the literal the assembler makes it an LEA on the ALU pipe, which
                                  // carries the byte extraction
};

// dynamic shared memory: two text stages (+ 35 columns read past the longest line, + the STOP byte behind the
// buffer), the class table, the tile's line starts (all of them; those of the grouped lines, padded to whole groups),
// the live entries
__host__ __device__ constexpr uint32_t k12_text_bytes(uint32_t ov) { return (kK1Tile + ov + 64u + 127u) & ~127u; }
__host__ __device__ constexpr uint32_t k12_smem_bytes(uint32_t ov, bool filter)
{
   return kFStages * k12_text_bytes(ov) + 1024u + (kFMaxEntries + 8u) * 2u + (kFMaxEntries + 32u) * 2u + (filter ? kFMaxEntries * 2u : 0u);
}

// (not volatile: the class table is constant for the life of the kernel, the loads may be scheduled freely)
__device__ __forceinline__ uint32_t lds_u32c(uint32_t addr)
{
   uint32_t v;
   asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
   return v;
}

// hi32(a * b) + c on the FMA pipe: a >> s for b = 2^(32 - s), added to c
__device__ __forceinline__ uint32_t mad_hi_u32(uint32_t a, uint32_t b, uint32_t c)
{
   uint32_t d;
   asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
   return d;
}

// the flag word of a 32-byte chunk (u-space, sqb_k12_arith.h): one multiply-add (high half) per text word
__device__ __forceinline__ uint32_t chunk_flags(const uint8_t *p)
{
   const uint4 va = *reinterpret_cast<const uint4 *>(p);
   const uint4 vb = *reinterpret_cast<const uint4 *>(p + 16);
   uint32_t acc = nl_flags(vb.w);                                  // k = 7: in place
   acc = mad_hi_u32(nl_flags(va.x), 1u << 25, acc);
   acc = mad_hi_u32(nl_flags(va.y), 1u << 26, acc);
   acc = mad_hi_u32(nl_flags(va.z), 1u << 27, acc);
   acc = mad_hi_u32(nl_flags(va.w), 1u << 28, acc);
   acc = mad_hi_u32(nl_flags(vb.x), 1u << 29, acc);
   acc = mad_hi_u32(nl_flags(vb.y), 1u << 30, acc);
   acc = mad_hi_u32(nl_flags(vb.z), 1u << 31, acc);
   return acc;
}

template <bool FILTER>
__global__ void __launch_bounds__(kFThreads, SQB_K12_CTAS) k12_scan_pack(const K12Args a, const __grid_constant__ ClassTable32 ct)
{
   extern __shared__ __align__(128) uint8_t dyn[];
   __shared__ uint64_t bar[2];
   __shared__ uint64_t bar_alloc;                              // the tile's plane / group allocation has been published
   __shared__ uint32_t s_tile[2], s_base, s_ovnl, s_pbase, s_gbase, s_skip, s_nlive, s_stores;
   __shared__ uint32_t s_wsum[32];                             // line starts per (pass, warp)
   __shared__ uint32_t s_gcols[kFMaxGroups];
   __shared__ uint32_t s_glong[kFMaxGroups];                   // longest line of every group
   __shared__ uint32_t s_glead[kFMaxGroups][3];                // lines of the group whose lead is >= 1, >= 2, == 3

   const uint32_t n = a.n, ov = a.ov;
   const uint32_t stage = kK1Tile + ov;                         // text bytes staged per tile
   const uint32_t ntiles = (n + kK1Tile - 1) / kK1Tile;
   const uint32_t n16 = (n + 15u) & ~15u;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const uint32_t tbytes = k12_text_bytes(ov);
   uint32_t *lut = reinterpret_cast<uint32_t *>(dyn + kFStages * tbytes);    // [256]
   uint16_t *lst = reinterpret_cast<uint16_t *>(lut + 256);     // [entries + 1] offset of every line start in the stage
   uint16_t *lsg = lst + kFMaxEntries + 8u;                     // [groups * 32] the same of the grouped lines, padded
   uint16_t *live = lsg + kFMaxEntries + 32u;                   // FILTER: dead flags, then the live entries
   const bool has_ov = (uint32_t)tid * 32u < ov;

   auto issue = [&](uint32_t st, uint32_t t) {   // (tid 0) TMA of tile t into text stage st
      const uint32_t start = t * kK1Tile;
      uint32_t bytes = n16 - start;
      if (bytes > stage) bytes = stage;
      mbar_expect_tx(&bar[st], bytes);
      bulk_g2s(dyn + st * tbytes, a.text + start, bytes, &bar[st]);
   };
   if (tid < 256) lut[tid] = ct.w[tid];
   if (tid == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      mbar_init(&bar_alloc, 1);
      mbar_fence_init();
      const uint32_t t = (uint32_t)atomicAdd(&a.ctr[C_TICKET_K1], 1ull);
      s_tile[0] = t;
      s_ovnl = 0xffffffffu;
      if (t < ntiles) issue(0u, t);
   }
   __syncthreads();
   uint32_t phases = 0;                   // bit st = parity to wait for on stage st
   uint32_t alloc_phase = 0;              // parity to wait for on bar_alloc

   // Two text stages.  Iteration j works on stage j & 1; behind its barrier A every warp is done with the tile
   // before (stage (j + 1) & 1), so thread 0 draws the next tile there and sends its TMA into that stage: the copy
   // runs under the rest of this iteration.  Three CTA barriers per tile (A, B, E); every shared scalar is rewritten
   // between two barriers that all its readers of the round before have passed.
   for (uint32_t iter = 0;; iter++) {
      const uint32_t st = iter & 1u, sb = kFStages == 2u ? st : 0u;          // tile-number slot, text stage
      const uint32_t tile = s_tile[st];
      if (tile >= ntiles) break;
      const uint32_t tile0 = tile * kK1Tile;
      uint8_t *buf = dyn + sb * tbytes;                         // the text of the tile
      mbar_wait(&bar[sb], (phases >> sb) & 1u);
      phases ^= 1u << sb;

      // the next tile's number: drawn now, needed behind barrier A (the atomic's round trip hides under the scan)
      unsigned long long nt64 = 0ull;
      if (tid == 0) nt64 = atomicAdd(&a.ctr[C_TICKET_K1], 1ull);
      // the bytes in front of the buffer are not looked at (thread 0 scans them itself, below)
      if (tid == 0 && tile == 0u)
         for (uint32_t i = 0; i < a.skip; i++) buf[i] = 'A';

      // ---- line scan: chunk tid of every pass; the overlap: one chunk per thread ----
      uint32_t c[kFPasses];
#pragma unroll
      for (int i = 0; i < kFPasses; i++) {
         const uint32_t o = (uint32_t)i * (kK1Tile / (uint32_t)kFPasses) + (uint32_t)tid * 32u;
         c[i] = chunk_flags(buf + o);
         // a newline at p opens a line at p + 1 only if p + 1 < n
         const uint32_t p = tile0 + o;
         if (p + 33u > n) c[i] &= chunk_before(p + 1u >= n ? 0u : n - 1u - p);
      }
      if (has_ov) {
         // only the first newline (= end of the tile's last line) is of interest; the last byte of the buffer counts
         const uint32_t o = kK1Tile + (uint32_t)tid * 32u;
         uint32_t x = chunk_flags(buf + o);
         const uint32_t p = tile0 + o;
         if (p + 32u > n) x &= chunk_before(p >= n ? 0u : n - p);
         uint32_t best = 32u;
         while (x) {
            const uint32_t u = (uint32_t)__ffs(x) - 1u;
            x &= x - 1u;
            best = min(best, chunk_byte_of(u));
         }
         if (best < 32u) atomicMin(&s_ovnl, o + best);
      }
      const uint32_t first = (tile == 0u && tid == 0 && n > a.skip) ? 1u : 0u;     // the line at the start of the buffer
      uint32_t cn[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int i = 0; i < kFPasses; i++) cn[i] = (uint32_t)__popc(c[i]) + (i == 0 ? first : 0u);
      // packed inclusive warp scans, two passes each (a lane has at most 33 starts per pass: 16 bits hold a warp's sum)
      uint32_t x01 = cn[0] | (cn[1] << 16), x23 = cn[2] | (cn[3] << 16);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         const uint32_t t01 = __shfl_up_sync(kFull, x01, d);
         if (lane >= d) x01 += t01;
         if (kFPasses > 2) {
            const uint32_t t23 = __shfl_up_sync(kFull, x23, d);
            if (lane >= d) x23 += t23;
         }
      }
      if (lane == 31) {
         s_wsum[warp] = x01 & 0xffffu;
         s_wsum[kFWarps + warp] = x01 >> 16;
         if (kFPasses > 2) {
            s_wsum[2 * kFWarps + warp] = x23 & 0xffffu;
            s_wsum[3 * kFWarps + warp] = x23 >> 16;
         }
      }
      fence_proxy_async();                   // this thread's reads of the OTHER stage (the tile before) come before its refill
      __syncthreads();                       // A: s_wsum, s_ovnl; every warp is done with the tile before
      // where the line starts of every (pass, warp) go: an exclusive scan of the 32 sums, by every warp for itself
      uint32_t base[kFPasses], tile_total;
      {
         const uint32_t v = s_wsum[lane];
         uint32_t x = v;
#pragma unroll
         for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, x, d);
            if (lane >= d) x += t;
         }
         tile_total = __shfl_sync(kFull, x, 31);
         const uint32_t ex = x - v;
#pragma unroll
         for (int i = 0; i < kFPasses; i++) base[i] = __shfl_sync(kFull, ex, i * kFWarps + warp);
      }
      base[0] += (x01 & 0xffffu) - cn[0];
      base[1] += (x01 >> 16) - cn[1];
      if (kFPasses > 2) {
         base[kFPasses > 2 ? 2 : 0] += (x23 & 0xffffu) - cn[2];
         base[kFPasses > 2 ? 3 : 0] += (x23 >> 16) - cn[3];
      }
      unsigned long long at64 = 0ull;        // (tid 0) the tile's place in ls_raw: used behind the emit loops
      if (tid == 0) {
         // the next tile: its TMA into the stage of the tile before
         const uint32_t nt = (uint32_t)nt64;
         s_tile[st ^ 1u] = nt;
         if (kFStages == 2u && nt < ntiles) issue(st ^ 1u, nt);
         at64 = atomicAdd(&a.ctr[C_LS_CURSOR], (unsigned long long)tile_total);
         const uint32_t ovnl = s_ovnl;
         s_ovnl = 0xffffffffu;
         // one STOP byte behind the end of the buffer closes a last line without a newline (every scanning
         // thread is past the stage; the filter and the classification read it after barrier B)
         if (n - tile0 < stage + 32u) buf[n - tile0] = 0;
         a.tile_cnt[tile] = tile_total;
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
            lst[tile_total] = (uint16_t)end;                  // <= kK1Tile + ov + 1 < 65536
         }
      }
      // ---- the tile's list of line starts, in text order ----
      if (first) {
         lst[0] = (uint16_t)a.skip;
         base[0] += 1u;
      }
#pragma unroll
      for (int i = 0; i < kFPasses; i++) {
         const uint32_t o = (uint32_t)i * (kK1Tile / (uint32_t)kFPasses) + (uint32_t)tid * 32u + 1u;
         uint32_t x = c[i];
         const bool several = (x & (x - 1u)) != 0u;                             // (lines of 32 bytes and more: never)
         while (x) {
            const uint32_t u = (uint32_t)__ffs(x) - 1u;
            x &= x - 1u;
            const uint32_t v = chunk_byte_of(u);                               // byte index of the newline in its chunk
            const uint32_t idx = base[i] + (several ? (uint32_t)__popc(c[i] & chunk_before(v)) : 0u);
            if (idx < kFMaxEntries) lst[idx] = (uint16_t)(o + v);
         }
      }
      if (tid == 0) {
         a.tile_off[tile] = (uint32_t)at64;
         s_base = (uint32_t)at64;
      }
      __syncthreads();                       // B: the list, s_base, s_skip
      const bool skip_tile = s_skip != 0u;   // (uniform) the host repeats the scan on the two-kernel path
      uint32_t nlive = tile_total;
      if (FILTER) {
         // a STOP among the first filter_k class codes of a line kills it; ls_raw (coalesced)
         if (!skip_tile) {
            const uint32_t gbase_ls = s_base;
            for (uint32_t j = (uint32_t)tid; j < tile_total; j += (uint32_t)kFThreads) {
               const uint32_t o = lst[j];
               uint32_t fl = 0u;
               for (uint32_t i = 0; i < a.filter_k; i++) {                     // filter_k <= 8
                  if ((lut[buf[o + i]] & 0x00010101u) == 0x00010001u) {        // STOP = 101
                     fl = kDeadBit;
                     break;
                  }
               }
               live[j] = fl ? 1 : 0;                                           // (flags for the compaction below)
               if (gbase_ls + j < a.ls_cap) a.ls_raw[gbase_ls + j] = (tile0 + o) | fl;
            }
         }
         __syncthreads();                    // C: the dead flags
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
               a.tile_alive[tile] = run;
               s_nlive = run;
            }
         }
         if (skip_tile && tid == 0) a.tile_alive[tile] = 0u;
         __syncthreads();                    // C2: the live entries
         if (!skip_tile) nlive = s_nlive;
      }
      const uint32_t ngroups = (nlive + 31u) >> 5;
      if (!skip_tile) {
         // ---- the tables of the groups, one group per warp and round: its columns (the longest line with the lead in
         //      front and the terminator), its longest line, its lead masks, the aligned starts of its lines ----
         for (uint32_t g = (uint32_t)warp; g < ngroups; g += (uint32_t)kFWarps) {
            const uint32_t j = g * 32u + (uint32_t)lane;
            uint32_t len = 0, lead = 0, begin;
            if (j < nlive) {
               const uint32_t i = FILTER ? (uint32_t)live[j] : j;
               begin = lst[i];
               len = (uint32_t)lst[i + 1u] - begin;
               lead = begin & 3u;
            } else {
               begin = lst[FILTER ? (uint32_t)live[g * 32u] : g * 32u];      // a slot without a line re-reads slot 0
            }
            lsg[j] = (uint16_t)(begin & ~3u);
            const uint32_t mx = __reduce_max_sync(kFull, len + lead), lg = __reduce_max_sync(kFull, len);
            const uint32_t m1 = __ballot_sync(kFull, lead >= 1u), m2 = __ballot_sync(kFull, lead >= 2u), m3 = __ballot_sync(kFull, lead == 3u);
            if (lane == 0) {
               s_gcols[g] = mx;
               s_glong[g] = lg;
               s_glead[g][0] = m1;
               s_glead[g][1] = m2;
               s_glead[g][2] = m3;
            }
         }
         if (!FILTER) {
            const uint32_t gbase_ls = s_base;
            for (uint32_t j = (uint32_t)tid; j < tile_total; j += (uint32_t)kFThreads)
               if (gbase_ls + j < a.ls_cap) a.ls_raw[gbase_ls + j] = tile0 + (uint32_t)lst[j];
         }
      }
      __syncthreads();                       // E: the group tables
      // one text stage: the next tile's copy starts when every warp is done with this one
      auto end_of_tile = [&]() {
         if (kFStages == 1u) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0 && (uint32_t)nt64 < ntiles) issue(0u, (uint32_t)nt64);
         }
      };
      if (skip_tile) {
         end_of_tile();
         continue;
      }
      // the blocks of 32 columns in front of every group (lane = group): every warp for itself
      const uint32_t my_cols = (uint32_t)lane < ngroups ? s_gcols[lane] : 0u;
      const uint32_t my_nb = (my_cols + 31u) >> 5;
      uint32_t my_gblk = my_nb;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         const uint32_t t = __shfl_up_sync(kFull, my_gblk, d);
         if (lane >= d) my_gblk += t;
      }
      const uint32_t nblocks = __shfl_sync(kFull, my_gblk, 31);
      my_gblk = (uint32_t)lane < ngroups ? my_gblk - my_nb : 0xffffffffu;
      if (__reduce_max_sync(kFull, (uint32_t)lane < ngroups ? s_glong[lane] : 0u) > ov) {
         // (uniform) a line longer than the overlap: its group would read beyond the staged text
         if (tid == 0) atomicMax(&a.ctr[C_FUSED_OVF], 1ull);
         end_of_tile();
         continue;
      }
      // The tile's room in the plane buffer and its group numbers: thread 0 asks for them now and publishes them
      // when it needs them itself, in front of its first store -- the round trip of the atomics hides under the
      // classification of the first blocks.  The other warps wait for the mbarrier bar_alloc in front of THEIR first store.
      unsigned long long pb64 = 0ull, gb64 = 0ull;
      if (tid == 0) {
         pb64 = atomicAdd(&a.ctr[C_PLANE_UNITS], (unsigned long long)nblocks * 24ull);
         gb64 = atomicAdd(&a.ctr[C_NGROUPS], (unsigned long long)ngroups);
      }
      bool ready = false;
      uint32_t pbase = 0, gbase = 0;
      bool stores = false;
      auto room = [&]() {                    // (warp-uniform call) pbase / stores are valid on return
         if (ready) return;
         if (warp == 0) {
            uint32_t pb = 0, gb = 0, ok = 0;
            if (lane == 0) {
               pb = (uint32_t)min(pb64, 0xffffffffull);
               gb = (uint32_t)gb64;
               // beyond a capacity nothing is stored; the counters go on counting and the host repeats the scan
               ok = (pb64 + (unsigned long long)nblocks * 24ull <= a.planes_cap && gb64 + ngroups <= a.gdesc_cap) ? 1u : 0u;
            }
            pb = __shfl_sync(kFull, pb, 0);
            gb = __shfl_sync(kFull, gb, 0);
            ok = __shfl_sync(kFull, ok, 0);
            if (ok && (uint32_t)lane < ngroups) {          // the descriptors of the groups: lane = group
               GroupDesc d;
               d.tile = tile;
               d.first = FILTER ? (gb + (uint32_t)lane) * 32u : (uint32_t)lane * 32u;
               d.meta = min(nlive - (uint32_t)lane * 32u, 32u) | (my_cols << 8);
               d.poff = pb + my_gblk * 24u;
               d.lead_lo = s_glead[lane][0] ^ s_glead[lane][1] ^ s_glead[lane][2];
               d.lead_hi = s_glead[lane][1];
               d.pad0 = d.pad1 = 0u;
               a.gdesc[gb + (uint32_t)lane] = d;
            }
            if (lane == 0) {
               s_pbase = pb;
               s_gbase = gb;
               s_stores = ok;
               mbar_arrive(&bar_alloc);                       // (release: the three words are visible to whoever sees the phase flip)
            }
            pbase = pb;
            gbase = gb;
            stores = ok != 0u;
         } else {
            mbar_wait(&bar_alloc, alloc_phase);              // (a spin on a flag took 13 % of the kernel's issue slots)
            pbase = s_pbase;
            gbase = s_gbase;
            stores = s_stores != 0u;
         }
         alloc_phase ^= 1u;                                  // every warp passes here exactly once per tile that gets this far
         ready = true;
      };
      const uint32_t lut_addr = smem_addr(lut);
      // ---- classify + transpose.  The blocks (32 lines x 32 columns) of the tile are numbered group by group.
      //      Whole rounds of 32 blocks: a warp takes four neighbouring blocks, eight lanes each, lane = (block, four
      //      columns).  The blocks that are left: one per warp, lane = column (a quarter of the work per round, so
      //      the warps that get none wait a quarter as long) ----
      const uint32_t nfull = nblocks & ~31u;
      const uint32_t cq = (uint32_t)lane & 7u;
      for (uint32_t b0 = (uint32_t)warp * 4u; b0 < nfull; b0 += (uint32_t)kFWarps * 4u) {
         const uint32_t blk = b0 + ((uint32_t)lane >> 3);
         // the group of the block: the last one that starts at or before it
         uint32_t g = 0;
#pragma unroll
         for (int q = 0; q < 4; q++) {
            const uint32_t m = __ballot_sync(kFull, my_gblk <= b0 + (uint32_t)q);
            if (((uint32_t)lane >> 3) == (uint32_t)q) g = 31u - (uint32_t)__clz(m);
         }
         const uint32_t cb = blk - __shfl_sync(kFull, my_gblk, g);
         const uint16_t *ls = lsg + g * 32u;
         const uint8_t *col = buf + cb * 32u + cq * 4u;
         uint32_t P[3][4];                                        // plane words of the lane's four columns
         {
            uint32_t A[4][4];                                     // [column][line octet]: bytes [- | p2 | p1 | p0], bit = line
#pragma unroll
            for (int q = 0; q < 4; q++) {
#pragma unroll
               for (int j = 0; j < 4; j++) A[j][q] = 0u;
#pragma unroll
               for (int i = 0; i < 8; i++) {
                  const uint32_t w = *reinterpret_cast<const uint32_t *>(col + ls[8 * q + i]);
                  const uint32_t x0 = w & 0xffu, x1 = __byte_perm(w, 0u, 0x4441u), x2 = __byte_perm(w, 0u, 0x4442u), x3 = w >> 24;
                  A[0][q] = mad_u32(lds_u32c(mad_u32(x0, a.four, lut_addr)), 1u << i, A[0][q]);
                  A[1][q] = mad_u32(lds_u32c(mad_u32(x1, a.four, lut_addr)), 1u << i, A[1][q]);
                  A[2][q] = mad_u32(lds_u32c(mad_u32(x2, a.four, lut_addr)), 1u << i, A[2][q]);
                  A[3][q] = mad_u32(lds_u32c(mad_u32(x3, a.four, lut_addr)), 1u << i, A[3][q]);
               }
            }
            // per column: the bytes p of the four octets make the word of plane p
#pragma unroll
            for (int j = 0; j < 4; j++) k12_planes_of_column(A[j], P[0][j], P[1][j], P[2][j]);
         }
         if (cb == 0u && cq == 0u) {
            // the columns in front of a line's first byte are NULL columns (111)
            const uint32_t m1 = s_glead[g][0], m2 = s_glead[g][1], m3 = s_glead[g][2];
#pragma unroll
            for (int p = 0; p < 3; p++) {
               P[p][0] |= m1;
               P[p][1] |= m2;
               P[p][2] |= m3;
            }
         }
         room();
         if (stores) {
            uint4 *out = a.planes + (size_t)pbase + (size_t)blk * 24u + cq;
            out[0] = make_uint4(P[0][0], P[0][1], P[0][2], P[0][3]);
            out[8] = make_uint4(P[1][0], P[1][1], P[1][2], P[1][3]);
            out[16] = make_uint4(P[2][0], P[2][1], P[2][2], P[2][3]);
         }
      }
      for (uint32_t blk = nfull + (uint32_t)warp; blk < nblocks; blk += (uint32_t)kFWarps) {
         const uint32_t g = 31u - (uint32_t)__clz(__ballot_sync(kFull, my_gblk <= blk));
         const uint32_t cb = blk - __shfl_sync(kFull, my_gblk, g);
         const uint16_t *ls = lsg + g * 32u;
         const uint8_t *col = buf + cb * 32u + (uint32_t)lane;
         uint32_t acc[4] = {0u, 0u, 0u, 0u};                      // [- | p2 | p1 | p0] bytes of the lines 8q .. 8q+7
#pragma unroll
         for (int r = 0; r < 32; r++) {
            const uint32_t e = lds_u32c(mad_u32((uint32_t)col[ls[r]], a.four, lut_addr));
            acc[r >> 3] = mad_u32(e, 1u << (r & 7), acc[r >> 3]);
         }
         uint32_t P0, P1, P2;
         k12_planes_of_column(acc, P0, P1, P2);
         if (cb == 0u && lane < 3) {
            const uint32_t m = s_glead[g][lane];
            P0 |= m;
            P1 |= m;
            P2 |= m;
         }
         room();
         if (stores) {
            uint32_t *out = reinterpret_cast<uint32_t *>(a.planes + (size_t)pbase + (size_t)blk * 24u) + lane;
            out[0] = P0;
            out[32] = P1;
            out[64] = P2;
         }
      }
      room();                                // (a warp without a block; thread 0 must publish in any case)
      if (FILTER && stores) {
         for (uint32_t j = (uint32_t)tid; j < ngroups * 32u; j += (uint32_t)kFThreads)
            a.gent[(size_t)gbase * 32u + j] = j < nlive ? live[j] : (uint16_t)0;
      }
      end_of_tile();
   }
}

}  // namespace sqb
