// sqb_kernels.cuh -- the hand-written sm_100a kernels of the matching path.
//
//   k1_scan_classify      K1  newline / line-offset scan + class nibbles (TMA-staged
//                             tiles, 16-byte loads, warp-aggregated prefix, line
//                             starts allocated per tile with one atomic)
//   k1_scan_tiles             exclusive scans over the K1 tiles (one CTA, staged)
//   k1_gather                 tile segments -> line order; line filter (dead-on-arrival
//                             flags or FASTQ record structure); results preset
//   k2_forward_thread     K2  Myers/Hyyro forward matcher, one read per thread
//                             (pattern <= 64 positions: 1 or 2 words)
//   k2_forward_lanes      K2  blocked multi-word automaton across warp lanes
//                             (pattern > 64 positions; carries via ballot)
//   k_tile_sums / _scan       records per 1024-line tile -> first record of each tile
//   k_seg_reduce              segment cuts: one result per line
//   k34_finish_lines      K3+K4 for SQ_FIRST / SQ_BEST: reverse pass + ordered
//                             ballot/popc compaction of one record per line
//   k34_finish_events     K3+K4 for SQ_ALL: reverse pass per event + ordered
//                             scatter to offs[line] + rank
//
// (The production matcher for patterns of up to 128 positions -- the bit-plane pack
// and the line-bit-sliced automaton -- lives in sqb_k2_bitslice.cuh / sqb_bitslice.h.)
//
// Semantics restated from /root/reference/src/libseeq.c:171-352.
#pragma once

#include "sqb_device.cuh"
#include "sqb_tables.h"

namespace sqb {

// ---------------------------------------------------------------------------
// parameters shared by host and device
// ---------------------------------------------------------------------------
constexpr int kMaxWords = 32;                 // pattern words per automaton (m <= 1024)

// byte classes: low 3 bits = base code 0..4 (A C G T N), bits 5:4 = kind
constexpr uint8_t kKindBase = 0x00;           // feeds the automaton
constexpr uint8_t kKindSkip = 0x10;           // invisible (SQ_IGNORE, '\n' in SQ_STREAM)
constexpr uint8_t kKindStop = 0x20;           // ends the line (NUL, '\n', SQ_FAIL)

struct Pattern {
   uint32_t eq[5][kMaxWords];                 // left-aligned match masks per base code
   uint8_t  cls[256];                         // byte -> class
   int      m;
   int      tau;
};

enum Mode { M_COUNT = 0, M_FIRST = 1, M_BEST = 2, M_ALL = 3, M_COUNTALL = 4 };

enum Counter {
   C_NLINES = 0,     // counted lines (K1)
   C_NMATCHED = 1,   // lines with >= 1 match
   C_NRECS = 2,      // records / events in total
   C_EVENTS = 3,     // events appended by K2 in SQ_ALL (may exceed the capacity)
   C_TICKET_K1 = 4,
   C_TICKET_SCAN = 5,
   C_TICKET_FIN = 6,
   C_LS_CURSOR = 7,  // K1: allocation cursor into the unordered line-start array
   C_BS_SELECTED = 8,// 1: the bit-sliced matcher serves this scan, 0: the word-parallel one, 2: planes too small
   C_BS_COLS = 9,    // tile columns the plane buffer must hold
   C_NPSEUDO = 10,   // entries of ls: counted lines + segment cuts (== C_NLINES without cuts)
   C_NCUTS = 11,     // segment cuts made by K1
   C_NZ_CORR = 12,   // SQ_ALL with cuts: segments with events beyond the first of their line
   C_NACTIVE = 13,   // line filter: entries of ls the matcher has to look at
   C_NGROUPS = 14,   // fused tokenise + pack: groups of 32 lines allocated (may exceed the capacity)
   C_PLANE_UNITS = 15,// fused: uint4 units of bit-planes allocated (may exceed the capacity)
   C_FUSED_OVF = 16, // fused: a tile met something the fused kernel does not handle (the host re-runs the scan)
   C_COUNT = 24
};

struct Event {        // SQ_ALL: one forward event, unordered
   uint32_t line;
   uint32_t rank;     // index among the events of its line
   uint32_t end;
   uint32_t dist;
};

struct Rec {          // == sqb_rec_t
   uint32_t line, start, end, dist;
};

constexpr unsigned long long kNoMatch = ~0ull;

// device-side choice between the bit-sliced matcher (sqb_k2_bitslice.cuh) and the
// word-parallel ones below: both are launched, exactly one of them does the scan
struct BsGate {
   uint32_t min_lines;          // fewer lines do not fill the bit-sliced warps (default 65536)
   uint32_t max_line;           // bytes; inputs with a longer line run thread-per-line (default 4096)
};

// ===========================================================================
// K1: line-offset scan + class coding ("tokenizer")
// ===========================================================================
// One pass over the text (HBM-bound in principle, issue-bound in practice: one
// table look-up per byte).  Tiles of 32 KiB are staged into shared memory by TMA
// bulk copies (two stages, mbarrier) and handed out in ticket order.  Inside a
// tile warp w owns the 4 KiB [w*4096, (w+1)*4096) and lane l the 128 CONTIGUOUS
// bytes [l*128, (l+1)*128) of it, read as eight 16-byte vectors in the rotated
// order (k + l) & 7 so that the quarter-warps hit distinct banks.
//
// Every byte goes through a 256-entry class table held in shared memory (PRMT
// extracts it, LDS.U8 looks it up, IMAD drops the nibble into place: one
// instruction on each of the ALU, LSU and FMA pipes).  Bit 3 of a nibble marks
// '\n', which is all the line scan needs: the newline flags of a lane's 128
// bytes are folded into four words (one per 32 text bytes), counted with popc,
// and ONE warp-shuffle scan + the block totals give every lane its place.
//
// The class nibbles (64 B per lane) are written IN PLACE over the warp's own
// text (all lanes hold their text in registers by then) and leave the SM as one
// 2 KiB bulk (TMA) store per warp -- coalesced and asynchronous.  A tile
// allocates room for its line starts with ONE atomicAdd on a cursor (no tile
// waits for another: a decoupled look-back here was measured to stall 58 % of
// the time) and writes them, ordered inside the tile, into `ls_raw`.  Two tiny
// kernels finish the job: k1_scan_tiles (exclusive scan of the per-tile counts =
// first line number of every tile) and k1_gather (moves the tile segments into
// line order, `ls`).
//
// Class nibble (bits 2:0): 0 A, 1 C, 2 G, 3 T/U, 4 N (or any other byte with
// SQ_CONVERT), 5 STOP (NUL, '\n', other bytes with SQ_FAIL), 6 SKIP (other
// bytes with SQ_IGNORE), 7 NULL (no byte: columns in front of a line; K2 only).
constexpr uint32_t kK1LaneBytes = 128;                        // contiguous text bytes per lane and tile
constexpr uint32_t kK1WarpBytes = 32 * kK1LaneBytes;          // 4 KiB of text per warp
constexpr uint32_t kK1Tile      = kWarps * kK1WarpBytes;      // 32 KiB of text per tile
constexpr uint32_t kK1Stage     = kK1Tile + 16;               // + look-ahead for the FASTA test
constexpr uint32_t kK1Smem      = 2 * kK1Stage + 256;         // stages + class table

struct K1Args {
   const uint8_t *text;
   uint32_t n;
   uint32_t *ls_raw;              // out: line starts, tile segments in allocation order
   uint32_t ls_cap;               // capacity of ls_raw / ls (entries); counting continues beyond it
   uint2 *codes;                  // out: 16 class nibbles per 16 text bytes (nullptr: not wanted)
   unsigned long long *ctr;
   uint32_t *tile_cnt;            // out: entries (line starts + cuts) of each tile
   uint32_t *tile_off;            // out: where the tile's segment starts in ls_raw
   uint32_t *tile_real;           // out (CUT): counted lines starting in each tile
   uint32_t *tile_last;           // out (CUT): 1 + the last line start inside the tile (0: none)
   uint32_t *tile_alive;          // out (FILTER): entries of the tile that are not dead on arrival
   uint32_t filter_k;             // FILTER: a STOP among the first filter_k (<= 8) bytes of a line kills it
   int fasta;
   uint32_t skip;                 // 0..15: the buffer proper starts at text + skip (text is the 16-byte aligned
                                  // address below it); the bytes in front are not looked at, byte skip-1 acts
                                  // as the '\n' in front of the first line
};

constexpr uint32_t kDeadBit = 0x80000000u;     // FILTER: flag in ls_raw (text < 2 GiB)

// shared -> global bulk (TMA) store of the calling thread's bulk group
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
   asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)),
                "r"(bytes)
                : "memory");
   asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// CUT: long lines are cut into segments (sqb_tables.h).  The candidates A = 256
// (mod 2048) are the first bytes of lanes 2 and 18 of every warp, and "the line
// has been running for 256 bytes" = the two lanes in front hold no line start:
// two shuffles of the warp scan decide it.
//
// FILTER: a line with a STOP byte among its first m - tau bytes (the '@', '+' and
// quality lines of FASTQ with -x 0; lines shorter than that) can never match
// (libseeq.c:267-270: the scan of a line ends at the first illegal byte).  K1
// flags such lines dead on arrival -- their first 8 class nibbles are at hand in
// shared memory when the line start is emitted -- and the matcher packs only the
// others into its tiles, which is a 4x denser tile for FASTQ.
template <bool CODES, bool CUT, bool FILTER>
__global__ void __launch_bounds__(kThreads, 3) k1_scan_classify(const K1Args a, const __grid_constant__ ClassTable ct)
{
   extern __shared__ __align__(128) uint8_t dyn[];            // 2 x kK1Stage, then the class table
   __shared__ uint64_t bar[2];
   __shared__ uint32_t s_tile[2];
   __shared__ uint32_t s_wsum[kWarps];
   __shared__ uint32_t s_wcut[kWarps];
   __shared__ uint32_t s_base[2];
   __shared__ uint32_t s_last[2];
   __shared__ uint32_t s_alive[2];

   static_assert(kCutWindow == 2 * kK1LaneBytes && kCutStride == 16 * kK1LaneBytes, "cut candidates = lanes 2 and 18");
   const uint32_t n = a.n;
   const uint32_t ntiles = (n + kK1Tile - 1) / kK1Tile;
   const uint32_t n16 = (n + 15u) & ~15u;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   uint8_t *lut = dyn + 2 * kK1Stage;                                    // [256] class nibbles
   const uint32_t rot = (uint32_t)lane & 7u;
   const uint32_t lane_off = (uint32_t)warp * kK1WarpBytes + (uint32_t)lane * kK1LaneBytes;   // text of this lane
   const uint32_t code_off = (uint32_t)warp * kK1WarpBytes + (uint32_t)lane * (kK1LaneBytes / 2);  // its nibbles, in place

   auto issue = [&](int stage, uint32_t tile) {
      const uint32_t start = tile * kK1Tile;
      uint32_t bytes = n16 - start;
      if (bytes > kK1Stage) bytes = kK1Stage;
      mbar_expect_tx(&bar[stage], bytes);
      bulk_g2s(dyn + stage * kK1Stage, a.text + start, bytes, &bar[stage]);
   };

   lut[tid] = ct.code[tid];
   if (tid == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      mbar_fence_init();
      const uint32_t t = (uint32_t)atomicAdd(&a.ctr[C_TICKET_K1], 1ull);
      s_tile[0] = t;
      if (t < ntiles) issue(0, t);
      s_last[0] = s_last[1] = 0u;
      s_alive[0] = s_alive[1] = 0u;
   }
   __syncthreads();

   uint32_t prev_tile = 0xffffffffu;      // tid 0: the tile whose last line start is still to be published
   uint32_t phases = 0;                   // bit s = parity to wait for on stage s
   bool store_pending = false;            // lane 0 of every warp: a bulk store of the other stage may still read it
   for (int stage = 0;; stage ^= 1) {
      const uint32_t tile = s_tile[stage];
      if (tile >= ntiles) break;
      mbar_wait(&bar[stage], (phases >> stage) & 1u);
      phases ^= 1u << stage;

      uint8_t *buf = dyn + stage * kK1Stage;
      const uint32_t pos0 = tile * kK1Tile + lane_off;        // text position of this lane's first byte

      // ---- classify: vector slot k holds text vector (k + rot) & 7 of the lane ----
      uint32_t lo[8], hi[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
         const uint4 v = *reinterpret_cast<const uint4 *>(buf + lane_off + ((((uint32_t)k + rot) & 7u) << 4));
         const uint32_t w[4] = {v.x, v.y, v.z, v.w};
         uint32_t l = 0, h = 0;
#pragma unroll
         for (int j = 0; j < 8; j++) {
            const uint32_t bx = __byte_perm(w[j >> 2], 0u, 0x4440u + (uint32_t)(j & 3));
            const uint32_t by = __byte_perm(w[2 + (j >> 2)], 0u, 0x4440u + (uint32_t)(j & 3));
            l = mad_u32((uint32_t)lut[bx], 1u << (4 * j), l);
            h = mad_u32((uint32_t)lut[by], 1u << (4 * j), h);
         }
         lo[k] = l;
         hi[k] = h;
      }
      if (a.skip != 0u && pos0 == 0u) {     // the bytes in front of the buffer: STOP, the last one a newline
         const uint32_t nl = a.skip - 1u;            // 0..14, in vector 0 (lane 0: rot == 0)
#pragma unroll
         for (int j = 0; j < 8; j++) {
            if ((uint32_t)j <= nl)
               lo[0] = (lo[0] & ~(0xFu << (4 * j))) | ((uint32_t)(kClsStop | ((uint32_t)j == nl ? kClsNewline : 0)) << (4 * j));
            if (8u + (uint32_t)j <= nl)
               hi[0] = (hi[0] & ~(0xFu << (4 * j))) | ((uint32_t)(kClsStop | (8u + (uint32_t)j == nl ? kClsNewline : 0)) << (4 * j));
         }
      }
      if (pos0 + kK1LaneBytes > n) {        // bytes at or beyond n are STOP and never newlines
#pragma unroll
         for (int k = 0; k < 8; k++) {
            const uint32_t p = pos0 + ((((uint32_t)k + rot) & 7u) << 4);
#pragma unroll
            for (int j = 0; j < 8; j++) {
               if (p + (uint32_t)j >= n) lo[k] = (lo[k] & ~(0xFu << (4 * j))) | ((uint32_t)kClsStop << (4 * j));
               if (p + 8u + (uint32_t)j >= n) hi[k] = (hi[k] & ~(0xFu << (4 * j))) | ((uint32_t)kClsStop << (4 * j));
            }
         }
      }

      // ---- newline flags: f[v] bit 4i+2 <=> byte i, bit 4i+3 <=> byte 8+i of vector v ----
      uint32_t f[8];
      {
         uint32_t g[8], h2[8];
#pragma unroll
         for (int k = 0; k < 8; k++) g[k] = ((lo[k] & 0x88888888u) >> 1) | (hi[k] & 0x88888888u);
         // slot k holds vector (k + rot) & 7: rotate the slots back into text order
#pragma unroll
         for (int v = 0; v < 8; v++) h2[v] = (rot & 1u) ? g[(v + 7) & 7] : g[v];
#pragma unroll
         for (int v = 0; v < 8; v++) g[v] = (rot & 2u) ? h2[(v + 6) & 7] : h2[v];
#pragma unroll
         for (int v = 0; v < 8; v++) f[v] = (rot & 4u) ? g[(v + 4) & 7] : g[v];
      }
      // c[q] bit 4i+e <=> byte 8e+i of the 32-byte chunk q of the lane
      uint32_t c[4];
#pragma unroll
      for (int q = 0; q < 4; q++) c[q] = (f[2 * q] >> 2) | f[2 * q + 1];

      // a newline at p opens a line at p+1 only if p+1 < n, and in FASTA mode only
      // if that line is not a header (the text is still intact here)
      if (a.fasta || pos0 + kK1LaneBytes + 1u > n) {
#pragma unroll
         for (int q = 0; q < 4; q++) {
            uint32_t bits = c[q];
            while (bits) {
               const int b = __ffs(bits) - 1;
               bits &= bits - 1;
               const uint32_t byte = 32u * (uint32_t)q + 8u * (uint32_t)(b & 3) + (uint32_t)(b >> 2);
               const uint32_t s = pos0 + byte + 1u;
               if (s >= n || (a.fasta && buf[lane_off + byte + 1u] == '>')) c[q] &= ~(1u << b);
            }
         }
      }
      // the first line of the buffer has no newline in front of it
      // (with a.skip > 0 the newline at skip-1 opens it like any other line)
      const uint32_t first = (tile == 0 && tid == 0 && n > 0 && a.skip == 0u && !(a.fasta && buf[0] == '>')) ? 1u : 0u;

      // ---- one prefix over the lanes of the warp, then over the warps ------------
      const uint32_t cnt = (uint32_t)(__popc(c[0]) + __popc(c[1]) + __popc(c[2]) + __popc(c[3])) + first;
      uint32_t inc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         const uint32_t t = __shfl_up_sync(kFull, inc, d);
         if (lane >= d) inc += t;
      }
      if (lane == 31) s_wsum[warp] = inc;
      uint32_t cut = 0, cutinc = 0;          // a cut at the first byte of this lane; cuts up to and including this lane
      if (CUT) {
         const uint32_t prev1 = __shfl_up_sync(kFull, inc, 1);
         const uint32_t prev3 = __shfl_up_sync(kFull, inc, 3);
         if ((lane == 2 || lane == 18) && pos0 < n) cut = (prev1 - (lane == 2 ? 0u : prev3)) == 0u ? 1u : 0u;
         const uint32_t cut2 = __shfl_sync(kFull, cut, 2), cut18 = __shfl_sync(kFull, cut, 18);
         cutinc = (lane >= 2 ? cut2 : 0u) + (lane >= 18 ? cut18 : 0u);
         if (lane == 31) s_wcut[warp] = cut2 + cut18;
      }
      // the bulk store this warp issued one tile ago has long read its stage; make
      // sure before the stage is refilled below
      if (CODES && store_pending) bulk_wait_read();
      __syncthreads();                       // A: every lane holds its text in registers, s_wsum is complete
      uint32_t before = 0, tile_total = 0, tile_cuts = 0;
#pragma unroll
      for (int w2 = 0; w2 < kWarps; w2++) {
         const uint32_t x = s_wsum[w2] + (CUT ? s_wcut[w2] : 0u);
         if (w2 < warp) before += x;
         tile_total += x;
         if (CUT) tile_cuts += s_wcut[w2];
      }
      if (tid == 0) {
         // prefetch the next ticket into the other stage (its last store has been waited for)
         const uint32_t t = (uint32_t)atomicAdd(&a.ctr[C_TICKET_K1], 1ull);
         s_tile[stage ^ 1] = t;
         if (t < ntiles) issue(stage ^ 1, t);
         const uint32_t at = (uint32_t)atomicAdd(&a.ctr[C_LS_CURSOR], (unsigned long long)tile_total);
         a.tile_cnt[tile] = tile_total;
         a.tile_off[tile] = at;
         s_base[stage] = at;
         if (CUT) {
            a.tile_real[tile] = tile_total - tile_cuts;
            if (tile_cuts) atomicAdd(&a.ctr[C_NCUTS], (unsigned long long)tile_cuts);
         }
         if (CUT || FILTER) {
            // every warp has passed A: the emit of the previous tile (other stage) is complete
            if (prev_tile != 0xffffffffu) {
               if (CUT) a.tile_last[prev_tile] = s_last[stage ^ 1];
               if (FILTER) a.tile_alive[prev_tile] = s_alive[stage ^ 1];
            }
            s_last[stage ^ 1] = 0u;
            s_alive[stage ^ 1] = 0u;
            prev_tile = tile;
         }
      }

      // ---- class nibbles: in place over the warp's own text, one bulk store per warp ----
      if (CODES) {
#pragma unroll
         for (int k = 0; k < 8; k++)
            *reinterpret_cast<uint2 *>(buf + code_off + ((((uint32_t)k + rot) & 7u) << 3)) = make_uint2(lo[k], hi[k]);
         fence_proxy_async();
         __syncwarp();
         if (lane == 0) {
            const uint32_t wpos = tile * kK1Tile + (uint32_t)warp * kK1WarpBytes;
            bulk_s2g(reinterpret_cast<uint8_t *>(a.codes) + (wpos >> 1), buf + (uint32_t)warp * kK1WarpBytes,
                     kK1WarpBytes / 2);
            store_pending = true;
         }
      }
      __syncthreads();                       // B: s_base, s_tile
      uint32_t idx = s_base[stage] + before + (inc - cnt) + (cutinc - cut);

      // ---- emit (ordered inside the tile) -------------------------------------
      uint32_t mylast = 0;                          // CUT: 1 + the last LINE start emitted by this lane
      uint32_t myalive = 0;                         // FILTER: entries of this lane that are not dead on arrival
      // FILTER: kDeadBit if the line that starts at text position s holds a STOP among its
      // first filter_k bytes.  The nibbles of the whole tile are in shared memory by now
      // (in place, warp by warp); a window that leaves the warp's 4 KiB is not looked at.
      auto dead_flag = [&](uint32_t s) -> uint32_t {
         if (!FILTER) return 0u;
         const uint32_t o = s - tile * kK1Tile;                           // <= kK1Tile
         const uint32_t i = o & (kK1WarpBytes - 1u);                      // nibble index inside its warp
         if (o >= kK1Tile || i + 8u > kK1WarpBytes) { myalive++; return 0u; }
         const uint8_t *wb = buf + (o & ~(kK1WarpBytes - 1u));
         const uint32_t a0 = (i >> 1) & ~3u;
         const uint32_t w0 = *reinterpret_cast<const uint32_t *>(wb + a0);
         const uint32_t w1 = *reinterpret_cast<const uint32_t *>(wb + a0 + 4u);
         uint32_t win = __funnelshift_r(w0, w1, (i & 7u) * 4u);           // 8 nibbles from the line start on
         if (a.filter_k < 8u) win &= (1u << (4u * a.filter_k)) - 1u;      // nibble 0 = A: never a STOP
         const uint32_t y = (win ^ 0x55555555u) & 0x77777777u;            // zero nibble <=> class 5 (STOP)
         const bool dead = ((y - 0x11111111u) & ~y & 0x88888888u) != 0u;
         myalive += dead ? 0u : 1u;
         return dead ? kDeadBit : 0u;
      };
      if (first) {
         if (idx < a.ls_cap) a.ls_raw[idx] = 0u | dead_flag(0u);
         idx++;
         mylast = 1u;
      }
      if (CUT && cut) {                             // the segment start comes before the line starts of the lane
         if (idx < a.ls_cap) a.ls_raw[idx] = pos0;
         idx++;
         myalive++;
      }
      // One line start per lane and round: the warp runs as many rounds as its busiest lane has
      // starts (FASTQ-like text: 3-4), all lanes in step.  The starts of a lane are taken in BIT
      // order of c[q], which is not text order (bit 4i+e <=> byte 8e+i); the place of each among
      // the starts of its lane comes from a popc over the bits that precede it in the text.
      {
         const uint32_t r1 = (uint32_t)__popc(c[0]), r2 = r1 + (uint32_t)__popc(c[1]), r3 = r2 + (uint32_t)__popc(c[2]);
         uint32_t w0 = c[0], w1 = c[1], w2 = c[2], w3 = c[3];
         for (uint32_t left = cnt - first; left != 0u; left--) {
            const uint32_t q = w0 ? 0u : (w1 ? 1u : (w2 ? 2u : 3u));
            const uint32_t cur = w0 ? w0 : (w1 ? w1 : (w2 ? w2 : w3));
            const uint32_t orig = q == 0u ? c[0] : (q == 1u ? c[1] : (q == 2u ? c[2] : c[3]));
            const uint32_t rbase = q == 0u ? 0u : (q == 1u ? r1 : (q == 2u ? r2 : r3));
            const uint32_t b = (uint32_t)__ffs(cur) - 1u, e = b & 3u;
            const uint32_t rest = cur & (cur - 1u);
            if (q == 0u) w0 = rest;
            else if (q == 1u) w1 = rest;
            else if (q == 2u) w2 = rest;
            else w3 = rest;
            // starts of this word in front of byte 8e+i: all of the bytes 8e'+i' with e' < e, and i' < i of e
            const uint32_t before_b = (0x11111111u * ((1u << e) - 1u)) | ((0x11111111u << e) & ((1u << (b & ~3u)) - 1u));
            const uint32_t at = idx + rbase + (uint32_t)__popc(orig & before_b);
            const uint32_t s = pos0 + 32u * q + 1u + 8u * e + (b >> 2);
            const uint32_t fl = dead_flag(s);
            if (at < a.ls_cap) a.ls_raw[at] = s | fl;
            mylast = max(mylast, s + 1u);
         }
      }
      if (CUT) {
         const uint32_t wl = __reduce_max_sync(kFull, mylast);
         if (lane == 0 && wl) atomicMax(&s_last[stage], wl);
      }
      if (FILTER) {
         const uint32_t wa = __reduce_add_sync(kFull, myalive);
         if (lane == 0 && wa) atomicAdd(&s_alive[stage], wa);
      }
   }
   if (CODES && store_pending) bulk_wait_all();
   if (CUT || FILTER) {
      __syncthreads();
      if (tid == 0 && prev_tile != 0xffffffffu) {                 // the slot of the other stage is 0
         if (CUT) a.tile_last[prev_tile] = s_last[0] | s_last[1];
         if (FILTER) a.tile_alive[prev_tile] = s_alive[0] + s_alive[1];
      }
   }
}

// exclusive scans over the K1 tiles (one CTA of 1024 threads):
//   tile_base[t]  = index in ls of the first entry of tile t;  total -> ctr[C_NPSEUDO]
// and, when K1 made segment cuts (tile_real != nullptr),
//   tile_rbase[t] = number of the first LINE of tile t;         total -> ctr[C_NLINES]
//   tile_lbeg[t]  = 1 + the last line start before tile t (running maximum of tile_last)
struct K1ScanArgs {
   const uint32_t *tile_cnt;
   uint32_t *tile_base;
   uint32_t ntiles;
   unsigned long long *ctr;
   const uint32_t *tile_real;     // nullptr: no cuts, every entry is a line
   uint32_t *tile_rbase;
   const uint32_t *tile_last;
   uint32_t *tile_lbeg;
   uint32_t *tile_alive;          // nullptr: no line filter
   uint32_t *tile_abase;          // index in act[] of the first live entry of tile t; total -> ctr[C_NACTIVE]
   int fastq;                     // the live entries are the sequence lines of 4-line records (entry index 1 mod 4):
                                  // tile_alive is computed here, from tile_base and tile_cnt
   // fused tokenise + pack (sqb_k12_fused.cuh): this kernel also makes the matcher decision k15_scan makes on
   // the two-kernel path -- ctr[C_BS_SELECTED] = 1 bit-sliced scan, 0 too few lines (word-parallel kernels),
   // 2 plane / group buffers too small (ctr[C_BS_COLS] = plane units needed), 4 the fused kernel gave up
   int fused;
   uint32_t planes_cap, gdesc_cap;
   BsGate gate;
};

static __global__ void __launch_bounds__(1024) k1_scan_tiles(const K1ScanArgs a)
{
   __shared__ unsigned long long s_warp[32], s_real[32], s_act[32];
   __shared__ uint32_t s_max[32];
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const bool cut = a.tile_real != nullptr && a.ctr[C_NCUTS] != 0ull;     // K1 is complete: the count is final
   if (!cut) {                            // the usual case: one or two plain prefix sums
      __shared__ CtaScanSmem cs;
      const unsigned long long total = cta_scan_u32(a.tile_cnt, a.tile_base, a.ntiles, cs);
      unsigned long long alive = total;
      if (a.tile_alive != nullptr && a.fastq) {
         // entries i < x with i = 1 (mod 4): (x + 2) / 4
         for (uint32_t t = threadIdx.x; t < a.ntiles; t += 1024u) {
            const uint32_t b = a.tile_base[t], e = b + a.tile_cnt[t];
            a.tile_alive[t] = (e + 2u) / 4u - (b + 2u) / 4u;
         }
         __syncthreads();
      }
      if (a.tile_alive != nullptr) alive = cta_scan_u32(a.tile_alive, a.tile_abase, a.ntiles, cs);
      if (threadIdx.x == 0) {
         a.ctr[C_NPSEUDO] = total;
         a.ctr[C_NLINES] = total;
         a.ctr[C_NACTIVE] = alive;
         if (a.fused) {
            const unsigned long long units = a.ctr[C_PLANE_UNITS], ng = a.ctr[C_NGROUPS];
            a.ctr[C_BS_COLS] = units;
            a.ctr[C_BS_SELECTED] = a.ctr[C_FUSED_OVF] != 0ull ? 4ull
                                   : (total < a.gate.min_lines ? 0ull : ((units > a.planes_cap || ng > a.gdesc_cap) ? 2ull : 1ull));
         }
      }
      return;
   }
   // warp w owns the contiguous run [w*per, (w+1)*per) and walks it 32 tiles at a
   // time (coalesced): first its totals, then -- after the 32 totals are scanned --
   // the exclusive prefix of every tile
   const uint32_t per = ((a.ntiles + 1023u) / 1024u) * 32u;
   const uint32_t t0 = min((uint32_t)warp * per, a.ntiles), t1 = min(t0 + per, a.ntiles);
   const bool filt = a.tile_alive != nullptr;
   unsigned long long sum = 0, rsum = 0, asum = 0;
   uint32_t mx = 0;
#pragma unroll 4
   for (uint32_t t = t0 + (uint32_t)lane; t < t1; t += 32) {
      sum += a.tile_cnt[t];
      if (cut) {
         rsum += a.tile_real[t];
         mx = max(mx, a.tile_last[t]);
      }
      if (filt) asum += a.tile_alive[t];
   }
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) {
      sum += __shfl_xor_sync(kFull, sum, d);
      rsum += __shfl_xor_sync(kFull, rsum, d);
      asum += __shfl_xor_sync(kFull, asum, d);
   }
   mx = __reduce_max_sync(kFull, mx);
   if (lane == 0) {
      s_warp[warp] = sum;
      s_real[warp] = rsum;
      s_act[warp] = asum;
      s_max[warp] = mx;
   }
   __syncthreads();
   unsigned long long run = 0, tot = 0, rrun = 0, rtot = 0, arun = 0, atot = 0;
   uint32_t mrun = 0;
   for (int w = 0; w < 32; w++) {
      const unsigned long long y = s_warp[w], z = s_real[w], u = s_act[w];
      if (w < warp) {
         run += y;
         rrun += z;
         arun += u;
         mrun = max(mrun, s_max[w]);
      }
      tot += y;
      rtot += z;
      atot += u;
   }
   // line numbers are u32 (a batch is < 4 GiB of text)
#pragma unroll 2
   for (uint32_t tb = t0; tb < t1; tb += 32) {
      const uint32_t t = tb + (uint32_t)lane;
      const uint32_t v = t < t1 ? a.tile_cnt[t] : 0u;
      const uint32_t rv = (cut && t < t1) ? a.tile_real[t] : 0u;
      const uint32_t lv = (cut && t < t1) ? a.tile_last[t] : 0u;
      const uint32_t av = (filt && t < t1) ? a.tile_alive[t] : 0u;
      uint32_t x = v, rx = rv, lx = lv, ax = av;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         const uint32_t y = __shfl_up_sync(kFull, x, d);
         const uint32_t ry = __shfl_up_sync(kFull, rx, d);
         const uint32_t ly = __shfl_up_sync(kFull, lx, d);
         const uint32_t ay = __shfl_up_sync(kFull, ax, d);
         if (lane >= d) {
            x += y;
            rx += ry;
            lx = max(lx, ly);
            ax += ay;
         }
      }
      // exclusive running maximum: the inclusive one of the lane in front
      uint32_t lprev = __shfl_up_sync(kFull, lx, 1);
      if (lane == 0) lprev = 0u;
      if (t < t1) {
         a.tile_base[t] = (uint32_t)run + x - v;
         if (cut) {
            a.tile_rbase[t] = (uint32_t)rrun + rx - rv;
            a.tile_lbeg[t] = max(mrun, lprev);
         }
         if (filt) a.tile_abase[t] = (uint32_t)arun + ax - av;
      }
      run += __shfl_sync(kFull, x, 31);
      rrun += __shfl_sync(kFull, rx, 31);
      arun += __shfl_sync(kFull, ax, 31);
      mrun = max(mrun, __shfl_sync(kFull, lx, 31));
   }
   if (tid == 0) {
      a.ctr[C_NPSEUDO] = tot;
      a.ctr[C_NLINES] = cut ? rtot : tot;
      a.ctr[C_NACTIVE] = filt ? atot : tot;
   }
}

// tile segments of ls_raw -> ls in order (one warp per tile) + sentinel.  With
// segment cuts also, for every entry p of ls:
//   lid[p]  = number of the line the entry belongs to
//   lbeg[p] = start of that line
// An entry is a line start iff it is 0 or follows a newline (bit 3 of the class
// nibble in front of it); a cut never does (there is no line start in the 256
// bytes before it).
struct K1GatherArgs {
   const uint32_t *ls_raw;
   uint32_t *ls;
   uint32_t ls_cap;
   const uint32_t *tile_cnt, *tile_off, *tile_base;
   uint32_t ntiles;
   uint32_t n;
   const unsigned long long *ctr;
   const uint8_t *codes;          // nullptr: no cuts
   const uint32_t *tile_rbase, *tile_lbeg;
   uint32_t *lid, *lbeg;
   const uint32_t *tile_abase;    // nullptr: no line filter
   uint32_t *act;                 // out (filter): the entries of ls the matcher looks at, in order
   uint8_t *lflags;               // out (filter): 1 = dead on arrival
   unsigned long long *res_init;  // out (or nullptr): per-entry result preset to kNoMatch -- the bit-sliced matcher
                                  // stores only the entries that match, and this saves a memset of 8 B per line
   int fastq;                     // filter by record structure: entry p is live iff p = 1 (mod 4)
};

static __global__ void __launch_bounds__(kThreads) k1_gather(const K1GatherArgs a)
{
   const int lane = threadIdx.x & 31;
   const uint32_t wid = (blockIdx.x * kThreads + threadIdx.x) >> 5;
   const uint32_t nwarps = (gridDim.x * kThreads) >> 5;
   const bool cut = a.codes != nullptr && a.ctr[C_NCUTS] != 0ull;
   const bool filt = a.tile_abase != nullptr;
   for (uint32_t t = wid; t < a.ntiles; t += nwarps) {
      const uint32_t cnt = a.tile_cnt[t], src = a.tile_off[t], dst = a.tile_base[t];
      if (!cut && !filt) {
         for (uint32_t j = lane; j < cnt; j += 32)
            if (dst + j < a.ls_cap && src + j < a.ls_cap) {
               a.ls[dst + j] = a.ls_raw[src + j];
               if (a.res_init && dst + j + 1u < a.ls_cap) a.res_init[dst + j] = kNoMatch;
            }
         continue;
      }
      uint32_t nact = filt ? a.tile_abase[t] : 0u;      // live entries before the current 32
      if (!cut) {                                        // line filter only
         for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
            const uint32_t j = j0 + (uint32_t)lane;
            const bool ok = j < cnt && dst + j < a.ls_cap && src + j < a.ls_cap;
            const uint32_t raw = ok ? a.ls_raw[src + j] : kDeadBit;
            const bool live = ok && (a.fastq ? ((dst + j) & 3u) == 1u : !(raw & kDeadBit));
            const uint32_t bal = __ballot_sync(kFull, live);
            if (ok) {
               a.ls[dst + j] = a.fastq ? raw : (raw & ~kDeadBit);
               if (a.res_init && dst + j + 1u < a.ls_cap) a.res_init[dst + j] = kNoMatch;
               if (a.act) {                 // (the fused path has its own slot tables: ls only)
                  a.lflags[dst + j] = live ? 0 : 1;
                  if (live) a.act[nact + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = dst + j;
               }
            }
            nact += (uint32_t)__popc(bal);
         }
         continue;
      }
      uint32_t nreal = a.tile_rbase[t];                 // line starts before the current 32 entries
      uint32_t lastp1 = a.tile_lbeg[t];                 // 1 + the last line start before them
      for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
         const uint32_t j = j0 + (uint32_t)lane;
         const bool ok = j < cnt && dst + j < a.ls_cap && src + j < a.ls_cap;
         const uint32_t raw = ok ? a.ls_raw[src + j] : 0u;
         const uint32_t pos = filt ? (raw & ~kDeadBit) : raw;
         const bool live = ok && !(filt && (raw & kDeadBit));
         const uint32_t lbal = __ballot_sync(kFull, live);
         bool real = false;
         if (ok) real = pos == 0u || ((a.codes[(pos - 1u) >> 1] >> ((((pos - 1u) & 1u) << 2) + 3u)) & 1u);
         const uint32_t bal = __ballot_sync(kFull, real);
         const uint32_t rb = nreal + (uint32_t)__popc(bal & ((1u << lane) - 1u));     // line starts before this entry
         uint32_t x = real ? pos + 1u : 0u;
#pragma unroll
         for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, x, d);
            if (lane >= d) x = max(x, y);
         }
         x = max(x, lastp1);
         if (ok) {
            a.ls[dst + j] = pos;
            if (a.res_init && dst + j + 1u < a.ls_cap) a.res_init[dst + j] = kNoMatch;
            a.lid[dst + j] = real ? rb : rb - 1u;
            a.lbeg[dst + j] = x - 1u;
            if (filt) {
               a.lflags[dst + j] = live ? 0 : 1;
               if (live) a.act[nact + (uint32_t)__popc(lbal & ((1u << lane) - 1u))] = dst + j;
            }
         }
         nact += (uint32_t)__popc(lbal);
         nreal += (uint32_t)__popc(bal);
         lastp1 = __shfl_sync(kFull, x, 31);
      }
   }
   if (blockIdx.x == 0 && threadIdx.x == 0) {
      const unsigned long long total = a.ctr[C_NPSEUDO];
      if (total < a.ls_cap) a.ls[total] = a.n;         // sentinel
   }
}

// ===========================================================================
// K2 (one read per thread)
// ===========================================================================
constexpr uint32_t kK2Stage = 48 * 1024;       // staged text per tile of kThreads lines

struct K2Args {
   const uint8_t *text;
   uint32_t n;
   const uint32_t *ls;
   uint32_t max_lines;          // capacity of the per-line arrays (line count is clamped to it)
   unsigned long long *ctr;
   unsigned long long *res;     // M_FIRST / M_BEST: per-line (dist << 32 | end) or kNoMatch
   uint32_t *cnt;               // M_ALL: per-line event count
   Event *ev;                   // M_ALL: unordered events
   uint32_t ev_cap;
   int gate;                    // 1: leave the scan to the bit-sliced kernel when it is selected
   int fastq;                   // 1: only the sequence lines of 4-line records (line index 1 mod 4) are scanned;
                                //    the others are treated as empty lines
};

template <int W> struct LutEntry;
template <> struct __align__(8) LutEntry<1> {
   uint32_t eq[1];
   uint32_t kind;
};
template <> struct __align__(16) LutEntry<2> {
   uint32_t eq[2];
   uint32_t kind;
   uint32_t pad;
};

// per-line state machine shared by the thread kernel; `rd(p)` returns text[p]
template <int W, int MODE, class Reader>
__device__ __forceinline__ void scan_line(const Reader &rd, const uint32_t line, const uint32_t begin,
                                          const uint32_t limit, const LutEntry<W> *lut, const int m,
                                          const int tau, const K2Args &a, uint32_t &matched,
                                          uint32_t &nevents)
{
   BitVec<W> bv;
   bv_reset(bv, m);
   int score = m;                       // uncapped search distance; see DESIGN.md "capping"
   bool flag = false;                   // the reference's `match` suppress flag
   int best_d = tau + 1;
   uint32_t best_end = 0;
   uint32_t nev = 0;
   bool hit = false;
   uint32_t p = begin;

   auto emit = [&](uint32_t end, int dist) {
      if (MODE == M_ALL) {
         const unsigned long long idx = atomicAdd(&a.ctr[C_EVENTS], 1ull);
         if (idx < a.ev_cap) a.ev[idx] = Event{line, nev, end, (uint32_t)dist};
      }
      best_d = dist;
      best_end = end;
      hit = true;
      nev++;
   };

   bool done = false;
   while (p < limit) {
      const uint8_t b = rd(p);
      const LutEntry<W> e = lut[b];
      if (e.kind != kKindBase) {
         if (e.kind == kKindSkip) { p++; continue; }
         break;
      }
      uint32_t rise, fall;
      bv_step<W>(bv, e.eq, rise, fall);
      const int streak = score;
      score += (int)rise - (int)fall;
      if (MODE == M_COUNT) {
         if (score <= tau) { hit = true; done = true; break; }
      } else {
         if (!rise) flag = false;                                   // libseeq.c:278
         bool evt = streak <= tau && !flag && (rise || streak == 0); // libseeq.c:286-288
         if (MODE == M_BEST) evt = evt && streak < best_d;
         if (evt) {
            flag = true;
            emit(p - begin, streak);
            if (MODE == M_FIRST) { done = true; break; }
         }
      }
      p++;
   }
   if (!done && MODE != M_COUNT) {
      // terminal step (NUL / '\n' / illegal byte / end of buffer): the distance
      // becomes tau+1, so any pending streak <= tau rises (libseeq.c:267-288)
      const int streak = score;
      bool evt = streak <= tau && !flag;
      if (MODE == M_BEST) evt = evt && streak < best_d;
      if (evt) emit(p - begin, streak);
   }
   if (MODE == M_FIRST || MODE == M_BEST)
      a.res[line] = hit ? (((unsigned long long)(uint32_t)best_d << 32) | best_end) : kNoMatch;
   if (MODE == M_ALL) a.cnt[line] = nev;
   matched = hit ? 1u : 0u;
   nevents = nev;
}

struct SmemReader {
   const uint8_t *base;          // stage - aligned tile start
   __device__ __forceinline__ uint8_t operator()(uint32_t p) const { return base[p]; }
};
struct GlobalReader {
   const uint8_t *base;
   __device__ __forceinline__ uint8_t operator()(uint32_t p) const { return __ldg(base + p); }
};

template <int W, int MODE>
__global__ void __launch_bounds__(kThreads) k2_forward_thread(const K2Args a, const __grid_constant__ Pattern pat)
{
   extern __shared__ __align__(128) uint8_t stage[];          // kK2Stage
   __shared__ LutEntry<W> lut[256];
   __shared__ uint64_t bar;
   __shared__ uint32_t s_red[2][kWarps];

   if (a.gate && a.ctr[C_BS_SELECTED] != 0ull) return;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   {  // byte -> (match mask words, kind)
      const uint8_t c = pat.cls[tid];
      LutEntry<W> e;
      e.kind = c & 0x30;
#pragma unroll
      for (int w = 0; w < W; w++) e.eq[w] = (c & 0x30) == kKindBase ? pat.eq[c & 7][w] : 0u;
      lut[tid] = e;
   }
   if (tid == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
   }
   __syncthreads();

   const uint32_t nlines = (uint32_t)min(a.ctr[C_NPSEUDO], (unsigned long long)a.max_lines);
   const uint32_t n = a.n;
   const uint32_t ntiles = (nlines + kThreads - 1) / kThreads;
   uint32_t phase = 0;
   uint32_t my_matched = 0, my_events = 0;

   for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const uint32_t l0 = tile * kThreads;
      const uint32_t l1 = min(l0 + kThreads, nlines);
      const uint32_t t_begin = a.ls[l0];
      const uint32_t t_end = a.ls[l1];                 // sentinel ls[nlines] == n
      const uint32_t a0 = t_begin & ~15u;
      const uint32_t bytes = ((t_end + 15u) & ~15u) - a0;
      const bool staged = bytes > 0 && bytes <= kK2Stage;
      if (staged) {
         if (tid == 0) {
            mbar_expect_tx(&bar, bytes);
            bulk_g2s(stage, a.text + a0, bytes, &bar);
         }
      }
      const uint32_t line = l0 + tid;
      uint32_t begin = 0;
      if (line < l1) begin = a.ls[line];
      if (staged) {
         mbar_wait(&bar, phase);
         phase ^= 1;
      }
      if (line < l1) {
         uint32_t mt, ne;
         const bool off = a.fastq && (line & 3u) != 1u;          // an empty line: no byte, no event (tau < m)
         const uint32_t limit = off ? begin : min(t_end, n);
         if (staged)
            scan_line<W, MODE>(SmemReader{stage - a0}, line, begin, limit, lut, pat.m, pat.tau, a, mt, ne);
         else
            scan_line<W, MODE>(GlobalReader{a.text}, line, begin, off ? begin : n, lut, pat.m, pat.tau, a, mt, ne);
         my_matched += mt;
         my_events += ne;
      }
      __syncthreads();      // everyone is done with the stage before it is refilled
   }

   if (MODE == M_COUNT || MODE == M_COUNTALL) {
      // block reduction, one atomic per CTA and counter
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
         unsigned long long sm = 0, se = 0;
         for (int w = 0; w < kWarps; w++) {
            sm += s_red[0][w];
            se += s_red[1][w];
         }
         if (sm) atomicAdd(&a.ctr[C_NMATCHED], sm);
         if (MODE == M_COUNTALL && se) atomicAdd(&a.ctr[C_NRECS], se);
      }
   }
}

// ===========================================================================
// K2 (blocked multi-word automaton across G lanes of a warp; m > 64)
// ===========================================================================
template <int G, int MODE>
__global__ void __launch_bounds__(kThreads) k2_forward_lanes(const K2Args a, const __grid_constant__ Pattern pat)
{
   __shared__ uint32_t s_eq[8][G];
   __shared__ uint8_t s_cls[256];
   __shared__ uint32_t s_red[2][kWarps];

   if (a.gate && a.ctr[C_BS_SELECTED] != 0ull) return;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int gl = lane & (G - 1);            // lane inside its group = word index
   const int top = (lane | (G - 1));         // lane holding the most significant word
   // bit l set <=> lane l is the top lane of its group
   uint32_t topmask = 0;
#pragma unroll
   for (int l = G - 1; l < 32; l += G) topmask |= 1u << l;

   s_cls[tid] = pat.cls[tid];
   if (tid < 5 * G) s_eq[tid / G][tid % G] = pat.eq[tid / G][tid % G];
   __syncthreads();

   const int m = pat.m, tau = pat.tau;
   const int pad = G * 32 - m;
   const int lo = gl * 32;
   const uint32_t pv0 = pad <= lo ? ~0u : (pad >= lo + 32 ? 0u : (~0u << (pad - lo)));

   constexpr int kLinesPerBlock = kThreads / G;
   const uint32_t nlines = (uint32_t)min(a.ctr[C_NPSEUDO], (unsigned long long)a.max_lines);
   const uint32_t n = a.n;
   const uint32_t ntiles = (nlines + kLinesPerBlock - 1) / kLinesPerBlock;
   uint32_t my_matched = 0, my_events = 0;

   for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const uint32_t line = tile * kLinesPerBlock + (uint32_t)(tid / G);
      bool done = line >= nlines || (a.fastq && (line & 3u) != 1u);      // not a sequence line: nothing to scan
      const uint32_t begin = done ? 0u : a.ls[line];
      uint32_t p = begin;
      uint32_t pv = pv0, mv = 0;
      int score = m;
      bool flag = false, hit = false;
      int best_d = tau + 1;
      uint32_t best_end = 0, nev = 0;

      auto emit = [&](uint32_t end, int dist) {
         if (MODE == M_ALL && gl == 0) {
            const unsigned long long idx = atomicAdd(&a.ctr[C_EVENTS], 1ull);
            if (idx < a.ev_cap) a.ev[idx] = Event{line, nev, end, (uint32_t)dist};
         }
         best_d = dist;
         best_end = end;
         hit = true;
         nev++;
      };

      while (__any_sync(kFull, !done)) {
         uint32_t kind = kKindStop;
         uint32_t code = 0;
         if (!done && p < n) {
            const uint8_t c = s_cls[__ldg(a.text + p)];
            kind = c & 0x30;
            code = c & 7;
         }
         const bool stepping = !done && kind == kKindBase;
         const uint32_t e = s_eq[code][gl];
         // ---- one column of the blocked automaton (all lanes, uniformly) ----
         const uint32_t xv = e | mv;
         const uint32_t x = e & pv;
         const uint32_t s0 = x + pv;
         // carry look-ahead over the lanes of each group: generate / propagate
         const uint32_t gen = __ballot_sync(kFull, s0 < x) & ~topmask;
         const uint32_t prop = __ballot_sync(kFull, s0 == ~0u) & ~topmask;
         const uint32_t carries = (((gen | prop) + gen) ^ prop);      // bit l = carry INTO lane l
         const uint32_t sum = s0 + ((carries >> lane) & 1u);
         const uint32_t xh = (sum ^ pv) | e;
         uint32_t ph = mv | ~(xh | pv);
         uint32_t mh = pv & xh;
         const uint32_t phm = __ballot_sync(kFull, ph >> 31);
         const uint32_t mhm = __ballot_sync(kFull, mh >> 31);
         const uint32_t ph_in = gl ? ((phm >> (lane - 1)) & 1u) : 0u;
         const uint32_t mh_in = gl ? ((mhm >> (lane - 1)) & 1u) : 0u;
         ph = (ph << 1) | ph_in;
         mh = (mh << 1) | mh_in;
         const uint32_t rise = (phm >> top) & 1u, fall = (mhm >> top) & 1u;
         if (stepping) {
            pv = mh | ~(xv | ph);
            mv = ph & xv;
            const int streak = score;
            score += (int)rise - (int)fall;
            if (MODE == M_COUNT) {
               if (score <= tau) { hit = true; done = true; }
            } else {
               if (!rise) flag = false;
               bool evt = streak <= tau && !flag && (rise || streak == 0);
               if (MODE == M_BEST) evt = evt && streak < best_d;
               if (evt) {
                  flag = true;
                  emit(p - begin, streak);
                  if (MODE == M_FIRST) done = true;
               }
            }
            p++;
         } else if (!done) {
            if (kind == kKindSkip) {
               p++;
            } else {
               if (MODE != M_COUNT) {
                  const int streak = score;
                  bool evt = streak <= tau && !flag;
                  if (MODE == M_BEST) evt = evt && streak < best_d;
                  if (evt) emit(p - begin, streak);
               }
               done = true;
            }
         }
      }
      if (line < nlines && gl == 0) {
         if (MODE == M_FIRST || MODE == M_BEST)
            a.res[line] = hit ? (((unsigned long long)(uint32_t)best_d << 32) | best_end) : kNoMatch;
         if (MODE == M_ALL) a.cnt[line] = nev;
         my_matched += hit ? 1u : 0u;
         my_events += nev;
      }
   }

   if (MODE == M_COUNT || MODE == M_COUNTALL) {
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
         unsigned long long sm = 0, se = 0;
         for (int w = 0; w < kWarps; w++) {
            sm += s_red[0][w];
            se += s_red[1][w];
         }
         if (sm) atomicAdd(&a.ctr[C_NMATCHED], sm);
         if (MODE == M_COUNTALL && se) atomicAdd(&a.ctr[C_NRECS], se);
      }
   }
}

// ===========================================================================
// K3: reverse start recovery (libseeq.c:290-316), W words per thread
// ===========================================================================
// rpat holds the masks of the REVERSED pattern and a class table in which
// every non-base byte is kKindSkip (the reference skips them all here).
// The byte-class and match-mask tables are staged in shared memory by the
// caller (rev_tables_load): indexed per thread, they would serialise in the
// constant bank.
template <int W> struct RevTables {
   uint8_t cls[256];
   uint32_t eq[8][W];
};

template <int W> __device__ __forceinline__ void rev_tables_load(RevTables<W> &t, const Pattern &rpat)
{
   for (int i = threadIdx.x; i < 256; i += blockDim.x) t.cls[i] = rpat.cls[i];
   for (int i = threadIdx.x; i < 8 * W; i += blockDim.x) t.eq[i / W][i % W] = (i / W) < 5 ? rpat.eq[i / W][i % W] : 0u;
   __syncthreads();
}

#ifndef SQB_REV_WINDOW
#define SQB_REV_WINDOW 1                   // r3a: k34_finish_lines 0.200 -> 0.179 ms on cfg2 (A/B against =0)
#endif

template <int W, int kChunk = 8>
__device__ __forceinline__ uint32_t reverse_start(const uint8_t *__restrict__ text, const uint32_t line_begin,
                                                  const uint32_t end, const int dist, const int m, const int tau,
                                                  const RevTables<W> &tab)
{
   BitVec<W> bv;
   bv_reset(bv, m);
   int score = m;
   int d = tau + 1, last_d = tau + 1;
   uint32_t j = 0, skipped = 0;
   const uint8_t *p = text + line_begin + end;           // the pass reads p[-1], p[-2], ...
   bool more = end > 0;
#if SQB_REV_WINDOW
   // (-DSQB_REV_WINDOW=0 removes it.)  The last 16 bytes in front of `end` in
   // ONE round trip -- the two aligned 16-byte vectors that hold them, realigned in registers -- instead of up to
   // four rounds of byte loads; a pass that is not over after 16 bytes carries on in the loop below.
   if (more && line_begin + end >= 16u) {
      const size_t a0 = (size_t)line_begin + end - 16u;                 // first byte of the window (offset in text)
      const uint4 *vec = reinterpret_cast<const uint4 *>(text + (a0 & ~(size_t)15));
      const uint4 v0 = __ldg(vec);                                      // (text is readable to the next multiple of 16:
      const uint4 v1 = (a0 & 15u) ? __ldg(vec + 1) : make_uint4(0u, 0u, 0u, 0u);   //  an aligned window ends with v0)
      const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      const uint32_t ws = ((uint32_t)a0 & 15u) >> 2, bs = ((uint32_t)a0 & 3u) * 8u;
      uint32_t u[5], win[4];
#pragma unroll
      for (int i = 0; i < 5; i++) {
         const uint32_t lo = (ws & 1u) ? w[i + 1] : w[i];
         const uint32_t hi = (ws & 1u) ? w[(i + 3) & 7] : w[i + 2];
         u[i] = (ws & 2u) ? hi : lo;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) win[i] = __funnelshift_r(u[i], u[i + 1], bs);   // win = bytes a0 .. a0+15
#pragma unroll
      for (int i = 0; i < 16; i++) {
         if (more) {
            j++;
            const uint8_t c = tab.cls[(win[3 - (i >> 2)] >> (8 * (3 - (i & 3)))) & 0xffu];   // byte end-1-i
            last_d = d;
            if ((c & 0x30) == kKindBase) {
               skipped = 0;
               uint32_t eq[W];
#pragma unroll
               for (int w2 = 0; w2 < W; w2++) eq[w2] = tab.eq[c & 7][w2];
               uint32_t rise, fall;
               bv_step<W>(bv, eq, rise, fall);
               score += (int)rise - (int)fall;
               d = min(score, tau + 1);
            } else {
               skipped++;
            }
            more = d > dist && j < end;
         }
      }
   }
#endif
   // the bytes are fetched eight at a time (independent loads) and then walked in
   // registers: a load per step would put the global latency on the dependency
   // chain of every step
   while (more) {
      const uint32_t left = end - j;                     // > 0
      uint8_t b[kChunk];
#pragma unroll
      for (int i = 0; i < kChunk; i++) b[i] = (uint32_t)i < left ? __ldg(p - j - 1 - i) : (uint8_t)0;
#pragma unroll
      for (int i = 0; i < kChunk; i++) {
         if (more) {
            j++;
            const uint8_t c = tab.cls[b[i]];
            last_d = d;
            if ((c & 0x30) == kKindBase) {
               skipped = 0;
               uint32_t eq[W];
#pragma unroll
               for (int w = 0; w < W; w++) eq[w] = tab.eq[c & 7][w];
               uint32_t rise, fall;
               bv_step<W>(bv, eq, rise, fall);
               score += (int)rise - (int)fall;
               d = min(score, tau + 1);
            } else {
               skipped++;
            }
            more = d > dist && j < end;
         }
      }
   }
   j = (last_d < d ? j - 1 : j) - skipped;
   return end - j;
}

// ===========================================================================
// K4: ordered compaction without inter-CTA waiting
// ===========================================================================
// Records must come out in line order.  Instead of a decoupled look-back (which
// was measured to leave the CTAs stalled on barriers most of the time), the
// lines are cut into tiles of 1024: k_tile_sums counts the records of every
// tile, k_tile_scan (one CTA) turns the counts into the first record index of
// every tile and the totals, and the finishing kernels then work tile by tile
// with nothing but a block-wide scan.
constexpr uint32_t kFinTile = 1024;

struct TileSumArgs {
   const unsigned long long *res;     // SQ_FIRST / SQ_BEST candidates (or nullptr)
   const uint32_t *cnt;               // SQ_ALL per-line event counts (or nullptr)
   uint32_t max_lines;
   unsigned long long *ctr;
   uint32_t *tile_sum;                // records per tile
   uint32_t *tile_nz;                 // lines with >= 1 record per tile
   uint32_t *tile_base;               // out of k_tile_scan
};

static __global__ void __launch_bounds__(kThreads) k_tile_sums(const TileSumArgs a)
{
   const int lane = threadIdx.x & 31;
   const uint32_t nlines = (uint32_t)min(a.ctr[C_NPSEUDO], (unsigned long long)a.max_lines);
   const uint32_t ntiles = (nlines + kFinTile - 1) / kFinTile;
   const uint32_t wid = (blockIdx.x * kThreads + threadIdx.x) >> 5, nw = (gridDim.x * kThreads) >> 5;
   for (uint32_t t = wid; t < ntiles; t += nw) {
      uint32_t sum = 0, nz = 0;
#pragma unroll 4
      for (int k = 0; k < 32; k++) {
         const uint32_t l = t * kFinTile + (uint32_t)k * 32u + (uint32_t)lane;
         if (l < nlines) {
            const uint32_t v = a.cnt ? a.cnt[l] : (a.res[l] != kNoMatch ? 1u : 0u);
            sum += v;
            nz += v != 0u;
         }
      }
      sum = __reduce_add_sync(kFull, sum);
      nz = __reduce_add_sync(kFull, nz);
      if (lane == 0) {
         a.tile_sum[t] = sum;
         a.tile_nz[t] = nz;
      }
   }
}

static __global__ void __launch_bounds__(1024) k_tile_scan(const TileSumArgs a)
{
   __shared__ CtaScanSmem cs;
   __shared__ unsigned long long s_nz[32];
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const uint32_t nlines = (uint32_t)min(a.ctr[C_NPSEUDO], (unsigned long long)a.max_lines);
   const uint32_t ntiles = (nlines + kFinTile - 1) / kFinTile;
   unsigned long long nz = 0;
   for (uint32_t t = (uint32_t)tid; t < ntiles; t += 1024u) nz += a.tile_nz[t];
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) nz += __shfl_xor_sync(kFull, nz, d);
   if (lane == 0) s_nz[warp] = nz;
   // records of a batch fit 32 bits (checked by the host)
   const unsigned long long tot = cta_scan_u32(a.tile_sum, a.tile_base, ntiles, cs);      // syncs: s_nz is visible
   if (tid == 0) {
      unsigned long long tnz = 0;
      for (int w = 0; w < 32; w++) tnz += s_nz[w];
      a.ctr[C_NRECS] = tot;
      a.ctr[C_NMATCHED] = tnz - a.ctr[C_NZ_CORR];        // SQ_ALL with cuts: lines, not segments, with records
   }
}

// ===========================================================================
// Segment cuts: one result per LINE
// ===========================================================================
// The matcher treats every segment of a cut line as a line of its own.  This
// kernel restores the line semantics before the compaction: one thread per cut
// line walks its segments in order.
//   SQ_FIRST  only the first segment with a candidate keeps it
//   SQ_BEST   the candidate with the smallest distance, the earliest on a tie
//   SQ_ALL    nothing to merge; counts the segments with events beyond the first
//             of their line (C_NZ_CORR: matched LINES = matched segments - that)
// In every mode the segments behind one that ran into a STOP byte (segstop; an
// illegal byte with SQ_FAIL, a NUL) are dead: the reference stops scanning the
// line there (libseeq.c:267-270).
struct SegReduceArgs {
   const uint32_t *lid;
   uint32_t max_lines;
   unsigned long long *ctr;
   unsigned long long *res;       // SQ_FIRST / SQ_BEST
   uint32_t *cnt;                 // SQ_ALL
   const uint8_t *segstop;
   uint8_t *deadseg;              // out (SQ_ALL): segments whose events are to be dropped
   int mode;
   const uint8_t *lflags;         // line filter: 1 = the line was dead on arrival and never scanned (or nullptr)
};

static __global__ void __launch_bounds__(kThreads) k_seg_reduce(const SegReduceArgs a)
{
   if (a.ctr[C_NCUTS] == 0ull) return;
   const uint32_t np = (uint32_t)min(a.ctr[C_NPSEUDO], (unsigned long long)a.max_lines);
   unsigned long long corr = 0;
   for (uint32_t p = blockIdx.x * kThreads + threadIdx.x; p < np; p += gridDim.x * kThreads) {
      const uint32_t me = a.lid[p];
      if ((p > 0u && a.lid[p - 1u] == me) || p + 1u >= np || a.lid[p + 1u] != me) continue;   // heads of cut lines only
      // a head segment the line filter dropped holds a STOP in its first bytes
      bool dead = a.lflags != nullptr && a.lflags[p] != 0, have = false;
      uint32_t best_q = 0, best_d = 0, with_events = 0;
      for (uint32_t q = p; q < np && a.lid[q] == me; q++) {
         if (a.mode == M_ALL) {
            if (dead) {
               a.cnt[q] = 0u;
               a.deadseg[q] = 1;
            } else if (a.cnt[q] != 0u) {
               with_events++;
            }
         } else {
            const unsigned long long key = a.res[q];
            if (key != kNoMatch) {
               const uint32_t d = (uint32_t)(key >> 32);
               bool keep = !dead;
               if (keep && have) {
                  if (a.mode == M_BEST && d < best_d) a.res[best_q] = kNoMatch;     // a strictly better one further on
                  else keep = false;
               }
               if (keep) {
                  have = true;
                  best_q = q;
                  best_d = d;
               } else {
                  a.res[q] = kNoMatch;
               }
            }
         }
         if (a.segstop[q]) dead = true;
      }
      if (with_events > 1u) corr += with_events - 1u;
   }
   if (a.mode == M_ALL) {
      corr = __reduce_add_sync(kFull, (uint32_t)corr);                  // < 2^32 per warp
      if ((threadIdx.x & 31) == 0 && corr) atomicAdd(&a.ctr[C_NZ_CORR], corr);
   }
}

// SQ_ALL: offs[line] = first record of the line (tile base + scan inside the tile)
struct OffsArgs {
   const uint32_t *cnt;
   uint32_t *offs;
   uint32_t max_lines;
   const unsigned long long *ctr;
   const uint32_t *tile_base;
};

static __global__ void __launch_bounds__(kThreads) k_offsets(const OffsArgs a)
{
   __shared__ BlockScanSmem sc;
   const uint32_t nlines = (uint32_t)min(a.ctr[C_NPSEUDO], (unsigned long long)a.max_lines);
   const uint32_t ntiles = (nlines + kFinTile - 1) / kFinTile;
   for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      uint32_t run = a.tile_base[tile];
#pragma unroll 1
      for (int k = 0; k < (int)(kFinTile / kThreads); k++) {
         const uint32_t line = tile * kFinTile + (uint32_t)k * kThreads + threadIdx.x;
         const uint32_t v = line < nlines ? a.cnt[line] : 0u;
         uint32_t total;
         const uint32_t excl = block_exclusive_scan(v, sc, &total);
         if (line < nlines) a.offs[line] = run + excl;
         run += total;
      }
   }
}

// ===========================================================================
// K3 + K4 for SQ_FIRST / SQ_BEST: one candidate per line
// ===========================================================================
struct FinArgs {
   const uint8_t *text;
   const uint32_t *ls;
   uint32_t max_lines;
   const unsigned long long *res;
   const uint32_t *offs;          // SQ_ALL only
   const Event *ev;               // SQ_ALL only
   uint32_t ev_cap;
   Rec *recs;
   uint32_t rec_cap;
   unsigned long long *ctr;
   const uint32_t *tile_base;     // first record of every 1024-line tile
   const uint32_t *lid, *lbeg;    // segment cuts: line and line start of every ls entry (or nullptr)
   const uint8_t *deadseg;        // SQ_ALL with cuts: segments whose events are dropped
   uint32_t wup;
};

// what K2 reports for entry p of ls (a line, or a segment of a cut line) in terms
// of the line: its number, its start and the offset of column 0 inside it
struct LineOf {
   uint32_t line, begin, col0;
};
__device__ __forceinline__ LineOf line_of(const FinArgs &a, uint32_t p)
{
   const uint32_t lp = a.ls[p];
   if (a.lid == nullptr || a.ctr[C_NCUTS] == 0ull) return LineOf{p, lp, 0u};
   const uint32_t line = a.lid[p], lb = a.lbeg[p];
   const bool cont = p > 0u && a.lid[p - 1u] == line;
   return LineOf{line, lb, lp - (cont ? a.wup : 0u) - lb};
}

// Per tile of 1024 lines: (1) the candidates are compacted, in line order, into
// shared memory (ballot/popc inside the warp, scan across the warps); (2) the
// threads then run the reverse pass over the DENSE candidate list -- half of the
// lines of a typical input have no match, and a thread per line would leave those
// lanes idle through the whole pass -- and write the records, already in order,
// behind the tile's first record (k_tile_sums / k_tile_scan).
template <int W, int CHUNK>
__global__ void __launch_bounds__(kThreads) k34_finish_lines(const FinArgs a, const __grid_constant__ Pattern rpat)
{
   __shared__ BlockScanSmem sc;
   __shared__ RevTables<W> tab;
   __shared__ uint4 cand[kFinTile];                       // line, line start, end, dist
   rev_tables_load(tab, rpat);
   const uint32_t nlines = (uint32_t)min(a.ctr[C_NPSEUDO], (unsigned long long)a.max_lines);
   const uint32_t ntiles = (nlines + kFinTile - 1) / kFinTile;
   const int tid = threadIdx.x;
   for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      uint32_t run = 0;
#pragma unroll 1
      for (int k = 0; k < (int)(kFinTile / kThreads); k++) {
         const uint32_t line = tile * kFinTile + (uint32_t)k * kThreads + (uint32_t)tid;
         unsigned long long key = kNoMatch;
         if (line < nlines) key = a.res[line];
         const bool valid = key != kNoMatch;
         uint32_t total;
         const uint32_t excl = block_exclusive_scan(valid ? 1u : 0u, sc, &total);
         if (valid) {
            const LineOf lo = line_of(a, line);
            cand[run + excl] = make_uint4(lo.line, lo.begin, (uint32_t)key + lo.col0, (uint32_t)(key >> 32));
         }
         run += total;
      }
      __syncthreads();
      const uint32_t base = a.tile_base[tile];
      for (uint32_t i = (uint32_t)tid; i < run; i += kThreads) {
         const uint4 c = cand[i];
         Rec r;
         r.line = c.x;
         r.end = c.z;
         r.dist = c.w;
         r.start = reverse_start<W, CHUNK>(a.text, c.y, c.z, (int)c.w, rpat.m, rpat.tau, tab);
         if (base + i < a.rec_cap) a.recs[base + i] = r;
      }
      __syncthreads();                                    // cand is reused by the next tile
   }
}

// K3 + K4 for SQ_ALL: one thread per event, destination offs[line] + rank
template <int W>
__global__ void __launch_bounds__(kThreads) k34_finish_events(const FinArgs a, const __grid_constant__ Pattern rpat)
{
   __shared__ RevTables<W> tab;
   rev_tables_load(tab, rpat);
   unsigned long long nev = a.ctr[C_EVENTS];
   if (nev > a.ev_cap) nev = a.ev_cap;
   for (unsigned long long i = (unsigned long long)blockIdx.x * kThreads + threadIdx.x; i < nev;
        i += (unsigned long long)gridDim.x * kThreads) {
      const Event e = a.ev[i];
      if (a.deadseg && a.deadseg[e.line]) continue;        // behind a STOP of its line
      const LineOf lo = line_of(a, e.line);
      Rec r;
      r.line = lo.line;
      r.end = e.end + lo.col0;
      r.dist = e.dist;
      r.start = reverse_start<W>(a.text, lo.begin, r.end, (int)e.dist, rpat.m, rpat.tau, tab);
      const unsigned long long dst = (unsigned long long)a.offs[e.line] + e.rank;
      if (dst < a.rec_cap) a.recs[dst] = r;
   }
}

}  // namespace sqb
