// sqb_k12_fused.cuh -- K1 + bit-plane pack in ONE kernel: text in, line starts and bit-planes out.
//
// The two-kernel path (k1_scan_classify -> class nibbles in HBM -> k15_pack) moves every text byte
// through DRAM three times (text in, nibbles out and in, planes out) and pays the pack's scattered
// 16-byte loads / 32-byte store pieces on the L1 data pipe.  Here a CTA stages one 32 KiB tile of text
// (plus `ov` bytes of overlap so that the tile's last line is complete) with ONE TMA bulk copy and
//
//   1. classifies it (one table look-up per byte, as K1) straight into TEXT-MAJOR BIT-PLANES in shared
//      memory: the table entry of a byte holds its three class bits and its newline bit in four BYTES,
//      one IMAD per text byte shifts it into place, and two rounds of PRMT turn the four accumulators of
//      a 32-byte chunk into the words p0, p1, p2 (bit i = class bit of byte i) and NL (bit i = byte i
//      is a newline).  NL is the line scan: popc, ONE warp-aggregated prefix, line starts allocated with
//      one atomic per tile (ls_raw, as K1);
//   2. forms GROUPS of 32 consecutive lines that start in the tile (tile-local: no CTA waits for
//      another); a warp takes a group, lane r follows line r through the three plane streams (one LDS and
//      one funnel shift per plane and 32 columns), the warp transposes 32 lines x 32 columns per plane
//      with the shuffle butterfly of k15_pack (three transposes, not four: the newline bit stays behind)
//      and stages the block in shared memory;
//   3. one bulk (TMA) store per 32 columns writes the planes of the group to HBM.
// The TMA of the next tile is issued as soon as every lane holds its text in registers and runs under
// phases 2 and 3.
//
// Plane layout of a group (allocated with one atomic per tile, any order across tiles):
//   [column block of 4][plane 0..2][4 columns] words = 48 bytes per 4 columns = 3 bits per text byte
// (the two-kernel path stores {p0,p1,p2,-}: 4 bits).  A group descriptor {K1 tile, first local entry,
// lines | columns << 8, plane offset} tells the matcher what its lane holds; the line NUMBER of slot r
// is tile_base[tile] + first + r, known after k1_scan_tiles like every line number of the scan.
// With the line filter (FASTQ-like input with -x 0) only the lines that are not dead on arrival are
// grouped; `gent` then holds the local entry index of every slot.
//
// What the kernel does not handle is detected on the device and sent back to the two-kernel path by
// the host (ctr[C_FUSED_OVF], one re-run, the engine remembers): a line that runs past the staged
// overlap, more than kFMaxEntries line starts in a tile.  Not used at all with segment cuts, FASTA
// headers, SQB_FASTQ, multi-part automata and pattern sets (sqb_engine.cu: use_fused).
//
// Replaces, like K1 + pack: the getline loop of seeqFileMatch (/root/reference/src/seeq.c:361-377)
// and the translate step of seeqStringMatch (libseeq.c:250-264).
#pragma once

#include "sqb_k2_bitslice.cuh"

namespace sqb {

constexpr uint32_t kFMaxEntries = 2048;                 // line starts per tile the fused path handles
constexpr uint32_t kFMaxGroups  = kFMaxEntries / 32;
constexpr uint32_t kFMaxOverlap = 4096;                 // bytes; one 32-byte chunk per thread at most (256 * 32 = 8192)
constexpr uint32_t kFWords      = kK1Tile / 32;         // plane words of a tile (1024)

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
   uint32_t four;                 // == 4, as a run-time value: table address = byte * four + base stays an IMAD (FMA
                                  // pipe); with the literal the assembler makes it an LEA on the ALU pipe, which
                                  // already carries one PRMT per text byte
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

// dynamic shared memory: the text stage, the plane words (3 planes), the class table, the tile's lists
__host__ __device__ constexpr uint32_t k12_text_bytes(uint32_t ov) { return (kK1Tile + ov + 16u + 127u) & ~127u; }
__host__ __device__ constexpr uint32_t k12_plane_words(uint32_t ov) { return kFWords + ov / 32u + 8u; }
__host__ __device__ constexpr uint32_t k12_smem_bytes(uint32_t ov)
{
   return k12_text_bytes(ov) + 3u * 4u * k12_plane_words(ov) + 1024u + (kFMaxEntries + 4u) * 4u + kFMaxEntries * 2u;
}

// Where text word w (32 text bytes) of the tile lives in a plane array: thread tid owns the words
// 4 tid .. 4 tid + 3 and stores word t of its four at t * 256 + tid -- lanes of a warp hit 32 banks; the
// words of the overlap follow in text order.
__device__ __forceinline__ uint32_t plane_slot(uint32_t w)
{
   return w < kFWords ? (((w & 3u) << 8) | (w >> 2)) : w;
}

__device__ __forceinline__ uint32_t lds_u32a(uint32_t addr)
{
   uint32_t v;
   asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
   return v;
}
// (not volatile: the class table is constant for the life of the kernel, the loads may be scheduled freely)
__device__ __forceinline__ uint32_t lds_u32c(uint32_t addr)
{
   uint32_t v;
   asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
   return v;
}

// 32 text bytes -> p0, p1, p2, nl (bit i = byte i).  first = the vector of bytes 0..15
// (lut = shared-memory ADDRESS of the table: the index is scaled and based by one IMAD on the FMA pipe --
// the compiler's LEA would sit on the ALU pipe, which carries the PRMTs)
__device__ __forceinline__ void classify32(const uint4 v0, const uint4 v1, const uint32_t lut, const uint32_t four,
                                           uint32_t &p0, uint32_t &p1, uint32_t &p2, uint32_t &nl)
{
   const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
   uint32_t acc[4] = {0u, 0u, 0u, 0u};                // [nl | p2 | p1 | p0] bytes of text bytes 8g .. 8g+7
#pragma unroll
   for (int g = 0; g < 4; g++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
         const uint32_t b = __byte_perm(w[2 * g + (j >> 2)], 0u, 0x4440u + (uint32_t)(j & 3));
         acc[g] = mad_u32(lds_u32c(mad_u32(b, four, lut)), 1u << j, acc[g]);
      }
   }
   const uint32_t t0 = __byte_perm(acc[0], acc[1], 0x5140u), t1 = __byte_perm(acc[2], acc[3], 0x5140u);
   const uint32_t t2 = __byte_perm(acc[0], acc[1], 0x7362u), t3 = __byte_perm(acc[2], acc[3], 0x7362u);
   p0 = __byte_perm(t0, t1, 0x5410u);
   p1 = __byte_perm(t0, t1, 0x7632u);
   p2 = __byte_perm(t2, t3, 0x5410u);
   nl = __byte_perm(t2, t3, 0x7632u);
}

// the bytes at positions >= n are STOP (p2 p1 p0 = 101) and never newlines; p = position of bit 0
__device__ __forceinline__ void stop_beyond32(uint32_t &p0, uint32_t &p1, uint32_t &p2, uint32_t &nl, uint32_t p, uint32_t n)
{
   if (p + 32u <= n) return;
   const uint32_t valid = p >= n ? 0u : ((1u << (n - p)) - 1u);       // n - p in 1..31
   p0 |= ~valid;
   p1 &= valid;
   p2 |= ~valid;
   nl &= valid;
}

__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <bool FILTER>
__global__ void __launch_bounds__(kThreads, 3) k12_scan_pack(const K12Args a, const __grid_constant__ ClassTable32 ct)
{
   extern __shared__ __align__(128) uint8_t dyn[];
   __shared__ uint64_t bar;
   __shared__ uint32_t s_tile[2], s_base, s_ovnl, s_nlive, s_pbase, s_gbase, s_skip, s_alive;
   __shared__ uint32_t s_wsum[kWarps];
   __shared__ uint32_t s_gcols[kFMaxGroups];
   __shared__ __align__(16) uint32_t s_out[kWarps][2][96];     // planes of 32 columns of one group, two buffers

   const uint32_t n = a.n, ov = a.ov;
   const uint32_t stage = kK1Tile + ov;                         // text bytes staged per tile
   const uint32_t ntiles = (n + kK1Tile - 1) / kK1Tile;
   const uint32_t n16 = (n + 15u) & ~15u;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const uint32_t pwords = k12_plane_words(ov);                 // words per plane array
   uint8_t *buf = dyn;                                          // the text of the tile (TMA)
   uint32_t *pl = reinterpret_cast<uint32_t *>(dyn + k12_text_bytes(ov));       // [3][pwords]
   uint32_t *lut = pl + 3u * pwords;                            // [256]
   uint32_t *lst = lut + 256;                                   // [entries + 1] offset in the stage | kDeadBit
   uint16_t *live = reinterpret_cast<uint16_t *>(lst + kFMaxEntries + 4u);   // FILTER: entries that are alive
   // the lane's 128 bytes are four 32-byte chunks; slot q holds chunk (q + lane) & 3, so that the 16-byte
   // loads of a quarter-warp fall into different bank groups two by two (one replay instead of seven)
   const uint32_t crot = (uint32_t)lane & 3u;
   const uint32_t lane_off = (uint32_t)warp * kK1WarpBytes + (uint32_t)lane * kK1LaneBytes;
   const bool has_ov = (uint32_t)tid * 32u < ov;
   const uint32_t last_word = kFWords + ov / 32u - 1u;          // last plane word with data
   const uint32_t pl_addr = smem_addr(pl), lut_addr = smem_addr(lut);

   // transpose constants (k15_pack)
   uint32_t keep[5], rotc[5];
   {
      const uint32_t m[5] = {0x0000FFFFu, 0x00FF00FFu, 0x0F0F0F0Fu, 0x33333333u, 0x55555555u};
#pragma unroll
      for (int s = 0; s < 5; s++) {
         const uint32_t d = 16u >> s;
         keep[s] = (lane & d) ? ~m[s] : m[s];
         rotc[s] = (lane & d) ? 32u - d : d;
      }
   }
   const uint32_t sel16 = (lane & 16) ? 0x3276u : 0x5410u, sel8 = (lane & 8) ? 0x3715u : 0x6240u;
   // where this lane's transposed words go in the staging block: lane = column; [column >> 2][plane][column & 3]
   const uint32_t out_idx = ((uint32_t)lane >> 2) * 12u + ((uint32_t)lane & 3u);

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
   bool out_pending = false;              // lane 0: bulk stores of this warp may still read s_out
   uint32_t oit = 0;                      // blocks of 32 columns this warp has staged so far (buffer = oit & 1)

   // The text stage is free again as soon as every lane holds its planes in registers (barrier A): the TMA
   // of the NEXT tile is issued there and runs under the emit and pack phases of this one.  The tile
   // numbers travel through s_tile[iteration parity]; every shared scalar is rewritten between two
   // barriers that all its readers of the round before have passed.
   for (uint32_t iter = 0;; iter++) {
      const uint32_t tile = s_tile[iter & 1u];
      if (tile >= ntiles) break;
      const uint32_t tile0 = tile * kK1Tile;
      mbar_wait(&bar, phase);
      phase ^= 1u;

      const uint32_t pos0 = tile0 + lane_off;                 // text position of this lane's first byte

      // ---- classify: slot q = chunk (q + crot) & 3 of the lane ----
      uint32_t P0[4], P1[4], P2[4], NL[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
         const uint32_t t = ((uint32_t)q + crot) & 3u;
         const uint4 *src = reinterpret_cast<const uint4 *>(buf + lane_off + (t << 5));
         classify32(src[0], src[1], lut_addr, a.four, P0[q], P1[q], P2[q], NL[q]);
      }
      if (a.skip != 0u && pos0 == 0u) {     // the bytes in front of the buffer: STOP, the last one a newline (lane 0: crot == 0)
         const uint32_t sm = (1u << a.skip) - 1u;
         P0[0] |= sm;
         P1[0] &= ~sm;
         P2[0] |= sm;
         NL[0] = (NL[0] & ~sm) | (1u << (a.skip - 1u));
      }
      if (pos0 + kK1LaneBytes > n) {
#pragma unroll
         for (int q = 0; q < 4; q++) stop_beyond32(P0[q], P1[q], P2[q], NL[q], pos0 + ((((uint32_t)q + crot) & 3u) << 5), n);
      }
      // ---- the overlap behind the tile: one 32-byte chunk per thread; only its planes and the first
      //      newline (= end of the tile's last line) are of interest ----
      uint32_t O0 = 0, O1 = 0, O2 = 0;
      if (has_ov) {
         const uint32_t o = kK1Tile + (uint32_t)tid * 32u;
         const uint4 *src = reinterpret_cast<const uint4 *>(buf + o);
         uint32_t onl;
         classify32(src[0], src[1], lut_addr, a.four, O0, O1, O2, onl);
         stop_beyond32(O0, O1, O2, onl, tile0 + o, n);
         if (onl) atomicMin(&s_ovnl, o + (uint32_t)(__ffs(onl) - 1));
      }

      // ---- line starts: c[t] = newline flags of chunk t in TEXT order (bit i = byte i) ----
      uint32_t c[4];
      {
         uint32_t g[4];
#pragma unroll
         for (int t = 0; t < 4; t++) g[t] = (crot & 1u) ? NL[(t + 3) & 3] : NL[t];
#pragma unroll
         for (int t = 0; t < 4; t++) c[t] = (crot & 2u) ? g[(t + 2) & 3] : g[t];
      }
      // a newline at p opens a line at p+1 only if p+1 < n
      if (pos0 + kK1LaneBytes + 1u > n) {
#pragma unroll
         for (int t = 0; t < 4; t++) {
            const uint32_t p = pos0 + 32u * (uint32_t)t;               // newline bit i opens a line at p + i + 1
            if (p + 33u > n) c[t] &= (p + 1u >= n) ? 0u : ((1u << (n - p - 1u)) - 1u);
         }
      }
      const uint32_t first = (tile == 0 && tid == 0 && n > 0 && a.skip == 0u) ? 1u : 0u;

      const uint32_t cnt = (uint32_t)(__popc(c[0]) + __popc(c[1]) + __popc(c[2]) + __popc(c[3])) + first;
      uint32_t inc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         const uint32_t t = __shfl_up_sync(kFull, inc, d);
         if (lane >= d) inc += t;
      }
      if (lane == 31) s_wsum[warp] = inc;
      fence_proxy_async();                   // this thread's reads of the text stage come before its refill by TMA
      __syncthreads();                       // A: the planes are in registers, s_wsum and s_ovnl are complete
      uint32_t before = 0, tile_total = 0;
#pragma unroll
      for (int w2 = 0; w2 < kWarps; w2++) {
         const uint32_t x = s_wsum[w2];
         if (w2 < warp) before += x;
         tile_total += x;
      }
      if (tid == 0) {
         const uint32_t ovnl = s_ovnl;
         s_ovnl = 0xffffffffu;
         s_alive = 0u;
         const uint32_t nt = (uint32_t)atomicAdd(&a.ctr[C_TICKET_K1], 1ull);
         s_tile[(iter + 1u) & 1u] = nt;
         if (nt < ntiles) issue(nt);
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
      // ---- the plane words of the tile (every reader of the round before has passed A) ----
#pragma unroll
      for (int q = 0; q < 4; q++) {
         const uint32_t slot = ((((uint32_t)q + crot) & 3u) << 8) + (uint32_t)tid;
         pl[slot] = P0[q];
         pl[pwords + slot] = P1[q];
         pl[2u * pwords + slot] = P2[q];
      }
      if (has_ov) {
         pl[kFWords + (uint32_t)tid] = O0;
         pl[pwords + kFWords + (uint32_t)tid] = O1;
         pl[2u * pwords + kFWords + (uint32_t)tid] = O2;
      }
      __syncthreads();                       // B: planes, s_base, s_skip
      const bool skip_tile = s_skip != 0u;

      // ---- emit: ls_raw (global, ordered inside the tile) and the tile's own list ----
      {
         uint32_t lidx = before + (inc - cnt);
         const uint32_t gbase = s_base;
         uint32_t myalive = 0;
         // FILTER: a STOP (101) among the first filter_k class codes of the line that starts at offset o
         auto dead_flag = [&](uint32_t o) -> uint32_t {
            if (!FILTER) return 0u;
            const uint32_t w = o >> 5, sh = o & 31u;
            const uint32_t i0 = plane_slot(w), i1 = plane_slot(min(w + 1u, last_word));
            const uint32_t w0 = __funnelshift_r(pl[i0], pl[i1], sh);
            const uint32_t w1 = __funnelshift_r(pl[pwords + i0], pl[pwords + i1], sh);
            const uint32_t w2 = __funnelshift_r(pl[2u * pwords + i0], pl[2u * pwords + i1], sh);
            const bool dead = (w0 & ~w1 & w2 & ((1u << a.filter_k) - 1u)) != 0u;          // filter_k <= 8
            myalive += dead ? 0u : 1u;
            return dead ? kDeadBit : 0u;
         };
         auto put = [&](uint32_t at, uint32_t o) {
            const uint32_t fl = dead_flag(o);
            if (gbase + at < a.ls_cap) a.ls_raw[gbase + at] = (tile0 + o) | fl;
            if (at < kFMaxEntries) lst[at] = o | fl;
         };
         if (first) {
            put(lidx, 0u);
            lidx++;
         }
         // one line start per lane and round, chunk by chunk in text order
         const uint32_t r1 = (uint32_t)__popc(c[0]), r2 = r1 + (uint32_t)__popc(c[1]), r3 = r2 + (uint32_t)__popc(c[2]);
         uint32_t w0 = c[0], w1 = c[1], w2 = c[2], w3 = c[3];
         for (uint32_t left = cnt - first; left != 0u; left--) {
            const uint32_t q = w0 ? 0u : (w1 ? 1u : (w2 ? 2u : 3u));
            const uint32_t cur = w0 ? w0 : (w1 ? w1 : (w2 ? w2 : w3));
            const uint32_t orig = q == 0u ? c[0] : (q == 1u ? c[1] : (q == 2u ? c[2] : c[3]));
            const uint32_t rbase = q == 0u ? 0u : (q == 1u ? r1 : (q == 2u ? r2 : r3));
            const uint32_t b = (uint32_t)__ffs(cur) - 1u;
            const uint32_t rest = cur & (cur - 1u);
            if (q == 0u) w0 = rest;
            else if (q == 1u) w1 = rest;
            else if (q == 2u) w2 = rest;
            else w3 = rest;
            put(lidx + rbase + (uint32_t)__popc(orig & ((1u << b) - 1u)), lane_off + 32u * q + b + 1u);
         }
         if (FILTER) {
            const uint32_t wa = __reduce_add_sync(kFull, myalive);
            if (lane == 0 && wa) atomicAdd(&s_alive, wa);
         }
      }
      __syncthreads();                       // C: the tile's list is complete
      if (FILTER && tid == 0) a.tile_alive[tile] = s_alive;
      if (skip_tile) continue;               // (uniform) the host repeats the scan on the two-kernel path

      // ---- line filter: the entries that are alive, in order (one warp; a few rounds) ----
      uint32_t nlive = tile_total;
      if (FILTER) {
         if (warp == 0) {
            uint32_t run = 0;
            for (uint32_t i0 = 0; i0 < tile_total; i0 += 32) {
               const uint32_t i = i0 + (uint32_t)lane;
               const bool lv = i < tile_total && !(lst[i] & kDeadBit);
               const uint32_t bal = __ballot_sync(kFull, lv);
               if (lv) live[run + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
               run += (uint32_t)__popc(bal);
            }
            if (lane == 0) s_nlive = run;
         }
         __syncthreads();                    // C2
         nlive = s_nlive;
      }
      const uint32_t ngroups = (nlive + 31u) >> 5;

      // ---- columns of every group = its longest line with the terminator ----
      for (uint32_t g = (uint32_t)warp; g < ngroups; g += kWarps) {
         const uint32_t j = g * 32u + (uint32_t)lane;
         uint32_t len = 0;
         if (j < nlive) {
            const uint32_t i = FILTER ? (uint32_t)live[j] : j;
            len = (lst[i + 1u] & ~kDeadBit) - (lst[i] & ~kDeadBit);
         }
         len = __reduce_max_sync(kFull, len);
         if (lane == 0) s_gcols[g] = len;
      }
      __syncthreads();                       // D
      // plane units (uint4) of the groups in front of group `lane` and `lane + 32` of the tile
      uint32_t u0 = (uint32_t)lane < ngroups ? 3u * ((s_gcols[lane] + 3u) >> 2) : 0u;
      uint32_t u1 = (uint32_t)lane + 32u < ngroups ? 3u * ((s_gcols[lane + 32] + 3u) >> 2) : 0u;
      uint32_t x0 = u0, x1 = u1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         const uint32_t t0 = __shfl_up_sync(kFull, x0, d), t1 = __shfl_up_sync(kFull, x1, d);
         if (lane >= d) {
            x0 += t0;
            x1 += t1;
         }
      }
      x1 += __shfl_sync(kFull, x0, 31);
      if (tid == 31) {                        // x1 of lane 31 = units of the whole tile
         const uint32_t pb = (uint32_t)min(atomicAdd(&a.ctr[C_PLANE_UNITS], (unsigned long long)x1), 0xffffffffull);
         const uint32_t gb = (uint32_t)atomicAdd(&a.ctr[C_NGROUPS], (unsigned long long)ngroups);
         s_pbase = pb;
         s_gbase = gb;
         // beyond a capacity nothing is stored; the counters go on counting and the host repeats the scan
         if ((unsigned long long)pb + x1 > a.planes_cap || gb + ngroups > a.gdesc_cap) s_skip = 1u;
      }
      __syncthreads();                       // E
      if (s_skip != 0u) continue;
      const uint32_t pbase = s_pbase, gbase = s_gbase;

      // ---- pack: one group per warp and round ----
      for (uint32_t g = (uint32_t)warp; g < ngroups; g += kWarps) {
         const uint32_t before_units = g < 32u ? __shfl_sync(kFull, x0 - u0, (int)g) : __shfl_sync(kFull, x1 - u1, (int)(g - 32u));
         const uint32_t ncols = s_gcols[g];
         const uint32_t nblk = (ncols + 3u) >> 2;                   // blocks of 4 columns
         const uint32_t j = g * 32u + (uint32_t)lane;
         const bool have = j < nlive;
         const uint32_t ent = have ? (FILTER ? (uint32_t)live[j] : j) : 0u;
         const uint32_t begin = have ? (lst[ent] & ~kDeadBit) : 0u;
         if (lane == 0) {
            GroupDesc d;
            d.tile = tile;
            d.first = FILTER ? (gbase + g) * 32u : g * 32u;
            d.meta = min(nlive - g * 32u, 32u) | (ncols << 8);
            d.poff = pbase + before_units;
            a.gdesc[gbase + g] = d;
         }
         if (FILTER) a.gent[(size_t)(gbase + g) * 32u + (uint32_t)lane] = (uint16_t)ent;
         uint4 *dst = a.planes + (size_t)pbase + before_units;
         // lane r follows line r through the plane streams: 32 columns = bits sh.. of word w and the next
         const uint32_t sh = begin & 31u;
         uint32_t w = begin >> 5;
         uint32_t a0, a1, a2;                                       // word w of the three planes
         {
            const uint32_t ad = pl_addr + (plane_slot(w) << 2);
            a0 = lds_u32a(ad);
            a1 = lds_u32a(ad + 4u * pwords);
            a2 = lds_u32a(ad + 8u * pwords);
         }
         for (uint32_t c0 = 0; c0 < ncols; c0 += 32, oit++) {
            w = min(w + 1u, last_word);
            const uint32_t ad = pl_addr + (plane_slot(w) << 2);
            const uint32_t b0 = lds_u32a(ad), b1 = lds_u32a(ad + 4u * pwords), b2 = lds_u32a(ad + 8u * pwords);
            // a lane without a line feeds STOP columns (101)
            const uint32_t x0c = have ? __funnelshift_r(a0, b0, sh) : ~0u;
            const uint32_t x1c = have ? __funnelshift_r(a1, b1, sh) : 0u;
            const uint32_t x2c = have ? __funnelshift_r(a2, b2, sh) : ~0u;
            a0 = b0;
            a1 = b1;
            a2 = b2;
            uint32_t *out = s_out[warp][oit & 1u];
            if (oit >= 2u) {                                        // the store that read this buffer two rounds ago
               if (lane == 0) bulk_wait_read_1();
               __syncwarp();
            }
            // after the transpose lane j holds column c0 + j of the plane, bit r = line r
            out[out_idx] = warp_transpose32(x0c, keep, rotc, sel16, sel8);
            out[out_idx + 4u] = warp_transpose32(x1c, keep, rotc, sel16, sel8);
            out[out_idx + 8u] = warp_transpose32(x2c, keep, rotc, sel16, sel8);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
               const uint32_t blocks = min(8u, nblk - (c0 >> 2));
               bulk_s2g(dst + (size_t)(c0 >> 2) * 3u, out, blocks * 48u);
               out_pending = true;
            }
         }
      }
   }
   if (out_pending) bulk_wait_all();
}

}  // namespace sqb
