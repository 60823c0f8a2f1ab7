// sqb_k12_fused.cuh -- K1 + bit-plane pack in ONE kernel: text in, line starts and bit-planes out.
//
// The two-kernel path (k1_scan_classify -> class nibbles in HBM -> k15_pack) moves every text byte
// through DRAM three times (text in, nibbles out and in, planes out) and pays the pack's scattered
// 16-byte loads / 32-byte store pieces on the L1 data pipe.  Here a CTA stages one 32 KiB tile of text
// (plus `ov` bytes of overlap so that the tile's last line is complete) with ONE TMA bulk copy and
//
//   1. classifies it exactly like K1 (one table look-up per byte, newline flags, warp-aggregated
//      prefix, line starts allocated with one atomic per tile) -- the class nibbles stay in SHARED
//      memory, written contiguously over the tile's own text;
//   2. forms GROUPS of 32 consecutive lines that start in the tile (tile-local: no CTA waits for
//      another); a warp takes a group, lane r streams the nibbles of line r out of shared memory
//      (16-byte LDS, funnel-shift realign), the warp transposes 32 lines x 32 columns with the
//      shuffle butterfly of k15_pack and stages the three planes of the block in shared memory;
//   3. one bulk (TMA) store per 32 columns writes the planes of the group to HBM.
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
constexpr uint32_t kFMaxOverlap = kThreads * 16;        // one 16-byte vector per thread

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
   uint32_t ov;                   // overlap bytes staged behind a tile: multiple of 16, 16 .. kFMaxOverlap
   GroupDesc *gdesc;              // out
   uint32_t gdesc_cap;
   uint16_t *gent;                // out (FILTER): [group * 32 + slot] local entry index
   uint4 *planes;                 // out
   uint32_t planes_cap;           // uint4 units
};

// dynamic shared memory: the text stage, the nibble array, the class table, the tile's lists
__host__ __device__ constexpr uint32_t k12_text_bytes(uint32_t ov) { return (kK1Tile + ov + 16u + 127u) & ~127u; }
__host__ __device__ constexpr uint32_t k12_nib_bytes(uint32_t ov) { return ((kK1Tile + ov) / 2u + 16u + 127u) & ~127u; }
__host__ __device__ constexpr uint32_t k12_smem_bytes(uint32_t ov)
{
   return k12_text_bytes(ov) + k12_nib_bytes(ov) + 256u + (kFMaxEntries + 4u) * 4u + kFMaxEntries * 2u;
}

__device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
   uint4 v;
   asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
   return v;
}

// nibble stream of one line out of the tile's nibble array in shared memory (cf. NibbleStream)
struct SmemNibbleStream {
   uint32_t base;                 // shared-memory address of the nibble array
   uint32_t chunk, last;          // next 16-byte chunk, last valid chunk
   uint32_t ws, bs;
   uint4 prev, cur;
   bool valid;

   __device__ __forceinline__ uint4 load()
   {
      uint4 v = cur;
      if (valid) v = lds_v4(base + (chunk << 4));
      chunk = min(chunk + 1u, last);
      return v;
   }
   __device__ __forceinline__ void open(uint32_t smem_base, uint32_t nchunks, uint32_t begin, bool ok)
   {
      base = smem_base;
      last = nchunks - 1u;
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

// class nibbles of one 16-byte vector: lo = bytes 0..7, hi = bytes 8..15 (nibble j = byte j)
__device__ __forceinline__ void classify16(const uint4 v, const uint8_t *lut, uint32_t &lo, uint32_t &hi)
{
   const uint32_t w[4] = {v.x, v.y, v.z, v.w};
   uint32_t l = 0, h = 0;
#pragma unroll
   for (int j = 0; j < 8; j++) {
      const uint32_t bx = __byte_perm(w[j >> 2], 0u, 0x4440u + (uint32_t)(j & 3));
      const uint32_t by = __byte_perm(w[2 + (j >> 2)], 0u, 0x4440u + (uint32_t)(j & 3));
      l = mad_u32((uint32_t)lut[bx], 1u << (4 * j), l);
      h = mad_u32((uint32_t)lut[by], 1u << (4 * j), h);
   }
   lo = l;
   hi = h;
}

// nibbles of the bytes at positions >= n become STOP (never a newline); p = position of byte 0
__device__ __forceinline__ void stop_beyond(uint32_t &lo, uint32_t &hi, uint32_t p, uint32_t n)
{
#pragma unroll
   for (int j = 0; j < 8; j++) {
      if (p + (uint32_t)j >= n) lo = (lo & ~(0xFu << (4 * j))) | ((uint32_t)kClsStop << (4 * j));
      if (p + 8u + (uint32_t)j >= n) hi = (hi & ~(0xFu << (4 * j))) | ((uint32_t)kClsStop << (4 * j));
   }
}

__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <bool FILTER>
__global__ void __launch_bounds__(kThreads, 3) k12_scan_pack(const K12Args a, const __grid_constant__ ClassTable ct)
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
   uint8_t *buf = dyn;                                          // the text of the tile (TMA)
   uint8_t *nib = dyn + k12_text_bytes(ov);                     // its class nibbles, one contiguous array
   uint8_t *lut = nib + k12_nib_bytes(ov);
   uint32_t *lst = reinterpret_cast<uint32_t *>(lut + 256);     // [entries + 1] offset in the stage | kDeadBit
   uint16_t *live = reinterpret_cast<uint16_t *>(lst + kFMaxEntries + 4u);   // FILTER: entries that are alive
   const uint32_t rot = (uint32_t)lane & 7u;
   const uint32_t lane_off = (uint32_t)warp * kK1WarpBytes + (uint32_t)lane * kK1LaneBytes;
   const uint32_t nib_off = lane_off >> 1;                      // this lane's nibbles in the contiguous array
   const bool has_ov = (uint32_t)tid * 16u < ov;
   const uint32_t nib_chunks = stage >> 5;                      // 16-byte chunks of the nibble array

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

   auto issue = [&](uint32_t t) {          // (tid 0) TMA of tile t into the text stage
      const uint32_t start = t * kK1Tile;
      uint32_t bytes = n16 - start;
      if (bytes > stage) bytes = stage;
      mbar_expect_tx(&bar, bytes);
      bulk_g2s(buf, a.text + start, bytes, &bar);
   };
   lut[tid] = ct.code[tid];
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

   // The text stage is free again as soon as every lane holds its text in registers (barrier A): the TMA
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

      // ---- classify: vector slot k holds text vector (k + rot) & 7 of the lane (as K1) ----
      uint32_t lo[8], hi[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
         const uint4 v = *reinterpret_cast<const uint4 *>(buf + lane_off + ((((uint32_t)k + rot) & 7u) << 4));
         classify16(v, lut, lo[k], hi[k]);
      }
      if (a.skip != 0u && pos0 == 0u) {     // the bytes in front of the buffer: STOP, the last one a newline
         const uint32_t nl = a.skip - 1u;
#pragma unroll
         for (int j = 0; j < 8; j++) {
            if ((uint32_t)j <= nl)
               lo[0] = (lo[0] & ~(0xFu << (4 * j))) | ((uint32_t)(kClsStop | ((uint32_t)j == nl ? kClsNewline : 0)) << (4 * j));
            if (8u + (uint32_t)j <= nl)
               hi[0] = (hi[0] & ~(0xFu << (4 * j))) | ((uint32_t)(kClsStop | (8u + (uint32_t)j == nl ? kClsNewline : 0)) << (4 * j));
         }
      }
      if (pos0 + kK1LaneBytes > n) {
#pragma unroll
         for (int k = 0; k < 8; k++) stop_beyond(lo[k], hi[k], pos0 + ((((uint32_t)k + rot) & 7u) << 4), n);
      }
      // ---- the overlap behind the tile: one vector per thread; only its nibbles and the first
      //      newline (= end of the tile's last line) are of interest ----
      uint32_t olo = 0, ohi = 0;
      if (has_ov) {
         const uint32_t o = kK1Tile + (uint32_t)tid * 16u;
         classify16(*reinterpret_cast<const uint4 *>(buf + o), lut, olo, ohi);
         if (tile0 + o + 16u > n) stop_beyond(olo, ohi, tile0 + o, n);
         const uint32_t fl = olo & 0x88888888u, fh = ohi & 0x88888888u;
         if (fl | fh) atomicMin(&s_ovnl, o + (fl ? (uint32_t)(__ffs(fl) - 1) >> 2 : 8u + ((uint32_t)(__ffs(fh) - 1) >> 2)));
      }

      // ---- newline flags (as K1) ----
      uint32_t f[8];
      {
         uint32_t g[8], h2[8];
#pragma unroll
         for (int k = 0; k < 8; k++) g[k] = ((lo[k] & 0x88888888u) >> 1) | (hi[k] & 0x88888888u);
#pragma unroll
         for (int v = 0; v < 8; v++) h2[v] = (rot & 1u) ? g[(v + 7) & 7] : g[v];
#pragma unroll
         for (int v = 0; v < 8; v++) g[v] = (rot & 2u) ? h2[(v + 6) & 7] : h2[v];
#pragma unroll
         for (int v = 0; v < 8; v++) f[v] = (rot & 4u) ? g[(v + 4) & 7] : g[v];
      }
      uint32_t c[4];
#pragma unroll
      for (int q = 0; q < 4; q++) c[q] = (f[2 * q] >> 2) | f[2 * q + 1];
      // a newline at p opens a line at p+1 only if p+1 < n
      if (pos0 + kK1LaneBytes + 1u > n) {
#pragma unroll
         for (int q = 0; q < 4; q++) {
            uint32_t bits = c[q];
            while (bits) {
               const int b = __ffs(bits) - 1;
               bits &= bits - 1;
               const uint32_t byte = 32u * (uint32_t)q + 8u * (uint32_t)(b & 3) + (uint32_t)(b >> 2);
               if (pos0 + byte + 1u >= n) c[q] &= ~(1u << b);
            }
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
      __syncthreads();                       // A: the text is in registers, s_wsum and s_ovnl are complete
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
      // ---- class nibbles: one contiguous array over the tile's own text ----
#pragma unroll
      for (int k = 0; k < 8; k++)
         *reinterpret_cast<uint2 *>(nib + nib_off + ((((uint32_t)k + rot) & 7u) << 3)) = make_uint2(lo[k], hi[k]);
      if (has_ov) *reinterpret_cast<uint2 *>(nib + (kK1Tile >> 1) + (uint32_t)tid * 8u) = make_uint2(olo, ohi);
      __syncthreads();                       // B: nibbles, s_base, s_skip
      const bool skip_tile = s_skip != 0u;

      // ---- emit: ls_raw (global, ordered inside the tile) and the tile's own list ----
      {
         uint32_t lidx = before + (inc - cnt);
         const uint32_t gbase = s_base;
         uint32_t myalive = 0;
         auto dead_flag = [&](uint32_t o) -> uint32_t {       // o = offset of the line start in the stage
            if (!FILTER) return 0u;
            const uint32_t a0 = (o >> 1) & ~3u;
            const uint32_t w0 = *reinterpret_cast<const uint32_t *>(nib + a0);
            const uint32_t w1 = *reinterpret_cast<const uint32_t *>(nib + a0 + 4u);
            uint32_t win = __funnelshift_r(w0, w1, (o & 7u) * 4u);
            if (a.filter_k < 8u) win &= (1u << (4u * a.filter_k)) - 1u;
            const uint32_t y = (win ^ 0x55555555u) & 0x77777777u;
            const bool dead = ((y - 0x11111111u) & ~y & 0x88888888u) != 0u;
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
            const uint32_t before_b = (0x11111111u * ((1u << e) - 1u)) | ((0x11111111u << e) & ((1u << (b & ~3u)) - 1u));
            put(lidx + rbase + (uint32_t)__popc(orig & before_b), lane_off + 32u * q + 1u + 8u * e + (b >> 2));
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
         SmemNibbleStream st;
         st.open(smem_addr(nib), nib_chunks, begin, have);
         for (uint32_t c0 = 0; c0 < ncols; c0 += 32, oit++) {
            uint32_t w[4];
            st.next(w);
            uint32_t *out = s_out[warp][oit & 1u];
            if (oit >= 2u) {                                        // the store that read this buffer two rounds ago
               if (lane == 0) bulk_wait_read_1();
               __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
               const uint32_t t = warp_transpose32(w[k], keep, rotc, sel16, sel8);
               // lane holds plane (lane & 3) of column c0 + 8k + (lane >> 2)
               const uint32_t cc = 8u * (uint32_t)k + ((uint32_t)lane >> 2);
               if ((lane & 3) != 3) out[(cc >> 2) * 12u + ((uint32_t)lane & 3u) * 4u + (cc & 3u)] = t;
            }
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
