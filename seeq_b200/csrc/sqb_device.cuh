// sqb_device.cuh -- device-side building blocks shared by the kernels:
//   * PTX wrappers for mbarrier + 1-D bulk (TMA) copies global -> shared
//   * block-wide exclusive scan and the decoupled look-back that chains tiles
//     into a device-wide ordered prefix (used for line numbering and for the
//     ordered write-back of records)
//   * the multi-word Myers/Hyyro bit-vector step
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace sqb {

constexpr int kThreads   = 256;          // threads per CTA in every kernel
constexpr int kWarps     = kThreads / 32;
constexpr uint32_t kFull = 0xffffffffu;

// a * b + c on the FMA pipe (IMAD), leaving the ALU pipe to the logic ops
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
   uint32_t d;
   asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
   return d;
}

// ---------------------------------------------------------------------------
// mbarrier + bulk async copy (SASS: SYNCS.*, UBLKCP)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
   return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals));
}

__device__ __forceinline__ void mbar_fence_init()
{
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                : "memory");
}

// one arrival (release at CTA scope)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// (A nanosleep back-off in this loop was tried: the retries are 15 % of K12's issued instructions, profiles/r4y, but they sit
// in otherwise idle issue slots -- 0.632 -> 0.626 ms -- and under compute-sanitizer racecheck the sleeping waiters of
// k2_forward_thread did not finish within 25 minutes, r5e.  try_wait already suspends the warp for a hardware time limit.)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
   asm volatile(
       "{\n"
       ".reg .pred p;\n"
       "WAIT_%=:\n"
       "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
       "@p bra DONE_%=;\n"
       "bra WAIT_%=;\n"
       "DONE_%=:\n"
       "}\n" ::"r"(smem_addr(bar)),
       "r"(parity)
       : "memory");
}

// dst, src 16-byte aligned; bytes a non-zero multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_addr(dst)),
                "l"(src), "r"(bytes), "r"(smem_addr(bar))
                : "memory");
}

// ---------------------------------------------------------------------------
// block-wide exclusive scan of one uint32 per thread (kThreads threads)
// ---------------------------------------------------------------------------
struct BlockScanSmem {
   uint32_t warp_sum[kWarps];
   uint32_t total;
   uint32_t base;      // result of the look-back, broadcast to the CTA
};

// returns the exclusive prefix of v inside the CTA; *total = CTA sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, BlockScanSmem &s, uint32_t *total)
{
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   uint32_t inc = v;
#pragma unroll
   for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(kFull, inc, d);
      if (lane >= d) inc += t;
   }
   if (lane == 31) s.warp_sum[warp] = inc;
   __syncthreads();
   uint32_t wbase = 0, tot = 0;
#pragma unroll
   for (int w = 0; w < kWarps; w++) {
      uint32_t x = s.warp_sum[w];
      if (w < warp) wbase += x;
      tot += x;
   }
   *total = tot;
   __syncthreads();
   return wbase + inc - v;
}

// ---------------------------------------------------------------------------
// exclusive prefix sums of a whole array by ONE CTA of 1024 threads
// ---------------------------------------------------------------------------
// The per-tile arrays of a scan (tens of thousands of entries) are turned into offsets by a
// single CTA between two grid-wide kernels, so what counts is latency: every round loads 8192
// entries with coalesced, independent loads into shared memory, each thread then owns 8
// CONTIGUOUS entries (padded index: conflict-free), one block scan of the 1024 thread sums
// places them, and the offsets leave through shared memory again, coalesced.
constexpr uint32_t kCtaScanRound = 8192;
struct CtaScanSmem {
   uint32_t stage[kCtaScanRound + kCtaScanRound / 32];
   unsigned long long warp_sum[32];
   uint32_t warp_max[32];
};

// out[i] = (uint32_t)(sum of in[0..i)); returns the total to every thread; *maxv (optional) = max in[]
__device__ __forceinline__ unsigned long long cta_scan_u32(const uint32_t *in, uint32_t *out, uint32_t n,
                                                           CtaScanSmem &sm, uint32_t *maxv = nullptr)
{
   const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;      // blockDim.x == 1024
   unsigned long long carry = 0;
   uint32_t mx = 0;
   for (uint32_t base = 0; base < n; base += kCtaScanRound) {
#pragma unroll
      for (uint32_t k = 0; k < 8; k++) {
         const uint32_t idx = k * 1024u + tid, i = base + idx;
         sm.stage[idx + (idx >> 5)] = i < n ? in[i] : 0u;
      }
      __syncthreads();
      uint32_t loc[8];
      unsigned long long sum = 0;
#pragma unroll
      for (uint32_t k = 0; k < 8; k++) {
         const uint32_t idx = tid * 8u + k;
         loc[k] = sm.stage[idx + (idx >> 5)];
         sum += loc[k];
         mx = max(mx, loc[k]);
      }
      unsigned long long x = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         const unsigned long long y = __shfl_up_sync(kFull, x, d);
         if (lane >= (uint32_t)d) x += y;
      }
      if (lane == 31u) sm.warp_sum[warp] = x;
      __syncthreads();
      unsigned long long before = 0, tot = 0;
#pragma unroll 8
      for (uint32_t w = 0; w < 32; w++) {
         const unsigned long long y = sm.warp_sum[w];
         if (w < warp) before += y;
         tot += y;
      }
      unsigned long long run = carry + before + x - sum;
#pragma unroll
      for (uint32_t k = 0; k < 8; k++) {
         const uint32_t idx = tid * 8u + k;
         sm.stage[idx + (idx >> 5)] = (uint32_t)run;
         run += loc[k];
      }
      __syncthreads();
#pragma unroll
      for (uint32_t k = 0; k < 8; k++) {
         const uint32_t idx = k * 1024u + tid, i = base + idx;
         if (i < n) out[i] = sm.stage[idx + (idx >> 5)];
      }
      carry += tot;
      __syncthreads();
   }
   if (maxv) {
      mx = __reduce_max_sync(kFull, mx);
      if (lane == 0u) sm.warp_max[warp] = mx;
      __syncthreads();
      uint32_t m2 = 0;
      for (uint32_t w = 0; w < 32; w++) m2 = max(m2, sm.warp_max[w]);
      *maxv = m2;
      __syncthreads();
   }
   return carry;
}

// ---------------------------------------------------------------------------
// decoupled look-back: tile `tile` publishes its aggregate and obtains the sum
// of the aggregates of all earlier tiles.  status[] must be zero before the
// launch.  Word layout: bits 63:62 = 0 empty / 1 aggregate / 2 inclusive
// prefix, bits 61:0 = value.  Tiles must be taken in ticket order (atomic
// counter) so that every predecessor is already running.
// Called by ALL threads of the CTA; returns the exclusive prefix of the tile.
// ---------------------------------------------------------------------------
constexpr unsigned long long kStAggregate = 1ull << 62;
constexpr unsigned long long kStPrefix    = 2ull << 62;
constexpr unsigned long long kStValue     = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
   unsigned long long v;
   asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}

__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v)
{
   asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long tile_lookback(unsigned long long *status, uint32_t tile,
                                                            unsigned long long aggregate,
                                                            unsigned long long *s_base)
{
   if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      if (lane == 0)
         st_status(status + tile, (tile == 0 ? kStPrefix : kStAggregate) | aggregate);
      unsigned long long excl = 0;
      if (tile > 0) {
         long long look = (long long)tile - 1;      // newest tile inspected by lane 0
         while (true) {
            const long long idx = look - lane;
            unsigned long long w = kStPrefix;        // tiles before 0 contribute nothing
            if (idx >= 0) {
               do {
                  w = ld_status(status + idx);
               } while ((w >> 62) == 0);
            }
            const uint32_t is_prefix = __ballot_sync(kFull, (w >> 62) == 2);
            // lanes up to and including the first prefix take part
            const int stop = is_prefix ? __ffs(is_prefix) - 1 : 31;
            unsigned long long val = lane <= stop ? (w & kStValue) : 0ull;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(kFull, val, d);
            excl += val;
            if (is_prefix) break;
            look -= 32;
         }
         if (lane == 0) st_status(status + tile, kStPrefix | (excl + aggregate));
      }
      if (lane == 0) *s_base = excl;
   }
   __syncthreads();
   const unsigned long long r = *s_base;
   __syncthreads();
   return r;
}

// Same, with ONE barrier: the caller provides a broadcast slot that is not
// reused before the CTA has passed another barrier (e.g. one slot per stage).
__device__ __forceinline__ unsigned long long tile_lookback1(unsigned long long *status, uint32_t tile,
                                                             unsigned long long aggregate,
                                                             unsigned long long *s_base)
{
   if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      if (lane == 0)
         st_status(status + tile, (tile == 0 ? kStPrefix : kStAggregate) | aggregate);
      unsigned long long excl = 0;
      if (tile > 0) {
         long long look = (long long)tile - 1;
         while (true) {
            const long long idx = look - lane;
            unsigned long long w = kStPrefix;
            if (idx >= 0) {
               do {
                  w = ld_status(status + idx);
               } while ((w >> 62) == 0);
            }
            const uint32_t is_prefix = __ballot_sync(kFull, (w >> 62) == 2);
            const int stop = is_prefix ? __ffs(is_prefix) - 1 : 31;
            unsigned long long val = lane <= stop ? (w & kStValue) : 0ull;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(kFull, val, d);
            excl += val;
            if (is_prefix) break;
            look -= 32;
         }
         if (lane == 0) st_status(status + tile, kStPrefix | (excl + aggregate));
      }
      if (lane == 0) *s_base = excl;
   }
   __syncthreads();
   return *s_base;
}

// ---------------------------------------------------------------------------
// Myers/Hyyro bit-vector automaton over W 32-bit words held by one thread.
//
// The pattern is LEFT-aligned in the W*32-bit vector: position j of the
// pattern is bit pad+j (pad = W*32 - m), so the horizontal delta of the last
// pattern row is always the top bit and falls off the vector when shifted.
// The pad low bits are wildcard rows (Eq = 1 for every base) whose vertical
// deltas start at 0; they stay exactly 0 (D = 0 along those rows), which makes
// row pad behave like the free-start row of the search recurrence.
// ---------------------------------------------------------------------------
template <int W> struct BitVec {
   uint32_t pv[W];
   uint32_t mv[W];
};

template <int W> __device__ __forceinline__ void bv_reset(BitVec<W> &s, int m)
{
   const int pad = W * 32 - m;
#pragma unroll
   for (int w = 0; w < W; w++) {
      const int lo = w * 32;
      s.pv[w] = pad <= lo ? ~0u : (pad >= lo + 32 ? 0u : (~0u << (pad - lo)));
      s.mv[w] = 0u;
   }
}

// One text byte.  eq[w] = match mask of the byte's class.  rise / fall are the
// +1 / -1 horizontal deltas of the last pattern row: score += rise - fall.
template <int W>
__device__ __forceinline__ void bv_step(BitVec<W> &s, const uint32_t *eq, uint32_t &rise, uint32_t &fall)
{
   uint32_t carry = 0, ph_in = 0, mh_in = 0;
#pragma unroll
   for (int w = 0; w < W; w++) {
      const uint32_t e = eq[w], pv = s.pv[w], mv = s.mv[w];
      const uint32_t xv = e | mv;
      const uint32_t a = e & pv;
      uint32_t sum;
      if (W == 1) {
         sum = a + pv;
      } else {
         const unsigned long long t = (unsigned long long)a + pv + carry;
         sum = (uint32_t)t;
         carry = (uint32_t)(t >> 32);
      }
      const uint32_t xh = (sum ^ pv) | e;
      uint32_t ph = mv | ~(xh | pv);
      uint32_t mh = pv & xh;
      const uint32_t ph_out = ph >> 31, mh_out = mh >> 31;
      ph = (ph << 1) | ph_in;
      mh = (mh << 1) | mh_in;
      ph_in = ph_out;
      mh_in = mh_out;
      s.pv[w] = mh | ~(xv | ph);
      s.mv[w] = ph & xv;
   }
   rise = ph_in;
   fall = mh_in;
}

}  // namespace sqb
