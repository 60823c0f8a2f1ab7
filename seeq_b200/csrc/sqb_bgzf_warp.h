// sqb_bgzf_warp.h -- two BGZF members inflated by one warp: the sequence of phases of k0_inflate_bgzf_pair, written
// over the warp's primitives (lane, shuffle, vote, barrier) as a template parameter.
//
// On the device W is the hardware (sqb_bgzf.cu: DeviceWarp -- __shfl_sync, __ballot_sync, __any_sync, __syncwarp);
// tests/host_inflate.cpp runs the SAME code with 32 host threads that meet at a barrier wherever the warp's lanes
// exchange something (tests/test_inflate_host.py::test_two_members_per_warp compares with zlib).  What a lane computes
// between two exchanges is sqb_inflate.h.  See sqb_bgzf.cu for the reasons behind the scheme.
#ifndef SQB_BGZF_WARP_H_
#define SQB_BGZF_WARP_H_

#include "sqb_inflate.h"

#ifdef __CUDACC__
#define SQB_WARP_FN __device__ __forceinline__
#else
#define SQB_WARP_FN static inline
#endif

namespace sqb {
namespace inf {

constexpr uint32_t kPairQueue = 16;                      // matches per queue: one per lane of the half-warp
enum : uint32_t { ST_HEADER = 0, ST_SYMBOLS = 1, ST_DONE = 2 };

SQB_WARP_FN uint32_t lowest_bit(uint32_t x)              // x != 0
{
#ifdef __CUDA_ARCH__
   return (uint32_t)__ffs((int)x) - 1u;
#else
   return (uint32_t)__builtin_ctz(x);
#endif
}

// Members 2 * pair and 2 * pair + 1 of members[first ...) (count of them), by the warp w; t2 / q2: the tables and the
// queue of the two halves.  The caller has made sure that 2 * pair < count.
template <class W>
SQB_WARP_FN void inflate_pair(W &w, const uint8_t *gz, const Member *members, uint32_t first, uint32_t count,
                              uint32_t pair, Tables *t2, MatchQueue *q2, uint8_t *text, uint32_t *status,
                              unsigned long long *first_error)
{
   const uint32_t lane = w.lane();
   const uint32_t hl = lane & 15u, hbase = lane & 16u;   // lane in the half, first lane of the half
   const uint32_t idx = pair * 2u + (lane >> 4);
   const bool have = idx < count;                        // the upper half of the last warp may have no member
   Member mb;
   mb.in_off = 0; mb.in_len = 0; mb.isize = 0; mb.out_off = 0;
   if (have) mb = members[first + idx];
   Tables &t = t2[lane >> 4];
   MatchQueue &q = q2[lane >> 4];
   uint8_t *out = text + mb.out_off;
   const uint32_t oend = mb.isize;

   BitReader br;
   br.init(gz + mb.in_off, mb.in_len);                  // every lane holds a reader; only the decoding lane's advances
   uint32_t pos = 0, err = OK, final_block = 0, type = 0;
   uint32_t st = have ? ST_HEADER : ST_DONE;            // the same in all lanes of a half

   while (w.any(st != ST_DONE)) {
      // ---- block header: the decoding lane; stored blocks and tables: the half ----
      const bool hdr = st == ST_HEADER;
      if (w.any(hdr)) {
         uint32_t len = 0;
         unsigned long long src = 0;
         if (hdr && hl == 0) {
            err = read_block_header(br, t, &type, &final_block);
            if (err == OK && type == 0) {           // stored: LEN, ~LEN, bytes
               br.align_byte();
               const bool over = br.refill();
               len = br.take(16);
               const uint32_t nlen = br.take(16);
               const uint8_t *sp = br.byte_ptr();
               if (over || len != (~nlen & 0xffffu)) err = ERR_HEADER;
               else if (sp + len > br.src_end()) err = ERR_INPUT;
               else if (len > oend - pos) err = ERR_OUTPUT;
               src = (unsigned long long)(uintptr_t)sp;
               if (err == OK) br.init(sp + len, (uint32_t)(br.src_end() - (sp + len)));
            }
         }
         w.sync();                                   // counts and sorted symbols are the decoding lane's: publish
         err = w.shfl(err, hbase);
         type = w.shfl(type, hbase);
         final_block = w.shfl(final_block, hbase);
         len = w.shfl(len, hbase);
         src = w.shfl(src, hbase);
         if (hdr && err == OK) {
            if (type == 0) {
               const uint8_t *sp = (const uint8_t *)(uintptr_t)src;
               for (uint32_t j = hl; j < len; j += 16) out[pos + j] = sp[j];
               pos += len;
               st = final_block ? ST_DONE : ST_HEADER;
            } else {
               // the scratch arrays of the header alias t.lit: every lane has passed the barrier, nobody reads them
               for (uint32_t e = hl; e < kLitN; e += 16) t.lit[e] = make_lit_entry(t.lcnt, t.lsym, e);
               for (uint32_t e = hl; e < kDistN; e += 16) t.dist[e] = make_dist_entry(t.dcnt, t.dsym, e);
               st = ST_SYMBOLS;
            }
         } else if (hdr) st = ST_DONE;
         w.sync();                                   // tables and stored bytes: visible
      }

      // ---- symbols: the decoding lane until its queue is full or the block ends; the queue: the half ----
      const bool sym = st == ST_SYMBOLS;
      if (w.any(sym)) {
         uint32_t nq = 0;
         int r = R_EOB;
         if (sym && hl == 0) r = run_symbols(br, t, out, pos, oend, q, kPairQueue, &nq);
         w.sync();                                   // literal stores and queue entries: visible
         r = w.shfl(r, hbase);
         nq = w.shfl(nq, hbase);
         pos = w.shfl(pos, hbase);
         // resolve the queues (sqb_inflate.h: match_ready): lane i of a half owns match i of its member
         uint32_t mp = 0, ml = 0, md = 0;
         const bool owner = sym && hl < nq;
         if (owner) { const MatchQueue::Entry qe = q.e[hl]; mp = qe.pos; ml = qe.ld & 0xffffu; md = qe.ld >> 16; }
         uint32_t pending = w.ballot(owner);
         while (pending) {
            const uint32_t hp = (pending >> hbase) & 0xffffu;
            const uint32_t f = hp ? lowest_bit(hp) + hbase : lane;
            const uint32_t P = w.shfl(mp, f);
            const bool ready = ((pending >> lane) & 1u) && (lane == f || match_ready(mp, ml, md, P));
            const bool mine = ready && match_by_lane(ml, md);
            const uint32_t rmask = w.ballot(ready);
            uint32_t wide = w.ballot(ready && !mine);
            if (mine) copy_by_lane(out, mp, ml, md);
            while (wide) {                               // long or self-overlapping: the half copies it
               const uint32_t hw = (wide >> hbase) & 0xffffu;
               const uint32_t i = hw ? lowest_bit(hw) + hbase : lane;
               const uint32_t bp = w.shfl(mp, i), bl = w.shfl(ml, i), bd = w.shfl(md, i);
               if (hw) {
                  if (bd >= bl) for (uint32_t j = hl; j < bl; j += 16) out[bp + j] = out[bp - bd + j];
                  else for (uint32_t j = hl; j < bl; j += 16) out[bp + j] = out[match_src(bp, bd, j)];
               }
               const uint32_t lo = wide & 0xffffu, hi = wide & 0xffff0000u;
               wide = (lo & (lo - 1u)) | (hi & (hi - 1u));          // the lowest bit of either half is done
            }
            pending &= ~rmask;
            w.sync();                                // this round's text is final for the next round
         }
         if (sym) {
            if (r >= R_ERR) { err = (uint32_t)(r - R_ERR); st = ST_DONE; }
            else if (r == R_EOB) st = final_block ? ST_DONE : ST_HEADER;
         }
      }
   }

   if (have && hl == 0) {
      if (err == OK && br.overrun() > 0) err = ERR_INPUT;
      if (err == OK && pos != oend) err = ERR_SHORT;
      status[first + idx] = err;
      if (err != OK) w.atomic_min(first_error, ((unsigned long long)(first + idx) << 8) | err);
   }
}

}  // namespace inf
}  // namespace sqb
#endif
