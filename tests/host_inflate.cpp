// host_inflate.cpp -- seeq_b200/csrc/sqb_inflate.h compiled for the host: one BGZF member inflated the way the warp of
// k0_inflate_bgzf does it (lane 0's code is the shared header; the lanes' loops run one after the other here).
// Test infrastructure (tests/test_inflate_host.py compares with zlib); not part of the product.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "sqb_inflate.h"

using namespace sqb;

static uint32_t g_qcap = inf::kQueue;       // matches per queue: 32 (one member per warp) or 16 (two members per warp)
extern "C" void host_set_queue(uint32_t n) { g_qcap = n; }

static uint32_t inflate_member(const uint8_t *gz, const inf::Member &mb, uint8_t *text, uint32_t *stats)
{
   static inf::Tables t;
   uint8_t *out = text + mb.out_off;
   const uint32_t oend = mb.isize;
   inf::BitReader br;
   br.init(gz + mb.in_off, mb.in_len);
   uint32_t pos = 0, err = inf::OK, final_block = 0;
   while (!final_block && err == inf::OK) {
      uint32_t type = 0;
      err = inf::read_block_header(br, t, &type, &final_block);
      if (err != inf::OK) break;
      if (stats) stats[type]++;
      if (type == 0) {
         br.align_byte();
         const bool over = br.refill();
         const uint32_t len = br.take(16);
         const uint32_t nlen = br.take(16);
         const uint8_t *sp = br.byte_ptr();
         if (over || len != (~nlen & 0xffffu)) { err = inf::ERR_HEADER; break; }
         if (sp + len > br.src_end()) { err = inf::ERR_INPUT; break; }
         if (len > oend - pos) { err = inf::ERR_OUTPUT; break; }
         for (uint32_t lane = 0; lane < 32; lane++)
            for (uint32_t j = lane; j < len; j += 32) out[pos + j] = sp[j];
         pos += len;
         br.init(sp + len, (uint32_t)(br.src_end() - (sp + len)));
         continue;
      }
      for (uint32_t lane = 0; lane < 32; lane++)
         for (uint32_t e = lane; e < inf::kLitN; e += 32) t.lit[e] = inf::make_lit_entry(t.lcnt, t.lsym, e);
      for (uint32_t lane = 0; lane < 32; lane++)
         for (uint32_t e = lane; e < inf::kDistN; e += 32) t.dist[e] = inf::make_dist_entry(t.dcnt, t.dsym, e);
      if (stats) for (uint32_t e = 0; e < inf::kLitN; e++) stats[4 + ((t.lit[e] >> 4) & 3u)]++;
      for (;;) {
         static inf::MatchQueue q;
         uint32_t nq = 0;
         const int r = inf::run_symbols(br, t, out, pos, oend, q, g_qcap, &nq);
         if (r >= inf::R_ERR) err = (uint32_t)(r - inf::R_ERR);
         // resolve: rounds as in the kernel; inside a round the lanes run in the WORST order (last lane first, every
         // copy through a temporary written back only at the end of the round would hide nothing: copies are done in
         // place, so a wrong readiness rule shows as wrong text)
         uint32_t pending = nq >= 32 ? 0xffffffffu : (1u << nq) - 1u;
         while (pending) {
            const int f = __builtin_ctz(pending);
            const uint32_t P = q.e[f].pos;
            uint32_t rmask = 0;
            for (int lane = 31; lane >= 0; lane--) {
               if (!((pending >> lane) & 1u)) continue;
               const uint32_t mp = q.e[lane].pos, ml = q.e[lane].ld & 0xffffu, md = q.e[lane].ld >> 16;
               if (!(lane == f || inf::match_ready(mp, ml, md, P))) continue;
               rmask |= 1u << lane;
               if (inf::match_by_lane(ml, md)) { inf::copy_by_lane(out, mp, ml, md); if (stats) stats[9]++; }
               else {
                  std::vector<uint8_t> tmp(ml);
                  for (int l2 = 31; l2 >= 0; l2--)
                     for (uint32_t j = (uint32_t)l2; j < ml; j += 32)
                        tmp[j] = out[md >= ml ? mp - md + j : inf::match_src(mp, md, j)];
                  memcpy(out + mp, tmp.data(), ml);
                  if (stats) stats[8]++;
               }
            }
            pending &= ~rmask;
            if (stats) stats[10]++;
         }
         if (stats) stats[11] += nq ? 1 : 0;
         if (r != inf::R_FULL) break;
      }
   }
   if (err == inf::OK && br.overrun() > 0) err = inf::ERR_INPUT;
   if (err == inf::OK && pos != oend) err = inf::ERR_SHORT;
   return err;
}

// Inflates a BGZF buffer.  Returns the text size, or -(1 + member index) * 16 - error for the first bad member,
// or -1 when the buffer is not BGZF.  stats (may be NULL): [0..2] blocks by type, [4..7] table entries by number of
// literals, [8] warp-wide copies, [9] one-lane copies, [10] rounds, [11] queues resolved.
extern "C" long long host_bgzf_inflate(const uint8_t *gz, uint64_t nbytes, uint8_t *text, uint64_t cap, uint32_t *stats)
{
   uint64_t off = 0, o = 0;
   long long idx = 0;
   // the device reads up to 11 bytes behind a member's stream: give the host copy the same slack
   std::vector<uint8_t> copy(nbytes + 16, 0);
   memcpy(copy.data(), gz, nbytes);
   while (off < nbytes) {
      inf::Member m;
      const uint64_t next = inf::parse_member(copy.data(), nbytes, off, &m);
      if (next == 0) return -1;
      m.out_off = o;
      if (o + m.isize > cap) return -2;
      if (m.isize) {
         const uint32_t err = inflate_member(copy.data(), m, text, stats);
         if (err) return -(idx + 1) * 16 - (long long)err;
      }
      o += m.isize;
      off = next;
      idx++;
   }
   return (long long)o;
}

extern "C" uint32_t host_lit_entry_of(uint32_t s, uint32_t l) { return inf::lit_entry_of(s, l); }
extern "C" uint32_t host_dist_entry_of(uint32_t s, uint32_t l) { return inf::dist_entry_of(s, l); }
extern "C" long long host_bgzf_index(const uint8_t *gz, uint64_t nbytes, uint64_t *text_bytes)
{
   uint64_t off = 0, o = 0;
   long long n = 0;
   while (off < nbytes) {
      inf::Member m;
      const uint64_t next = inf::parse_member(gz, nbytes, off, &m);
      if (next == 0) return -1;
      o += m.isize;
      off = next;
      n++;
   }
   *text_bytes = o;
   return n;
}
