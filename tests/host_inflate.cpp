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

// ---- two members per warp: sqb_bgzf_warp.h (the body of k0_inflate_bgzf_pair) run by 32 host threads -------------
// The warp's primitives over a barrier: every lane deposits its value, all meet, every lane reads what it needs, all
// meet again (the slots are free for the next exchange).  Between two exchanges the lanes run truly side by side, in
// whatever order the host schedules them: an ordering the kernel relies on without a barrier shows as wrong text.
#include <pthread.h>

#include <thread>

#include "sqb_bgzf_warp.h"

namespace {
struct WarpShared {
   pthread_barrier_t bar;
   uint64_t slot[32];
};
struct HostWarp {
   WarpShared *sh;
   uint32_t id;
   uint32_t lane() const { return id; }
   void wait() { pthread_barrier_wait(&sh->bar); }
   template <class T> T shfl(T v, uint32_t src)
   {
      sh->slot[id] = (uint64_t)v;
      wait();
      const T r = (T)sh->slot[src & 31u];
      wait();
      return r;
   }
   uint32_t ballot(bool p)
   {
      sh->slot[id] = p ? 1u : 0u;
      wait();
      uint32_t m = 0;
      for (int i = 0; i < 32; i++) m |= (uint32_t)(sh->slot[i] & 1u) << i;
      wait();
      return m;
   }
   bool any(bool p) { return ballot(p) != 0; }
   void sync() { wait(); }
   void atomic_min(unsigned long long *p, unsigned long long v)
   {
      unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
      while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
   }
};
}  // namespace

// Inflates a BGZF buffer the way the grid of k0_inflate_bgzf_pair does, one warp after the other.  Returns the text
// size, -1 (not BGZF), -2 (no room), or -(16 * (1 + member) + error) for the first bad member by index; status (may be
// NULL) gets one word per member that carries text.
extern "C" long long host_bgzf_inflate_pair(const uint8_t *gz, uint64_t nbytes, uint8_t *text, uint64_t cap, uint32_t *status)
{
   std::vector<uint8_t> copy(nbytes + 16, 0);
   memcpy(copy.data(), gz, nbytes);
   std::vector<inf::Member> mem;
   uint64_t off = 0, o = 0;
   while (off < nbytes) {
      inf::Member m;
      const uint64_t next = inf::parse_member(copy.data(), nbytes, off, &m);
      if (next == 0) return -1;
      m.out_off = o;
      if (o + m.isize > cap) return -2;
      if (m.isize) mem.push_back(m);
      o += m.isize;
      off = next;
   }
   const uint32_t count = (uint32_t)mem.size();
   std::vector<uint32_t> st(count + 1, 0xffffffffu);
   unsigned long long first_error = ~0ull;
   static inf::Tables tables[2];
   static inf::MatchQueue queues[2];
   WarpShared sh;
   pthread_barrier_init(&sh.bar, NULL, 32);
   std::vector<std::thread> lanes;
   for (uint32_t id = 0; id < 32; id++)
      lanes.emplace_back([&, id] {
         HostWarp w{&sh, id};
         for (uint32_t pair = 0; pair * 2u < count; pair++)
            inf::inflate_pair(w, copy.data(), mem.data(), 0u, count, pair, tables, queues, text, st.data(), &first_error);
      });
   for (std::thread &t : lanes) t.join();
   pthread_barrier_destroy(&sh.bar);
   if (status) memcpy(status, st.data(), count * sizeof(uint32_t));
   if (first_error != ~0ull) return -(long long)(16ull * (1ull + (first_error >> 8)) + (first_error & 0xffull));
   for (uint32_t i = 0; i < count; i++) if (st[i] != inf::OK) return -3;             // a member nobody reported
   return (long long)o;
}
