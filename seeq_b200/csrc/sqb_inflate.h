// sqb_inflate.h -- DEFLATE (RFC 1951) decoding of one BGZF block, in a header that compiles for the host as well.
//
// The reference reads plain text only (seeq.c:201-256 opens the file, :361 getline); compressed read sets are
// SURVEY §8f row 3 ("input formats").  BGZF (bgzip, the blocked gzip of htslib; SAM/BAM specification §4.1) is a
// series of gzip members of at most 64 KiB of text each, every member carrying its compressed size in a "BC" extra
// sub-field: the members are found without decoding and are inflated independently -- one WARP per member on the
// device (sqb_bgzf.cu: k0_inflate_bgzf), which writes the text straight into the HBM buffer K12 / K1 scan.
//
// What is here is everything a lane computes: the bit reader, the block headers, the canonical-code decoder, the
// construction of the look-up tables entry by entry (on the device the lanes of the warp take entries 32 apart), the
// symbol loop that lane 0 runs until its queue of matches is full, and the rule by which the warp copies them.  tests/host_inflate.cpp drives the same code on the
// CPU against zlib (tests/test_inflate_host.py); the kernel adds the shuffles and the warp-wide copies only.
//
// Decoding follows RFC 1951 §3.2; the canonical decoder (count of codes per length + symbols sorted by length) is
// the published scheme of RFC 1951 §3.2.2.  The tables are this kernel's own: a literal/length table of 2^TB
// 32-bit entries that yields UP TO THREE LITERALS per look-up (DNA text under a dynamic code is 2-3 bits per base:
// one shared-memory load decodes three bases), and a distance table of 2^TBD entries with base and extra-bit count
// resolved at build time.
#ifndef SQB_INFLATE_H_
#define SQB_INFLATE_H_

#include <stdint.h>

#ifdef __CUDACC__
#define SQB_INF_HD __host__ __device__ __forceinline__
#define SQB_INF_M __host__ __device__ __forceinline__
#else
#define SQB_INF_HD static inline
#define SQB_INF_M inline
#endif
#ifdef __CUDA_ARCH__
#define SQB_INF_LOAD(p) __ldg(p)              /* the compressed bytes are never written by the kernel */
#else
#define SQB_INF_LOAD(p) (*(p))
#endif
#define SQB_INF_RARE(x) __builtin_expect(!!(x), 0)

// Table look-ups and queue entries of the symbol loop.  On the device they go through 32-bit shared-memory addresses
// taken once in front of the loop: a generic pointer to shared memory is rebuilt from the CTA's shared window when
// registers are short (an S2R in the loop).
#ifdef __CUDA_ARCH__
namespace sqb { namespace inf {
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
   uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
   asm volatile("" : "+r"(a));                // opaque: kept in a register, not recomputed in the loop
   return a;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
   uint32_t v;
   asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
   return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y)
{
   asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
} }
#endif

#ifndef SQB_INF_TB
#define SQB_INF_TB 9       /* 512 entries: 2 KiB.  Measured on the CPU (tests/host_inflate.cpp counts the symbols that leave
                              the fast path): 1.5 % of the symbols of deflated DNA and FASTQ text with 9 bits as with 10 */
#endif

namespace sqb {
namespace inf {

constexpr int TB = SQB_INF_TB;                      // index bits of the literal/length table
constexpr int TBD = 8;                       // index bits of the distance table
constexpr uint32_t kLitN = 1u << TB;
constexpr uint32_t kDistN = 1u << TBD;

// status of a member (0 = inflated, ISIZE bytes written)
enum : uint32_t {
   OK = 0,
   ERR_INPUT = 1,      // the bit stream runs past the member's compressed data
   ERR_OUTPUT = 2,     // more text than ISIZE announces
   ERR_CODE = 3,       // a bit pattern that is no code word / an invalid symbol / an over-subscribed code
   ERR_DIST = 4,       // a distance that reaches in front of the member's text
   ERR_HEADER = 5,     // block type 3, stored-block length check, repeat with nothing to repeat, no end-of-block code
   ERR_SHORT = 6       // end of the stream with less text than ISIZE announces
};

// what the symbol loop returns: end of block, queue full, or R_ERR + one of the errors above
enum : int { R_EOB = 0, R_FULL = 1, R_ERR = 16 };

// ---- literal/length entries -------------------------------------------------------------------------------------
//   [3:0]   bits to drop: all code words of the literals of the entry, or the code word of the length / end symbol
//   [5:4]   literals in the entry (1..3); 0 = not a literal:
//   [7:6]      0 length  1 end of block  2 code word longer than TB bits (canonical decoder)  3 invalid
//   literals: [15:8] [23:16] [31:24];  length: [16:8] base (3..258), [19:17] extra bits
constexpr uint32_t K_LEN = 0u << 6, K_EOB = 1u << 6, K_LONG = 2u << 6, K_BAD = 3u << 6;
// ---- distance entries:  [3:0] bits of the code word  [7:4] extra bits  [8] long  [9] invalid  [31:16] base
constexpr uint32_t D_LONG = 1u << 8, D_BAD = 1u << 9;

// The tables of one warp (shared memory on the device).  While a dynamic block's header is decoded the first
// bytes of lit[] hold the scratch arrays (code lengths etc.): they are dead before the first entry is written.
struct Tables {
   uint32_t lit[kLitN];
   uint32_t dist[kDistN];
   uint16_t lcnt[16], dcnt[16];             // codes per length
   uint16_t lsym[288];                       // symbols sorted by (length, symbol)
   uint16_t dsym[32];
};

struct Scratch {                             // aliases Tables::lit
   uint8_t  lens[320];                       // code lengths: HLIT literal/length codes, then HDIST distance codes
   uint8_t  cl[20];                          // lengths of the 19 code-length codes
   uint16_t clcnt[16];
   uint16_t clsym[20];
   uint16_t offs[16];
};

static_assert(sizeof(Scratch) <= sizeof(Tables::lit), "the header's scratch arrays alias the literal/length table");

// ---- bit reader: LSB-first, 64-bit window, aligned 32-bit loads, the next word already on its way ---------------
struct BitReader {
   const uint32_t *w0;                       // the aligned word that holds the first byte of the stream
   uint32_t wi;                              // index (from w0) of the word AFTER `next`
   uint32_t wend;                            // words of this index or beyond lie wholly behind the compressed data
   uint32_t end_off;                         // bytes from w0 to one past the last byte of compressed data
   uint64_t buf;
   uint32_t next;                            // word w0[wi - 1], not yet in buf
   int cnt;                                  // valid bits in buf

   // src[0 .. nbytes) is the stream; up to 11 bytes behind it are READ (never used): in a BGZF file the member's
   // trailer and the next header lie there
   SQB_INF_M void init(const uint8_t *src, uint32_t nbytes)
   {
      const uintptr_t a = (uintptr_t)src;
      const uint32_t lead = (uint32_t)(a & 3u);
      w0 = (const uint32_t *)(a - lead);
      end_off = lead + nbytes;
      wend = (end_off + 3u) >> 2;
      buf = (uint64_t)(SQB_INF_LOAD(w0) >> (8u * lead));
      cnt = 32 - 8 * (int)lead;
      next = SQB_INF_LOAD(w0 + 1);
      wi = 2;
   }
   SQB_INF_M const uint8_t *src_end() const { return (const uint8_t *)w0 + end_off; }
   // at least 33 valid bits afterwards; true when the refill wanted a word two words or more behind the data: the
   // stream is exhausted and every caller stops at once.  A well-formed stream never asks for more once the window
   // holds bits of words behind the data only; bits of the one garbage word that can enter the window before that
   // are caught by overrun() at the end of the member.
   SQB_INF_M bool refill()
   {
      if (cnt <= 32) {
         buf |= (uint64_t)next << cnt;
         cnt += 32;
         if (SQB_INF_RARE(wi >= wend + 2u)) return true;
         next = SQB_INF_LOAD(w0 + wi);
         wi++;
      }
      return false;
   }
   SQB_INF_M void drop(uint32_t n) { buf >>= n; cnt -= (int)n; }
   SQB_INF_M uint32_t take(uint32_t n)
   {
      const uint32_t v = (uint32_t)buf & ((1u << n) - 1u);
      drop(n);
      return v;
   }
   // 32 bits of the window from bit n on (one funnel shift on the device)
   SQB_INF_M uint32_t peek_from(uint32_t n) const { return (uint32_t)(buf >> n); }
   // bits consumed beyond the end of the compressed data (> 0: the stream is truncated or corrupt)
   SQB_INF_M long overrun() const { return ((long)(wi - 1u) * 4L - (long)end_off) * 8L - (long)cnt; }
   SQB_INF_M void align_byte() { drop((uint32_t)cnt & 7u); }
   // the byte the window starts at (after align_byte)
   SQB_INF_M const uint8_t *byte_ptr() const { return (const uint8_t *)(w0 + (wi - 1u)) - (cnt >> 3); }
};

// ---- canonical decoder (RFC 1951 §3.2.2): one symbol from the low bits of `bits`, code words up to maxlen bits ----
// returns symbol | length << 16, or -1 when no code word of at most maxlen bits matches
SQB_INF_HD int canon_decode(const uint16_t *cnt, const uint16_t *sym, uint32_t bits, int maxlen)
{
   int code = 0, first = 0, index = 0;
   for (int len = 1; len <= maxlen; len++) {
      code |= (int)(bits & 1u);
      bits >>= 1;
      const int count = cnt[len];
      if (code - count < first) return (int)sym[index + (code - first)] | (len << 16);
      index += count;
      first += count;
      first <<= 1;
      code <<= 1;
   }
   return -1;
}

// counts and sorted symbols of a code given by its lengths; false if the code is over-subscribed
SQB_INF_HD bool build_canon(const uint8_t *lens, int n, uint16_t *cnt, uint16_t *sym, uint16_t *offs)
{
   for (int l = 0; l < 16; l++) cnt[l] = 0;
   for (int s = 0; s < n; s++) cnt[lens[s]]++;
   int left = 1;
   for (int l = 1; l < 16; l++) {
      left <<= 1;
      left -= (int)cnt[l];
      if (left < 0) return false;
   }
   offs[1] = 0;
   for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + cnt[l]);
   for (int s = 0; s < n; s++)
      if (lens[s]) sym[offs[lens[s]]++] = (uint16_t)s;
   cnt[0] = 0;                               // canon_decode never looks at length 0
   return true;
}

// entry of ONE literal/length symbol whose code word has l bits
SQB_INF_HD uint32_t lit_entry_of(uint32_t s, uint32_t l)
{
   if (s < 256u) return l | (1u << 4) | (s << 8);
   if (s == 256u) return l | K_EOB;
   if (s > 285u) return l | K_BAD;
   uint32_t base, extra;
   if (s < 265u) { base = s - 254u; extra = 0; }
   else if (s == 285u) { base = 258u; extra = 0; }
   else { extra = (s - 261u) >> 2; base = ((4u + ((s - 265u) & 3u)) << extra) + 3u; }
   return l | K_LEN | (base << 8) | (extra << 17);
}

SQB_INF_HD uint32_t dist_entry_of(uint32_t s, uint32_t l)
{
   if (s > 29u) return l | D_BAD;
   uint32_t base, extra;
   if (s < 4u) { base = s + 1u; extra = 0; }
   else { extra = (s >> 1) - 1u; base = ((2u + (s & 1u)) << extra) + 1u; }
   return l | (extra << 4) | (base << 16);
}

// entry e of the literal/length table: what the TB bits e decode to -- up to three literals
SQB_INF_HD uint32_t make_lit_entry(const uint16_t *cnt, const uint16_t *sym, uint32_t e)
{
   const int r = canon_decode(cnt, sym, e, TB);
   if (r < 0) return K_LONG;
   uint32_t ent = lit_entry_of((uint32_t)r & 0xffffu, (uint32_t)r >> 16);
   if (((ent >> 4) & 3u) == 0) return ent;
   uint32_t total = ent & 15u, n = 1;
   while (n < 3u && total < (uint32_t)TB) {
      const int r2 = canon_decode(cnt, sym, e >> total, TB - (int)total);
      if (r2 < 0 || ((uint32_t)r2 & 0xffffu) > 255u) break;
      ent |= ((uint32_t)r2 & 0xffu) << (8u + 8u * n);
      total += (uint32_t)r2 >> 16;
      n++;
   }
   return (ent & ~0x3fu) | total | (n << 4);
}

SQB_INF_HD uint32_t make_dist_entry(const uint16_t *cnt, const uint16_t *sym, uint32_t e)
{
   const int r = canon_decode(cnt, sym, e, TBD);
   if (r < 0) return D_LONG;
   return dist_entry_of((uint32_t)r & 0xffffu, (uint32_t)r >> 16);
}

// ---- block header (RFC 1951 §3.2.3, §3.2.7): the decoding lane ---------------------------------------------------
// type 0: stored (the caller copies), 1: fixed code, 2: dynamic code.  For 1 and 2 the counts and sorted symbols of
// both codes are in t afterwards (the entries are built by all lanes).  Returns OK or an error.
SQB_INF_HD uint32_t read_block_header(BitReader &br, Tables &t, uint32_t *type, uint32_t *final_block)
{
   Scratch &s = *(Scratch *)(void *)t.lit;
   if (br.refill()) return ERR_INPUT;
   *final_block = br.take(1);
   *type = br.take(2);
   if (*type == 3u) return ERR_HEADER;
   if (*type == 0u) return OK;
   int hlit, hdist;
   if (*type == 1u) {
      hlit = 288;
      hdist = 32;
      for (int i = 0; i < 144; i++) s.lens[i] = 8;
      for (int i = 144; i < 256; i++) s.lens[i] = 9;
      for (int i = 256; i < 280; i++) s.lens[i] = 7;
      for (int i = 280; i < 288; i++) s.lens[i] = 8;
      for (int i = 288; i < 320; i++) s.lens[i] = 5;
   } else {
      hlit = (int)br.take(5) + 257;
      hdist = (int)br.take(5) + 1;
      const int hclen = (int)br.take(4) + 4;
      if (hlit > 286 || hdist > 30) return ERR_HEADER;
      for (int i = 0; i < 19; i++) s.cl[i] = 0;
      const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      for (int i = 0; i < hclen; i++) {
         if (br.refill()) return ERR_INPUT;
         s.cl[order[i]] = (uint8_t)br.take(3);
      }
      if (!build_canon(s.cl, 19, s.clcnt, s.clsym, s.offs)) return ERR_CODE;
      int i = 0;
      const int n = hlit + hdist;
      while (i < n) {
         if (br.refill()) return ERR_INPUT;
         const int r = canon_decode(s.clcnt, s.clsym, (uint32_t)br.buf, 7);
         if (r < 0) return ERR_CODE;
         br.drop((uint32_t)r >> 16);
         const uint32_t sym = (uint32_t)r & 0xffffu;
         if (sym < 16u) { s.lens[i++] = (uint8_t)sym; continue; }
         uint32_t rep, val = 0;
         if (sym == 16u) {
            if (i == 0) return ERR_HEADER;
            val = s.lens[i - 1];
            rep = 3u + br.take(2);
         } else if (sym == 17u) rep = 3u + br.take(3);
         else rep = 11u + br.take(7);
         if (i + (int)rep > n) return ERR_HEADER;
         while (rep--) s.lens[i++] = (uint8_t)val;
      }
      if (s.lens[256] == 0) return ERR_HEADER;
   }
   // the distance lengths move to a fixed place first: build_canon of the literal code overwrites nothing of lens,
   // but lsym / dsym live outside the scratch, so the order is free
   if (!build_canon(s.lens, hlit, t.lcnt, t.lsym, s.offs)) return ERR_CODE;
   if (!build_canon(s.lens + hlit, hdist, t.dcnt, t.dsym, s.offs)) return ERR_CODE;
   return OK;
}

// ---- the symbol loop of the decoding lane -----------------------------------------------------------------------
// Decoding never READS the text: literals are stored as they are decoded, matches are only noted -- (position,
// length, distance) in a queue of kQueue entries -- and copied by the whole warp afterwards (resolve, below).  A
// copy is a round trip to L2 (a store does not allocate in L1): thirty-two of them in flight at once instead of
// one behind the other is what keeps match-heavy text (random DNA deflates to matches of 6-10 bases) off the latency
// floor.
constexpr uint32_t kQueue = 32;              // one match per lane
constexpr uint32_t kLaneCopy = 16;           // matches up to this length that do not overlap themselves: one lane each

struct MatchQueue {
   struct alignas(8) Entry {
      uint32_t pos;
      uint32_t ld;                           // length | distance << 16
   } e[kQueue];
};

// Decodes into out[pos ...) (pos counts from the start of the member's text, oend = ISIZE) until the end of the
// block (R_EOB), a queue of qcap matches (R_FULL) or an error (R_ERR + which).  pos runs ahead over the queued
// matches; *nq entries of q are to be copied in either case.
//
// The literals of an entry are stored as three bytes whatever their number: the bytes behind the last one are
// rewritten by the next symbol (a literal: at once; a match: when the queue is resolved, in front of every match that
// reads them -- match_ready), and the last three bytes of the member take the general path.
SQB_INF_HD int run_symbols(BitReader &br, const Tables &t, uint8_t *out, uint32_t &pos_io, const uint32_t oend,
                           MatchQueue &q, const uint32_t qcap, uint32_t *nq)
{
   uint32_t n = 0, pos = pos_io;
   int ret;
#ifdef __CUDA_ARCH__
   const uint32_t s_lit = smem_addr(t.lit), s_dist = smem_addr(t.dist), s_q = smem_addr(q.e);
#define SQB_INF_LIT(i) lds32(s_lit + ((i) << 2))
#define SQB_INF_DIST(i) lds32(s_dist + ((i) << 2))
#define SQB_INF_PUSH(i, p, v) sts64(s_q + ((i) << 3), (p), (v))
#else
#define SQB_INF_LIT(i) t.lit[i]
#define SQB_INF_DIST(i) t.dist[i]
#define SQB_INF_PUSH(i, p, v) (q.e[i].pos = (p), q.e[i].ld = (v))
#endif
   for (;;) {
      if (SQB_INF_RARE(br.refill())) { ret = R_ERR + (int)ERR_INPUT; goto done; }
      {
         // ---- the usual symbols, decoded out of one 32-bit view of the window without touching the reader: three
         // literals with room for three bytes, or a match whose code words are in the tables and whose bits are all
         // in the window (33 at least: a length takes up to TB + 5 here, a distance up to 21).  Whatever else -- long code
         // words, end of block, the last bytes of the member, anything invalid -- is left to the general path below.
         const uint32_t w = (uint32_t)br.buf;
         const uint32_t e = SQB_INF_LIT(w & (kLitN - 1u));
         if (e & 0x30u) {
            if (SQB_INF_RARE(pos + 3u > oend)) goto general;
            uint8_t *o = out + pos;
            o[0] = (uint8_t)(e >> 8);
            o[1] = (uint8_t)(e >> 16);
            o[2] = (uint8_t)(e >> 24);
            br.drop(e & 15u);
            pos += (e >> 4) & 3u;
            continue;
         }
         if (SQB_INF_RARE(e & 0xc0u)) goto general;
         const uint32_t l1 = e & 15u, x1 = (e >> 17) & 7u, c1 = l1 + x1;
         const uint32_t len = ((e >> 8) & 0x1ffu) + ((w >> l1) & ~(~0u << x1));
         const uint32_t d = SQB_INF_DIST((w >> c1) & (kDistN - 1u));
         const uint32_t c2 = c1 + (d & 15u), x2 = (d >> 4) & 15u, total = c2 + x2;
         const uint32_t dist = (d >> 16) + (br.peek_from(c2) & ~(~0u << x2));
         if (SQB_INF_RARE((d & (D_LONG | D_BAD)) != 0u || total > (uint32_t)br.cnt || dist > pos || len > oend - pos))
            goto general;
         br.drop(total);
         SQB_INF_PUSH(n, pos, len | dist << 16);
         n++;
         pos += len;
         if (n == qcap) { ret = R_FULL; goto done; }
         continue;
      }
   general:
      {
#ifdef SQB_INF_COUNT_GENERAL
         SQB_INF_COUNT_GENERAL++;                        /* host harness only: symbols off the fast path */
#endif
         // ---- one symbol, step by step (RFC 1951 3.2.3), with every check
         uint32_t e = SQB_INF_LIT((uint32_t)br.buf & (kLitN - 1u));
         if ((e & 0xf0u) == K_LONG) {
            const int r = canon_decode(t.lcnt, t.lsym, (uint32_t)br.buf, 15);
            if (r < 0) { ret = R_ERR + (int)ERR_CODE; goto done; }
            e = lit_entry_of((uint32_t)r & 0xffffu, (uint32_t)r >> 16);
         }
         const uint32_t k = (e >> 4) & 3u;
         if (k) {
            if (pos + k > oend) { ret = R_ERR + (int)ERR_OUTPUT; goto done; }
            br.drop(e & 15u);
            out[pos] = (uint8_t)(e >> 8);
            if (k > 1u) out[pos + 1] = (uint8_t)(e >> 16);
            if (k > 2u) out[pos + 2] = (uint8_t)(e >> 24);
            pos += k;
            continue;
         }
         const uint32_t kind = e & (3u << 6);
         if (kind == K_BAD) { ret = R_ERR + (int)ERR_CODE; goto done; }
         br.drop(e & 15u);
         if (kind == K_EOB) { ret = R_EOB; goto done; }
         const uint32_t len = ((e >> 8) & 0x1ffu) + br.take((e >> 17) & 7u);
         if (br.refill()) { ret = R_ERR + (int)ERR_INPUT; goto done; }
         uint32_t d = SQB_INF_DIST((uint32_t)br.buf & (kDistN - 1u));
         if (d & D_LONG) {
            const int r = canon_decode(t.dcnt, t.dsym, (uint32_t)br.buf, 15);
            if (r < 0) { ret = R_ERR + (int)ERR_CODE; goto done; }
            d = dist_entry_of((uint32_t)r & 0xffffu, (uint32_t)r >> 16);
         }
         if (d & D_BAD) { ret = R_ERR + (int)ERR_CODE; goto done; }
         br.drop(d & 15u);
         const uint32_t dist = (d >> 16) + br.take((d >> 4) & 15u);
         if (dist > pos) { ret = R_ERR + (int)ERR_DIST; goto done; }
         if (len > oend - pos) { ret = R_ERR + (int)ERR_OUTPUT; goto done; }
         SQB_INF_PUSH(n, pos, len | dist << 16);
         n++;
         pos += len;
         if (n == qcap) { ret = R_FULL; goto done; }
      }
   }
done:
#undef SQB_INF_LIT
#undef SQB_INF_DIST
#undef SQB_INF_PUSH
   pos_io = pos;
   *nq = n;
   return ret;
}

// ---- resolving the queue: what ONE lane decides and copies ---------------------------------------------------------
// The queued matches lie at increasing positions.  Everything in front of P, the position of the first match still
// pending, is final text (literals are stored before the queue is resolved; earlier matches are done).  A match
// whose source ends at or before P can be copied now, together with every other such match: none of them reads what
// another one writes.  The first pending match always can (its source lies in front of itself).  A match that
// overlaps itself (distance < length) has its source end at its own position: it goes when it is the first.
SQB_INF_HD bool match_ready(uint32_t mp, uint32_t ml, uint32_t md, uint32_t P)
{
   return mp - md + (ml < md ? ml : md) <= P;
}
SQB_INF_HD bool match_by_lane(uint32_t ml, uint32_t md) { return ml <= kLaneCopy && md >= ml; }
// all loads in front of all stores: one round trip
SQB_INF_HD void copy_by_lane(uint8_t *out, uint32_t mp, uint32_t ml, uint32_t md)
{
   uint8_t b[kLaneCopy];
   const uint8_t *src = out + mp - md;
   uint8_t *dst = out + mp;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
   for (uint32_t k = 0; k < kLaneCopy; k++) if (k < ml) b[k] = src[k];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
   for (uint32_t k = 0; k < kLaneCopy; k++) if (k < ml) dst[k] = b[k];
}
// byte j of a match copied by the whole warp: the source index (an overlapping match repeats its last md bytes)
SQB_INF_HD uint32_t match_src(uint32_t mp, uint32_t md, uint32_t j)
{
   return mp - md + (j < md ? j : j % md);
}

// ---- BGZF member header (SAM specification §4.1; RFC 1952 §2.3) ---------------------------------------------------
struct Member {
   uint64_t in_off;      // first byte of the deflate stream, from the start of the buffer
   uint32_t in_len;      // bytes of deflate stream
   uint32_t isize;       // bytes of text
   uint64_t out_off;     // where the text goes
};

// Parses the gzip member at gz[off ...): fills in_off / in_len / isize and returns the offset of the next member,
// 0 when this is not a complete BGZF member (no "BC" sub-field: plain gzip cannot be cut without decoding it).
static inline uint64_t parse_member(const uint8_t *gz, uint64_t nbytes, uint64_t off, Member *m)
{
   if (off + 18 > nbytes) return 0;
   const uint8_t *h = gz + off;
   if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
   const uint32_t xlen = h[10] | (uint32_t)h[11] << 8;
   if (off + 12 + xlen > nbytes) return 0;
   uint32_t bsize = 0;
   bool found = false;
   for (uint32_t x = 0; x + 4 <= xlen;) {
      const uint8_t *f = h + 12 + x;
      const uint32_t slen = f[2] | (uint32_t)f[3] << 8;
      if (f[0] == 'B' && f[1] == 'C' && slen == 2 && x + 6 <= xlen) { bsize = f[4] | (uint32_t)f[5] << 8; found = true; }
      x += 4 + slen;
   }
   if (!found) return 0;
   uint64_t hdr = 12 + xlen;
   if (h[3] & 8) { while (off + hdr < nbytes && h[hdr]) hdr++; hdr++; }      // FNAME
   if (h[3] & 16) { while (off + hdr < nbytes && h[hdr]) hdr++; hdr++; }     // FCOMMENT
   if (h[3] & 2) hdr += 2;                                                   // FHCRC
   const uint64_t total = (uint64_t)bsize + 1;
   if (total < hdr + 8 || off + total > nbytes) return 0;
   m->in_off = off + hdr;
   m->in_len = (uint32_t)(total - hdr - 8);
   const uint8_t *tail = h + total - 4;
   m->isize = tail[0] | (uint32_t)tail[1] << 8 | (uint32_t)tail[2] << 16 | (uint32_t)tail[3] << 24;
   return off + total;
}

}  // namespace inf
}  // namespace sqb
#endif
