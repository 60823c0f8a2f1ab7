// sqb_tables.h -- host-side construction of the tables the kernels consume:
// the byte -> class nibble table of the tokenizer (K1) and the row / slot
// description of the bit-sliced matcher (K2).  Plain C++, no CUDA: shared by
// sqb_engine.cu and the host test harness tests/host_bitslice.cpp.
#ifndef SQB_TABLES_H_
#define SQB_TABLES_H_

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "sqb_bitslice.h"

namespace sqb {

// option bits of libseeq.h
enum { OPT_MATCH = 0x03, OPT_BEST = 0x01, OPT_ALL = 0x02, OPT_CONVERT = 0x04, OPT_IGNORE = 0x08,
       OPT_NONDNA = 0x0C, OPT_STREAM = 0x10 };

constexpr uint8_t kClsN = 4, kClsStop = 5, kClsSkip = 6, kClsNull = 7, kClsNewline = 8;

struct ClassTable {
   uint8_t code[256];              // class nibble per byte value ('\n' has bit 3 set)
};

static inline int base_code(int ch)
{
   switch (ch) {
   case 'A': case 'a': return 0;
   case 'C': case 'c': return 1;
   case 'G': case 'g': return 2;
   case 'T': case 't': case 'U': case 'u': return 3;
   case 'N': case 'n': return 4;
   default: return -1;
   }
}

// Follows seeqcore.h:89-111 + libseeq.c:255-270: NUL and '\n' end the line,
// any other non-base byte ends it (SQ_FAIL), is a text N (SQ_CONVERT) or is
// invisible (SQ_IGNORE).  Bytes >= 0x80 are "other" (the reference indexes its
// table with a signed char there: undefined).
static inline void build_class_table(int options, ClassTable *t)
{
   const int nondna = options & OPT_NONDNA;
   for (int b = 0; b < 256; b++) {
      const int code = b < 128 ? base_code(b) : -1;
      uint8_t c;
      if (code >= 0) c = (uint8_t)code;
      else if (b == 0) c = kClsStop;
      else if (b == '\n') c = kClsStop | kClsNewline;
      else if (nondna == OPT_CONVERT) c = kClsN;
      else if (nondna == OPT_IGNORE) c = kClsSkip;
      else c = kClsStop;
      t->code[b] = c;
   }
}

// the kernel instance (R rows per part, G parts) that serves a pattern of m
// positions; false if there is none (m > 128)
struct BsShape { int rows, parts; };
static const BsShape kBsShapes[] = {{8, 1}, {10, 1}, {12, 1}, {16, 1}, {20, 1}, {24, 1}, {32, 1},   // m <= 32: one lane per group
                                    {20, 2}, {24, 2}, {32, 2},                       // m <= 64: two lanes
                                    {20, 4}, {24, 4}, {26, 4}, {28, 4}, {32, 4}};    // m <= 128: four lanes
static inline bool bs_shape_for(int m, BsShape *out)
{
   for (const BsShape &b : kBsShapes)
      if (m <= b.rows * b.parts) { *out = b; return true; }
   return false;
}

// Long lines are cut into SEGMENTS so that they fill the lanes of the bit-sliced
// matcher like short lines do.  A cut is made at every text offset A = kCutWindow
// (mod kCutStride) whose line has been running for at least kCutWindow bytes (no
// line start in the kCutWindow bytes before A).  The window only has to hold the
// warm-up; a short one keeps the first segment of a line (kCutWindow .. kCutWindow
// + kCutStride bytes) close to the length of the others, and the lanes of a tile
// run as long as its longest segment.  The segment that starts at A
// scans from A - warm-up with a fresh automaton and reports the events that end
// after A; the segment before it scans up to and including byte A and reports
// the events that end at or before A.  A warm-up of m + 2 tau + 2 automaton
// inputs reproduces both the capped distance (it depends on the last m + tau
// inputs) and the suppress flag (a rising run of capped distances has at most
// tau + 2 members) -- SURVEY.md 3.3, pinned by tests/test_bitslice_host.py.
constexpr uint32_t kCutStride = 2048;
constexpr uint32_t kCutWindow = 256;
static inline uint32_t bs_warmup(int m, int tau) { return (uint32_t)(m + 2 * tau + 2); }

// keys: one class byte per pattern position (bit0 A .. bit3 T, 0x1F = N).
// Returns false if the pattern needs more than kBsMaxCustom custom classes or does not
// fit (the caller then uses the word-parallel kernels).
static inline bool build_bs_pattern(const unsigned char *keys, int m, int tau, BsPattern *p)
{
   memset(p, 0, sizeof *p);
   BsShape shape;
   if (!bs_shape_for(m, &shape) || tau + 1 > 15) return false;
   p->m = m;
   p->tau = tau;
   p->rows = shape.rows;
   p->parts = shape.parts;
   const int R = shape.rows * shape.parts;        // rows of the whole automaton
   const int pad = R - m;
   int ncustom = 0;
   unsigned char custom_key[kBsMaxCustom] = {0};
   for (int j = 0; j < R; j++) {
      if (j < pad) { p->slot[j] = BS_ONES; continue; }     // entries beyond R stay 0 (unused)
      const unsigned char k = keys[j - pad] & 0x1F;
      int slot = -1;
      switch (k) {
      case 0x01: slot = BS_A; break;
      case 0x02: slot = BS_C; break;
      case 0x04: slot = BS_G; break;
      case 0x08: slot = BS_T; break;
      case 0x10: slot = BS_N; break;
      case 0x1F: slot = BS_ANY; break;
      default:
         for (int c = 0; c < ncustom; c++) if (custom_key[c] == k) slot = BS_CUSTOM0 + c;
         if (slot < 0) {
            if (ncustom == kBsMaxCustom) return false;
            custom_key[ncustom] = k;
            for (int b = 0; b < 5; b++) p->custom[ncustom][b] = (k >> b) & 1 ? ~0u : 0u;
            slot = BS_CUSTOM0 + ncustom++;
         }
      }
      p->slot[j] = (uint8_t)slot;
   }
   p->ncustom = ncustom;
   for (int j = 0; j < kBsMaxRows; j++) p->slot_off[j] = (uint32_t)p->slot[j] * 32u * 4u;
   for (int k = 0; k < 8; k++) {
      p->tau_plane[k] = (tau >> k) & 1 ? ~0u : 0u;
      p->m_plane[k] = (m >> k) & 1 ? ~0u : 0u;
   }
   for (int k = 0; k < kBsBestBits; k++) p->best_plane[k] = ((tau + 1) >> k) & 1 ? ~0u : 0u;
   return true;
}

// sysfs cpulist ("0-31,64-95\n") -> one byte per CPU (1 = listed); returns the number of CPUs listed
static inline int parse_cpulist(const char *s, unsigned char *cpus, int maxcpus)
{
   int count = 0;
   memset(cpus, 0, (size_t)maxcpus);
   while (*s) {
      while (*s == ',' || *s == ' ' || *s == '\n' || *s == '\t') s++;
      if (*s < '0' || *s > '9') break;
      long a = 0, b;
      while (*s >= '0' && *s <= '9') a = a * 10 + (*s++ - '0');
      b = a;
      if (*s == '-') {
         s++;
         if (*s < '0' || *s > '9') break;
         b = 0;
         while (*s >= '0' && *s <= '9') b = b * 10 + (*s++ - '0');
      }
      for (long c = a; c <= b && c < maxcpus; c++)
         if (!cpus[c]) { cpus[c] = 1; count++; }
   }
   return count;
}

// (memrchr is a GNU extension; a plain loop keeps this header portable -- the search is a few lines long)
static inline const void *sqb_memrchr(const void *s, int c, size_t n)
{
   const unsigned char *p = (const unsigned char *)s + n;
   while (n--) if (*--p == (unsigned char)c) return p;
   return nullptr;
}

// SQB_FASTQ: chunks must start at record boundaries (the matcher takes the lines 1 mod 4 of a chunk).
// The start of the last record that begins at or before `cut` (a line start) inside text[lo, nbytes): a
// line that starts with '@' whose next-but-one line starts with '+'.  Exact for well-formed 4-line
// records: a quality line may start with '@', but two lines on comes a sequence line, never a '+'.
// Looks at the 64 lines in front of `cut` at most; (size_t)-1 if there is no such line.
static inline size_t fastq_record_start(const char *text, size_t lo, size_t cut, size_t nbytes)
{
   size_t q = cut;
   for (int tries = 0; tries < 64; tries++) {
      if (q < nbytes && text[q] == '@') {
         const char *nl1 = (const char *)memchr(text + q, '\n', nbytes - q);
         const char *nl2 = nl1 ? (const char *)memchr(nl1 + 1, '\n', (size_t)(text + nbytes - (nl1 + 1))) : NULL;
         if (nl2 && (size_t)(nl2 + 1 - text) < nbytes && nl2[1] == '+') return q;
      }
      if (q <= lo) break;
      const char *prev = q >= lo + 2 ? (const char *)sqb_memrchr(text + lo, '\n', q - 1 - lo) : NULL;
      q = prev ? (size_t)(prev - text) + 1 : lo;
   }
   return (size_t)-1;
}

}  // namespace sqb
#endif
