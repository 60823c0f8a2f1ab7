/* sqb_gen.h -- counter-based synthetic read generator (tests and benchmarks).
 *
 * One function, compiled for host (gcc / nvcc host pass) and device alike, so
 * that the oracle on the host and the kernels on the GPU see the same bytes
 * (SURVEY.md 8d).  Every read is a pure function of (seed, read index):
 * bases i.i.d. uniform over ACGT, optional N / non-DNA noise, and optionally a
 * mutated copy of a fixed sequence planted at a uniform offset.
 * All records have a fixed size so that any read can be generated in place.
 */
#ifndef SQB_GEN_H_
#define SQB_GEN_H_

#include <stdint.h>
#include "seeq_b200.h"

#ifdef __CUDACC__
#define SQB_HD __host__ __device__ __forceinline__
#else
#define SQB_HD static inline
#endif

SQB_HD uint64_t sqb_mix64(uint64_t x)
{
   x += 0x9E3779B97F4A7C15ull;
   x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
   x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
   return x ^ (x >> 31);
}

#define SQB_FASTQ_IDW 12   /* "@r" + 12 digits */

SQB_HD size_t sqb_gen_record_bytes(const sqb_gen_t *g)
{
   const size_t seq = (size_t)g->line_len + 1;
   if (!g->fastq) return seq;
   return (size_t)(2 + SQB_FASTQ_IDW + 1) + seq + 2 + seq;
}

/* writes exactly sqb_gen_record_bytes(g) bytes */
SQB_HD void sqb_gen_record(const sqb_gen_t *g, uint64_t read, char *dst)
{
   const uint32_t L = g->line_len;
   char *seq = dst;
   if (g->fastq) {
      *dst++ = '@';
      *dst++ = 'r';
      uint64_t v = read;
      for (int k = SQB_FASTQ_IDW - 1; k >= 0; k--) {
         dst[k] = (char)('0' + (int)(v % 10));
         v /= 10;
      }
      dst += SQB_FASTQ_IDW;
      *dst++ = '\n';
      seq = dst;
   }
   const uint64_t key = sqb_mix64(g->seed ^ (read * 0xD1B54A32D192ED03ull));
   for (uint32_t p = 0; p < L; p++) {
      const uint64_t h = sqb_mix64(key + p);
      char c = "ACGT"[h & 3];
      if (((h >> 2) & 1023) < g->n_per_1024) c = 'N';
      if (((h >> 12) & 1023) < g->junk_per_1024) c = "RYKMSW.-"[(h >> 22) & 7];
      seq[p] = c;
   }
   /* plant a mutated copy of g->plant */
   const uint64_t d = sqb_mix64(key ^ 0xA5A5A5A5A5A5A5A5ull);
   if (g->plant_len > 0 && (d & 1023) < g->plant_per_1024) {
      char buf[320];
      uint32_t len = g->plant_len;
      for (uint32_t k = 0; k < len; k++) buf[k] = g->plant[k];
      const uint32_t nedits = (uint32_t)((d >> 10) % (uint64_t)(g->max_edits + 1));
      for (uint32_t e = 0; e < nedits; e++) {
         const uint64_t r = sqb_mix64(d + e + 1);
         const uint32_t kind = (uint32_t)(r % 3);
         const char base = "ACGT"[(r >> 8) & 3];
         if (kind == 0 || len <= 1) {                 /* substitution */
            buf[(r >> 16) % len] = base;
         } else if (kind == 1 && len < 300) {         /* insertion */
            const uint32_t at = (uint32_t)((r >> 16) % (len + 1));
            for (uint32_t k = len; k > at; k--) buf[k] = buf[k - 1];
            buf[at] = base;
            len++;
         } else {                                     /* deletion */
            const uint32_t at = (uint32_t)((r >> 16) % len);
            for (uint32_t k = at; k + 1 < len; k++) buf[k] = buf[k + 1];
            len--;
         }
      }
      if (len <= L) {
         const uint32_t off = (uint32_t)((d >> 24) % (uint64_t)(L - len + 1));
         for (uint32_t k = 0; k < len; k++) seq[off + k] = buf[k];
      }
   }
   seq[L] = '\n';
   if (g->fastq) {
      char *q = seq + L + 1;
      *q++ = '+';
      *q++ = '\n';
      for (uint32_t p = 0; p < L; p++) {
         const uint64_t h = sqb_mix64(key + 0x51ED270B0000ull + p);
         q[p] = (char)('#' + (int)(h % 39));          /* '#'..'I' */
      }
      q[L] = '\n';
   }
}

#endif
