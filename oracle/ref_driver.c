/*
 * ref_driver.c -- thin driver around the UNMODIFIED reference (TEST
 * INFRASTRUCTURE ONLY).
 *
 * This file is ours; it is compiled together with /root/reference/src/libseeq.c
 * and /root/reference/src/seeq.c *where they lie* (oracle/Makefile) into
 * oracle/_ref/libseeq_ref.so.  No reference source is copied into the repo.
 * It exposes the reference through flat C entry points so that the tests can
 * fuzz the oracle against it and bench.py can time it (cpu_baseline.kind =
 * "reference") without Python overhead inside the timed loop.
 */
#define _GNU_SOURCE
#include "seeq.h"      /* the reference's own header, via -I/root/reference/src */

#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

static double now_s(void)
{
   struct timeval tv;
   gettimeofday(&tv, NULL);
   return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

/* One string through seeqNew/seeqStringMatch/seeqMatchIter.
 * out receives (start,end,dist) triples in iterator order.
 * Returns the hit count, -1 on matcher error, -100-seeqerr if seeqNew failed. */
long ref_string_match(const char *pattern, int tau, const char *text,
                      int options, uint64_t *out, long cap)
{
   seeq_t *sq = seeqNew(pattern, tau, 0);
   if (sq == NULL) return -100 - seeqerr;
   long hits = seeqStringMatch(text, sq, options);
   if (hits >= 0) {
      long k = 0;
      match_t *mt;
      while ((mt = seeqMatchIter(sq)) != NULL && k < cap) {
         out[3 * k + 0] = mt->start;
         out[3 * k + 1] = mt->end;
         out[3 * k + 2] = mt->dist;
         k++;
      }
   }
   seeqFree(sq);
   return hits;
}

/* seeqOpen() restated for an in-memory file (seeq.c:224-255). */
static seeqfile_t *open_mem(const char *buf, size_t n)
{
   seeqfile_t *f = calloc(1, sizeof(seeqfile_t));
   if (f == NULL) return NULL;
   f->fdi = n ? fmemopen((void *)buf, n, "r") : fopen("/dev/null", "r");
   if (f->fdi == NULL) { free(f); return NULL; }
   if (n && buf[0] == '>') {
      f->flags = 1;
      f->info = calloc(32, 1);
   }
   return f;
}

/* All records of all lines via seeqFileMatch(.., SQ_ANY).
 * out receives (line,start,end,dist) quadruples. */
long ref_buffer_scan(const char *buf, size_t n, const char *pattern, int tau,
                     int options, uint64_t *out, long cap, uint64_t *nlines,
                     uint64_t *nmatched)
{
   seeq_t *sq = seeqNew(pattern, tau, 0);
   if (sq == NULL) return -100 - seeqerr;
   seeqfile_t *f = open_mem(buf, n);
   if (f == NULL) { seeqFree(sq); return -2; }
   long total = 0, rv;
   uint64_t matched = 0;
   while ((rv = seeqFileMatch(f, sq, options, SQ_ANY)) > 0) {
      if (sq->hits > 0) matched++;
      match_t *mt;
      while ((mt = seeqMatchIter(sq)) != NULL) {
         if (total < cap) {
            out[4 * total + 0] = f->line;
            out[4 * total + 1] = mt->start;
            out[4 * total + 2] = mt->end;
            out[4 * total + 3] = mt->dist;
         }
         total++;
      }
   }
   if (nlines) *nlines = f->line;
   if (nmatched) *nmatched = matched;
   seeqClose(f);
   seeqFree(sq);
   return rv < 0 ? -1 : total;
}

/* Whole-file counts: file_opt = SQ_COUNTLINES (3) or SQ_COUNTMATCH (4). */
long ref_buffer_count(const char *buf, size_t n, const char *pattern, int tau,
                      int options, int file_opt)
{
   seeq_t *sq = seeqNew(pattern, tau, 0);
   if (sq == NULL) return -100 - seeqerr;
   seeqfile_t *f = open_mem(buf, n);
   if (f == NULL) { seeqFree(sq); return -2; }
   long rv = seeqFileMatch(f, sq, options, file_opt);
   seeqClose(f);
   seeqFree(sq);
   return rv;
}

/* Checksum of a record list that is LINEAR in the line number, so that the sums of newline-aligned
 * shards (local line numbers) combine into the sum of the whole buffer:
 *   sum over records of  line * P0 + start * P1 + end * P2 + dist * P3 + P4   (mod 2^64)
 * bench.py computes the same from the GPU's records of the same sample. */
#define CK_P0 0x9E3779B97F4A7C15ull
#define CK_P1 0xC2B2AE3D27D4EB4Full
#define CK_P2 0x165667B19E3779F9ull
#define CK_P3 0x27D4EB2F165667C5ull
#define CK_P4 0x85EBCA77C2B2AE63ull

typedef struct {
   long     result;      /* mode 0 / 1: the count; mode 2: number of records */
   uint64_t ck_rest;     /* sum of start * P1 + end * P2 + dist * P3 + P4     */
   uint64_t line_sum;    /* sum of the (shard-local, 1-based) line numbers    */
   uint64_t nlines;      /* counted lines of the shard                        */
} shard_out_t;

/* work of one shard for ref_bench */
static void shard_work(const char *buf, size_t n, const char *pattern, int tau,
                       int options, int mode, shard_out_t *out)
{
   memset(out, 0, sizeof *out);
   if (mode == 0) { out->result = ref_buffer_count(buf, n, pattern, tau, options, SQ_COUNTLINES); return; }
   if (mode == 1) { out->result = ref_buffer_count(buf, n, pattern, tau, options, SQ_COUNTMATCH); return; }
   /* mode 2: what the CLI does without -c: iterate matching lines and their
    * records (seeq.c:131-176), folding them into a checksum instead of printf */
   seeq_t *sq = seeqNew(pattern, tau, 0);
   if (sq == NULL) { out->result = -100 - seeqerr; return; }
   seeqfile_t *f = open_mem(buf, n);
   if (f == NULL) { seeqFree(sq); out->result = -2; return; }
   long recs = 0;
   uint64_t rest = 0, lines = 0;
   while (seeqFileMatch(f, sq, options, SQ_MATCH) > 0) {
      match_t *mt;
      while ((mt = seeqMatchIter(sq)) != NULL) {
         rest += mt->start * CK_P1 + mt->end * CK_P2 + mt->dist * CK_P3 + CK_P4;
         lines += f->line;
         recs++;
      }
   }
   out->result = recs;
   out->ck_rest = rest;
   out->line_sum = lines;
   out->nlines = f->line;          /* at end of input: every line read was counted (seeq.c:377) */
   seeqClose(f);
   seeqFree(sq);
}

/* Time the reference on `nproc` host processes over newline-aligned shards of
 * buf (SURVEY.md 8d "N cores").  Returns wall seconds (fork to last reap), the
 * summed result in *total and (mode 2) the checksum of all records with
 * buffer-global 1-based line numbers in *checksum.  nproc <= 1 runs in-process. */
double ref_bench_ck(const char *buf, size_t n, const char *pattern, int tau,
                    int options, int mode, int nproc, long *total, uint64_t *checksum)
{
   if (nproc < 1) nproc = 1;
   shard_out_t *res = mmap(NULL, sizeof(shard_out_t) * (size_t)nproc, PROT_READ | PROT_WRITE,
                           MAP_SHARED | MAP_ANONYMOUS, -1, 0);
   if (res == MAP_FAILED) return -1.0;
   size_t *cut = malloc(sizeof(size_t) * (size_t)(nproc + 1));
   cut[0] = 0;
   for (int k = 1; k < nproc; k++) {
      size_t p = (size_t)((double)n * k / nproc);
      if (p < cut[k - 1]) p = cut[k - 1];
      const char *nl = p < n ? memchr(buf + p, '\n', n - p) : NULL;
      cut[k] = nl ? (size_t)(nl - buf) + 1 : n;
   }
   cut[nproc] = n;
   double t0 = now_s();
   if (nproc == 1) {
      shard_work(buf, n, pattern, tau, options, mode, &res[0]);
   } else {
      for (int k = 0; k < nproc; k++) {
         pid_t pid = fork();
         if (pid == 0) {
            shard_work(buf + cut[k], cut[k + 1] - cut[k], pattern, tau, options, mode, &res[k]);
            _exit(0);
         }
         if (pid < 0) res[k].result = -3;
      }
      while (wait(NULL) > 0) {}
   }
   double t1 = now_s();
   long sum = 0;
   uint64_t ck = 0, base = 0;
   for (int k = 0; k < nproc; k++) {
      sum += res[k].result;
      ck += res[k].ck_rest + (res[k].line_sum + base * (uint64_t)(res[k].result > 0 ? res[k].result : 0)) * CK_P0;
      base += res[k].nlines;
   }
   if (total) *total = sum;
   if (checksum) *checksum = mode == 2 ? ck : (uint64_t)sum;
   munmap(res, sizeof(shard_out_t) * (size_t)nproc);
   free(cut);
   return t1 - t0;
}

double ref_bench(const char *buf, size_t n, const char *pattern, int tau,
                 int options, int mode, int nproc, long *total)
{
   return ref_bench_ck(buf, n, pattern, tau, options, mode, nproc, total, NULL);
}
