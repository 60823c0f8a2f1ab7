/*
 * seeq_file.c -- file driver and CLI formatter on top of the batch engine
 * (host side, C99, written from scratch).
 *
 * Mirrors, for the same names / arguments / return values / error codes:
 *
 *   seeqOpen       /root/reference/src/seeq.c:201-256
 *   seeqClose      :258-291
 *   seeqFileMatch  :293-392   line iterator: getline, '\n' strip, FASTA header
 *                             rule, 1-based line counter, SQ_ANY / SQ_MATCH /
 *                             SQ_NOMATCH / SQ_COUNTLINES / SQ_COUNTMATCH
 *   seeq           :32-199    output formatter of the CLI
 *
 * The reference matches one line per getline().  Here the stream is read in
 * large chunks that end on a line boundary; each chunk is matched on the GPU
 * in ONE batch (K1..K4) and seeqFileMatch then hands its lines out one call at
 * a time from the ordered record list, so callers see the same iterator.
 */
#define _GNU_SOURCE
#include "sqb_private.h"

#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

/* ------------------------------------------------------------------------ */
/* open / close                                                              */
/* ------------------------------------------------------------------------ */
seeqfile_t *seeqOpen(const char *file)
{
   seeqerr = 0;
   sqb_file_t *f = calloc(1, sizeof *f);
   if (f == NULL) { seeqerr = errno; return NULL; }

   FILE *in = file ? fopen(file, "r") : stdin;
   if (in == NULL) {
      seeqerr = errno;          /* the raw errno, as the reference does (seeq.c:234) */
      free(f);
      return NULL;
   }
   f->magic = SQB_FILE_MAGIC;
   f->pub.line = 0;
   f->pub.fdi = in;

   /* FASTA iff the very first byte of the stream is '>' (seeq.c:243-253) */
   const int first = getc(in);
   if (first == '>') {
      f->pub.flags = 1;
      f->pub.info = calloc(32, 1);
      if (f->pub.info == NULL) { seeqerr = errno; free(f); return NULL; }
   }
   if (first != EOF) ungetc(first, in);
   return &f->pub;
}

static void release_buffers(sqb_file_t *f)
{
   if (f->thread_started) {
      /* the reader may sit in fread() on a pipe that never delivers: cancel it there */
      pthread_mutex_lock(&f->mu);
      f->stop = 1;
      pthread_cond_broadcast(&f->cv);
      pthread_mutex_unlock(&f->mu);
      pthread_cancel(f->thread);
      pthread_join(f->thread, NULL);
      pthread_mutex_destroy(&f->mu);
      pthread_cond_destroy(&f->cv);
      f->thread_started = 0;
   }
   for (int k = 0; k < 2; k++) {
      if (f->chunk[k].buf) sqbHostFree(f->chunk[k].buf);
      f->chunk[k].buf = NULL;
      f->chunk[k].cap = 0;
   }
   free(f->carry);
   free(f->last_header);
   f->carry = NULL;
   f->buf = NULL;
   f->recs = NULL;
   f->lines = NULL;
   f->last_header = NULL;
}

int seeqClose(seeqfile_t *sqfile)
{
   seeqerr = 0;
   sqb_file_t *f = (sqb_file_t *)sqfile;
   FILE *in = sqfile->fdi;
   free(sqfile->info);
   sqfile->info = NULL;
   if (f->magic == SQB_FILE_MAGIC) release_buffers(f);
   f->magic = 0;
   free(f);
   if (in != NULL && in != stdin) {
      if (fclose(in) != 0) { seeqerr = errno; return -1; }
   }
   return 0;
}

/* ------------------------------------------------------------------------ */
/* chunk reader                                                              */
/* ------------------------------------------------------------------------ */
/* bytes per chunk: $SEEQ_B200_FILE_CHUNK_MB (64) MiB per GPU that sqbScanHost will use ($SEEQ_B200_DEVICES) */
static size_t chunk_target(FILE *in)
{
   const char *env = getenv("SEEQ_B200_FILE_CHUNK_MB");
   size_t mb = env ? (size_t)atol(env) : 64;
   if (mb < 1) mb = 1;
   if (mb > 1024) mb = 1024;
   const char *devs = getenv("SEEQ_B200_DEVICES");
   if (devs && *devs) {
      int nd = strcmp(devs, "all") == 0 ? sqbDeviceCount() : atoi(devs);
      if (nd > sqbDeviceCount()) nd = sqbDeviceCount();
      if (nd > 1) mb *= (size_t)nd;
      if (mb > 2048) mb = 2048;
   }
   size_t target = mb << 20;
   /* do not pin 64 MiB for a tiny regular file */
   struct stat st;
   if (fstat(fileno(in), &st) == 0 && S_ISREG(st.st_mode)) {
      const long at = ftell(in);
      size_t left = (size_t)st.st_size > (size_t)(at > 0 ? at : 0) ? (size_t)st.st_size - (size_t)(at > 0 ? at : 0) : 0;
      left += 4096;
      if (left < target) target = left;
   }
   return target;
}

/* (reader thread) room for `need` bytes in a chunk buffer, contents kept */
static int chunk_reserve(sqb_chunk_t *c, size_t keep, size_t need)
{
   if (need <= c->cap) return 0;
   size_t cap = c->cap ? c->cap : 4096;
   while (cap < need) cap *= 2;
   char *nb = sqbHostAlloc(cap);
   if (nb == NULL) return -1;
   if (keep) memcpy(nb, c->buf, keep);
   if (c->buf) sqbHostFree(c->buf);
   c->buf = nb;
   c->cap = cap;
   return 0;
}

static void reader_fail(sqb_file_t *f, int err)
{
   pthread_mutex_lock(&f->mu);
   f->reader_errno = err ? err : EIO;
   pthread_cond_broadcast(&f->cv);
   pthread_mutex_unlock(&f->mu);
}

/* The reader: fills chunk 0, 1, 0, ... -- each starts with the partial line left behind the chunk before
 * it, then takes `target` bytes of the stream (more while no newline has been seen: a line longer than
 * the chunk) and ends after its last newline; the rest is carried to the next chunk. */
static void *reader_main(void *arg)
{
   sqb_file_t *f = arg;
   for (int k = 0;; k ^= 1) {
      pthread_mutex_lock(&f->mu);
      while (f->state[k] != SQB_CH_FREE && !f->stop) pthread_cond_wait(&f->cv, &f->mu);
      const int stop = f->stop;
      pthread_mutex_unlock(&f->mu);
      if (stop) break;
      sqb_chunk_t *c = &f->chunk[k];
      size_t fill = f->carry_len;
      if (chunk_reserve(c, 0, fill + f->target + 16)) { reader_fail(f, ENODEV); break; }
      if (fill) memcpy(c->buf, f->carry, fill);
      f->carry_len = 0;
      int eof = 0, bad = 0;
      for (;;) {
         if (chunk_reserve(c, fill, fill + f->target + 16)) { bad = ENODEV; break; }
         const size_t got = fread(c->buf + fill, 1, f->target, f->in);
         const int newline = got && memrchr(c->buf + fill, '\n', got) != NULL;
         fill += got;
         if (got < f->target) {
            if (ferror(f->in)) { bad = errno ? errno : EIO; break; }
            eof = 1;
            break;
         }
         if (newline) break;
      }
      if (bad) { reader_fail(f, bad); break; }
      size_t len = fill;
      if (!eof) {
         const char *nl = memrchr(c->buf, '\n', fill);
         len = (size_t)(nl - c->buf) + 1;
         const size_t tail = fill - len;
         if (tail > f->carry_cap) {
            char *nc = realloc(f->carry, tail + tail / 2 + 64);
            if (nc == NULL) { reader_fail(f, ENOMEM); break; }
            f->carry = nc;
            f->carry_cap = tail + tail / 2 + 64;
         }
         if (tail) memcpy(f->carry, c->buf + len, tail);
         f->carry_len = tail;
      }
      c->len = len;
      c->last = eof;
      pthread_mutex_lock(&f->mu);
      f->state[k] = SQB_CH_READY;
      pthread_cond_broadcast(&f->cv);
      pthread_mutex_unlock(&f->mu);
      if (eof) break;
   }
   return NULL;
}

/* start of the line that ends just before position `end` (end > 0) */
static size_t line_start_before(const char *buf, size_t end)
{
   if (end == 0) return 0;
   const char *nl = memrchr(buf, '\n', end - 1);
   return nl ? (size_t)(nl - buf) + 1 : 0;
}

/* FASTA: remember the last header of the chunk that is about to be dropped */
static int remember_last_header(sqb_file_t *f)
{
   size_t end = f->len;
   while (end > 0) {
      const size_t s = line_start_before(f->buf, end);
      if (f->buf[s] == '>') {
         size_t n = end - s;
         if (n && f->buf[s + n - 1] == '\n') n--;
         char *h = malloc(n + 1);
         if (h == NULL) { seeqerr = 666; return -1; }    /* seeq.c:370 */
         memcpy(h, f->buf + s, n);
         h[n] = 0;
         free(f->last_header);
         f->last_header = h;
         return 0;
      }
      end = s;
   }
   return 0;
}

/* Makes the next chunk resident.  1 = a chunk is resident, 0 = end of input. */
static int chunk_load(sqb_file_t *f)
{
   if (f->started) {
      if (f->buf != NULL) {
         if ((f->pub.flags & 1) && remember_last_header(f)) return -1;
         f->line_base += f->nlines;
         /* hand the buffer back to the reader */
         pthread_mutex_lock(&f->mu);
         f->state[f->cur] = SQB_CH_FREE;
         pthread_cond_broadcast(&f->cv);
         pthread_mutex_unlock(&f->mu);
         f->cur ^= 1;
      }
   } else {
      f->in = f->pub.fdi;
      f->target = chunk_target(f->in);
      f->cur = 0;
      pthread_mutex_init(&f->mu, NULL);
      pthread_cond_init(&f->cv, NULL);
      if (pthread_create(&f->thread, NULL, reader_main, f) != 0) {
         pthread_mutex_destroy(&f->mu);
         pthread_cond_destroy(&f->cv);
         return -1;                                       /* errno tells */
      }
      f->thread_started = 1;
      f->started = 1;
   }
   f->buf = NULL;
   f->len = 0;
   f->cur_line = f->cur_rec = 0;
   f->nlines = f->nrecs = 0;
   f->recs = NULL;
   f->lines = NULL;
   f->res_valid = 0;
   f->hdr_scan = f->hdr_off = 0;
   f->hdr_seen = 0;
   f->hdr_dirty = 1;
   if (f->eof) return 0;

   pthread_mutex_lock(&f->mu);
   while (f->state[f->cur] != SQB_CH_READY && f->reader_errno == 0) pthread_cond_wait(&f->cv, &f->mu);
   const int err = f->state[f->cur] == SQB_CH_READY ? 0 : f->reader_errno;
   pthread_mutex_unlock(&f->mu);
   if (err) {
      if (err == ENODEV) fprintf(stderr, "seeq-b200: %s\n", sqbLastError());
      seeqerr = 0;
      errno = err;                                        /* errno tells */
      return -1;
   }
   f->buf = f->chunk[f->cur].buf;
   f->len = f->chunk[f->cur].len;
   f->eof = f->chunk[f->cur].last;
   return f->len > 0 ? 1 : 0;
}

/* ------------------------------------------------------------------------ */
/* batch results of the resident chunk                                       */
/* ------------------------------------------------------------------------ */
static int ensure_results(sqb_file_t *f, seeq_t *sq, int opt)
{
   const sqb_seeq_t *p = (const sqb_seeq_t *)sq;
   sqb_engine_t *eng = seeqEngine(sq);
   if (eng == NULL) return -1;
   if (f->res_valid && f->res_uid == p->uid && f->res_opt == opt && f->res_eng == eng &&
       f->res_gen == sqbScanGeneration(eng)) return 0;
   const int flags = opt | SQB_KEEP_LINES | ((f->pub.flags & 1) ? SQB_FASTA : 0);
   sqb_stats_t st;
   if (sqbScanHost(eng, f->buf, f->len, flags, &st)) {
      fprintf(stderr, "seeq-b200: %s\n", sqbLastError());
      errno = EIO;
      return -1;
   }
   /* the engine's own arrays serve the iterator: they stay as they are until the next scan of this engine,
    * which changes its generation (then the chunk is scanned again) */
   uint64_t nr = 0, nl = 0;
   f->recs = sqbHostRecords(eng, &nr);
   sqbHostLineStarts(eng, &f->lines, &nl);
   f->nrecs = (size_t)nr;
   f->nlines = (size_t)nl;
   f->res_uid = p->uid;
   f->res_opt = opt;
   f->res_eng = eng;
   f->res_gen = sqbScanGeneration(eng);
   f->res_valid = 1;
   /* first record at or after the line to hand out next */
   size_t lo = 0, hi = f->nrecs;
   while (lo < hi) {
      const size_t mid = (lo + hi) / 2;
      if (f->recs[mid].line < f->cur_line) lo = mid + 1;
      else hi = mid;
   }
   f->cur_rec = lo;
   return 0;
}

/* copy bytes [off, end of line) of the chunk into sq->string (getline's buffer) */
static int set_string(seeq_t *sq, const char *src, size_t n)
{
   if (sq->string == NULL || sq->bufsz < n + 1) {
      size_t cap = sq->bufsz ? sq->bufsz : 128;
      while (cap < n + 1) cap *= 2;
      char *s = realloc(sq->string, cap);
      if (s == NULL) return -1;
      sq->string = s;
      sq->bufsz = cap;
   }
   memcpy(sq->string, src, n);
   sq->string[n] = 0;
   return 0;
}

static size_t line_length(const sqb_file_t *f, size_t off)
{
   const char *nl = memchr(f->buf + off, '\n', f->len - off);
   return nl ? (size_t)(nl - (f->buf + off)) : f->len - off;
}

/* FASTA: sqfile->info = last header line before chunk offset `off` (a line start).  The cursor walks
 * forward from the line served before; info is rewritten only when the header in force changes. */
static int update_info(sqb_file_t *f, size_t off)
{
   if (off < f->hdr_scan) {                         /* (not on the iterator's path: start over) */
      f->hdr_scan = 0;
      f->hdr_seen = 0;
      f->hdr_dirty = 1;
   }
   size_t p = f->hdr_scan;
   while (p < off) {
      if (f->buf[p] == '>') {
         f->hdr_off = p;
         f->hdr_seen = 1;
         f->hdr_dirty = 1;
      }
      const char *nl = memchr(f->buf + p, '\n', off - p);
      if (nl == NULL) break;                        /* off is a line start: not reached */
      p = (size_t)(nl - f->buf) + 1;
   }
   f->hdr_scan = off;
   if (!f->hdr_dirty) return 0;
   char *h = NULL;
   if (f->hdr_seen) {
      const size_t n = line_length(f, f->hdr_off);
      h = malloc(n + 1);
      if (h == NULL) { seeqerr = 666; return -1; }  /* seeq.c:370 */
      memcpy(h, f->buf + f->hdr_off, n);
      h[n] = 0;
   } else if (f->last_header) {
      h = strdup(f->last_header);
      if (h == NULL) { seeqerr = 666; return -1; }
   }
   if (h) {
      free(f->pub.info);
      f->pub.info = h;
   }
   f->hdr_dirty = 0;
   return 0;
}

/* hand out counted line k of the chunk with its matches */
static int serve_line(sqb_file_t *f, seeq_t *sq, size_t k, size_t rec0, size_t nrec)
{
   const size_t off = (size_t)f->lines[k];
   if (set_string(sq, f->buf + off, line_length(f, off))) return -1;
   if (sqb_store_matches(sq, f->recs + rec0, nrec)) return -1;
   f->pub.line = f->line_base + k + 1;
   if ((f->pub.flags & 1) && update_info(f, off)) return -1;
   return 0;
}

/* the reference leaves the LAST line read in sq->string when it runs into the
 * end of the input (getline buffer); do the same */
static int leave_last_line(sqb_file_t *f, seeq_t *sq)
{
   if (f->len == 0) return 0;
   size_t end = f->len;
   if (f->buf[end - 1] == '\n') end--;                 /* getline's '\n' is stripped */
   const char *nl = end ? memrchr(f->buf, '\n', end) : NULL;
   const size_t start = nl ? (size_t)(nl - f->buf) + 1 : 0;
   return set_string(sq, f->buf + start, end - start);
}

/* ------------------------------------------------------------------------ */
/* seeqFileMatch                                                             */
/* ------------------------------------------------------------------------ */
long seeqFileMatch(seeqfile_t *sqfile, seeq_t *sq, int match_opt, int file_opt)
{
   seeqerr = 0;
   sqb_file_t *f = (sqb_file_t *)sqfile;

   if (file_opt == SQ_COUNTMATCH) match_opt = (match_opt & ~MASK_MATCH) | SQ_ALL;
   else if (file_opt == SQ_COUNTLINES) match_opt = (match_opt & ~MASK_MATCH) | SQ_FIRST;

   if (sqfile->fdi == NULL) { seeqerr = 10; return -1; }
   if (f->magic != SQB_FILE_MAGIC) { errno = EINVAL; return -1; }

   const int opt = match_opt & (MASK_MATCH | MASK_NONDNA);
   const size_t startline = sqfile->line;
   long count = 0;

   /* ---- whole-input counts: one count-only batch per chunk ---------------- */
   if (file_opt == SQ_COUNTLINES || file_opt == SQ_COUNTMATCH) {
      sqb_engine_t *eng = seeqEngine(sq);
      if (eng == NULL) return -1;
      const int flags = opt | SQB_COUNT_ONLY | ((sqfile->flags & 1) ? SQB_FASTA : 0);
      sq->hits = 0;
      for (;;) {
         /* rest of the resident chunk, from the next line to hand out */
         size_t off = f->len;
         if (f->started && f->len > 0) {
            if (f->cur_line == 0) off = 0;
            else if (f->res_valid && f->cur_line < f->nlines) off = (size_t)f->lines[f->cur_line];
         }
         if (off < f->len) {
            sqb_stats_t st;
            if (sqbScanHost(eng, f->buf + off, f->len - off, flags, &st)) {
               fprintf(stderr, "seeq-b200: %s\n", sqbLastError());
               errno = EIO;
               return -1;
            }
            count += (long)(file_opt == SQ_COUNTLINES ? st.nmatched : st.nrecs);
            sqfile->line += (size_t)st.nlines;
            if (leave_last_line(f, sq)) return -1;
            /* mark the chunk as consumed */
            if (!f->res_valid) f->nlines = f->cur_line + (size_t)st.nlines;
            f->cur_line = f->nlines;
            f->cur_rec = f->nrecs;
         }
         const int more = chunk_load(f);
         if (more < 0) return -1;
         if (more == 0) break;
      }
      return sqfile->line == startline ? 0 : count;
   }

   /* ---- line iterator ----------------------------------------------------- */
   for (;;) {
      if (!f->started || f->len == 0 || (f->res_valid && f->cur_line >= f->nlines)) {
         /* the resident chunk (if any) is used up */
         if (f->started && f->len > 0 && sqfile->line != startline) {
            if (leave_last_line(f, sq)) return -1;
         }
         const int more = chunk_load(f);
         if (more < 0) return -1;
         if (more == 0) return sqfile->line == startline ? 0 : count;
      }
      if (ensure_results(f, sq, opt)) return -1;
      if (f->cur_line >= f->nlines) continue;              /* chunk without counted lines */

      if (file_opt == SQ_MATCH) {
         /* jump to the next line that owns a record */
         if (f->cur_rec >= f->nrecs) {
            sqfile->line = f->line_base + f->nlines;
            f->cur_line = f->nlines;
            sq->hits = 0;
            continue;
         }
         const size_t k = f->recs[f->cur_rec].line;
         size_t r1 = f->cur_rec;
         while (r1 < f->nrecs && f->recs[r1].line == k) r1++;
         if (serve_line(f, sq, k, f->cur_rec, r1 - f->cur_rec)) return -1;
         f->cur_rec = r1;
         f->cur_line = k + 1;
         return 1;
      }

      /* SQ_ANY and SQ_NOMATCH walk line by line */
      while (f->cur_line < f->nlines) {
         const size_t k = f->cur_line;
         size_t r1 = f->cur_rec;
         while (r1 < f->nrecs && f->recs[r1].line == k) r1++;
         const size_t nrec = r1 - f->cur_rec;
         if (file_opt == SQ_NOMATCH && nrec > 0) {
            /* a matching line is consumed silently; its hits add to the count
             * returned if the input ends here (seeq.c:382-391) */
            count += (long)nrec;
            sqfile->line = f->line_base + k + 1;
            sq->hits = nrec;          /* refreshed below if this was the last line */
            if (k + 1 == f->nlines) {
               if (serve_line(f, sq, k, f->cur_rec, nrec)) return -1;
            }
            f->cur_rec = r1;
            f->cur_line = k + 1;
            continue;
         }
         if (serve_line(f, sq, k, f->cur_rec, nrec)) return -1;
         f->cur_rec = r1;
         f->cur_line = k + 1;
         return 1;
      }
   }
}

/* ------------------------------------------------------------------------ */
/* seeq(): the CLI's formatter                                               */
/* ------------------------------------------------------------------------ */
static void put_range(const char *s, size_t from, size_t to)
{
   if (to > from) fwrite(s + from, 1, to - from, stdout);
}

/* ---- the formatter's fast path -----------------------------------------------------------------------
 * seeq()'s match loop below goes through the public iterator: one seeqFileMatch call, one copy of the line
 * into sq->string, one match stack and three or four fprintf per record -- 280 ns per record, 1.3 s for the
 * 4.8 M records of BASELINE config 2, more than reading and matching the 1.5 GB file together (r4k).  For
 * plain-text input (no FASTA header rule), no colour and no -i the same bytes are produced here straight
 * from the chunk's ordered records into one large buffer: decimal conversion by hand, one fwrite per MiB. */
#define SQB_OB_CAP ((size_t)1 << 20)
typedef struct {
   char  *p;
   size_t n;
} sqb_obuf_t;

static void ob_flush(sqb_obuf_t *o)
{
   if (o->n) fwrite(o->p, 1, o->n, stdout);
   o->n = 0;
}
static void ob_room(sqb_obuf_t *o, size_t need)
{
   if (o->n + need > SQB_OB_CAP) ob_flush(o);
}
static void ob_bytes(sqb_obuf_t *o, const char *s, size_t n)
{
   if (n > SQB_OB_CAP / 2) {                    /* a very long line goes out on its own */
      ob_flush(o);
      fwrite(s, 1, n, stdout);
      return;
   }
   ob_room(o, n);
   memcpy(o->p + o->n, s, n);
   o->n += n;
}
static void ob_char(sqb_obuf_t *o, char c)
{
   ob_room(o, 1);
   o->p[o->n++] = c;
}
static void ob_long(sqb_obuf_t *o, long v)     /* "%ld" */
{
   char t[24];
   int k = 0;
   unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
   do { t[k++] = (char)('0' + u % 10); u /= 10; } while (u);
   ob_room(o, (size_t)k + 1);
   if (v < 0) o->p[o->n++] = '-';
   while (k) o->p[o->n++] = t[--k];
}

/* 0: the input is used up; -1: error (errno / seeqerr tell, as after seeqFileMatch) */
static long fast_match_output(sqb_file_t *f, seeq_t *sq, int opt, const struct seeqarg_t *args)
{
   const int mopt = opt & (MASK_MATCH | MASK_NONDNA);
   sqb_obuf_t o = {malloc(SQB_OB_CAP), 0};
   if (o.p == NULL) return -1;
   long rv = 0;
   for (;;) {
      if (!f->started || f->len == 0 || (f->res_valid && f->cur_line >= f->nlines)) {
         const int more = chunk_load(f);
         if (more < 0) { rv = -1; break; }
         if (more == 0) break;
      }
      if (ensure_results(f, sq, mopt)) { rv = -1; break; }
      for (size_t r = f->cur_rec; r < f->nrecs; r++) {
         const sqb_rec_t *rc = &f->recs[r];
         const size_t k = rc->line, off = (size_t)f->lines[k];
         /* every line of plain text is a counted line: the next start is one behind this line's '\n' */
         const size_t len = k + 1 < f->nlines ? (size_t)f->lines[k + 1] - off - 1 : line_length(f, off);
         const char *s = f->buf + off;
         const long line = (long)(f->line_base + k + 1);
         if (args->compact) {
            ob_long(&o, line); ob_char(&o, ':');
            ob_long(&o, (long)rc->start); ob_char(&o, '-');
            ob_long(&o, (long)rc->end - 1); ob_char(&o, ':');
            ob_long(&o, (long)rc->dist);
         } else {
            if (args->showline) { ob_long(&o, line); ob_char(&o, ' '); }
            if (args->showpos) { ob_long(&o, (long)rc->start); ob_char(&o, '-'); ob_long(&o, (long)rc->end - 1); ob_char(&o, ' '); }
            if (args->showdist) { ob_long(&o, (long)rc->dist); ob_char(&o, ' '); }
            /* strings are printed up to their first NUL, like "%s" */
            const char *nul = memchr(s, 0, len);
            const size_t slen = nul ? (size_t)(nul - s) : len;
            const size_t a = rc->start < slen ? rc->start : slen;
            const size_t b = rc->end < slen ? rc->end : slen;
            if (args->matchonly) {
               if (b > a) ob_bytes(&o, s + a, b - a);
            } else if (args->prefix) {
               ob_bytes(&o, s, a);
            } else if (args->endline) {
               if (slen > b) ob_bytes(&o, s + b, slen - b);
            } else if (args->split) {
               ob_bytes(&o, s, a);
               ob_char(&o, '\t');
               if (b > a) ob_bytes(&o, s + a, b - a);
               ob_char(&o, '\t');
               if (slen > b) ob_bytes(&o, s + b, slen - b);
            } else if (args->printline) {
               ob_bytes(&o, s, slen);
            }
         }
         ob_char(&o, '\n');
      }
      f->cur_rec = f->nrecs;
      f->cur_line = f->nlines;
      f->pub.line = f->line_base + f->nlines;
   }
   ob_flush(&o);
   free(o.p);
   return rv;
}

int seeq(char *expression, char *input, struct seeqarg_t args)
{
   seeq_t *sq = seeqNew(expression, args.dist, args.memory);
   if (sq == NULL) {
      fprintf(stderr, "error in 'seeqNew()'; %s\n:", seeqPrintError());
      return EXIT_FAILURE;
   }
   if (args.verbose) fprintf(stderr, "opening input file... ");
   seeqfile_t *in = seeqOpen(input);
   if (in == NULL) {
      fprintf(stderr, "error in 'seeqOpen()': %s\n", seeqPrintError());
      seeqFree(sq);
      return EXIT_FAILURE;
   }
   const int fasta = in->flags & 1;
   clock_t t0 = 0;
   if (args.verbose) {
      fprintf(stderr, "\nmatching...\n");
      t0 = clock();
   }

   int opt = 0;
   if (args.non_dna == 1) opt |= SQ_CONVERT;
   else if (args.non_dna == 2) opt |= SQ_IGNORE;

   if (args.count) {
      const long n = seeqFileMatch(in, sq, opt, SQ_COUNTLINES);
      if (n < 0) fprintf(stderr, "error in 'seeqFileMatch()': %s\n", seeqPrintError());
      else fprintf(stdout, "%ld\n", n);
   } else {
      if (args.all) {                 /* -a wins over -b and implies match-only */
         opt |= SQ_ALL;
         args.matchonly = 1;
      } else if (args.best) {
         opt |= SQ_BEST;
      }
      /* the FASTA header is echoed only when nothing is printed in front of
       * the sequence (seeq.c:117-121) */
      const int echo_header = fasta && !args.split && !args.showline && !args.showpos && !args.showdist;
      const int colour = COLOR_TERMINAL && isatty(fileno(stdout));
      long rv;
      if (args.invert) {
         while ((rv = seeqFileMatch(in, sq, opt, SQ_NOMATCH)) > 0) {
            if (args.showline) fprintf(stdout, "%ld ", (long)in->line);
            if (echo_header) fprintf(stdout, "%s\n", in->info);
            fprintf(stdout, "%s\n", sq->string);
         }
      } else if (!fasta && !colour && in->fdi != NULL && getenv("SEEQ_B200_SLOW_FORMATTER") == NULL) {
         seeqerr = 0;
         rv = fast_match_output((sqb_file_t *)in, sq, opt, &args);
      } else {
         while ((rv = seeqFileMatch(in, sq, opt, SQ_MATCH)) > 0) {
            const char *s = sq->string;
            match_t *mt;
            while ((mt = seeqMatchIter(sq)) != NULL) {
               if (args.compact) {
                  fprintf(stdout, "%ld:%ld-%ld:%ld", (long)in->line, (long)mt->start, (long)mt->end - 1, (long)mt->dist);
               } else {
                  if (args.showline) fprintf(stdout, "%ld ", (long)in->line);
                  if (args.showpos) fprintf(stdout, "%ld-%ld ", (long)mt->start, (long)mt->end - 1);
                  if (args.showdist) fprintf(stdout, "%ld ", (long)mt->dist);
                  if (echo_header) fprintf(stdout, "%s\n", in->info);
                  /* strings are printed up to their first NUL, like "%s" */
                  const size_t slen = strlen(s);
                  const size_t a = mt->start < slen ? mt->start : slen;
                  const size_t b = mt->end < slen ? mt->end : slen;
                  if (args.matchonly) {
                     put_range(s, a, b);
                  } else if (args.prefix) {
                     put_range(s, 0, a);
                     /* the reference cuts the line here for good (seeq.c:148) */
                     sq->string[a] = 0;
                  } else if (args.endline) {
                     put_range(s, b, slen);
                  } else if (args.split) {
                     put_range(s, 0, a);
                     fputc('\t', stdout);
                     put_range(s, a, b);
                     fputc('\t', stdout);
                     put_range(s, b, slen);
                  } else if (args.printline) {
                     if (colour) {
                        put_range(s, 0, a);
                        fputs(mt->dist ? BOLDRED : BOLDGREEN, stdout);
                        put_range(s, a, b);
                        fputs(RESET, stdout);
                        put_range(s, b, slen);
                     } else {
                        put_range(s, 0, slen);
                     }
                  }
               }
               fputc('\n', stdout);
            }
         }
      }
      if (rv == -1) fprintf(stderr, "error in 'seeqFileMatch()': %s\n", seeqPrintError());
   }

   if (args.verbose) {
      fprintf(stderr, "memory: %.2f MB (DFA: %.2f MB, trie: %.2f MB)\n", 0.0, 0.0, 0.0);
      fprintf(stderr, "done in %.3fs\n", (double)(clock() - t0) / CLOCKS_PER_SEC);
   }
   seeqFree(sq);
   seeqClose(in);
   return EXIT_SUCCESS;
}
