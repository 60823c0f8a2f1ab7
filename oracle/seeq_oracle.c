/*
 * seeq_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see seeq_oracle.h).
 *
 * Independent restatement of the reference's per-line matcher with a plain
 * capped dynamic-programming column; every function cites the reference lines
 * (under /root/reference/src/) whose behaviour it restates.
 */
#include "seeq_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* pattern parser: libseeq.c:511-603                                          */
/* ------------------------------------------------------------------------- */
int orc_parse(const char *pattern, unsigned char *keys, int *err)
{
   int dummy;
   if (err == NULL) err = &dummy;
   *err = 0;
   const int limit = (int)strlen(pattern);
   memset(keys, 0, (size_t)limit);

   int npos = 0;        /* completed positions                     */
   int open = 0;        /* inside a [...] class                    */
   char prev = 0;
   for (int i = 0; i < limit && npos < limit; i++) {
      const char ch = pattern[i];
      switch (ch) {
      case 'A': case 'a': keys[npos] |= 0x01; break;
      case 'C': case 'c': keys[npos] |= 0x02; break;
      case 'G': case 'g': keys[npos] |= 0x04; break;
      case 'T': case 't': case 'U': case 'u': keys[npos] |= 0x08; break;
      case 'N': case 'n': keys[npos] |= 0x1F; break;
      case '[':
         if (open) { *err = 2; return -1; }      /* libseeq.c:573-579 */
         open = 1;
         break;
      case ']':
         if (!open) { *err = 3; return -1; }     /* libseeq.c:580-584 */
         if (prev == '[') npos--;                /* "[]" adds nothing, :585 */
         open = 0;
         break;
      default:
         *err = 4;                               /* libseeq.c:588-591 */
         return -1;
      }
      if (!open) npos++;
      prev = ch;
   }
   if (open) { *err = 5; return -1; }            /* libseeq.c:598-601 */
   return npos;
}

/* ------------------------------------------------------------------------- */
/* text translation: seeqcore.h:89-111                                        */
/* ------------------------------------------------------------------------- */
int orc_code(unsigned char byte, int convert)
{
   switch (byte) {
   case 'A': case 'a': return 0;
   case 'C': case 'c': return 1;
   case 'G': case 'g': return 2;
   case 'T': case 't': case 'U': case 'u': return 3;
   case 'N': case 'n': return 4;
   case '\0': return 5;
   case '\n': return 6;
   default:   return convert ? 4 : 7;
   }
}

/* ------------------------------------------------------------------------- */
/* one column update: libseeq.c:766-789                                       */
/* col[0..m], values capped at tau+1.  Returns m - last_active, the           */
/* reference's min_to_match.                                                  */
/* ------------------------------------------------------------------------- */
static int col_step(int *col, const unsigned char *keys, int m, int tau, int code)
{
   const int cap = tau + 1;
   const int bit = 1 << code;
   int diag = col[0];
   int up = 0;
   int last_active = 1;
   col[0] = 0;
   for (int r = 1; r <= m; r++) {
      const int left = col[r];
      int v = diag + ((keys[r - 1] & bit) ? 0 : 1);
      if (up + 1 < v)   v = up + 1;
      if (left + 1 < v) v = left + 1;
      if (v > cap)      v = cap;
      if (v <= tau) last_active = r;
      col[r] = v;
      up = v;
      diag = left;
   }
   return m - last_active;
}

/* root column: libseeq.c:681-682 ([0,1,..,tau,tau+1,tau+1,...]) */
static void col_root(int *col, int m, int tau)
{
   for (int r = 0; r <= m; r++) col[r] = r <= tau ? r : tau + 1;
}

int orc_distances(const char *text, const unsigned char *keys, int m, int tau,
                  int *dist, int *min_to_match)
{
   int *col = malloc((size_t)(m + 1) * sizeof(int));
   if (col == NULL) return -1;
   col_root(col, m, tau);
   int n = 0;
   for (const unsigned char *p = (const unsigned char *)text; ; p++) {
      const int code = orc_code(*p, 0);
      if (code > 4) break;
      const int mtm = col_step(col, keys, m, tau, code);
      dist[n] = col[m];
      if (min_to_match) min_to_match[n] = mtm;
      n++;
   }
   free(col);
   return n;
}

/* ------------------------------------------------------------------------- */
/* record list helper (seeqAddMatch, libseeq.c:427-443)                       */
/* ------------------------------------------------------------------------- */
static int push_rec(orc_rec_t **recs, size_t *cap, size_t pos, orc_rec_t r)
{
   if (pos >= *cap) {
      size_t ncap = *cap ? *cap * 2 : 16;
      orc_rec_t *p = realloc(*recs, ncap * sizeof(orc_rec_t));
      if (p == NULL) return -1;
      *recs = p;
      *cap = ncap;
   }
   (*recs)[pos] = r;
   return 0;
}

/* ------------------------------------------------------------------------- */
/* the matcher proper: libseeq.c:216-352                                      */
/*                                                                            */
/* Steps w..slen are executed with an automaton that is fresh at step w;      */
/* events are recorded only for steps in [own_lo, own_hi).  The plain         */
/* seeqStringMatch is w = own_lo = 0, own_hi = slen + 1.                      */
/* ------------------------------------------------------------------------- */
static long scan_steps(const char *data, int slen, const unsigned char *keys,
                       const unsigned char *rkeys, int m, int tau, int options,
                       int w, int own_lo, int own_hi, orc_rec_t **recs,
                       size_t *cap, size_t base)
{
   const int mode    = options & 0x03;
   const int best    = mode == ORC_BEST;
   const int keep_on = best || mode == ORC_ALL;        /* libseeq.c:219-221 */
   const int nondna  = options & 0x0C;
   const int ignore  = nondna == ORC_IGNORE;           /* libseeq.c:223-226 */
   const int convert = nondna == ORC_CONVERT;
   const int stream  = options & ORC_STREAM;           /* libseeq.c:228     */
   const int cap_d   = tau + 1;

   int *col  = malloc((size_t)(m + 1) * sizeof(int));
   int *rcol = malloc((size_t)(m + 1) * sizeof(int));
   if (col == NULL || rcol == NULL) { free(col); free(rcol); return -1; }
   col_root(col, m, tau);

   size_t hits = 0;
   int best_d = cap_d;                                  /* libseeq.c:240-243 */
   int streak = cap_d;
   int flag = 0;
   int stop = 0;

   for (int i = w; i <= slen && i < own_hi; i++) {      /* libseeq.c:250     */
      const int code = orc_code((unsigned char)data[i], convert);
      int cur = cap_d;
      int min_to_match = 0;
      if (code < 5) {                                   /* libseeq.c:255-264 */
         min_to_match = col_step(col, keys, m, tau, code);
         cur = col[m];
      }
      else if (code == 6 && stream) continue;           /* libseeq.c:265     */
      else if (code == 7 && ignore) continue;           /* libseeq.c:266     */
      else stop = 1;                                    /* libseeq.c:267-270 */

      if (slen - i - 1 < min_to_match) {                /* libseeq.c:272-275 */
         cur = cap_d;
         stop = 1;
      }

      if (streak >= cur) flag = 0;                      /* libseeq.c:278     */

      const int perfect = streak == 0;                  /* libseeq.c:286-288 */
      const int rising  = streak <= tau && streak < cur;
      if ((perfect || rising) && !flag && (!best || streak < best_d)) {
         flag = 1;
         /* reverse pass, libseeq.c:290-315 */
         int j = 0, d = cap_d, last_d, skipped = 0;
         col_root(rcol, m, tau);
         do {
            j++;
            const int rc = orc_code((unsigned char)data[i - j], convert);
            last_d = d;
            if (rc < 5) {
               skipped = 0;
               col_step(rcol, rkeys, m, tau, rc);
               d = rcol[m];
            } else {
               skipped++;
            }
         } while (d > streak && j < i);
         j = (last_d < d ? j - 1 : j) - skipped;
         if (i >= own_lo) {
            orc_rec_t r = { 0, (uint64_t)(i - j), (uint64_t)i, (uint64_t)streak };
            if (best) {                                 /* libseeq.c:321-325 */
               if (push_rec(recs, cap, base, r)) { hits = (size_t)-1; break; }
               hits = 1;
               best_d = streak;
            } else {
               if (push_rec(recs, cap, base + hits, r)) { hits = (size_t)-1; break; }
               hits++;
            }
            if (!keep_on) stop = 1;                     /* libseeq.c:330     */
         }
      }
      if (stop) break;                                  /* libseeq.c:334     */
      streak = cur;                                     /* libseeq.c:337     */
   }
   free(col);
   free(rcol);
   return (long)hits;
}

static unsigned char *reversed_keys(const unsigned char *keys, int m)
{
   unsigned char *rk = malloc((size_t)m);
   if (rk == NULL) return NULL;
   for (int i = 0; i < m; i++) rk[i] = keys[m - 1 - i];   /* libseeq.c:89 */
   return rk;
}

long orc_string_match(const char *data, const unsigned char *keys, int m,
                      int tau, int options, orc_rec_t **recs, size_t *cap)
{
   unsigned char *rk = reversed_keys(keys, m);
   if (rk == NULL) return -1;
   const int slen = (int)strlen(data);                  /* libseeq.c:245 */
   long n = scan_steps(data, slen, keys, rk, m, tau, options, 0, 0, slen + 1,
                       recs, cap, 0);
   free(rk);
   return n;
}

long orc_string_match_segmented(const char *data, const unsigned char *keys,
                                int m, int tau, int options, int seg, int warm,
                                orc_rec_t **recs, size_t *cap)
{
   unsigned char *rk = reversed_keys(keys, m);
   if (rk == NULL) return -1;
   const int slen = (int)strlen(data);
   const int nondna  = options & 0x0C;
   const int convert = nondna == ORC_CONVERT;
   size_t total = 0;
   for (int lo = 0; lo <= slen; lo += seg) {
      /* walk back over `warm` automaton-visible bytes */
      int w = lo, seen = 0;
      while (w > 0 && seen < warm) {
         w--;
         if (orc_code((unsigned char)data[w], convert) < 5) seen++;
      }
      long n = scan_steps(data, slen, keys, rk, m, tau, options, w, lo, lo + seg,
                          recs, cap, total);
      if (n < 0) { free(rk); return -1; }
      total += (size_t)n;
   }
   free(rk);
   return (long)total;
}

/* ------------------------------------------------------------------------- */
/* file driver: seeq.c:293-392 with file_opt == SQ_ANY, plus the FASTA probe  */
/* of seeqOpen (seeq.c:243-253)                                               */
/* ------------------------------------------------------------------------- */
long orc_buffer_scan(const char *buf, size_t n, const unsigned char *keys,
                     int m, int tau, int options, orc_rec_t **recs,
                     size_t *cap, uint64_t *nlines, uint64_t *nmatched)
{
   unsigned char *rk = reversed_keys(keys, m);
   if (rk == NULL) return -1;
   const int fasta = n > 0 && buf[0] == '>';
   uint64_t line = 0, matched = 0;
   size_t total = 0;
   char *tmp = NULL;
   size_t tmpcap = 0;
   size_t pos = 0;
   while (pos < n) {
      /* getline: up to and including the next '\n' (seeq.c:361) */
      const char *nl = memchr(buf + pos, '\n', n - pos);
      size_t len = nl ? (size_t)(nl - (buf + pos)) : n - pos;
      if (len + 1 > tmpcap) {
         tmpcap = (len + 1) * 2;
         char *t = realloc(tmp, tmpcap);
         if (t == NULL) { free(tmp); free(rk); return -1; }
         tmp = t;
      }
      memcpy(tmp, buf + pos, len);
      tmp[len] = 0;                                      /* seeq.c:364 */
      pos += len + (nl ? 1 : 0);

      if (fasta && tmp[0] == '>') continue;              /* seeq.c:367-374 */
      line++;                                            /* seeq.c:377 */

      const int slen = (int)strlen(tmp);
      long h = scan_steps(tmp, slen, keys, rk, m, tau, options, 0, 0, slen + 1,
                          recs, cap, total);
      if (h < 0) { free(tmp); free(rk); return -1; }
      for (long k = 0; k < h; k++) (*recs)[total + (size_t)k].line = line;
      total += (size_t)h;
      if (h > 0) matched++;
   }
   free(tmp);
   free(rk);
   if (nlines) *nlines = line;
   if (nmatched) *nmatched = matched;
   return (long)total;
}
