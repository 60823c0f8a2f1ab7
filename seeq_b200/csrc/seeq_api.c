/*
 * seeq_api.c -- the libseeq C API on top of the GPU engine (host side, C99).
 *
 * Mirrors the operator interface of the reference library (same names,
 * argument meaning, return values and error codes) for the matching path:
 *
 *   seeqNew          /root/reference/src/libseeq.c:43-138
 *   seeqFree         :140-168
 *   seeqStringMatch  :171-352     (the matching itself runs in the CUDA kernels)
 *   seeqAddMatch     :427-443
 *   seeqMatchIter    :446-465
 *   seeqGetString    :467-486
 *   seeqPrintError   :488-507
 *   pattern parser   :511-603
 *
 * Written from scratch.  There is no CPU matcher here: if no CUDA device is
 * usable, the matching calls fail loudly (-1, errno = ENODEV, message on
 * stderr).
 */
#define _GNU_SOURCE
#include "sqb_private.h"

#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int seeqerr = 0;

/* texts of seeqPrintError(), indexed by seeqerr (libseeq.c:28-41) */
static const char *const error_text[] = {
   "Check errno",
   "Illegal matching distance value",
   "Incorrect pattern (double opening brackets)",
   "Incorrect pattern (double closing brackets)",
   "Incorrect pattern (illegal character)",
   "Incorrect pattern (missing closing bracket)",
   "Illegal path value passed to 'trie_search'",
   "Illegal nodeid passed to 'trie_getrow' (node is not a leaf).\n",
   "Illegal path value passed to 'trie_insert'",
   "Pattern length must be larger than matching distance",
   "Passed seeq_t struct does not contain a valid file pointer",
   "End of line reached.",
};

/* --------------------------------------------------------------------------
 * Pattern text -> one class byte per position.  A position is either a single
 * letter or a bracketed set; 'N' is every class including text-N (0x1F); an
 * empty set "[]" contributes no position.  Errors: 2 "[[", 3 stray "]",
 * 4 illegal character, 5 unterminated set.
 * ------------------------------------------------------------------------ */
int sqb_parse_pattern(const char *text, char *keys)
{
   seeqerr = 0;
   const size_t limit = strlen(text);
   memset(keys, 0, limit);

   size_t pos = 0;            /* position being filled   */
   int in_set = 0;
   int set_is_empty = 0;
   for (size_t i = 0; i < limit; i++) {
      unsigned char bits = 0;
      switch (text[i]) {
      case 'a': case 'A': bits = 0x01; break;
      case 'c': case 'C': bits = 0x02; break;
      case 'g': case 'G': bits = 0x04; break;
      case 't': case 'T': case 'u': case 'U': bits = 0x08; break;
      case 'n': case 'N': bits = 0x1F; break;
      case '[':
         if (in_set) { seeqerr = 2; return -1; }
         in_set = 1;
         set_is_empty = 1;
         continue;
      case ']':
         if (!in_set) { seeqerr = 3; return -1; }
         in_set = 0;
         if (!set_is_empty) pos++;
         continue;
      default:
         seeqerr = 4;
         return -1;
      }
      keys[pos] |= (char)bits;
      if (in_set) set_is_empty = 0;
      else pos++;
   }
   if (in_set) { seeqerr = 5; return -1; }
   return (int)pos;
}

/* -------------------------------------------------------------------------- */
seeq_t *seeqNew(const char *pattern, int mismatches, size_t maxmemory)
{
   if (mismatches < 0) { seeqerr = 1; return NULL; }

   const size_t plen = strlen(pattern);
   sqb_seeq_t *p = calloc(1, sizeof *p);
   char *keys = malloc(plen ? plen : 1);
   char *rkeys = malloc(plen ? plen : 1);
   match_t *stack = malloc(INITIAL_MATCH_STACK_SIZE * sizeof(match_t));
   if (p == NULL || keys == NULL || rkeys == NULL || stack == NULL) goto fail;

   const int wlen = sqb_parse_pattern(pattern, keys);
   if (wlen < 0) goto fail;                      /* seeqerr set by the parser */
   for (int i = 0; i < wlen; i++) rkeys[i] = keys[wlen - 1 - i];
   if (mismatches >= wlen) { seeqerr = 9; goto fail; }

   static unsigned long long next_uid = 0;
   p->magic = SQB_SEEQ_MAGIC;
   p->uid = ++next_uid;
   p->maxmemory = maxmemory;                     /* accepted, unused: no DFA */
   /* dfa / rdfa: opaque, non-NULL; laid out so that a caller peeking at the
    * reference's dfa_t header {pos,size,maxmemory,state_size,trie*} reads
    * harmless numbers (seeq.c:184-189 does that for -z) */
   for (int k = 0; k < 2; k++) {
      p->shadow[k].pos = 2;
      p->shadow[k].size = 2;
      p->shadow[k].maxmemory = maxmemory;
      p->shadow[k].state_size = 24 + (size_t)(wlen + 4) / 5;
      p->shadow[k].trie = &p->shadow[k].trie_head;
      p->shadow[k].trie_head[0] = 1;
      p->shadow[k].trie_head[1] = 1;
      p->shadow[k].trie_head[2] = (size_t)wlen;
   }
   seeq_t *sq = &p->pub;
   sq->hits = 0;
   sq->stacksize = INITIAL_MATCH_STACK_SIZE;
   sq->match = stack;
   sq->bufsz = 0;
   sq->string = NULL;
   sq->tau = mismatches;
   sq->wlen = wlen;
   sq->keys = keys;
   sq->rkeys = rkeys;
   sq->dfa = &p->shadow[0];
   sq->rdfa = &p->shadow[1];
   return sq;

fail:
   free(p); free(keys); free(rkeys); free(stack);
   return NULL;
}

void seeqFree(seeq_t *sq)
{
   if (sq == NULL) return;
   sqb_seeq_t *p = (sqb_seeq_t *)sq;
   if (p->magic == SQB_SEEQ_MAGIC && p->engine) sqbEngineFree(p->engine);
   free(sq->string);
   free(sq->match);
   free(sq->keys);
   free(sq->rkeys);
   p->magic = 0;
   free(p);
}

sqb_engine_t *seeqEngine(seeq_t *sq)
{
   sqb_seeq_t *p = (sqb_seeq_t *)sq;
   if (p->magic != SQB_SEEQ_MAGIC) { errno = EINVAL; return NULL; }
   if (p->engine == NULL) {
      /* seeqNew accepts a pattern of any length, as the reference does (libseeq.c:43-138); the blocked
       * automaton serves up to sqbMaxPatternLength() positions.  A longer pattern is not a missing
       * device: say so with an errno of its own (E2BIG) instead of ENODEV */
      if (sq->wlen > sqbMaxPatternLength()) {
         fprintf(stderr, "seeq-b200: pattern of %d positions exceeds the %d supported by the GPU matcher\n",
                 sq->wlen, sqbMaxPatternLength());
         errno = E2BIG;
         return NULL;
      }
      p->engine = sqbEngineNew((const unsigned char *)sq->keys, sq->wlen, sq->tau, -1);
      if (p->engine == NULL) {
         fprintf(stderr, "seeq-b200: %s\n", sqbLastError());
         errno = ENODEV;
      }
   }
   return p->engine;
}

int seeqAddMatch(seeq_t *sq, match_t match)
{
   if (sq->hits >= sq->stacksize) {
      const size_t grown = sq->stacksize ? 2 * sq->stacksize : 1;
      match_t *m = realloc(sq->match, grown * sizeof(match_t));
      if (m == NULL) return -1;
      sq->match = m;
      sq->stacksize = grown;
   }
   sq->match[sq->hits++] = match;
   return 0;
}

/* Store n records in sq->match right-to-left, which is how the reference
 * leaves them (libseeq.c:345-349): seeqMatchIter pops from the back and so
 * yields them left-to-right, and match[0] is the LAST match of the line. */
int sqb_store_matches(seeq_t *sq, const sqb_rec_t *recs, size_t n)
{
   sq->hits = 0;
   for (size_t k = n; k-- > 0;) {
      match_t m = { recs[k].start, recs[k].end, recs[k].dist };
      if (seeqAddMatch(sq, m)) return -1;
   }
   return 0;
}

long seeqStringMatch(const char *data, seeq_t *sq, int options)
{
   seeqerr = 0;
   sq->hits = 0;
   sqb_engine_t *eng = seeqEngine(sq);
   if (eng == NULL) return -1;
   const int opt = (options & (MASK_MATCH | MASK_NONDNA | MASK_INPUT)) | SQB_SINGLE_LINE;
   sqb_stats_t st;
   if (sqbScanHost(eng, data, strlen(data), opt, &st)) {
      fprintf(stderr, "seeq-b200: %s\n", sqbLastError());
      errno = EIO;
      return -1;
   }
   uint64_t n = 0;
   const sqb_rec_t *recs = sqbHostRecords(eng, &n);
   if (sqb_store_matches(sq, recs, (size_t)n)) return -1;
   return (long)sq->hits;
}

long seeqBatchMatch(seeq_t *sq, const char *text, size_t nbytes, int match_opt, int file_opt,
                    const sqb_rec_t **recs, sqb_stats_t *stats)
{
   seeqerr = 0;
   sqb_engine_t *eng = seeqEngine(sq);
   if (eng == NULL) return -1;
   int opt = match_opt & (MASK_MATCH | MASK_NONDNA | SQB_FASTA | SQB_FASTQ | SQB_TIMING | SQB_KEEP_LINES);
   if (file_opt == 4)      opt = (opt & ~MASK_MATCH) | SQ_ALL | SQB_COUNT_ONLY;     /* SQ_COUNTMATCH */
   else if (file_opt == 3) opt = (opt & ~MASK_MATCH) | SQ_FIRST | SQB_COUNT_ONLY;   /* SQ_COUNTLINES */
   else if (file_opt != 0) { errno = EINVAL; return -1; }
   sqb_stats_t st;
   if (sqbScanHost(eng, text, nbytes, opt, &st)) {
      fprintf(stderr, "seeq-b200: %s\n", sqbLastError());
      errno = EIO;
      return -1;
   }
   if (stats) *stats = st;
   if (file_opt == 4) return (long)st.nrecs;
   if (file_opt == 3) return (long)st.nmatched;
   uint64_t n = 0;
   const sqb_rec_t *r = sqbHostRecords(eng, &n);
   if (recs) *recs = r;
   return (long)n;
}

match_t *seeqMatchIter(seeq_t *sq)
{
   if (sq->hits == 0) return NULL;
   sq->hits--;
   return sq->match + sq->hits;
}

char *seeqGetString(seeq_t *sq)
{
   return sq->string;
}

const char *seeqPrintError(void)
{
   if (seeqerr == 0) return strerror(errno);
   if (seeqerr > 0 && seeqerr < (int)(sizeof error_text / sizeof error_text[0])) return error_text[seeqerr];
   return "Unknown error";
}

/* ---- legacy exports, unused by the library (libseeq.c:355-424) ------------ */
mstack_t *stackNew(size_t size)
{
   if (size < 1) size = 1;
   mstack_t *st = malloc(sizeof(mstack_t) + size * sizeof(match_t));
   if (st == NULL) return NULL;
   st->size = size;
   st->pos = 0;
   return st;
}

int stackAddMatch(mstack_t **stackp, match_t match)
{
   mstack_t *st = *stackp;
   if (st->pos >= st->size) {
      const size_t grown = 2 * st->size;
      st = realloc(st, sizeof(mstack_t) + grown * sizeof(match_t));
      if (st == NULL) return -1;
      st->size = grown;
      *stackp = st;
   }
   st->match[st->pos++] = match;
   return 0;
}

int recursive_merge(size_t start, size_t end, int tau, seeq_t *sq, mstack_t **stackp)
{
   /* The reference keeps this symbol but never calls it (its call site is
    * commented out, libseeq.c:340).  Kept as a link-compatibility stub. */
   (void)start; (void)end; (void)tau; (void)sq; (void)stackp;
   return 0;
}
