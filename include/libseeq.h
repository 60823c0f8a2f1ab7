/*
 * libseeq.h -- public interface of the B200-native libseeq replacement.
 *
 * This header is written from scratch.  It declares, with the same names,
 * values, argument meaning and struct layouts, the interface that the
 * reference declares in /root/reference/src/libseeq.h (option macros :34-48,
 * INITIAL_MATCH_STACK_SIZE :52, seeqerr :56, match_t :62-66, seeq_t :68-80,
 * mstack_t :82-86, prototypes :89-100, colour macros :103-105), so that
 * existing callers (the reference CLI seeq-main.c, the CPython module
 * seeqmodule.c, user programs) compile and link against this library
 * unchanged.  Behind it the matching runs on the GPU (see seeq_b200.h and
 * DESIGN.md); there is no CPU matcher in this library.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif

#ifndef SEEQ_B200_LIBSEEQ_H_
#define SEEQ_B200_LIBSEEQ_H_
/* guard of the reference header: a translation unit sees one of the two */
#ifndef _SEEQLIB_H_
#define _SEEQLIB_H_

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIBSEEQ_VERSION "libseeq-1.1"          /* same ABI generation      */
#define LIBSEEQ_BACKEND "seeq-b200 (CUDA sm_100a)"
#define COLOR_TERMINAL 1

/* ---- options of seeqStringMatch / seeqFileMatch (bitwise OR of one value
 *      per group) ---------------------------------------------------------- */
/* which matches of a line are reported */
#define SQ_FIRST      0x00   /* stop at the first match (default)            */
#define SQ_BEST       0x01   /* the first match of minimum distance          */
#define SQ_ALL        0x02   /* every match                                  */
#define SQ_COUNT      0x03   /* accepted; behaves as SQ_FIRST                */
/* bytes that are not A C G T U N (any case) */
#define SQ_FAIL       0x00   /* the scan of the line ends there (default)    */
#define SQ_CONVERT    0x04   /* treated as a text 'N' (a mismatch)           */
#define SQ_IGNORE     0x08   /* invisible to the matcher, still counted in
                                the reported byte offsets                    */
/* where a line ends */
#define SQ_LINES      0x00   /* at '\n' or NUL (default)                     */
#define SQ_STREAM     0x10   /* at NUL only, '\n' is skipped                 */

#define MASK_MATCH    0x03
#define MASK_NONDNA   0x0C
#define MASK_INPUT    0x10

#define INITIAL_MATCH_STACK_SIZE 16

/* Error of the last library call: 0 = consult errno, 1 illegal distance,
 * 2..5 pattern syntax, 9 distance >= pattern length, 10 no file pointer. */
extern int seeqerr;

typedef struct seeq_t   seeq_t;
typedef struct match_t  match_t;
typedef struct mstack_t mstack_t;

/* One match: bytes [start, end) of the line, at edit distance dist. */
struct match_t {
   size_t start;
   size_t end;
   size_t dist;
};

/* Pattern object.  Field order and types are ABI (callers read hits, match[],
 * string, tau, wlen, keys).  dfa / rdfa are opaque, non-NULL handles. */
struct seeq_t {
   size_t    hits;        /* matches left in match[] for seeqMatchIter       */
   size_t    stacksize;   /* capacity of match[]                             */
   match_t * match;       /* matches, stored right-to-left                   */
   size_t    bufsz;       /* capacity of string                              */
   char    * string;      /* last line handed out by seeqFileMatch           */
   int       tau;         /* distance threshold                              */
   int       wlen;        /* pattern length in positions                     */
   char    * keys;        /* one base-class byte per position                */
   char    * rkeys;       /* keys reversed                                   */
   void    * dfa;         /* opaque (device context)                         */
   void    * rdfa;        /* opaque                                          */
};

struct mstack_t {
   size_t  size;
   size_t  pos;
   match_t match[];
};

seeq_t     * seeqNew         (const char * pattern, int mismatches, size_t maxmemory);
void         seeqFree        (seeq_t * sq);
match_t    * seeqMatchIter   (seeq_t * sq);
char       * seeqGetString   (seeq_t * sq);
long         seeqStringMatch (const char * data, seeq_t * sq, int options);
const char * seeqPrintError  (void);
int          seeqAddMatch    (seeq_t * sq, match_t match);

/* legacy helpers kept for link compatibility (unused by the library) */
mstack_t   * stackNew        (size_t size);
int          stackAddMatch   (mstack_t ** stackp, match_t match);
int          recursive_merge (size_t start, size_t end, int tau, seeq_t * sq, mstack_t ** stackp);

#define RESET       "\033[0m"
#define BOLDRED     "\033[1m\033[31m"
#define BOLDGREEN   "\033[1m\033[32m"

#ifdef __cplusplus
}
#endif
#endif /* _SEEQLIB_H_ */
#endif /* SEEQ_B200_LIBSEEQ_H_ */
