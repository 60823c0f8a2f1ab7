/*
 * seeq.h -- file driver and formatter interface of the B200-native seeq.
 *
 * Written from scratch; same names, values and layouts as the interface of
 * /root/reference/src/seeq.h (seeqarg_t :35-53, seeqfile_t :55-60, SQ_ANY..
 * SQ_COUNTMATCH :64-68, prototypes :70-73) so that the reference CLI
 * (seeq-main.c) and programs written against seeq.h re-link unchanged.
 *
 * seeqFileMatch keeps its line-by-line iterator contract, but underneath the
 * file is read in large newline-aligned chunks that are matched on the GPU in
 * one batch each (DESIGN.md "iterator over a batch engine").
 */
#ifndef SEEQ_B200_SEEQ_H_
#define SEEQ_B200_SEEQ_H_
#ifndef _SEEQ_H_
#define _SEEQ_H_

#define SEEQ_VERSION "seeq-1.2"

#include "libseeq.h"
#include <stdlib.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct seeqfile_t seeqfile_t;

/* Output switches of seeq(); all are 0/1 except dist (the distance
 * threshold), non_dna (0 fail, 1 convert, 2 ignore) and memory (accepted and
 * ignored: there is no DFA cache to bound). */
struct seeqarg_t {
   int showdist;
   int showpos;
   int showline;
   int printline;
   int matchonly;
   int count;
   int compact;
   int dist;
   int verbose;
   int endline;
   int prefix;
   int split;
   int invert;
   int best;
   int non_dna;
   int all;
   size_t memory;
};

/* Open input.  Only these four fields are public (callers read line and
 * info, and may overwrite fdi); the library allocates a larger private
 * object behind them. */
struct seeqfile_t {
   int     flags;   /* bit 0: FASTA (first byte of the stream is '>')        */
   size_t  line;    /* 1-based number of the current non-header line         */
   char  * info;    /* FASTA: last header line seen                          */
   FILE  * fdi;     /* input stream                                          */
};

/* file_opt of seeqFileMatch: when does a call return */
#define SQ_ANY        0   /* after every line                                */
#define SQ_MATCH      1   /* at the next line with at least one match        */
#define SQ_NOMATCH    2   /* at the next line without a match                */
#define SQ_COUNTLINES 3   /* at end of input: number of matching lines       */
#define SQ_COUNTMATCH 4   /* at end of input: number of matches              */

int          seeq          (char * expression, char * input, struct seeqarg_t args);
long         seeqFileMatch (seeqfile_t * sqfile, seeq_t * sq, int match_opt, int file_opt);
seeqfile_t * seeqOpen      (const char * file);
int          seeqClose     (seeqfile_t * sqfile);

#ifdef __cplusplus
}
#endif
#endif /* _SEEQ_H_ */
#endif /* SEEQ_B200_SEEQ_H_ */
