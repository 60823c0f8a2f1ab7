/* sqb_private.h -- private layouts behind the public seeq_t / seeqfile_t.
 *
 * Both public structs are the FIRST member of a larger object allocated by
 * seeqNew / seeqOpen, so the ABI seen by callers (field offsets of libseeq.h /
 * seeq.h) is unchanged while the library keeps its own state behind them.
 */
#ifndef SQB_PRIVATE_H_
#define SQB_PRIVATE_H_

#include "seeq.h"
#include "seeq_b200.h"

#define SQB_SEEQ_MAGIC 0x5351423230305351ull   /* "SQB200SQ" */
#define SQB_FILE_MAGIC 0x5351423230304649ull   /* "SQB200FI" */

/* what seeq_t.dfa / .rdfa point to: mimics the head of the reference's dfa_t
 * {pos, size, maxmemory, state_size, trie*} and trie_t {pos, size, height}
 * (seeqcore.h:63-84) for callers that peek at it; carries no matcher state */
typedef struct {
   size_t  pos, size, maxmemory, state_size;
   void  * trie;
   size_t  trie_head[3];
} sqb_shadow_t;

typedef struct {
   seeq_t         pub;          /* must be first */
   unsigned long long magic;
   sqb_engine_t * engine;       /* created on first use */
   unsigned long long uid;      /* distinguishes seeq_t objects reusing an address */
   size_t         maxmemory;
   sqb_shadow_t   shadow[2];
} sqb_seeq_t;

/* The input file is read in large chunks that end on a line boundary, by a READER THREAD that fills one
 * pinned buffer while the batch of the other is matched and handed out (seeq.c:361 reads line by line with
 * getline).  The resident chunk and its batch results: */
#include <pthread.h>

typedef struct {
   char         * buf;          /* pinned; buf[0..len) holds whole lines (the last one may lack '\n' only at end of input) */
   size_t         cap, len;
   int            last;         /* no chunk follows */
} sqb_chunk_t;

enum { SQB_CH_FREE = 0, SQB_CH_READY = 1 };

typedef struct {
   seeqfile_t     pub;          /* must be first */
   unsigned long long magic;
   /* the resident chunk (one of chunk[]) */
   char         * buf;
   size_t         len;
   int            eof;          /* the resident chunk is the last one */
   int            started;      /* at least one chunk was loaded */
   /* batch results for (res_uid, res_opt) over the resident chunk: the engine's own arrays, valid while the
    * engine's scan generation is res_gen (a seeqStringMatch on the same seeq_t in between re-scans the chunk) */
   unsigned long long res_uid, res_gen;
   sqb_engine_t * res_eng;
   int            res_opt;
   int            res_valid;
   const sqb_rec_t * recs;      size_t nrecs;
   const uint64_t * lines;      size_t nlines;             /* offsets of counted lines */
   size_t         cur_line;     /* next counted line of the chunk to hand out */
   size_t         cur_rec;      /* first record with line >= cur_line         */
   size_t         line_base;    /* counted lines before the resident chunk    */
   char         * last_header;  /* FASTA: last header seen before the chunk   */
   /* FASTA: forward cursor over the headers of the resident chunk.  Lines are handed out in increasing
    * order, so the header in force for a line is found by walking on from the previous line served:
    * O(chunk) per chunk in total (the reference keeps the last header as it reads, seeq.c:367-374) */
   size_t         hdr_scan;     /* a line start: everything before it has been examined    */
   size_t         hdr_off;      /* offset of the last header line before hdr_scan          */
   int            hdr_seen;     /* 1: hdr_off is valid (else the header is last_header)    */
   int            hdr_dirty;    /* 1: pub.info does not show the header in force           */
   /* reader thread */
   pthread_t      thread;
   int            thread_started;
   pthread_mutex_t mu;
   pthread_cond_t cv;
   sqb_chunk_t    chunk[2];
   int            state[2];     /* SQB_CH_* (under mu) */
   int            cur;          /* index of the resident chunk */
   int            stop;         /* seeqClose: the reader is to give up */
   int            reader_errno; /* != 0: the reader failed (under mu) */
   FILE         * in;
   size_t         target;       /* bytes to read per chunk */
   char         * carry;        /* partial line behind the last chunk (reader only) */
   size_t         carry_len, carry_cap;
} sqb_file_t;

int  sqb_parse_pattern (const char * text, char * keys);
int  sqb_store_matches (seeq_t * sq, const sqb_rec_t * recs, size_t n);

#endif
