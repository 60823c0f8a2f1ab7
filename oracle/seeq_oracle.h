/*
 * seeq_oracle.h -- CPU oracle for the seeq per-line Levenshtein matching path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker.
 *
 * The oracle is an independent restatement, in plain C, of what the reference
 * computes (a capped Needleman-Wunsch column per text byte plus the report
 * state machine and the reverse start-recovery pass).  It deliberately does
 * NOT use the bit-parallel Myers/Hyyro formulation of the CUDA path, so that a
 * GPU-vs-oracle comparison also checks "Myers == capped DP column".
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks it against every
 * known-answer vector of the reference test-suite (test/testset.c) and
 * tests/test_oracle_vs_ref.py fuzzes it against the compiled reference
 * (oracle/_ref/libseeq_ref.so, built by oracle/Makefile from /root/reference).
 *
 * Reference lines cited below are relative to /root/reference/src/.
 */
#ifndef SEEQ_ORACLE_H_
#define SEEQ_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* option bits, numerically identical to libseeq.h:34-48 */
#define ORC_FIRST   0x00
#define ORC_BEST    0x01
#define ORC_ALL     0x02
#define ORC_FAIL    0x00
#define ORC_CONVERT 0x04
#define ORC_IGNORE  0x08
#define ORC_STREAM  0x10

typedef struct {
   uint64_t line;   /* 1-based number among counted (non-header) lines */
   uint64_t start;  /* first byte of the match, offset in the line      */
   uint64_t end;    /* one past the last byte of the match              */
   uint64_t dist;   /* Levenshtein distance of the match                */
} orc_rec_t;

/* Pattern text -> one class byte per position (libseeq.c:511-603).
 * Returns the number of positions, or -1 with *err in 2..5. */
int orc_parse(const char *pattern, unsigned char *keys, int *err);

/* Text byte -> code 0..7 (seeqcore.h:89-111). convert != 0 selects the
 * SQ_CONVERT table (every "other" byte becomes 4 = N). */
int orc_code(unsigned char byte, int convert);

/* Capped search distance after every byte of a NUL-terminated text, i.e.
 * dist[i] = min(tau+1, min_s ed(P, T[s..i])) and the reference's
 * "min_to_match" word (libseeq.c:779-789).  Non-base bytes stop the scan.
 * Returns the number of entries written. */
int orc_distances(const char *text, const unsigned char *keys, int m, int tau,
                  int *dist, int *min_to_match);

/* seeqStringMatch restated (libseeq.c:171-352).  Matches are returned in
 * left-to-right order (the order seeqMatchIter yields them); line is set to 0.
 * *recs is grown with realloc, *cap is its capacity in records.
 * Returns the number of matches, -1 on allocation failure. */
long orc_string_match(const char *data, const unsigned char *keys, int m,
                      int tau, int options, orc_rec_t **recs, size_t *cap);

/* The seeqFileMatch(.., SQ_ANY) loop (seeq.c:293-392) over an in-memory copy
 * of a file: getline splitting, '\n' stripping, FASTA header rule, 1-based
 * line numbers.  All records of all lines are appended to *recs in file order.
 * Outputs (each may be NULL): number of counted lines, number of lines with at
 * least one match.  Returns the number of records, -1 on allocation failure. */
long orc_buffer_scan(const char *buf, size_t n, const unsigned char *keys,
                     int m, int tau, int options, orc_rec_t **recs,
                     size_t *cap, uint64_t *nlines, uint64_t *nmatched);

/* Same events, but computed segment-wise: the line is cut every `seg` bytes
 * and each segment is re-started `warm` valid bytes earlier with a fresh
 * automaton.  Used by the CPU tests to pin the warm-up length the
 * segment-parallel CUDA kernel relies on. SQ_ALL only. */
long orc_string_match_segmented(const char *data, const unsigned char *keys,
                                int m, int tau, int options, int seg, int warm,
                                orc_rec_t **recs, size_t *cap);

#ifdef __cplusplus
}
#endif
#endif
