/*
 * seeq_b200.h -- batch / device interface of the B200-native matcher.
 *
 * This is the thin C-ABI between the C host library (libseeq API, file
 * driver) and the hand-written sm_100a kernels: plain pointers and sizes, no
 * C++ or torch types.  It is also what a binding written in another language
 * (ctypes, cgo, JNI ...) would bind, see INTEGRATION.md.
 *
 * The reference has no batch entry point: its unit of work is one line
 * (seeqStringMatch, libseeq.c:171-352) driven by the getline loop of
 * seeqFileMatch (seeq.c:293-392).  A batch call here computes exactly what
 * that loop computes for every line of a buffer, in one pass on the GPU:
 *
 *   K1  newline / line-offset scan            replaces getline + '\n' strip +
 *                                             FASTA header rule + line++
 *                                             (seeq.c:361-377)
 *   K2  forward bit-parallel matcher          replaces the forward loop of
 *                                             seeqStringMatch + dfa_step
 *                                             (libseeq.c:250-338, :779-786)
 *   K3  reverse start-recovery pass           replaces libseeq.c:290-316
 *   K4  ordered compaction of records         replaces seeqAddMatch + final
 *                                             reversal (libseeq.c:427-443,
 *                                             :345-349)
 */
#ifndef SEEQ_B200_H_
#define SEEQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sqb_engine sqb_engine_t;

/* One match record, 16 bytes, in file order (by line, then by end).
 * line is the 0-based index of the line among the COUNTED lines (FASTA
 * headers excluded) of the scanned buffer; start/end/dist are match_t's. */
typedef struct {
   uint32_t line;
   uint32_t start;
   uint32_t end;
   uint32_t dist;
} sqb_rec_t;

typedef struct {
   uint64_t nbytes;     /* bytes scanned                                     */
   uint64_t nlines;     /* counted lines                                     */
   uint64_t nmatched;   /* lines with at least one match                     */
   uint64_t nrecs;      /* records (SQ_FIRST/SQ_BEST: == nmatched)           */
   double   device_ms;  /* CUDA-event time of the kernels of this scan       */
   double   kernel_ms[8];/* only with SQB_TIMING.  Stages: [0] K1 (tokenizer, tile
                         * scan, gather)  [1] K2 (bit-plane pack + matcher)  [2]
                         * scan + K3/K4.  Single kernels: [3] the matcher
                         * (k2_bitslice / k2_forward_*)  [4] k15_pack  [5]
                         * k1_scan_classify                                    */
   uint32_t launches;   /* kernels launched by this scan                     */
   uint32_t reruns;     /* scans repeated because a capacity guess was low   */
   uint32_t path;       /* which kernels served the (last chunk of the) scan: SQB_PATH_* bits */
   uint32_t devices;    /* GPUs that took part (sqbScanHost with SEEQ_B200_DEVICES)          */
} sqb_stats_t;

#define SQB_PATH_BITSLICE 0x1   /* line-bit-sliced matcher (k2_bitslice), else the word-parallel kernels  */
#define SQB_PATH_FUSED    0x2   /* fused tokenise + bit-plane pack (k12_scan_pack), else K1 + k15_pack    */
#define SQB_PATH_CUTS     0x4   /* long lines cut into segments                                          */
#define SQB_PATH_FILTER   0x8   /* line filter on (dead-on-arrival lines / FASTQ records not packed)     */

/* flags, OR-ed into `options` next to the SQ_* bits of libseeq.h */
#define SQB_COUNT_ONLY   0x0100  /* no records: nmatched and nrecs only      */
#define SQB_FASTA        0x0200  /* lines starting with '>' are headers      */
#define SQB_SINGLE_LINE  0x0400  /* the buffer is ONE line (string API)      */
#define SQB_TIMING       0x0800  /* fill kernel_ms[] (adds event records)    */
#define SQB_KEEP_LINES   0x1000  /* sqbScanHost: also return line offsets    */
#define SQB_FASTQ        0x4000  /* the buffer holds 4-line records (@id, sequence, +, quality) and starts at
                                   * one: only the SEQUENCE line of every record (line index 1 mod 4) is matched,
                                   * the other lines count as lines and never match -- no false hits in quality
                                   * strings, and the matcher's tiles hold sequence lines only with every -x mode.
                                   * Chunked scans cut at record boundaries ('@' line, '+' two lines on).
                                   * The reference has no such mode (it scans every line, seeq.c:358-391).     */
#define SQB_DEVICE_RESULTS 0x2000 /* chunked scans (sqbScanDeviceLarge, sqbScanHost, pattern sets): the records
                                   * of all chunks stay in HBM (sqbDeviceRecordsAll) instead of travelling to
                                   * the pinned host array (sqbHostRecords)                                     */

/* ---- engine life cycle --------------------------------------------------- */
/* keys: one class byte per pattern position as produced by the parser
 * (bit0 A, bit1 C, bit2 G, bit3 T/U, 0x1F N); 0 <= tau < m.
 * device < 0 selects $SEEQ_B200_DEVICE, else $LOCAL_RANK, else 0.
 * Returns NULL on failure (no CUDA device, unsupported pattern length);
 * sqbLastError() says why.  There is no CPU fallback. */
sqb_engine_t * sqbEngineNew   (const unsigned char * keys, int m, int tau, int device);
void           sqbEngineFree  (sqb_engine_t * e);
const char   * sqbLastError   (void);
int            sqbDeviceCount (void);
int            sqbMaxPatternLength (void);

/* ---- device-resident input ----------------------------------------------- */
/* d_text: device pointer, 16-byte aligned, readable up to the next multiple of
 * 16 bytes past nbytes; nbytes < 4 GiB.  stream: a cudaStream_t (NULL = the
 * engine's own stream).  Blocks until the results are ready.
 * Returns 0, or -1 with sqbLastError(). */
int sqbScanDevice (sqb_engine_t * e, const void * d_text, size_t nbytes,
                   int options, void * stream, sqb_stats_t * stats);

/* The same in two halves, for callers that keep the device busy: up to two scans
 * may be in flight (slot 0 and 1, each with its own result arrays); Issue returns
 * as soon as the kernels are queued on `stream` (NULL: the slot's own stream),
 * Wait blocks until the scan of that slot is complete and fills stats.  A slot
 * must be waited for before it is issued again. */
int sqbScanDeviceIssue (sqb_engine_t * e, int slot, const void * d_text, size_t nbytes,
                        int options, void * stream);
int sqbScanDeviceWait  (sqb_engine_t * e, int slot, sqb_stats_t * stats);

/* A device-resident buffer of ANY size (a 12.5 GB shard of a 100 GB read set): cut into
 * newline-aligned chunks below 2 GiB ($SEEQ_B200_DEVICE_CHUNK_MB, default 1536) that are scanned
 * where they lie, back to back on `stream` (NULL: a stream of the engine); counts are summed and
 * the records of all chunks are delivered like sqbScanHost's, with buffer-global line indices
 * (sqbHostRecords; sqbHostLineStarts with SQB_KEEP_LINES).  d_text needs no alignment (the
 * 16-byte vector it lies in must be readable, as inside any CUDA allocation). */
int sqbScanDeviceLarge (sqb_engine_t * e, const void * d_text, size_t nbytes,
                        int options, void * stream, sqb_stats_t * stats);
/* with SQB_DEVICE_RESULTS: all records of the last chunked scan, in order, on the device */
const sqb_rec_t * sqbDeviceRecordsAll (sqb_engine_t * e, uint64_t * count);

/* results of the last sqbScanDevice / sqbScanDeviceWait, resident on the device */
const sqb_rec_t * sqbDeviceRecords    (sqb_engine_t * e);
const uint32_t  * sqbDeviceLineStarts (sqb_engine_t * e);   /* nlines+1 entries */
int sqbFetchRecords    (sqb_engine_t * e, sqb_rec_t * dst, uint64_t first, uint64_t count);
int sqbFetchLineStarts (sqb_engine_t * e, uint32_t * dst, uint64_t first, uint64_t count);

/* ---- host-resident input (the end-to-end path) ---------------------------- */
/* Scans a host buffer of any size: newline-aligned chunks are copied to the
 * device and matched in a software pipeline (copy of chunk k+1 overlaps the
 * kernels of chunk k); records come back with buffer-global line indices.
 * text may be pageable or pinned (sqbHostAlloc); pinned is faster. */
int sqbScanHost (sqb_engine_t * e, const char * text, size_t nbytes,
                 int options, sqb_stats_t * stats);
const sqb_rec_t * sqbHostRecords    (sqb_engine_t * e, uint64_t * count);
/* $SEEQ_B200_DEVICES = "all" or a count: sqbScanHost -- and with it seeqBatchMatch, seeqFileMatch, seeq() --
 * shards the buffer by newline-aligned byte ranges over that many GPUs of the box, one host thread and
 * engine per device inside this ONE process; results are those of the single-device scan (stats.devices
 * says how many took part).  Shards below $SEEQ_B200_MIN_SHARD_MB (16) MiB are not made.
 * A number that changes whenever a scan rewrites the arrays sqbHostRecords / sqbHostLineStarts return: */
unsigned long long sqbScanGeneration (sqb_engine_t * e);
/* byte offset of every counted line of the last sqbScanHost (on request) */
int               sqbHostLineStarts (sqb_engine_t * e, const uint64_t ** starts, uint64_t * count);

void * sqbHostAlloc (size_t nbytes);      /* pinned host memory               */
void   sqbHostFree  (void * p);
void * sqbDeviceAlloc (size_t nbytes);    /* plain device memory (benchmarks) */
void   sqbDeviceFree  (void * p);
int    sqbMemcpyH2D (void * dst, const void * src, size_t nbytes);
int    sqbMemcpyD2H (void * dst, const void * src, size_t nbytes);

/* ---- BGZF (bgzip) input, inflated on the device -------------------------------------- */
/* The reference reads plain text (seeq.c:201-256, getline at :361).  A BGZF file (bgzip: gzip members of at most
 * 64 KiB of text, each with its compressed size in a "BC" extra sub-field) is indexed on the host without decoding,
 * crosses the PCIe link COMPRESSED, is inflated by one warp per member (k0_inflate_bgzf) into one text buffer in
 * HBM and scanned where it lies.  Plain gzip (no "BC" sub-field) is refused: it cannot be cut without decoding.
 * CRC-32 is not checked; ISIZE, code validity and the bounds of every member are (sqbLastError names the member). */
typedef struct {
   uint64_t in_off;     /* first byte of the member's deflate stream in the buffer   */
   uint32_t in_len;     /* bytes of deflate stream                                   */
   uint32_t isize;      /* bytes of text (ISIZE)                                     */
   uint64_t out_off;    /* offset of the member's text in the inflated buffer        */
} sqb_bgzf_member_t;
/* Host only: the members of gz[0..nbytes) that carry text (the empty end-of-file member is dropped) and the size
 * of the inflated text.  members may be NULL (count only).  -1: not BGZF / truncated. */
int sqbBgzfIndex (const void * gz, size_t nbytes, sqb_bgzf_member_t * members, uint64_t cap,
                  uint64_t * count, uint64_t * text_bytes);
/* Inflates `count` members of a DEVICE-resident compressed buffer (d_gz: readable 16 bytes past its end) into
 * d_text; members is a HOST array.  Blocks until done; kernel_ms (may be NULL) = CUDA-event time of the kernel. */
int sqbBgzfInflateDevice (int device, const void * d_gz, const sqb_bgzf_member_t * members, uint64_t count,
                          void * d_text, void * stream, double * kernel_ms);
/* sqbScanHost for a BGZF buffer in host memory (pinned is faster): compressed slices of $SEEQ_B200_BGZF_SLICE_MB
 * (32) MiB leave for the device at once, the members are indexed while they travel (by several threads:
 * $SEEQ_B200_BGZF_INDEX_THREADS), the members of a slice are inflated as soon as it has arrived (k0_inflate_bgzf_pair,
 * the slices side by side on 32 streams), then the text is scanned like sqbScanDeviceLarge does.  stats->nbytes counts
 * TEXT bytes; records / line starts as after sqbScanHost.  One device (that of the engine).
 * $SEEQ_B200_BGZF_TRACE=1: host wall time of the phases on stderr. */
int sqbScanHostBgzf (sqb_engine_t * e, const void * gz, size_t nbytes, int options, sqb_stats_t * stats);
/* the inflated text of the last sqbScanHostBgzf on the engine's device (valid until the next one there) */
const void * sqbBgzfDeviceText (sqb_engine_t * e, uint64_t * nbytes);
int sqbEngineDevice (sqb_engine_t * e);

/* ---- seeq_t level batch entry (the batched analogue of seeqStringMatch) ---- */
struct seeq_t;
/* Matches every line of a host buffer against the pattern of `sq` in one GPU
 * pass.  match_opt: SQ_FIRST/SQ_BEST/SQ_ALL | SQ_FAIL/SQ_CONVERT/SQ_IGNORE
 * [| SQB_FASTA].  file_opt: SQ_ANY (0) returns the number of records and points
 * *recs at them (file order, 0-based line indices, valid until the next call
 * on `sq`); SQ_COUNTLINES (3) / SQ_COUNTMATCH (4) return the counts only.
 * Returns -1 on error with seeqerr / errno set as for seeqStringMatch. */
long seeqBatchMatch (struct seeq_t * sq, const char * text, size_t nbytes,
                     int match_opt, int file_opt, const sqb_rec_t ** recs,
                     sqb_stats_t * stats);
/* the engine behind a seeq_t (created on first use); NULL if no device */
sqb_engine_t * seeqEngine (struct seeq_t * sq);

/* ---- pattern sets: several patterns over ONE pass of the text -------------------------- */
/* Adapter / barcode sets (the extension the reference's authors name, doc/response.tex:358-360).
 * The text crosses PCIe once and is tokenized (K1) and packed into bit-planes once; every
 * pattern then runs its own matcher (K2) and finishing kernels (K3/K4) on the staged planes.
 * Results are those of npatterns independent scans: one stats entry per pattern (caller's
 * order), records of pattern p via sqbHostRecords(sqbMultiEngine(mp, p), &n), with
 * buffer-global line indices.  Any buffer size; device text needs no alignment. */
typedef struct sqb_multi sqb_multi_t;
sqb_multi_t  * sqbMultiNew    (int npatterns, const unsigned char * const * keys, const int * m,
                               const int * tau, int device);
void           sqbMultiFree   (sqb_multi_t * mp);
int            sqbMultiCount  (sqb_multi_t * mp);
sqb_engine_t * sqbMultiEngine (sqb_multi_t * mp, int pattern);
int sqbMultiScanHost   (sqb_multi_t * mp, const char * text, size_t nbytes, int options,
                        sqb_stats_t * stats /* [npatterns] */);
int sqbMultiScanDevice (sqb_multi_t * mp, const void * d_text, size_t nbytes, int options,
                        void * stream, sqb_stats_t * stats /* [npatterns] */);

/* ---- multi-GPU helper ------------------------------------------------------ */
/* Newline-aligned byte range of shard `rank` of `world` over a buffer: the
 * boundary k*nbytes/world is moved forward to just after the next '\n'. */
void sqbShardRange (const char * text, size_t nbytes, int rank, int world,
                    size_t * begin, size_t * end);

/* ---- synthetic reads (tests and benchmarks; same bytes on host and device) - */
typedef struct {
   uint64_t seed;
   uint32_t line_len;      /* bases per line (without '\n')                  */
   uint32_t plant_per_1024;/* lines that receive a mutated copy, per 1024    */
   uint32_t max_edits;     /* 0..max_edits random edits per planted copy     */
   uint32_t n_per_1024;    /* bases replaced by 'N', per 1024                */
   uint32_t junk_per_1024; /* bases replaced by a byte of "RYKMSW.-"         */
   uint32_t fastq;         /* 1: 4-line records (@id, seq, +, quality)       */
   uint32_t plant_len;     /* length of plant[]                              */
   char     plant[256];    /* sequence to plant (plain ACGT)                 */
} sqb_gen_t;

/* bytes produced for `nreads` reads starting at read index `first` */
size_t sqbGenBytes  (const sqb_gen_t * g, uint64_t first, uint64_t nreads);
int    sqbGenHost   (const sqb_gen_t * g, uint64_t first, uint64_t nreads, char * dst);
int    sqbGenDevice (const sqb_gen_t * g, uint64_t first, uint64_t nreads, void * d_dst, void * stream);

#ifdef __cplusplus
}
#endif
#endif
