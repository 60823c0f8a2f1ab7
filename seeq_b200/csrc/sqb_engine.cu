// sqb_engine.cu -- host side of the GPU matcher and its C-ABI (seeq_b200.h).
//
// One engine = one pattern on one device.  A scan runs K1 -> K2 -> (scan) ->
// K3/K4 on one stream without any host round trip in between: kernels size
// themselves from device-resident counters, the host only reads the final
// counters.  Capacities that cannot be known up front (lines per byte, events
// per byte in SQ_ALL) are guessed generously, detected exactly by the kernels
// (which keep counting past the capacity) and, if ever exceeded, the scan is
// repeated once with exact sizes (stats->reruns).
//
// On top of the single scan (slot_enqueue / slot_issue / slot_finish): graph replay of
// repeated scans, the chunk pipeline (scan_chunks) behind sqbScanHost, sqbScanDeviceLarge
// and the pattern sets (sqbMulti*), NUMA-aware pinned allocations.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "seeq_b200.h"
#include "sqb_gen.h"
#include "sqb_kernels.cuh"
#include "sqb_k2_bitslice.cuh"
#include "sqb_k12_fused.cuh"

using namespace sqb;

#define SQB_KEEP_LINES_INTERNAL SQB_KEEP_LINES

static thread_local char g_err[512] = "";
static const unsigned long long kMaxBatch = 0xfff00000ull;     // bytes per device batch (offsets are u32)

static void set_err(const char *fmt, ...)
{
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(g_err, sizeof g_err, fmt, ap);
   va_end(ap);
}

#define CU(call)                                                                       \
   do {                                                                                \
      cudaError_t e_ = (call);                                                         \
      if (e_ != cudaSuccess) {                                                         \
         set_err("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__,      \
                 __LINE__, #call);                                                     \
         return -1;                                                                    \
      }                                                                                \
   } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE attribute of a kernel: the opt-in
// is made once per (kernel, device), from whichever host thread gets there first (engines of several
// devices may live in one process, one host thread each: sqbScanHost with SEEQ_B200_DEVICES)
bool sqb_first_use(const void *fn);
static bool first_use(const void *fn) { return sqb_first_use(fn); }
bool sqb_first_use(const void *fn)
{
   static std::mutex mu;
   static std::set<std::pair<const void *, int>> seen;
   int dev = 0;
   if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
   std::lock_guard<std::mutex> lock(mu);
   return seen.insert(std::make_pair(fn, dev)).second;
}

// ---------------------------------------------------------------------------
// CUDA events of a scan (SQB_TIMING records all of them, otherwise only E_DONE)
enum { E_BEGIN = 0, E_K1C_END, E_K1_END, E_PACK_BEGIN, E_PACK_END, E_MATCH_END, E_K2_END, E_FIN_END, E_DONE, E_COUNT };

struct Slot {
   cudaStream_t stream = nullptr;
   // device
   uint8_t *d_text = nullptr;   size_t text_cap = 0;      // host path only
   uint32_t *d_ls = nullptr;    size_t line_cap = 0;      // entries (incl. sentinel)
   uint8_t *d_codes = nullptr;  size_t codes_cap = 0;     // class nibbles (K1 -> bit-sliced K2)
   uint32_t *d_ls_raw = nullptr; size_t ls_raw_cap = 0;   // line starts in tile-allocation order
   uint32_t *d_tiles = nullptr; size_t tiles_cap = 0;     // per K1 tile: count, offset, base (+ 4 arrays of the segment cuts)
   uint32_t *d_lid = nullptr;   size_t lid_cap = 0;       // segment cuts: line of every ls entry
   uint32_t *d_lbeg = nullptr;  size_t lbeg_cap = 0;      //               start of that line
   uint32_t *d_gmask = nullptr; size_t gmask_cap = 0;     //               per group of 32 entries: continuations, followed
   uint8_t *d_segflags = nullptr; size_t segflags_cap = 0; //              segstop[], deadseg[]
   uint32_t *d_act = nullptr;   size_t act_cap = 0;       // line filter: entries of ls the matcher looks at
   uint8_t *d_lflags = nullptr; size_t lflags_cap = 0;    //              1 = dead on arrival
   bool cur_filter = false;
   bool cur_bitslice = false;     // this scan produced class nibbles and bit-planes (a front others may share)
   const Slot *cur_front = nullptr; // multi-pattern: the slot whose K1 / pack output this scan reads (or nullptr)
   uint4 *d_planes = nullptr;   size_t planes_cap = 0;    // bit-planes, 32 uint4 per tile column
   uint32_t *d_bstiles = nullptr; size_t bstiles_cap = 0; // per match tile: columns, offset
   GroupDesc *d_gdesc = nullptr; size_t gdesc_cap = 0;    // fused tokenise + pack: one descriptor per group of 32 lines
   uint16_t *d_gent = nullptr;  size_t gent_cap = 0;      //   line filter: local entry index of every slot
   bool cur_fused = false;
   uint32_t *d_fintiles = nullptr; size_t fintiles_cap = 0; // per 1024-line tile: records, matched lines, first record
   unsigned long long *d_res = nullptr; size_t res_cap = 0;
   uint32_t *d_cnt = nullptr;   size_t cnt_cap = 0;
   uint32_t *d_offs = nullptr;  size_t offs_cap = 0;
   Event *d_ev = nullptr;       size_t ev_cap = 0;
   Rec *d_recs = nullptr;       size_t rec_cap = 0;
   unsigned long long *d_ctl = nullptr; size_t ctl_cap = 0;   // counters + look-back words
   // pinned host
   unsigned long long *h_ctr = nullptr;
   Rec *h_recs = nullptr;       size_t h_rec_cap = 0;
   uint32_t *h_ls = nullptr;    size_t h_ls_cap = 0;
   uint32_t *h_init = nullptr;                              // single-line ls init
   cudaEvent_t ev[E_COUNT] = {};
   // description of the scan in flight
   const uint8_t *cur_text = nullptr;
   uint32_t cur_n = 0;
   uint32_t cur_skip = 0;       // the buffer proper starts cur_skip (< 16) bytes into cur_text (K1Args::skip)
   int cur_options = 0;
   cudaStream_t cur_stream = nullptr;
   uint32_t launches = 0;
   bool busy = false;
   // host-path bookkeeping
   size_t chunk_off = 0;
   // the scan as a CUDA graph: a scan that repeats the one before it (same text, size, options,
   // stream and engine state) is captured once and replayed with one launch from then on
   struct GraphKey {
      const uint8_t *text = nullptr;
      uint32_t n = 0, skip = 0;
      int options = -1;
      cudaStream_t st = nullptr;
      unsigned long long version = 0;
      bool operator==(const GraphKey &o) const
      { return text == o.text && n == o.n && skip == o.skip && options == o.options && st == o.st && version == o.version; }
   } gkey;
   cudaGraphExec_t gexec = nullptr;
   uint32_t glaunches = 0, grepeats = 0;
   bool gbroken = false;          // a capture failed once: this slot stays eager
};

struct sqb_engine {
   int device = 0;
   int sms = 148;
   int m = 0, tau = 0;
   int words = 1;                 // automaton words: 1, 2, 4, 8, 16 or 32
   unsigned char keys[kMaxWords * 32];
   bool bs_ok = false;            // the pattern fits the bit-sliced matcher
   int cuts = 1;                  // long lines are cut into segments: 0 never, 1 once a scan met lines longer
                                  // than bs_max_line (then from that scan on), 2 always  (SEEQ_B200_CUTS)
   bool cuts_wanted = false;      // a scan met long lines
   int filter = 1;                // line filter (lines with a STOP in their first m - tau bytes are not packed):
                                  // 0 never, 1 if the first filtered scan drops >= 25 % of the lines, 2 always
   bool nfa_levels = true;        // tau <= 2: NFA-level automaton instead of Myers' (SEEQ_B200_NFA=0 disables)
   int filter_state = -1;         // filter == 1: -1 undecided (probe with the next scan), 0 off, 1 on
   int filter_k = 1;              // a STOP among the first filter_k bytes kills a line: min(8, m - tau); the leader
                                  // of a pattern set uses the smallest of the set
   BsGate bs_gate{65536u, 4096u};
   uint32_t bs_min_bytes = 1u << 20;
   double cols_per_byte = 1.3 / 1024.0;   // tile columns per text byte (plane buffer guess)
   // fused tokenise + pack (sqb_k12_fused.cuh): 1 = use it where it applies, 0 = never (SEEQ_B200_FUSED=0, a
   // pattern set, or a scan that met something the fused kernel does not handle: from then on the
   // two-kernel path serves this engine)
   int fused = 1;
   uint32_t fused_ov = 512;               // overlap staged behind a tile; 4096 after a scan met longer lines
   double units_per_byte = 0.028;         // plane units (16 B) per text byte: 3 bits per byte + padding to 32 columns
   BsPattern bs_pat;
   Slot slot[2];
   // lines / events per byte seen so far (capacity guesses)
   double lines_per_byte = 1.0 / 24.0;
   double events_per_byte = 1.0 / 64.0;
   // results of the last sqbScanHost
   sqb_rec_t *host_recs = nullptr;     // pinned; the records of the last sqbScanHost, all chunks, in order
   size_t host_recs_cap = 0, host_recs_n = 0;
   std::vector<uint64_t> host_lines;
   int last_slot = 0;
   sqb_stats_t last_stats;
   Rec *d_all_recs = nullptr;          // SQB_DEVICE_RESULTS: the records of all chunks of the last chunked scan
   size_t d_all_cap = 0, d_all_n = 0;
   unsigned long long *d_word = nullptr, *h_word = nullptr;     // device_cuts: one word each
   cudaStream_t big_stream = nullptr;                           // sqbScanDeviceLarge: the kernels of all chunks
   bool graphs = true;                                          // SEEQ_B200_GRAPHS=0 disables graph replay
   // sqbScanHost over several GPUs from ONE process ($SEEQ_B200_DEVICES): engines of the same pattern on the
   // other devices, created on first use and owned by this engine; one host thread drives each
   std::vector<sqb_engine *> peers;
   unsigned long long scan_generation = 0;                      // bumped by every scan that rewrites the host results
   unsigned long long version = 1;                              // bumped whenever a capacity guess or a mode changes
};

// ---------------------------------------------------------------------------
// pattern tables
// ---------------------------------------------------------------------------
// Forward tables follow seeqcore.h:89-111 + libseeq.c:255-270; in the reverse
// pass every non-base byte is skipped (libseeq.c:297-311).
static void build_pattern(const sqb_engine *e, int options, bool reverse, Pattern *p)
{
   memset(p, 0, sizeof *p);
   p->m = e->m;
   p->tau = e->tau;
   const int W = e->words;
   const int pad = W * 32 - e->m;
   for (int c = 0; c < 5; c++) {
      for (int b = 0; b < pad; b++) p->eq[c][b >> 5] |= 1u << (b & 31);      // wildcard rows
      for (int j = 0; j < e->m; j++) {
         const unsigned char k = reverse ? e->keys[e->m - 1 - j] : e->keys[j];
         if (k & (1u << c)) p->eq[c][(pad + j) >> 5] |= 1u << ((pad + j) & 31);
      }
   }
   const int nondna = options & OPT_NONDNA;
   for (int b = 0; b < 256; b++) {
      const int code = b < 128 ? base_code(b) : -1;
      uint8_t cls;
      if (code >= 0) cls = kKindBase | (uint8_t)code;
      else if (reverse) cls = (nondna == OPT_CONVERT && b != 0 && b != '\n') ? (kKindBase | 4) : kKindSkip;
      else if (b == 0) cls = kKindStop;
      else if (b == '\n') cls = (options & OPT_STREAM) ? kKindSkip : kKindStop;
      else if (nondna == OPT_CONVERT) cls = kKindBase | 4;
      else if (nondna == OPT_IGNORE) cls = kKindSkip;
      else cls = kKindStop;
      p->cls[b] = cls;
   }
}

// ---------------------------------------------------------------------------
// NUMA placement of pinned host memory
// ---------------------------------------------------------------------------
// Pinned buffers (the text a caller stages with sqbHostAlloc, the record arrays) are the
// two ends of every PCIe transfer.  On a two-socket host with 8 GPUs a buffer that lives
// on the other socket makes every DMA cross the inter-socket link, and with one process
// per GPU all of them do at once.  For the duration of an allocation the calling thread
// is therefore moved onto the CPUs of the GPU's NUMA node (sysfs: numa_node of the PCI
// device, cpulist of the node) and its memory policy set to prefer that node; the affinity
// mask and the caller's policy (get_mempolicy) are put back afterwards.  No-op where sysfs has no answer (VMs), or with SEEQ_B200_NUMA=0.
struct NumaScope {
   cpu_set_t old_set;
   bool moved = false, policy = false;
   int old_mode = 0;                       // the caller's memory policy (numactl --membind / --interleave ...)
   unsigned long old_mask[16] = {};        // 1024 nodes

   static int node_of(int device)
   {
      static int cache[64];
      static bool known[64];
      static std::mutex mu;
      if (device < 0 || device >= 64) return -1;
      std::lock_guard<std::mutex> lock(mu);
      if (known[device]) return cache[device];
      int node = -1;
      const char *env = getenv("SEEQ_B200_NUMA");
      char bus[64] = "";
      if (!(env && atoi(env) == 0) && cudaDeviceGetPCIBusId(bus, sizeof bus, device) == cudaSuccess) {
         for (char *c = bus; *c; c++) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
         char path[160];
         snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
         if (FILE *f = fopen(path, "r")) {
            if (fscanf(f, "%d", &node) != 1) node = -1;
            fclose(f);
         }
      } else {
         cudaGetLastError();
      }
      cache[device] = node;
      known[device] = true;
      return node;
   }

   explicit NumaScope(int device)
   {
      const int node = node_of(device);
      if (node < 0) return;
      char path[96];
      snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
      FILE *f = fopen(path, "r");
      if (f == nullptr) return;
      cpu_set_t want, cur;
      CPU_ZERO(&want);
      char line[4096] = "";
      const bool got = fgets(line, sizeof line, f) != nullptr;
      fclose(f);
      if (!got) return;
      unsigned char listed[CPU_SETSIZE];
      if (parse_cpulist(line, listed, CPU_SETSIZE) == 0) return;          // "0-31,64-95"
      for (int c = 0; c < CPU_SETSIZE; c++) if (listed[c]) CPU_SET(c, &want);
      if (sched_getaffinity(0, sizeof old_set, &old_set) != 0) return;
      CPU_AND(&cur, &want, &old_set);
      if (CPU_COUNT(&cur) > 0 && !CPU_EQUAL(&cur, &old_set)) moved = sched_setaffinity(0, sizeof cur, &cur) == 0;
      if (node < 64) {
         unsigned long mask = 1ul << node;
         if (syscall(SYS_get_mempolicy, &old_mode, old_mask, sizeof old_mask * 8ul, nullptr, 0ul) != 0) return;
         policy = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, 65ul) == 0;
      }
   }
   ~NumaScope()
   {
      if (policy) {                        // exactly what the caller had (MPOL_DEFAULT takes no node mask)
         if (old_mode == 0) syscall(SYS_set_mempolicy, 0, nullptr, 0ul);
         else syscall(SYS_set_mempolicy, old_mode, old_mask, sizeof old_mask * 8ul);
      }
      if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
   }
};

static cudaError_t pinned_alloc(void **p, size_t bytes)
{
   int device = 0;
   if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; }
   NumaScope scope(device);
   // portable: every device of the process may DMA out of / into it (multi-GPU sqbScanHost)
   return cudaHostAlloc(p, bytes, cudaHostAllocPortable);
}

// ---------------------------------------------------------------------------
// memory helpers
// ---------------------------------------------------------------------------
template <class T> static int dev_reserve(T **p, size_t *cap, size_t need, size_t slack_div = 8)
{
   if (need <= *cap) return 0;
   if (*p) CU(cudaFree(*p));
   *p = nullptr;
   *cap = 0;
   const size_t want = need + need / slack_div + 256;
   CU(cudaMalloc((void **)p, want * sizeof(T)));
   *cap = want;
   return 0;
}

template <class T> static int pin_reserve(T **p, size_t *cap, size_t need)
{
   if (need <= *cap) return 0;
   if (*p) CU(cudaFreeHost(*p));
   *p = nullptr;
   *cap = 0;
   const size_t want = need + need / 4 + 1024;
   CU(pinned_alloc((void **)p, want * sizeof(T)));
   *cap = want;
   return 0;
}

static void slot_free(Slot &s)
{
   cudaFree(s.d_text); cudaFree(s.d_ls); cudaFree(s.d_codes); cudaFree(s.d_ls_raw); cudaFree(s.d_tiles); cudaFree(s.d_lid); cudaFree(s.d_lbeg); cudaFree(s.d_gmask); cudaFree(s.d_segflags); cudaFree(s.d_act); cudaFree(s.d_lflags); cudaFree(s.d_planes); cudaFree(s.d_bstiles); cudaFree(s.d_gdesc); cudaFree(s.d_gent); cudaFree(s.d_fintiles); cudaFree(s.d_res); cudaFree(s.d_cnt); cudaFree(s.d_offs);
   cudaFree(s.d_ev); cudaFree(s.d_recs); cudaFree(s.d_ctl);
   cudaFreeHost(s.h_ctr); cudaFreeHost(s.h_recs); cudaFreeHost(s.h_ls); cudaFreeHost(s.h_init);
   for (auto &ev : s.ev) if (ev) cudaEventDestroy(ev);
   if (s.gexec) cudaGraphExecDestroy(s.gexec);
   if (s.stream) cudaStreamDestroy(s.stream);
   s = Slot();
}

static int slot_init(Slot &s)
{
   if (s.stream) return 0;
   CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
   CU(pinned_alloc((void **)&s.h_ctr, C_COUNT * sizeof(unsigned long long)));
   CU(pinned_alloc((void **)&s.h_init, 4 * sizeof(uint32_t)));
   for (auto &ev : s.ev) CU(cudaEventCreate(&ev));
   return 0;
}

// ---------------------------------------------------------------------------
// kernel dispatch
// ---------------------------------------------------------------------------
static int mode_of(int options)
{
   const int match = options & OPT_MATCH;
   const bool count = options & SQB_COUNT_ONLY;
   if (match == OPT_ALL) return count ? M_COUNTALL : M_ALL;
   if (match == OPT_BEST) return count ? M_COUNT : M_BEST;
   return count ? M_COUNT : M_FIRST;         // SQ_FIRST and SQ_COUNT (libseeq.c:219-221)
}

template <int W> static int launch_k2_thread(int mode, int grid, cudaStream_t st, const K2Args &a, const Pattern &p)
{
#define SQB_CASE(M)                                                                              \
   case M: {                                                                                     \
      if (first_use((const void *)k2_forward_thread<W, M>))                                      \
         CU(cudaFuncSetAttribute(k2_forward_thread<W, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)kK2Stage));                                                \
      k2_forward_thread<W, M><<<grid, kThreads, kK2Stage, st>>>(a, p);                            \
      break;                                                                                     \
   }
   switch (mode) {
      SQB_CASE(M_COUNT)
      SQB_CASE(M_FIRST)
      SQB_CASE(M_BEST)
      SQB_CASE(M_ALL)
      SQB_CASE(M_COUNTALL)
   }
#undef SQB_CASE
   CU(cudaGetLastError());
   return 0;
}

template <int G> static int launch_k2_lanes(int mode, int grid, cudaStream_t st, const K2Args &a, const Pattern &p)
{
   switch (mode) {
   case M_COUNT: k2_forward_lanes<G, M_COUNT><<<grid, kThreads, 0, st>>>(a, p); break;
   case M_FIRST: k2_forward_lanes<G, M_FIRST><<<grid, kThreads, 0, st>>>(a, p); break;
   case M_BEST: k2_forward_lanes<G, M_BEST><<<grid, kThreads, 0, st>>>(a, p); break;
   case M_ALL: k2_forward_lanes<G, M_ALL><<<grid, kThreads, 0, st>>>(a, p); break;
   case M_COUNTALL: k2_forward_lanes<G, M_COUNTALL><<<grid, kThreads, 0, st>>>(a, p); break;
   }
   CU(cudaGetLastError());
   return 0;
}

static int launch_k2(const sqb_engine *e, int mode, int grid, cudaStream_t st, const K2Args &a, const Pattern &p)
{
   switch (e->words) {
   case 1: return launch_k2_thread<1>(mode, grid, st, a, p);
   case 2: return launch_k2_thread<2>(mode, grid, st, a, p);
   case 4: return launch_k2_lanes<4>(mode, grid, st, a, p);
   case 8: return launch_k2_lanes<8>(mode, grid, st, a, p);
   case 16: return launch_k2_lanes<16>(mode, grid, st, a, p);
   default: return launch_k2_lanes<32>(mode, grid, st, a, p);
   }
}

static int launch_finish(const sqb_engine *e, bool all, int grid, cudaStream_t st, const FinArgs &a, const Pattern &rp)
{
   // bytes fetched per round of the reverse pass (tuning knob)
   static const int chunk = getenv("SEEQ_B200_REV_CHUNK") ? atoi(getenv("SEEQ_B200_REV_CHUNK")) : 4;
#define SQB_FIN(W)                                                                  \
   case W:                                                                          \
      if (all) k34_finish_events<W><<<grid, kThreads, 0, st>>>(a, rp);              \
      else if (chunk <= 1) k34_finish_lines<W, 1><<<grid, kThreads, 0, st>>>(a, rp); \
      else if (chunk <= 4) k34_finish_lines<W, 4><<<grid, kThreads, 0, st>>>(a, rp); \
      else k34_finish_lines<W, 8><<<grid, kThreads, 0, st>>>(a, rp);                \
      break;
   switch (e->words) {
      SQB_FIN(1) SQB_FIN(2) SQB_FIN(4) SQB_FIN(8) SQB_FIN(16)
   default:
      if (all) k34_finish_events<32><<<grid, kThreads, 0, st>>>(a, rp);
      else k34_finish_lines<32, 4><<<grid, kThreads, 0, st>>>(a, rp);
   }
#undef SQB_FIN
   CU(cudaGetLastError());
   return 0;
}

// ---------------------------------------------------------------------------
// one scan: issue (asynchronous) and finish (synchronise, verify capacities)
// ---------------------------------------------------------------------------
static size_t div_up(size_t a, size_t b) { return (a + b - 1) / b; }

// the line-bit-sliced matcher covers short patterns over many lines; everything
// else (one string, long patterns) runs the thread-per-line / lane-blocked kernels
static bool use_bitslice(const sqb_engine *e, int options, uint32_t n)
{
   return e->bs_ok && !(options & (SQB_SINGLE_LINE | OPT_STREAM)) && n >= e->bs_min_bytes;
}

// long lines are cut into segments (sqb_tables.h) when the scan is bit-sliced, returns
// records and runs in a mode in which every byte in front of a STOP feeds the
// automaton.  Not with SQ_IGNORE (a warm-up could hold too few automaton inputs),
// not in FASTA mode (a cut inside a long header line would be scanned), not when
// the caller wants the line starts back, not in the count-only modes.
static bool cuts_allowed(const sqb_engine *e, int options, uint32_t n)
{
   if (!use_bitslice(e, options, n) || e->cuts == 0) return false;
   // (SQB_KEEP_LINES -- every seeqFileMatch call -- takes cuts too: the line starts handed back are the entries of
   // ls that open a line, see scan_one_chunk)
   // ... and so do the count-only scans: slot_enqueue runs them as SQ_FIRST / SQ_ALL scans without the finish kernels
   // (matched lines and events are counted per LINE behind k_seg_reduce, not per segment in the matcher)
   if (options & (SQB_FASTA | SQB_FASTQ)) return false;
   if ((options & OPT_NONDNA) == OPT_IGNORE) return false;
   return bs_warmup(e->m, e->tau) <= kCutWindow;
}
static bool use_filter(const sqb_engine *e, int options, uint32_t n)
{
   if (!use_bitslice(e, options, n) || e->filter == 0 || n >= 0x80000000u) return false;
   return e->filter == 2 || e->filter_state != 0;
}
static bool use_cuts(const sqb_engine *e, int options, uint32_t n)
{
   return cuts_allowed(e, options, n) && (e->cuts == 2 || e->cuts_wanted);
}
// the fused tokenise + pack kernel serves the bit-sliced scans of single-part automata over plain lines:
// no FASTA header rule, no record structure, no segment cuts (an engine that has met long lines cuts
// them on the two-kernel path), not inside a pattern set
static bool use_fused(const sqb_engine *e, int options, uint32_t n)
{
   if (!e->fused || !use_bitslice(e, options, n) || e->bs_pat.parts != 1) return false;
   if (options & (SQB_FASTA | SQB_FASTQ)) return false;
   // with the line filter the two-kernel path is the faster one (cfg5, r4r: 1145 against 1092 GB/s): K12's groups are local
   // to a tile, and the 104 lines of a 32 KiB tile that survive the filter fill the matcher's lanes to 81 %
   // (SEEQ_B200_FUSED=2: fused all the same -- the tests of that path)
   // (a scan that only PROBES the filter -- the first of an engine -- stays fused: the fused kernel counts the live lines too)
   if (e->fused < 2 && use_filter(e, options, n) && (e->filter == 2 || e->filter_state == 1)) return false;
   return !use_cuts(e, options, n);
}

template <int R, int G, int MODE> static int launch_bs2(bool skip, int grid, cudaStream_t st, const K2BsArgs &a, const BsPattern &p)
{
   const size_t smem = (MODE == BS_ALL ? sizeof(BsWarpSmemAll) : sizeof(BsWarpSmem)) * kBsWarps;
   if (first_use((const void *)k2_bitslice<R, G, MODE, true>)) {
      CU(cudaFuncSetAttribute(k2_bitslice<R, G, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CU(cudaFuncSetAttribute(k2_bitslice<R, G, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   }
   if (skip) k2_bitslice<R, G, MODE, true><<<grid, kBsThreads, smem, st>>>(a, p);
   else k2_bitslice<R, G, MODE, false><<<grid, kBsThreads, smem, st>>>(a, p);
   CU(cudaGetLastError());
   return 0;
}

template <int R, int G> static int launch_bs1(int bsmode, bool skip, int grid, cudaStream_t st, const K2BsArgs &a, const BsPattern &p)
{
   switch (bsmode) {
   case BS_FIRST: return launch_bs2<R, G, BS_FIRST>(skip, grid, st, a, p);
   case BS_BEST: return launch_bs2<R, G, BS_BEST>(skip, grid, st, a, p);
   default: return launch_bs2<R, G, BS_ALL>(skip, grid, st, a, p);
   }
}

// sqb_engine_wm.cu (compiled twice: planes of k15_pack / group planes of the fused kernel), sqb_engine_bsf.cu
cudaError_t sqb_launch_bitslice_wm(int rows, int levels, int bsmode, bool skip, int grid, cudaStream_t st,
                                   const K2BsArgs &a, const BsPattern &p);
cudaError_t sqb_launch_bitslice_wm_fused(int rows, int levels, int bsmode, bool skip, int grid, cudaStream_t st,
                                         const K2BsArgs &a, const BsPattern &p);
cudaError_t sqb_launch_bitslice_myers_fused(int rows, int bsmode, bool skip, int grid, cudaStream_t st,
                                            const K2BsArgs &a, const BsPattern &p);

static int launch_bitslice(const sqb_engine *e, int mode, int options, size_t max_lines, cudaStream_t st, const K2BsArgs &a)
{
   const int bsmode = (mode == M_ALL || mode == M_COUNTALL) ? BS_ALL : (mode == M_BEST ? BS_BEST : BS_FIRST);
   const bool skip = (options & OPT_NONDNA) == OPT_IGNORE;
   const int R = e->bs_pat.rows, G = e->bs_pat.parts;
   const bool nfa = G == 1 && e->tau <= 2 && e->nfa_levels;
   int per_sm = G > 1 ? (R <= 24 ? SQB_G2_CTAS : 2) : (R <= 16 ? 4 : 3);
   if (nfa) per_sm = R * (e->tau + 1) <= 24 ? 6 : (R * (e->tau + 1) <= SQB_WM_4CTA_ROWS ? 4 : 3);      // = the kernels' launch bounds
   if (const char *c = getenv("SEEQ_B200_BS_CTAS")) per_sm = std::max(1, atoi(c));
   // work items = (tile, 1/G of its groups), one warp each
   const int grid = (int)std::max<size_t>(1, std::min<size_t>(div_up(div_up(max_lines, kBsTileLines) * G, kBsWarps),
                                                               (size_t)e->sms * per_sm * 2));
   const bool fused = a.gdesc != nullptr;                 // (single-part automata only: use_fused)
   if (nfa) {                                             // small tau: the NFA-level automaton is cheaper
      if (fused) CU(sqb_launch_bitslice_wm_fused(R, e->tau + 1, bsmode, skip, grid, st, a, e->bs_pat));
      else CU(sqb_launch_bitslice_wm(R, e->tau + 1, bsmode, skip, grid, st, a, e->bs_pat));
      return 0;
   }
   if (fused) {
      CU(sqb_launch_bitslice_myers_fused(R, bsmode, skip, grid, st, a, e->bs_pat));
      return 0;
   }
#define SQB_SHAPE(RR, GG) if (R == RR && G == GG) return launch_bs1<RR, GG>(bsmode, skip, grid, st, a, e->bs_pat);
   SQB_SHAPE(8, 1) SQB_SHAPE(10, 1) SQB_SHAPE(12, 1) SQB_SHAPE(16, 1) SQB_SHAPE(20, 1) SQB_SHAPE(24, 1) SQB_SHAPE(32, 1)
   SQB_SHAPE(20, 2) SQB_SHAPE(24, 2) SQB_SHAPE(32, 2)
   SQB_SHAPE(20, 4) SQB_SHAPE(24, 4) SQB_SHAPE(26, 4) SQB_SHAPE(28, 4) SQB_SHAPE(32, 4)
#undef SQB_SHAPE
   set_err("no bit-sliced kernel for %d rows x %d parts", R, G);
   return -1;
}

// cudaEventRecord that also works while `st` is being captured into a graph: there a plain record
// is only a dependency edge of the capture; the EXTERNAL flavour becomes an event-record node
// that really stamps the event every time the graph runs (SQB_TIMING inside a replayed scan)
static cudaError_t record_event(cudaEvent_t ev, cudaStream_t st)
{
   cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
   if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive)
      return cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal);
   return cudaEventRecord(ev, st);
}

// front != nullptr: the slot of ANOTHER engine that has scanned (or is scanning, earlier on the same
// stream) the same text with the same options; its line starts, line filter and bit-planes are read
// instead of being computed again (several patterns over one pass of the text).
static int slot_enqueue(sqb_engine *e, Slot &s, const uint8_t *d_text, uint32_t n, const int options_asked, cudaStream_t st,
                        uint32_t skip, const Slot *front)
{
   // A count-only scan of long lines is cut into segments like any other; the matcher would then count segments, so the
   // scan runs as a plain SQ_FIRST / SQ_ALL scan up to the per-tile sums (which leave the counts of matched LINES and of
   // events in the counters) and stops in front of K3 / K4.
   int options = options_asked;
   const bool count_cut = (options & SQB_COUNT_ONLY) && !(options & SQB_SINGLE_LINE) && front == nullptr &&
                          use_cuts(e, options, n);
   if (count_cut) options &= ~SQB_COUNT_ONLY;
   const int mode = mode_of(options);
   const bool single = options & SQB_SINGLE_LINE;
   const bool timing = options & SQB_TIMING;
   s.cur_text = d_text;
   s.cur_n = n;
   s.cur_skip = skip;
   s.cur_options = options_asked;
   s.cur_stream = st;
   s.cur_front = front;
   s.launches = 0;
   s.busy = true;
   const Slot &f = front ? *front : s;                      // owner of ls, act, lflags, planes, bstiles

   // ---- capacities ----------------------------------------------------------
   if (!front) {
      size_t want_lines = single ? 2 : (size_t)((double)n * e->lines_per_byte) + 1024;
      if (want_lines > (size_t)n + 2) want_lines = (size_t)n + 2;
      if (dev_reserve(&s.d_ls, &s.line_cap, want_lines)) return -1;
      if (!single && dev_reserve(&s.d_ls_raw, &s.ls_raw_cap, s.line_cap, 64)) return -1;
   }
   const size_t lines_cap = f.line_cap - 1;                 // one entry is the sentinel
   if (mode == M_FIRST || mode == M_BEST) {
      if (dev_reserve(&s.d_res, &s.res_cap, lines_cap, 64)) return -1;
      if (dev_reserve(&s.d_recs, &s.rec_cap, lines_cap, 64)) return -1;
   }
   if (mode == M_ALL) {
      if (dev_reserve(&s.d_cnt, &s.cnt_cap, lines_cap, 64)) return -1;
      if (dev_reserve(&s.d_offs, &s.offs_cap, lines_cap, 64)) return -1;
      size_t want_ev = (size_t)((double)n * e->events_per_byte) + 4096;
      if (dev_reserve(&s.d_ev, &s.ev_cap, want_ev)) return -1;
      if (dev_reserve(&s.d_recs, &s.rec_cap, s.ev_cap, 64)) return -1;
   }
   // control block: the counters; per-tile arrays of K1 and of the compaction
   const size_t k1_tiles = div_up(n, kK1Tile) + 1;
   const size_t fin_tiles = div_up(lines_cap, kFinTile) + 1;
   const size_t ctl_words = C_COUNT;
   if (dev_reserve(&s.d_ctl, &s.ctl_cap, ctl_words)) return -1;
   if (dev_reserve(&s.d_fintiles, &s.fintiles_cap, 3 * fin_tiles)) return -1;
   unsigned long long *ctr = s.d_ctl;

   if (timing) CU(record_event(s.ev[E_BEGIN], st));
   CU(cudaMemsetAsync(s.d_ctl, 0, ctl_words * sizeof(unsigned long long), st));
   const bool cut = !single && !front && use_cuts(e, options, n);
   // SQB_FASTQ: the line filter by record structure -- the matcher is handed the sequence lines only
   // (entry index 1 mod 4; k1_scan_tiles / k1_gather), K1 itself computes no dead-on-arrival flags
   const bool fastq = !single && (options & SQB_FASTQ);
   const bool filter = front ? front->cur_filter
                             : (!single && (fastq ? use_bitslice(e, options, n) : use_filter(e, options, n)));
   const bool k1_filter = filter && !fastq;
   s.cur_filter = filter;
   const bool fused = !single && !front && use_fused(e, options, n);
   s.cur_fused = fused;

   // ---- K1 ------------------------------------------------------------------
   if (front) {
      // the front's counters: lines, entries, the matcher decision, live entries
      CU(cudaMemcpyAsync(ctr + C_NLINES, front->d_ctl + C_NLINES, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
      CU(cudaMemcpyAsync(ctr + C_BS_SELECTED, front->d_ctl + C_BS_SELECTED,
                         (C_NACTIVE + 1 - C_BS_SELECTED) * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
   } else if (single) {
      s.h_init[0] = 0;
      s.h_init[1] = n;
      CU(cudaMemcpyAsync(s.d_ls, s.h_init, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
      s.h_ctr[C_NLINES] = 1;      // reuse pinned word as the source of the line count
      CU(cudaMemcpyAsync(ctr + C_NLINES, s.h_ctr + C_NLINES, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(ctr + C_NPSEUDO, s.h_ctr + C_NLINES, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
   } else {
      const bool want_codes = use_bitslice(e, options, n);
      if (want_codes && !fused) {
         const size_t need = k1_tiles * (kK1Tile / 2) + 256;
         if (dev_reserve(&s.d_codes, &s.codes_cap, need)) return -1;
      }
      if (dev_reserve(&s.d_tiles, &s.tiles_cap, 9 * k1_tiles)) return -1;
      uint32_t *tile_cnt = s.d_tiles, *tile_off = tile_cnt + k1_tiles, *tile_base = tile_off + k1_tiles;
      uint32_t *tile_real = tile_base + k1_tiles, *tile_rbase = tile_real + k1_tiles;
      uint32_t *tile_last = tile_rbase + k1_tiles, *tile_lbeg = tile_last + k1_tiles;
      uint32_t *tile_alive = tile_lbeg + k1_tiles, *tile_abase = tile_alive + k1_tiles;
      if (filter && !fused) {
         if (dev_reserve(&s.d_act, &s.act_cap, s.line_cap, 64)) return -1;
         if (dev_reserve(&s.d_lflags, &s.lflags_cap, s.line_cap, 64)) return -1;
      }
      if (cut) {
         if (dev_reserve(&s.d_lid, &s.lid_cap, s.line_cap, 64)) return -1;
         if (dev_reserve(&s.d_lbeg, &s.lbeg_cap, s.line_cap, 64)) return -1;
         if (dev_reserve(&s.d_gmask, &s.gmask_cap, 2 * (div_up(s.line_cap, 64) * 2 + 64), 64)) return -1;
         if (dev_reserve(&s.d_segflags, &s.segflags_cap, 2 * s.line_cap, 64)) return -1;
         CU(cudaMemsetAsync(s.d_segflags, 0, 2 * s.line_cap, st));
      }
      const uint32_t ntiles = (uint32_t)div_up(n, kK1Tile);
      ClassTable ct;
      build_class_table(options, &ct);
      const int grid = (int)std::min<size_t>(div_up(n, kK1Tile), (size_t)e->sms * (fused ? SQB_K12_CTAS : 3));
      if (fused) {
         // groups: 32 lines each, plus one partial group per K1 tile at most
         if (dev_reserve(&s.d_gdesc, &s.gdesc_cap, s.line_cap / 32 + k1_tiles + 64, 64)) return -1;
         if (filter && dev_reserve(&s.d_gent, &s.gent_cap, s.gdesc_cap * 32, 64)) return -1;
         const size_t want_units = (size_t)((double)n * e->units_per_byte) + 4096;
         if (dev_reserve(&s.d_planes, &s.planes_cap, want_units, 16)) return -1;
         K12Args ka{d_text, n, s.d_ls_raw, (uint32_t)s.line_cap, ctr, tile_cnt, tile_off, tile_alive,
                    (uint32_t)e->filter_k, skip, e->fused_ov, s.d_gdesc, (uint32_t)std::min<size_t>(s.gdesc_cap, 0xffffffffu),
                    s.d_gent, s.d_planes, (uint32_t)std::min<size_t>(s.planes_cap, 0xffffffffu), 4u};
         ClassTable32 ct32;
         build_class_table32(ct, &ct32);
         if (first_use((const void *)k12_scan_pack<true>)) {
            CU(cudaFuncSetAttribute(k12_scan_pack<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k12_smem_bytes(kFMaxOverlap, true)));
            CU(cudaFuncSetAttribute(k12_scan_pack<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k12_smem_bytes(kFMaxOverlap, false)));
         }
         const size_t smem = k12_smem_bytes(e->fused_ov, filter);
         if (filter) k12_scan_pack<true><<<grid, kFThreads, smem, st>>>(ka, ct32);
         else k12_scan_pack<false><<<grid, kFThreads, smem, st>>>(ka, ct32);
      } else {
      K1Args k1{d_text, n, s.d_ls_raw, (uint32_t)s.line_cap, want_codes ? (uint2 *)s.d_codes : nullptr, ctr,
                tile_cnt, tile_off, tile_real, tile_last, tile_alive,
                (uint32_t)e->filter_k, (options & SQB_FASTA) ? 1 : 0, skip};
      if (first_use((const void *)k1_scan_classify<true, true, true>)) {
         CU(cudaFuncSetAttribute(k1_scan_classify<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kK1Smem));
         CU(cudaFuncSetAttribute(k1_scan_classify<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kK1Smem));
         CU(cudaFuncSetAttribute(k1_scan_classify<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kK1Smem));
         CU(cudaFuncSetAttribute(k1_scan_classify<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kK1Smem));
         CU(cudaFuncSetAttribute(k1_scan_classify<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kK1Smem));
      }
      if (cut && k1_filter) k1_scan_classify<true, true, true><<<grid, kThreads, kK1Smem, st>>>(k1, ct);
      else if (cut) k1_scan_classify<true, true, false><<<grid, kThreads, kK1Smem, st>>>(k1, ct);
      else if (k1_filter) k1_scan_classify<true, false, true><<<grid, kThreads, kK1Smem, st>>>(k1, ct);
      else if (want_codes) k1_scan_classify<true, false, false><<<grid, kThreads, kK1Smem, st>>>(k1, ct);
      else k1_scan_classify<false, false, false><<<grid, kThreads, kK1Smem, st>>>(k1, ct);
      }
      CU(cudaGetLastError());
      if (timing) CU(record_event(s.ev[E_K1C_END], st));
      K1ScanArgs ks{tile_cnt, tile_base, ntiles, ctr, cut ? tile_real : nullptr, tile_rbase, tile_last, tile_lbeg,
                    filter ? tile_alive : nullptr, tile_abase, fastq ? 1 : 0,
                    fused ? 1 : 0, (uint32_t)std::min<size_t>(s.planes_cap, 0xffffffffu),
                    (uint32_t)std::min<size_t>(s.gdesc_cap, 0xffffffffu), e->bs_gate};
      k1_scan_tiles<<<1, 1024, 0, st>>>(ks);
      K1GatherArgs kg{s.d_ls_raw, s.d_ls, (uint32_t)s.line_cap, tile_cnt, tile_off, tile_base, ntiles, n, ctr,
                      cut ? s.d_codes : nullptr, tile_rbase, tile_lbeg, s.d_lid, s.d_lbeg,
                      filter ? tile_abase : nullptr, fused ? nullptr : s.d_act, s.d_lflags,
                      (want_codes && (mode == M_FIRST || mode == M_BEST)) ? s.d_res : nullptr, fastq ? 1 : 0};
      k1_gather<<<(int)std::min<size_t>(div_up(ntiles, kWarps), (size_t)e->sms * 8), kThreads, 0, st>>>(kg);
      CU(cudaGetLastError());
      s.launches += 3;
   }
   if (timing && (single || front)) CU(record_event(s.ev[E_K1C_END], st));
   if (timing) CU(record_event(s.ev[E_K1_END], st));

   // ---- K2 ------------------------------------------------------------------
   Pattern fwd, rev;
   build_pattern(e, options, false, &fwd);
   build_pattern(e, options, true, &rev);
   const size_t max_lines = single ? 1 : std::min<size_t>(lines_cap, n);
   const int lines_per_cta = e->words <= 2 ? kThreads : kThreads / e->words;
   const bool bitslice = !single && use_bitslice(e, options, n) && (!front || front->cur_bitslice);
   s.cur_bitslice = bitslice && !front;
   K2Args k2{d_text, n, f.d_ls, (uint32_t)lines_cap, ctr, s.d_res, s.d_cnt, s.d_ev,
             (uint32_t)std::min<size_t>(s.ev_cap, 0xffffffffu), bitslice ? 1 : 0, fastq ? 1 : 0};
   if (timing && !bitslice) {
      CU(record_event(s.ev[E_PACK_BEGIN], st));
      CU(record_event(s.ev[E_PACK_END], st));
   }
   if (bitslice) {
      // the bit-sliced kernel stores only the lines that match: k1_gather has preset the results of
      // its own scan, a scan that rides on another engine's front presets its own
      if ((mode == M_FIRST || mode == M_BEST) && front)
         CU(cudaMemsetAsync(s.d_res, 0xFF, std::min<size_t>(lines_cap, (size_t)n + 1) * sizeof(unsigned long long), st));
      if (mode == M_ALL && filter)      // the lines the filter drops are never written
         CU(cudaMemsetAsync(s.d_cnt, 0, std::min<size_t>(lines_cap, (size_t)n + 1) * sizeof(uint32_t), st));
      const size_t max_tiles = div_up(lines_cap, kBsTileLines) + 1;
      if (!front && !fused) {
         if (dev_reserve(&s.d_bstiles, &s.bstiles_cap, 2 * max_tiles)) return -1;
         const size_t want_cols = (size_t)((double)n * e->cols_per_byte) + 4096;
         if (dev_reserve(&s.d_planes, &s.planes_cap, want_cols * 32, 16)) return -1;
      }
      uint32_t *tile_cols = f.d_bstiles, *tile_off = tile_cols + max_tiles;
      const uint32_t wup = bs_warmup(e->m, e->tau);
      uint32_t *gmask = s.d_gmask, *gfollow = cut ? s.d_gmask + s.gmask_cap / 2 : nullptr;
      uint8_t *segstop = s.d_segflags;
      if (timing && (front || fused)) CU(record_event(s.ev[E_PACK_BEGIN], st));
      if (!front && !fused) {
         BsPrepArgs bp{s.d_ls, (uint32_t)lines_cap, n, ctr, tile_cols, tile_off, (uint32_t)max_tiles,
                       (unsigned long long)(s.planes_cap / 32), e->bs_gate, cut ? s.d_lid : nullptr, wup,
                       (!cut && cuts_allowed(e, options, n)) ? 1 : 0, filter ? s.d_act : nullptr};
         k15_tile_cols<<<(int)std::min<size_t>(div_up(max_tiles, kWarps), (size_t)e->sms * 8), kThreads, 0, st>>>(bp);
         k15_scan<<<1, 1024, 0, st>>>(bp);
         if (timing) CU(record_event(s.ev[E_PACK_BEGIN], st));
         BsPackArgs pk{(const uint4 *)s.d_codes, (uint32_t)(div_up(n, kK1Tile) * (kK1Tile / 32)), s.d_ls,
                       (uint32_t)lines_cap, ctr, tile_cols, tile_off, s.d_planes, cut ? s.d_lid : nullptr, wup,
                       gmask, gfollow, filter ? s.d_act : nullptr};
         k15_pack<<<(int)std::max<size_t>(1, std::min<size_t>(div_up(div_up(max_lines, 64), kWarps), (size_t)e->sms * 16)),
                    kThreads, 0, st>>>(pk);
         CU(cudaGetLastError());
         s.launches += 3;
      }
      if (timing) CU(record_event(s.ev[E_PACK_END], st));
      K2BsArgs kb{f.d_planes, tile_cols, tile_off, (uint32_t)lines_cap, ctr, s.d_res, s.d_cnt, s.d_ev, k2.ev_cap,
                  (mode == M_COUNT || mode == M_COUNTALL) ? 1 : 0, cut ? gmask : nullptr, gfollow, segstop, wup,
                  (filter && !fused) ? f.d_act : nullptr,
                  fused ? s.d_gdesc : nullptr, (uint32_t)std::min<size_t>(s.gdesc_cap, 0xffffffffu),
                  (fused && filter) ? s.d_gent : nullptr, fused ? s.d_tiles + 2 * k1_tiles : nullptr};
      if (launch_bitslice(e, mode, options, max_lines, st, kb)) return -1;
      s.launches++;
   }
   {
      const int grid = (int)std::min<size_t>(div_up(max_lines, lines_per_cta), (size_t)e->sms * 8);
      if (launch_k2(e, mode, grid, st, k2, fwd)) return -1;
      s.launches++;
   }
   if (timing) CU(record_event(s.ev[E_MATCH_END], st));
   if (cut) {
      // one result per line out of the results per segment
      SegReduceArgs sr{s.d_lid, (uint32_t)lines_cap, ctr, s.d_res, s.d_cnt, s.d_segflags, s.d_segflags + s.line_cap, mode,
                       filter ? s.d_lflags : nullptr};
      k_seg_reduce<<<(int)std::max<size_t>(1, std::min<size_t>(div_up(max_lines, kThreads), (size_t)e->sms * 8)),
                     kThreads, 0, st>>>(sr);
      CU(cudaGetLastError());
      s.launches++;
   }
   if (timing) CU(record_event(s.ev[E_K2_END], st));

   // ---- scan + K3/K4 --------------------------------------------------------
   uint32_t *tile_sum = s.d_fintiles, *tile_nz = tile_sum + fin_tiles, *tile_recbase = tile_nz + fin_tiles;
   FinArgs fa{d_text, f.d_ls, (uint32_t)lines_cap, s.d_res, s.d_offs, s.d_ev, k2.ev_cap, s.d_recs,
              (uint32_t)std::min<size_t>(s.rec_cap, 0xffffffffu), ctr, tile_recbase,
              cut ? s.d_lid : nullptr, cut ? s.d_lbeg : nullptr,
              (cut && mode == M_ALL) ? s.d_segflags + s.line_cap : nullptr, bs_warmup(e->m, e->tau)};
   if (mode == M_FIRST || mode == M_BEST || mode == M_ALL) {
      const bool all = mode == M_ALL;
      TileSumArgs ts{all ? nullptr : s.d_res, all ? s.d_cnt : nullptr, (uint32_t)lines_cap, ctr, tile_sum, tile_nz,
                     tile_recbase};
      const int gsum = (int)std::max<size_t>(1, std::min<size_t>(div_up(div_up(max_lines, kFinTile), kWarps), (size_t)e->sms * 8));
      k_tile_sums<<<gsum, kThreads, 0, st>>>(ts);
      k_tile_scan<<<1, 1024, 0, st>>>(ts);
      const int gtile = (int)std::max<size_t>(1, std::min<size_t>(div_up(max_lines, kFinTile), (size_t)e->sms * 8));
      if (count_cut) {
         s.launches += 2;                     // the counts are in place: no records wanted
      } else if (!all) {
         if (launch_finish(e, false, gtile, st, fa, rev)) return -1;
         s.launches += 3;
      } else {
         OffsArgs oa{s.d_cnt, s.d_offs, (uint32_t)lines_cap, ctr, tile_recbase};
         k_offsets<<<gtile, kThreads, 0, st>>>(oa);
         CU(cudaGetLastError());
         const int grid2 = (int)std::min<size_t>(div_up(std::max<size_t>(s.ev_cap, 1), kThreads), (size_t)e->sms * 8);
         if (launch_finish(e, true, single ? 1 : grid2, st, fa, rev)) return -1;
         s.launches += 4;
      }
   }
   if (timing) CU(record_event(s.ev[E_FIN_END], st));
   CU(cudaMemcpyAsync(s.h_ctr, ctr, C_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
   return 0;
}

// Queue one scan on `st`.  A scan is enqueued call by call; when the same scan has been asked for
// three times in a row, the same calls are recorded into a CUDA graph (stream capture: every
// capacity is in place by then, so no allocation happens underneath) and from then on a scan is
// ONE graph launch -- the ~15 launches and memsets of a step cost the host nothing, and the
// device runs them back to back.  With SQB_TIMING the event records around the single kernels
// are nodes of the graph.  Not for single strings.
static int slot_issue(sqb_engine *e, Slot &s, const uint8_t *d_text, uint32_t n, int options, cudaStream_t st,
                      uint32_t skip = 0, bool may_replay = false, const Slot *front = nullptr)
{
   Slot::GraphKey key;
   key.text = d_text; key.n = n; key.skip = skip; key.options = options; key.st = st; key.version = e->version;
   // only the direct device scans replay (the chunk pipelines scan a different chunk every time, and
   // two chunks of equal size in a row would pay for an instantiation that is used once), and only
   // from the third identical scan on
   const bool eligible = may_replay && !front && e->graphs && !s.gbroken && !(options & SQB_SINGLE_LINE) && st != nullptr;
   if (eligible && key == s.gkey && ++s.grepeats >= 2) {
      if (s.gexec == nullptr) {
         cudaGraph_t graph = nullptr;
         bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess;
         if (ok) {
            const int rc = slot_enqueue(e, s, d_text, n, options, st, skip, nullptr);
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            ok = rc == 0 && ce == cudaSuccess && graph != nullptr;
            if (ok) ok = cudaGraphInstantiate(&s.gexec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
         }
         if (!ok) {
            cudaGetLastError();
            if (s.gexec) { cudaGraphExecDestroy(s.gexec); s.gexec = nullptr; }
            s.gbroken = true;
         } else {
            s.glaunches = s.launches;
         }
      }
      if (s.gexec != nullptr) {
         s.cur_text = d_text; s.cur_n = n; s.cur_skip = skip; s.cur_options = options; s.cur_stream = st;
         s.cur_front = nullptr;
         s.launches = s.glaunches;
         s.busy = true;
         CU(cudaGraphLaunch(s.gexec, st));
         CU(cudaEventRecord(s.ev[E_DONE], st));
         return 0;
      }
   } else if (s.gexec != nullptr) {
      cudaGraphExecDestroy(s.gexec);
      s.gexec = nullptr;
   }
   if (!(eligible && key == s.gkey)) s.grepeats = 0;
   s.gkey = eligible ? key : Slot::GraphKey();
   if (slot_enqueue(e, s, d_text, n, options, st, skip, front)) return -1;
   CU(cudaEventRecord(s.ev[E_DONE], st));
   return 0;
}

// waits for the scan in `s`; repeats it if a capacity was exceeded
static int slot_finish(sqb_engine *e, Slot &s, sqb_stats_t *stats)
{
   uint32_t reruns = 0;
   for (;;) {
      CU(cudaEventSynchronize(s.ev[E_DONE]));
      const int mode = mode_of(s.cur_options);
      const unsigned long long nlines = std::max(s.h_ctr[C_NLINES], s.h_ctr[C_NPSEUDO]);   // entries of ls
      const unsigned long long nev = s.h_ctr[C_EVENTS];
      bool again = false;
      const bool follower = s.cur_front != nullptr;         // the front's owner repeats a scan whose front was short
      if (!follower && nlines + 1 > s.line_cap) {
         e->lines_per_byte = (double)(nlines + 2) / (double)std::max<uint32_t>(s.cur_n, 1) * 1.05;
         again = true;
      }
      if (!follower && s.h_ctr[C_BS_SELECTED] == 3ull) {          // long lines: cut them, in this scan and from now on
         e->cuts_wanted = true;
         again = true;
      }
      if (!follower && s.h_ctr[C_BS_SELECTED] == 2ull) {          // plane buffer too small for the bit-sliced scan
         if (s.cur_fused) e->units_per_byte = (double)(s.h_ctr[C_BS_COLS] + 64) / (double)std::max<uint32_t>(s.cur_n, 1) * 1.05;
         else e->cols_per_byte = (double)(s.h_ctr[C_BS_COLS] + 64) / (double)std::max<uint32_t>(s.cur_n, 1) * 1.05;
         again = true;
      }
      if (!follower && s.h_ctr[C_BS_SELECTED] == 4ull) {
         // the fused kernel met a line beyond its overlap (or a tile with too many line starts): a wider
         // overlap once, then the two-kernel path for this engine
         if (e->fused_ov < kFMaxOverlap) e->fused_ov = kFMaxOverlap;
         else e->fused = 0;
         again = true;
      }
      if (mode == M_ALL && nev > s.ev_cap) {
         e->events_per_byte = (double)(nev + 1) / (double)std::max<uint32_t>(s.cur_n, 1) * 1.05;
         again = true;
      }
      if (!again) break;
      e->version++;
      if (++reruns > 6) { set_err("capacity re-run did not converge"); return -1; }
      if (slot_issue(e, s, s.cur_text, s.cur_n, s.cur_options, s.cur_stream, s.cur_skip, false, s.cur_front)) return -1;
   }
   s.busy = false;
   if (!s.cur_front && s.cur_filter && !(s.cur_options & SQB_FASTQ) && e->filter == 1 && e->filter_state < 0 &&
       s.h_ctr[C_NPSEUDO] > 0) {
      // the probe: keep filtering if it drops a quarter of the lines or more
      const double live = (double)s.h_ctr[C_NACTIVE] / (double)s.h_ctr[C_NPSEUDO];
      e->filter_state = live <= 0.75 ? 1 : 0;
      e->version++;
   }
   if (stats) {
      memset(stats, 0, sizeof *stats);
      stats->nbytes = s.cur_n;
      stats->nlines = s.h_ctr[C_NLINES];
      stats->nmatched = s.h_ctr[C_NMATCHED];
      stats->nrecs = s.h_ctr[C_NRECS];
      stats->launches = s.launches;
      stats->reruns = reruns;
      stats->devices = 1;
      stats->path = (s.h_ctr[C_BS_SELECTED] == 1ull ? SQB_PATH_BITSLICE : 0u) | (s.h_ctr[C_NCUTS] ? SQB_PATH_CUTS : 0u) |
                    ((s.cur_fused && s.h_ctr[C_BS_SELECTED] == 1ull) ? SQB_PATH_FUSED : 0u) |
                    (s.cur_filter ? SQB_PATH_FILTER : 0u);
      if (s.cur_options & SQB_TIMING) {
         static const int span[6][2] = {{E_BEGIN, E_K1_END}, {E_K1_END, E_K2_END}, {E_K2_END, E_FIN_END},
                                        {E_PACK_END, E_MATCH_END}, {E_PACK_BEGIN, E_PACK_END}, {E_BEGIN, E_K1C_END}};
         float ms = 0;
         CU(cudaEventElapsedTime(&ms, s.ev[E_BEGIN], s.ev[E_FIN_END]));
         stats->device_ms = ms;
         for (int k = 0; k < 6; k++) {
            // (a follower of a pattern set has no K1 / pack of its own: empty spans)
            if (cudaEventElapsedTime(&ms, s.ev[span[k][0]], s.ev[span[k][1]]) != cudaSuccess) { cudaGetLastError(); ms = 0; }
            stats->kernel_ms[k] = ms;
         }
      }
   }
   return 0;
}

// records of a chunk -> buffer-global line numbers (in place, or into the array of all chunks)
static __global__ void k_add_line_base(Rec *dst, const Rec *src, unsigned long long n, uint32_t base)
{
   for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
        i += (unsigned long long)gridDim.x * blockDim.x) {
      Rec r = src[i];
      r.line += base;
      dst[i] = r;
   }
}

// last (or first) '\n' of text[lo, hi): *out = max (min) over 1 + its position
static __global__ void k_find_newline(const uint8_t *text, unsigned long long lo, unsigned long long hi, int last,
                                      unsigned long long *out)
{
   unsigned long long best = last ? 0ull : ~0ull;
   const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
   for (unsigned long long p = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < hi; p += stride)
      if (text[p] == '\n') {
         if (last) best = p + 1ull;
         else { best = p + 1ull; break; }
      }
   if (last) { if (best) atomicMax(out, best); }
   else if (best != ~0ull) atomicMin(out, best);
}

// ---------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------
extern "C" {

const char *sqbLastError(void) { return g_err; }
void sqb_set_error(const char *msg) { set_err("%s", msg); }      // for the other translation units (sqb_bgzf.cu)
int sqbEngineDevice(sqb_engine_t *e) { return e ? e->device : -1; }

int sqbDeviceCount(void)
{
   int n = 0;
   if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
   return n;
}

int sqbMaxPatternLength(void) { return kMaxWords * 32; }

sqb_engine_t *sqbEngineNew(const unsigned char *keys, int m, int tau, int device)
{
   if (keys == NULL || m < 1 || tau < 0 || tau >= m) {
      set_err("invalid pattern (m=%d, tau=%d)", m, tau);
      return NULL;
   }
   if (m > kMaxWords * 32) {
      set_err("pattern of %d positions exceeds the %d supported by the blocked automaton", m, kMaxWords * 32);
      return NULL;
   }
   int ndev = 0;
   cudaError_t ce = cudaGetDeviceCount(&ndev);
   if (ce != cudaSuccess || ndev == 0) {
      set_err("no CUDA device available (%s); this library has no CPU fallback",
              ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
      cudaGetLastError();
      return NULL;
   }
   if (device < 0) {
      const char *env = getenv("SEEQ_B200_DEVICE");
      if (env == NULL) env = getenv("LOCAL_RANK");
      device = env ? atoi(env) % ndev : 0;
   }
   if (device >= ndev) { set_err("device %d out of range (%d devices)", device, ndev); return NULL; }
   if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice(%d) failed", device); return NULL; }
   sqb_engine *e = new sqb_engine();
   e->device = device;
   cudaDeviceGetAttribute(&e->sms, cudaDevAttrMultiProcessorCount, device);
   e->m = m;
   e->tau = tau;
   memcpy(e->keys, keys, (size_t)m);
   e->filter_k = std::min(8, std::max(1, m - tau));
   const int words = (m + 31) / 32;
   e->words = 1;
   while (e->words < words) e->words *= 2;
   if (e->words > 2 && e->words < 4) e->words = 4;
   memset(&e->last_stats, 0, sizeof e->last_stats);
   e->bs_ok = build_bs_pattern(e->keys, m, tau, &e->bs_pat);
   // test / tuning knobs: SEEQ_B200_MATCHER=word forces the word-parallel kernels,
   // =bitslice lifts the size thresholds of the bit-sliced one
   if (const char *mk = getenv("SEEQ_B200_MATCHER")) {
      if (!strcmp(mk, "word")) e->bs_ok = false;
      if (!strcmp(mk, "bitslice")) { e->bs_gate = BsGate{1u, 1u << 30}; e->bs_min_bytes = 1; }
   }
   if (const char *c = getenv("SEEQ_B200_CUTS")) e->cuts = atoi(c);
   if (const char *c = getenv("SEEQ_B200_FILTER")) e->filter = atoi(c);
   if (const char *c = getenv("SEEQ_B200_NFA")) e->nfa_levels = atoi(c) != 0;
   if (const char *c = getenv("SEEQ_B200_GRAPHS")) e->graphs = atoi(c) != 0;
   if (const char *c = getenv("SEEQ_B200_FUSED")) e->fused = std::max(0, std::min(2, atoi(c)));
   if (const char *c = getenv("SEEQ_B200_FUSED_OV")) e->fused_ov = std::min<uint32_t>(kFMaxOverlap, std::max(32, atoi(c)) & ~31u);
   return e;
}

void sqbEngineFree(sqb_engine_t *e)
{
   if (e == NULL) return;
   for (sqb_engine *p : e->peers) sqbEngineFree(p);
   e->peers.clear();
   cudaSetDevice(e->device);
   for (auto &s : e->slot) slot_free(s);
   if (e->host_recs) cudaFreeHost(e->host_recs);
   if (e->d_all_recs) cudaFree(e->d_all_recs);
   if (e->d_word) cudaFree(e->d_word);
   if (e->h_word) cudaFreeHost(e->h_word);
   if (e->big_stream) cudaStreamDestroy(e->big_stream);
   delete e;
}

int sqbScanDevice(sqb_engine_t *e, const void *d_text, size_t nbytes, int options, void *stream,
                  sqb_stats_t *stats)
{
   if (nbytes >= kMaxBatch) { set_err("sqbScanDevice: %zu bytes exceed the batch limit of %zu", nbytes, (size_t)kMaxBatch); return -1; }
   if (((uintptr_t)d_text & 15) != 0) { set_err("sqbScanDevice: text pointer must be 16-byte aligned"); return -1; }
   CU(cudaSetDevice(e->device));
   Slot &s = e->slot[0];
   if (slot_init(s)) return -1;
   e->last_slot = 0;
   if (nbytes == 0 && !(options & SQB_SINGLE_LINE)) {
      memset(&e->last_stats, 0, sizeof e->last_stats);
      if (stats) memset(stats, 0, sizeof *stats);
      memset(s.h_ctr, 0, C_COUNT * sizeof(unsigned long long));
      return 0;
   }
   cudaStream_t st = stream ? (cudaStream_t)stream : s.stream;
   if (slot_issue(e, s, (const uint8_t *)d_text, (uint32_t)nbytes, options, st, 0, true)) return -1;
   if (slot_finish(e, s, &e->last_stats)) return -1;
   if (stats) *stats = e->last_stats;
   return 0;
}

int sqbScanDeviceIssue(sqb_engine_t *e, int slot, const void *d_text, size_t nbytes, int options, void *stream)
{
   if (slot < 0 || slot > 1) { set_err("sqbScanDeviceIssue: slot %d out of range", slot); return -1; }
   if (nbytes == 0 || nbytes >= kMaxBatch) { set_err("sqbScanDeviceIssue: %zu bytes are outside the batch limits", nbytes); return -1; }
   if (((uintptr_t)d_text & 15) != 0) { set_err("sqbScanDeviceIssue: text pointer must be 16-byte aligned"); return -1; }
   CU(cudaSetDevice(e->device));
   Slot &s = e->slot[slot];
   if (slot_init(s)) return -1;
   if (s.busy) { set_err("sqbScanDeviceIssue: slot %d has a scan in flight", slot); return -1; }
   return slot_issue(e, s, (const uint8_t *)d_text, (uint32_t)nbytes, options, stream ? (cudaStream_t)stream : s.stream, 0, true);
}

int sqbScanDeviceWait(sqb_engine_t *e, int slot, sqb_stats_t *stats)
{
   if (slot < 0 || slot > 1 || !e->slot[slot].busy) { set_err("sqbScanDeviceWait: no scan in flight in slot %d", slot); return -1; }
   CU(cudaSetDevice(e->device));
   e->last_slot = slot;
   if (slot_finish(e, e->slot[slot], &e->last_stats)) return -1;
   if (stats) *stats = e->last_stats;
   return 0;
}

const sqb_rec_t *sqbDeviceRecords(sqb_engine_t *e) { return (const sqb_rec_t *)e->slot[e->last_slot].d_recs; }
const uint32_t *sqbDeviceLineStarts(sqb_engine_t *e) { return e->slot[e->last_slot].d_ls; }

int sqbFetchRecords(sqb_engine_t *e, sqb_rec_t *dst, uint64_t first, uint64_t count)
{
   Slot &s = e->slot[e->last_slot];
   if (first + count > s.h_ctr[C_NRECS]) { set_err("sqbFetchRecords: range beyond %llu records", s.h_ctr[C_NRECS]); return -1; }
   if (count == 0) return 0;
   CU(cudaMemcpy(dst, s.d_recs + first, count * sizeof(Rec), cudaMemcpyDeviceToHost));
   return 0;
}

int sqbFetchLineStarts(sqb_engine_t *e, uint32_t *dst, uint64_t first, uint64_t count)
{
   Slot &s = e->slot[e->last_slot];
   if (first + count > s.h_ctr[C_NLINES] + 1) { set_err("sqbFetchLineStarts: range beyond %llu lines", s.h_ctr[C_NLINES]); return -1; }
   if (count == 0) return 0;
   CU(cudaMemcpy(dst, s.d_ls + first, count * sizeof(uint32_t), cudaMemcpyDeviceToHost));
   return 0;
}

// ---- host path ----------------------------------------------------------------
static size_t host_chunk_bytes(void)
{
   const char *env = getenv("SEEQ_B200_CHUNK_MB");
   size_t mb = env ? (size_t)atol(env) : 64;
   if (mb < 1) mb = 1;
   if (mb > 2048) mb = 2048;
   return mb << 20;
}

// device-resident buffers: chunks below 2 GiB keep the line filter available (kDeadBit)
static size_t device_chunk_bytes(void)
{
   const char *env = getenv("SEEQ_B200_DEVICE_CHUNK_MB");
   size_t mb = env ? (size_t)atol(env) : 1536;
   if (mb < 1) mb = 1;
   if (mb > 2040) mb = 2040;
   return mb << 20;
}

// collect the results of the chunk in flight in slot s (in chunk order)
static int host_collect(sqb_engine *e, Slot &s, int options, uint64_t *line_base, sqb_stats_t *acc)
{
   sqb_stats_t st;
   if (slot_finish(e, s, &st)) return -1;
   const bool count_only = options & SQB_COUNT_ONLY;
   const bool on_device = options & SQB_DEVICE_RESULTS;
   if (!count_only && on_device && st.nrecs > 0) {
      // the records stay in HBM: appended, rebased, to the device array of the whole scan
      const size_t need = e->d_all_n + (size_t)st.nrecs;
      if (need > e->d_all_cap) {
         Rec *bigger = nullptr;
         const size_t cap = need + need / 2 + (1u << 16);
         CU(cudaMalloc((void **)&bigger, cap * sizeof(Rec)));
         if (e->d_all_n) CU(cudaMemcpy(bigger, e->d_all_recs, e->d_all_n * sizeof(Rec), cudaMemcpyDeviceToDevice));
         if (e->d_all_recs) CU(cudaFree(e->d_all_recs));
         e->d_all_recs = bigger;
         e->d_all_cap = cap;
      }
      k_add_line_base<<<(int)std::min<size_t>(div_up((size_t)st.nrecs, 256), (size_t)e->sms * 8), 256, 0, s.stream>>>(
         e->d_all_recs + e->d_all_n, s.d_recs, st.nrecs, (uint32_t)*line_base);
      CU(cudaGetLastError());
      e->d_all_n += (size_t)st.nrecs;
   }
   // the records go straight into the (pinned) result array of the scan
   if (!count_only && !on_device && st.nrecs > 0) {
      const size_t need = e->host_recs_n + (size_t)st.nrecs;
      if (need > e->host_recs_cap) {
         sqb_rec_t *bigger = nullptr;
         const size_t cap = need + need / 2 + (1u << 16);
         CU(pinned_alloc((void **)&bigger, cap * sizeof(sqb_rec_t)));
         if (e->host_recs_n) memcpy(bigger, e->host_recs, e->host_recs_n * sizeof(sqb_rec_t));
         if (e->host_recs) CU(cudaFreeHost(e->host_recs));
         e->host_recs = bigger;
         e->host_recs_cap = cap;
      }
      // lines of a chunk are numbered from 0 on the device: rebase them there, before the copy
      // (the scan is complete; the slot's stream carries only this)
      if (*line_base) {
         k_add_line_base<<<(int)std::min<size_t>(div_up((size_t)st.nrecs, 256), (size_t)e->sms * 8), 256, 0, s.stream>>>(
            s.d_recs, s.d_recs, st.nrecs, (uint32_t)*line_base);
         CU(cudaGetLastError());
      }
      CU(cudaMemcpyAsync(e->host_recs + e->host_recs_n, s.d_recs, st.nrecs * sizeof(Rec), cudaMemcpyDeviceToHost, s.stream));
   }
   // with segment cuts the entries of ls are segments: the line starts are the entries whose line differs from the one
   // before (lid), picked out on the host
   const bool keep_cut = (options & SQB_KEEP_LINES_INTERNAL) && s.h_ctr[C_NCUTS] != 0ull;
   const size_t nent = keep_cut ? (size_t)s.h_ctr[C_NPSEUDO] : (size_t)st.nlines;
   if ((options & SQB_KEEP_LINES_INTERNAL) && st.nlines > 0) {
      if (pin_reserve(&s.h_ls, &s.h_ls_cap, nent * (keep_cut ? 2 : 1))) return -1;
      CU(cudaMemcpyAsync(s.h_ls, s.d_ls, nent * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
      if (keep_cut) CU(cudaMemcpyAsync(s.h_ls + nent, s.d_lid, nent * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
   }
   CU(cudaStreamSynchronize(s.stream));
   if (!count_only && !on_device && st.nrecs > 0) e->host_recs_n += (size_t)st.nrecs;
   if ((options & SQB_KEEP_LINES_INTERNAL) && st.nlines > 0) {
      const size_t old = e->host_lines.size();
      e->host_lines.resize(old + st.nlines);
      if (!keep_cut) {
         for (uint64_t k = 0; k < st.nlines; k++) e->host_lines[old + k] = (uint64_t)s.chunk_off + s.h_ls[k];
      } else {
         const uint32_t *lid = s.h_ls + nent;
         size_t w = 0;
         for (size_t k = 0; k < nent; k++)
            if ((k == 0 || lid[k] != lid[k - 1]) && w < (size_t)st.nlines) e->host_lines[old + w++] = (uint64_t)s.chunk_off + s.h_ls[k];
         if (w != (size_t)st.nlines) { set_err("segment cuts: %zu line starts for %llu lines", w, (unsigned long long)st.nlines); return -1; }
      }
   }
   *line_base += st.nlines;
   acc->nbytes += st.nbytes;
   acc->nlines += st.nlines;
   acc->nmatched += st.nmatched;
   acc->nrecs += st.nrecs;
   acc->launches += st.launches;
   acc->reruns += st.reruns;
   acc->path = st.path;
   acc->devices = 1;
   acc->device_ms += st.device_ms;
   for (int k = 0; k < 8; k++) acc->kernel_ms[k] += st.kernel_ms[k];
   return 0;
}

// Chunk boundaries of a DEVICE-resident buffer: cuts[0] = 0 < cuts[1] < ... = nbytes, every inner cut
// just behind a '\n'.  The last newline of a window is looked for in its final 1 MiB, then 64 MiB,
// then all of it; a window without any is extended to the next newline.
static int device_cuts(sqb_engine *e, const uint8_t *d_text, size_t nbytes, size_t chunk, int options,
                       std::vector<size_t> &cuts)
{
   std::vector<char> win;
   Slot &s = e->slot[0];
   if (e->d_word == nullptr) {
      CU(cudaMalloc((void **)&e->d_word, sizeof(unsigned long long)));
      CU(pinned_alloc((void **)&e->h_word, sizeof(unsigned long long)));
   }
   auto find = [&](size_t lo, size_t hi, bool last, size_t *where) -> int {
      *e->h_word = last ? 0ull : ~0ull;
      CU(cudaMemcpyAsync(e->d_word, e->h_word, sizeof(unsigned long long), cudaMemcpyHostToDevice, s.stream));
      const int grid = (int)std::max<size_t>(1, std::min<size_t>(div_up(hi - lo, 256 * 16), (size_t)e->sms * 8));
      k_find_newline<<<grid, 256, 0, s.stream>>>(d_text, lo, hi, last ? 1 : 0, e->d_word);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(e->h_word, e->d_word, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
      CU(cudaStreamSynchronize(s.stream));
      *where = (size_t)*e->h_word;            // 1 + position; 0 / ~0: none
      return 0;
   };
   cuts.assign(1, 0);
   size_t pos = 0;
   while (nbytes - pos > chunk) {
      size_t at = 0;
      const size_t hi = pos + chunk;
      for (size_t win : {(size_t)1 << 20, (size_t)64 << 20, chunk}) {
         const size_t lo = hi - std::min(win, chunk);
         if (find(lo, hi, true, &at)) return -1;
         if (at != 0 || lo == pos) break;
      }
      if (at == 0) {
         if (find(hi, nbytes, false, &at)) return -1;
         if (at == (size_t)~0ull) break;                 // no newline left: the rest is one chunk
      }
      if (at >= nbytes) break;
      if (options & SQB_FASTQ) {
         // back to a record boundary: the heuristic runs on the host over a window around the cut
         const size_t half = 32u << 10;
         const size_t w0 = at - pos > half ? at - half : pos, w1 = std::min(nbytes, at + half);
         win.resize(w1 - w0);
         CU(cudaMemcpy(win.data(), d_text + w0, w1 - w0, cudaMemcpyDeviceToHost));
         const size_t q = fastq_record_start(win.data(), 0, at - w0, w1 - w0);
         if (q == (size_t)-1 || w0 + q <= pos) {
            set_err("SQB_FASTQ: no record boundary ('@' line, '+' two lines on) in the 32 KiB in front of byte %zu", at);
            return -1;
         }
         at = w0 + q;
      }
      cuts.push_back(at);
      pos = at;
   }
   cuts.push_back(nbytes);
   return 0;
}

// The chunk pipeline behind sqbScanHost, sqbScanDeviceLarge and the multi-pattern scans:
// newline-aligned chunks, two slots, the results of chunk k are collected while chunk k+1 runs.
// Host text is copied into the slot's device buffer; device text is scanned where it lies (a chunk
// that does not start on a 16-byte boundary starts `skip` bytes into its aligned address,
// K1Args::skip).  With P > 1 engines (a pattern set), engs[0] scans every chunk in full and the
// others read its line starts, line filter and bit-planes (slot_enqueue: front) on the same stream.
static int scan_chunks(sqb_engine **engs, int P, const char *text, size_t nbytes, int options, sqb_stats_t *stats,
                       bool on_device, cudaStream_t user_stream)
{
   sqb_engine *e = engs[0];
   CU(cudaSetDevice(e->device));
   std::vector<sqb_stats_t> acc((size_t)P);
   std::vector<uint64_t> line_base((size_t)P, 0);
   for (int p = 0; p < P; p++) {
      for (auto &s : engs[p]->slot) if (slot_init(s)) return -1;
      engs[p]->host_recs_n = 0;
      engs[p]->d_all_n = 0;
      engs[p]->host_lines.clear();
      memset(&acc[(size_t)p], 0, sizeof(sqb_stats_t));
   }
   const bool single = options & SQB_SINGLE_LINE;
   const size_t chunk = on_device ? device_chunk_bytes() : host_chunk_bytes();
   std::vector<size_t> cuts;
   if (on_device && !single && device_cuts(e, (const uint8_t *)text, nbytes, chunk, options, cuts)) return -1;
   // device text: the kernels of all chunks follow each other on ONE stream (the caller's, or one of
   // the engine's own) while the slots' streams carry the record copies of the chunk before
   if (on_device && user_stream == nullptr) {
      if (e->big_stream == nullptr) CU(cudaStreamCreateWithFlags(&e->big_stream, cudaStreamNonBlocking));
      user_stream = e->big_stream;
   }
   // results of the chunk in flight in slot k, leader first: if the leader had to repeat its scan
   // (a capacity guess was low) the others have read a front that was cut short and go again
   auto collect = [&](int k) -> int {
      Slot &ls = e->slot[k];
      const uint32_t before = acc[0].reruns;
      if (host_collect(e, ls, options, &line_base[0], &acc[0])) return -1;
      for (int p = 1; p < P; p++) {
         Slot &fs = engs[p]->slot[k];
         if (acc[0].reruns != before) {
            if (slot_finish(engs[p], fs, nullptr)) return -1;
            if (slot_issue(engs[p], fs, ls.cur_text, ls.cur_n, options, ls.cur_stream, ls.cur_skip, false, &ls)) return -1;
         }
         if (host_collect(engs[p], fs, options, &line_base[(size_t)p], &acc[(size_t)p])) return -1;
      }
      return 0;
   };
   size_t pos = 0;
   int c = 0;
   int pending[2] = {0, 0};
   while (pos < nbytes || (single && c == 0)) {
      size_t len;
      if (single) {
         len = nbytes;
      } else if (on_device) {
         len = cuts[(size_t)c + 1] - pos;
      } else {
         len = nbytes - pos;
         if (len > chunk) {
            // cut after the last newline inside the window, else extend to the next one
            const char *nl = (const char *)memrchr(text + pos, '\n', chunk);
            if (nl == NULL) nl = (const char *)memchr(text + pos + chunk, '\n', nbytes - pos - chunk);
            len = nl ? (size_t)(nl - (text + pos)) + 1 : nbytes - pos;
            if ((options & SQB_FASTQ) && pos + len < nbytes) {
               const size_t q = fastq_record_start(text, pos, pos + len, nbytes);
               if (q == (size_t)-1 || q <= pos) {
                  set_err("SQB_FASTQ: no record boundary ('@' line, '+' two lines on) in front of byte %zu", pos + len);
                  return -1;
               }
               len = q - pos;
            }
         }
      }
      if (len + 16 >= kMaxBatch) { set_err("a single line of %zu bytes exceeds the batch limit of %zu", len, (size_t)kMaxBatch); return -1; }
      const int k = c & 1;
      Slot &s = e->slot[k];
      if (pending[k]) {
         if (collect(k)) return -1;
         pending[k] = 0;
      }
      const uint8_t *d_chunk = nullptr;
      uint32_t skip = 0;
      cudaStream_t st = s.stream;
      if (on_device) {
         skip = single ? 0u : (uint32_t)((uintptr_t)(text + pos) & 15u);
         if (single && ((uintptr_t)text & 15u)) { set_err("device text must be 16-byte aligned"); return -1; }
         d_chunk = (const uint8_t *)text + pos - skip;
         st = user_stream;
      } else {
         if (dev_reserve(&s.d_text, &s.text_cap, len + 64)) return -1;
         if (len) CU(cudaMemcpyAsync(s.d_text, text + pos, len, cudaMemcpyHostToDevice, s.stream));
         d_chunk = s.d_text;
      }
      if (len || single) {
         for (int p = 0; p < P; p++) {
            Slot &ps = engs[p]->slot[k];
            ps.chunk_off = pos - skip;
            if (slot_issue(engs[p], ps, d_chunk, (uint32_t)(len + skip), options, st, skip, false, p ? &s : nullptr)) return -1;
         }
         pending[k] = 1;
      }
      pos += len;
      c++;
      if (single) break;
   }
   // drain in chunk order
   for (int j = 0; j < 2; j++) {
      const int k = (c + j) & 1;
      if (pending[k]) {
         if (collect(k)) return -1;
         pending[k] = 0;
      }
   }
   for (int p = 0; p < P; p++) {
      acc[(size_t)p].nbytes = nbytes;          // (a device chunk counts its alignment bytes)
      engs[p]->last_stats = acc[(size_t)p];
      if (stats) stats[p] = acc[(size_t)p];
   }
   return 0;
}

// ---- several GPUs from one process ------------------------------------------------
// $SEEQ_B200_DEVICES = "all" or a count: sqbScanHost (and with it seeqBatchMatch, seeqFileMatch and seeq())
// cuts the buffer into newline-aligned shards (sqbShardRange), one per device, and drives every device from
// a host thread of its own -- each with its own engine, streams and chunk pipeline, so every PCIe link
// carries its shard at the same time.  No collective: the line bases are prefix-summed on the host and
// the records of the shards are concatenated in shard order.  The reference is single-threaded
// (seeq.c:293-392); its "N cores" baseline runs N processes over the same kind of shards.
static int host_devices(const sqb_engine *e, size_t nbytes, int options)
{
   if (options & (SQB_SINGLE_LINE | SQB_DEVICE_RESULTS)) return 1;
   const char *env = getenv("SEEQ_B200_DEVICES");
   if (env == nullptr || *env == 0) return 1;
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return 1; }
   int want = strcmp(env, "all") == 0 ? ndev : atoi(env);
   want = std::max(1, std::min(want, ndev));
   // a shard below 16 MiB does not pay for its launches
   const size_t min_shard = getenv("SEEQ_B200_MIN_SHARD_MB") ? (size_t)atol(getenv("SEEQ_B200_MIN_SHARD_MB")) << 20 : (size_t)16 << 20;
   while (want > 1 && nbytes / (size_t)want < std::max<size_t>(min_shard, 1)) want--;
   (void)e;
   return want;
}

static int scan_host_multi(sqb_engine *e, int nd, const char *text, size_t nbytes, int options, sqb_stats_t *stats)
{
   int ndev = 0;
   CU(cudaGetDeviceCount(&ndev));
   while ((int)e->peers.size() < nd - 1) {
      const int dev = (e->device + 1 + (int)e->peers.size()) % ndev;
      sqb_engine *p = sqbEngineNew(e->keys, e->m, e->tau, dev);
      if (p == nullptr) return -1;
      p->cuts = e->cuts; p->filter = e->filter; p->fused = e->fused; p->bs_ok = e->bs_ok;
      p->bs_gate = e->bs_gate; p->bs_min_bytes = e->bs_min_bytes;
      e->peers.push_back(p);
   }
   CU(cudaSetDevice(e->device));
   std::vector<sqb_engine *> eng((size_t)nd);
   eng[0] = e;
   for (int r = 1; r < nd; r++) eng[(size_t)r] = e->peers[(size_t)r - 1];
   // shards: newline-aligned, and at record boundaries with SQB_FASTQ
   std::vector<size_t> cut((size_t)nd + 1, 0);
   for (int r = 0; r < nd; r++) {
      size_t b, en;
      sqbShardRange(text, nbytes, r, nd, &b, &en);
      cut[(size_t)r] = b;
      cut[(size_t)r + 1] = en;
   }
   if (options & SQB_FASTQ) {
      for (int r = 1; r < nd; r++) {
         if (cut[(size_t)r] >= nbytes) continue;
         const size_t q = fastq_record_start(text, cut[(size_t)r - 1], cut[(size_t)r], nbytes);
         if (q == (size_t)-1 || q < cut[(size_t)r - 1]) { set_err("SQB_FASTQ: no record boundary near byte %zu", cut[(size_t)r]); return -1; }
         cut[(size_t)r] = q;
      }
   }
   std::vector<sqb_stats_t> st((size_t)nd);
   std::vector<int> rc((size_t)nd, 0);
   std::vector<std::string> errs((size_t)nd);
   auto work = [&](int r) {
      sqb_engine *er = eng[(size_t)r];
      rc[(size_t)r] = scan_chunks(&er, 1, text + cut[(size_t)r], cut[(size_t)r + 1] - cut[(size_t)r], options, &st[(size_t)r],
                                  false, nullptr);
      if (rc[(size_t)r]) errs[(size_t)r] = g_err;
   };
   {
      std::vector<std::thread> th;
      for (int r = 1; r < nd; r++) th.emplace_back(work, r);
      work(0);
      for (auto &t : th) t.join();
   }
   CU(cudaSetDevice(e->device));
   for (int r = 0; r < nd; r++)
      if (rc[(size_t)r]) { set_err("device %d: %s", eng[(size_t)r]->device, errs[(size_t)r].c_str()); return -1; }
   // ---- merge: line bases, records in shard order, line offsets -------------------------
   std::vector<uint64_t> line_base((size_t)nd, 0), rec_off((size_t)nd, 0);
   uint64_t lines = 0, recs = 0;
   for (int r = 0; r < nd; r++) {
      line_base[(size_t)r] = lines;
      rec_off[(size_t)r] = recs;
      lines += st[(size_t)r].nlines;
      recs += (options & SQB_COUNT_ONLY) ? 0 : eng[(size_t)r]->host_recs_n;
   }
   if (!(options & SQB_COUNT_ONLY) && recs > e->host_recs_n) {
      if (recs > e->host_recs_cap) {
         sqb_rec_t *bigger = nullptr;
         const size_t cap = (size_t)recs + (size_t)recs / 4 + (1u << 16);
         CU(pinned_alloc((void **)&bigger, cap * sizeof(sqb_rec_t)));
         if (e->host_recs_n) memcpy(bigger, e->host_recs, e->host_recs_n * sizeof(sqb_rec_t));
         if (e->host_recs) CU(cudaFreeHost(e->host_recs));
         e->host_recs = bigger;
         e->host_recs_cap = cap;
      }
      // every shard's records are copied (line numbers rebased) by a thread of their own
      auto copy = [&](int r) {
         const sqb_engine *er = eng[(size_t)r];
         sqb_rec_t *dst = e->host_recs + rec_off[(size_t)r];
         const uint32_t base = (uint32_t)line_base[(size_t)r];
         for (size_t i = 0; i < er->host_recs_n; i++) {
            sqb_rec_t x = er->host_recs[i];
            x.line += base;
            dst[i] = x;
         }
      };
      std::vector<std::thread> th;
      for (int r = 2; r < nd; r++) th.emplace_back(copy, r);
      copy(1);
      for (auto &t : th) t.join();
      e->host_recs_n = (size_t)recs;
   }
   if (options & SQB_KEEP_LINES_INTERNAL) {
      for (int r = 1; r < nd; r++) {
         const sqb_engine *er = eng[(size_t)r];
         const size_t old = e->host_lines.size();
         e->host_lines.resize(old + er->host_lines.size());
         for (size_t i = 0; i < er->host_lines.size(); i++) e->host_lines[old + i] = er->host_lines[i] + cut[(size_t)r];
      }
   }
   sqb_stats_t acc;
   memset(&acc, 0, sizeof acc);
   for (int r = 0; r < nd; r++) {
      const sqb_stats_t &x = st[(size_t)r];
      acc.nlines += x.nlines;
      acc.nmatched += x.nmatched;
      acc.nrecs += x.nrecs;
      acc.launches += x.launches;
      acc.reruns += x.reruns;
      acc.device_ms = std::max(acc.device_ms, x.device_ms);
      for (int k = 0; k < 8; k++) acc.kernel_ms[k] = std::max(acc.kernel_ms[k], x.kernel_ms[k]);
      acc.path |= x.path;
   }
   acc.nbytes = nbytes;
   acc.devices = (uint32_t)nd;
   e->last_stats = acc;
   if (stats) *stats = acc;
   return 0;
}

int sqbScanHost(sqb_engine_t *e, const char *text, size_t nbytes, int options, sqb_stats_t *stats)
{
   e->scan_generation++;
   const int nd = host_devices(e, nbytes, options);
   if (nd > 1) return scan_host_multi(e, nd, text, nbytes, options, stats);
   return scan_chunks(&e, 1, text, nbytes, options, stats, false, nullptr);
}

unsigned long long sqbScanGeneration(sqb_engine_t *e) { return e->scan_generation; }

int sqbScanDeviceLarge(sqb_engine_t *e, const void *d_text, size_t nbytes, int options, void *stream,
                       sqb_stats_t *stats)
{
   if (options & SQB_SINGLE_LINE) { set_err("sqbScanDeviceLarge: not for single lines"); return -1; }
   return scan_chunks(&e, 1, (const char *)d_text, nbytes, options, stats, true, (cudaStream_t)stream);
}

// ---- pattern sets ---------------------------------------------------------------
struct sqb_multi {
   std::vector<sqb_engine *> engs;     // engs[0] is the leader (its K1 / pack output is shared)
   std::vector<int> order;             // order[i] = caller's index of engs[i]
   std::vector<sqb_stats_t> tmp;
};

sqb_multi_t *sqbMultiNew(int npatterns, const unsigned char *const *keys, const int *m, const int *tau, int device)
{
   if (npatterns < 1 || keys == NULL || m == NULL || tau == NULL) { set_err("sqbMultiNew: invalid arguments"); return NULL; }
   sqb_multi *mp = new sqb_multi();
   int fk = 8;
   for (int i = 0; i < npatterns; i++) {
      sqb_engine *e = sqbEngineNew(keys[i], m[i], tau[i], device);
      if (e == NULL) { sqbMultiFree(mp); return NULL; }
      e->cuts = 0;                     // segment cuts depend on the pattern (warm-up): long lines run un-cut
      e->fused = 0;                    // the followers read the leader's planes[tile][column][group]
      fk = std::min(fk, e->filter_k);
      mp->engs.push_back(e);
      mp->order.push_back(i);
   }
   // the leader must produce bit-planes if anyone is to read them
   for (size_t i = 0; i < mp->engs.size(); i++)
      if (mp->engs[i]->bs_ok) { std::swap(mp->engs[0], mp->engs[i]); std::swap(mp->order[0], mp->order[i]); break; }
   mp->engs[0]->filter_k = fk;         // dead for the shortest pattern of the set = dead for all of them
   mp->tmp.resize(mp->engs.size());
   return mp;
}

void sqbMultiFree(sqb_multi_t *mp)
{
   if (mp == NULL) return;
   for (sqb_engine *e : mp->engs) sqbEngineFree(e);
   delete mp;
}

int sqbMultiCount(sqb_multi_t *mp) { return (int)mp->engs.size(); }

sqb_engine_t *sqbMultiEngine(sqb_multi_t *mp, int pattern)
{
   for (size_t i = 0; i < mp->engs.size(); i++) if (mp->order[i] == pattern) return mp->engs[i];
   set_err("sqbMultiEngine: pattern %d out of range", pattern);
   return NULL;
}

static int multi_scan(sqb_multi *mp, const char *text, size_t nbytes, int options, sqb_stats_t *stats, bool on_device,
                      cudaStream_t st)
{
   if (options & SQB_SINGLE_LINE) { set_err("pattern sets are for buffers of lines"); return -1; }
   const int P = (int)mp->engs.size();
   if (scan_chunks(mp->engs.data(), P, text, nbytes, options, mp->tmp.data(), on_device, st)) return -1;
   if (stats) for (int i = 0; i < P; i++) stats[mp->order[(size_t)i]] = mp->tmp[(size_t)i];
   return 0;
}

int sqbMultiScanHost(sqb_multi_t *mp, const char *text, size_t nbytes, int options, sqb_stats_t *stats)
{
   return multi_scan(mp, text, nbytes, options, stats, false, nullptr);
}

int sqbMultiScanDevice(sqb_multi_t *mp, const void *d_text, size_t nbytes, int options, void *stream, sqb_stats_t *stats)
{
   return multi_scan(mp, (const char *)d_text, nbytes, options, stats, true, (cudaStream_t)stream);
}

const sqb_rec_t *sqbDeviceRecordsAll(sqb_engine_t *e, uint64_t *count)
{
   if (count) *count = e->d_all_n;
   return (const sqb_rec_t *)e->d_all_recs;
}

const sqb_rec_t *sqbHostRecords(sqb_engine_t *e, uint64_t *count)
{
   if (count) *count = e->host_recs_n;
   return e->host_recs;
}

int sqbHostLineStarts(sqb_engine_t *e, const uint64_t **starts, uint64_t *count)
{
   if (starts) *starts = e->host_lines.data();
   if (count) *count = e->host_lines.size();
   return 0;
}

void *sqbHostAlloc(size_t nbytes)
{
   void *p = NULL;
   if (pinned_alloc(&p, nbytes ? nbytes : 1) != cudaSuccess) {
      set_err("cudaMallocHost(%zu) failed", nbytes);
      cudaGetLastError();
      return NULL;
   }
   return p;
}
void sqbHostFree(void *p) { if (p) cudaFreeHost(p); }

void *sqbDeviceAlloc(size_t nbytes)
{
   void *p = NULL;
   if (cudaMalloc(&p, nbytes + 64) != cudaSuccess) {
      set_err("cudaMalloc(%zu) failed", nbytes);
      cudaGetLastError();
      return NULL;
   }
   return p;
}
void sqbDeviceFree(void *p) { if (p) cudaFree(p); }

int sqbMemcpyH2D(void *dst, const void *src, size_t nbytes)
{
   CU(cudaMemcpy(dst, src, nbytes, cudaMemcpyHostToDevice));
   return 0;
}

int sqbMemcpyD2H(void *dst, const void *src, size_t nbytes)
{
   CU(cudaMemcpy(dst, src, nbytes, cudaMemcpyDeviceToHost));
   return 0;
}

void sqbShardRange(const char *text, size_t nbytes, int rank, int world, size_t *begin, size_t *end)
{
   size_t cut[2];
   for (int k = 0; k < 2; k++) {
      const int r = rank + k;
      size_t p = r >= world ? nbytes : (size_t)((unsigned __int128)nbytes * (unsigned)r / (unsigned)world);
      if (r > 0 && r < world && p > 0) {
         // move forward to just after the next newline (a boundary that already
         // follows a newline stays)
         if (text[p - 1] != '\n') {
            const char *nl = (const char *)memchr(text + p, '\n', nbytes - p);
            p = nl ? (size_t)(nl - text) + 1 : nbytes;
         }
      }
      cut[k] = p;
   }
   *begin = cut[0];
   *end = cut[1];
}

// ---- synthetic reads --------------------------------------------------------
size_t sqbGenBytes(const sqb_gen_t *g, uint64_t first, uint64_t nreads)
{
   (void)first;
   return sqb_gen_record_bytes(g) * (size_t)nreads;
}

int sqbGenHost(const sqb_gen_t *g, uint64_t first, uint64_t nreads, char *dst)
{
   const size_t rb = sqb_gen_record_bytes(g);
#pragma omp parallel for schedule(static)
   for (long long r = 0; r < (long long)nreads; r++) sqb_gen_record(g, first + (uint64_t)r, dst + (size_t)r * rb);
   return 0;
}

}  // extern "C"

__global__ void k_gen_reads(const __grid_constant__ sqb_gen_t g, uint64_t first, uint64_t nreads, char *dst)
{
   const size_t rb = sqb_gen_record_bytes(&g);
   for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nreads; r += (uint64_t)gridDim.x * blockDim.x)
      sqb_gen_record(&g, first + r, dst + (size_t)r * rb);
}

extern "C" int sqbGenDevice(const sqb_gen_t *g, uint64_t first, uint64_t nreads, void *d_dst, void *stream)
{
   if (nreads == 0) return 0;
   const int grid = (int)std::min<uint64_t>((nreads + 127) / 128, 148 * 16);
   k_gen_reads<<<grid, 128, 0, (cudaStream_t)stream>>>(*g, first, nreads, (char *)d_dst);
   CU(cudaGetLastError());
   CU(cudaStreamSynchronize((cudaStream_t)stream));
   return 0;
}
