// sqb_bgzf.cu -- BGZF (bgzip) read sets inflated on the device, in front of the scan.
//
// The reference opens plain text (seeq.c:201-256) and reads it line by line (seeq.c:361); a compressed read set
// has to be inflated by another process first.  Here the COMPRESSED bytes cross the PCIe link (DNA text deflates to
// 0.25-0.3 of its size), one warp per BGZF member inflates them into one HBM text buffer (k0_inflate_bgzf), and the
// scan kernels (K12 / K1, matcher, finish) run over that buffer where it lies: sqbScanHostBgzf is sqbScanHost for
// a .gz buffer.  Everything a lane computes is in sqb_inflate.h and is pinned on the CPU against zlib
// (tests/test_inflate_host.py); this file adds the warp: shuffles, table entries 32 apart, warp-wide match copies.
//
// Only the public C-ABI of the engine is used from here (sqbScanDeviceLarge): the scan path does not change.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "seeq_b200.h"
#include "sqb_inflate.h"
#include "sqb_bgzf_warp.h"

using namespace sqb;

extern "C" void sqb_set_error(const char *msg);          // sqb_engine.cu: the text behind sqbLastError()
extern "C" int sqbEngineDevice(sqb_engine_t *e);

namespace {

constexpr uint32_t kFull = 0xffffffffu;

// Two members per warp, one per half-warp.  Lanes 0 and 16 decode symbols (sqb_inflate.h: run_symbols) side by side:
// with one decoding lane per warp the kernel is bound by the issue slots of the SM -- a warp instruction costs its
// slot whatever the number of active lanes -- and two lanes that run the same loop share theirs.  Literals are stored
// as they are decoded; matches are queued, 16 per member, and copied by the sixteen lanes of the half -- one lane per
// short match, all sixteen on a long one, in as few rounds as their dependencies allow (match_ready).  The halves
// also build their tables (entries 16 apart) and copy stored blocks.  The warp runs as ONE sequence of phases (block
// header, tables, symbols, queue): every shuffle, vote and barrier is warp-wide, a half with nothing to do in a
// phase sits it out.  Text bytes are written straight to HBM: the 126 MB L2 merges the byte stores of the ~4700
// members in flight into full sectors before they reach DRAM.
constexpr int kPairWarps = 4;                            // 8 members per CTA: 8 x 3.9 KB of tables and queue, 72 registers:
constexpr int kPairCtas = 7;                             // 7 CTAs = 56 members per SM (the kernel is bound by the latency
                                                         // of the decoding lanes' chains: the more of them, the better)

struct DeviceWarp {                                      // the warp's primitives as sqb_bgzf_warp.h names them
   __device__ __forceinline__ uint32_t lane() const { return threadIdx.x & 31u; }
   template <class T> __device__ __forceinline__ T shfl(T v, uint32_t src) const { return __shfl_sync(kFull, v, src); }
   __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(kFull, p); }
   __device__ __forceinline__ bool any(bool p) const { return __any_sync(kFull, p) != 0; }
   __device__ __forceinline__ void sync() const { __syncwarp(); }
   __device__ __forceinline__ void atomic_min(unsigned long long *p, unsigned long long v) const { atomicMin(p, v); }
};

__global__ void __launch_bounds__(kPairWarps * 32, kPairCtas)
k0_inflate_bgzf_pair(const uint8_t *__restrict__ gz, const inf::Member *__restrict__ members, uint32_t first,
                     uint32_t count, uint8_t *text, uint32_t *__restrict__ status,
                     unsigned long long *__restrict__ first_error)
{
   __shared__ inf::Tables s_tables[2 * kPairWarps];
   __shared__ inf::MatchQueue s_queue[2 * kPairWarps];
   const uint32_t warp = threadIdx.x >> 5;
   const uint32_t pair = blockIdx.x * kPairWarps + warp;
   if (pair * 2u >= count) return;                       // the whole warp
   DeviceWarp w;
   inf::inflate_pair(w, gz, members, first, count, pair, s_tables + 2 * warp, s_queue + 2 * warp, text, status, first_error);
}

// One member per warp: lane 0 decodes, the warp resolves queues of 32 matches.  The first form of the kernel, kept
// as the plain statement of the scheme and for A/B runs (SEEQ_B200_BGZF_KERNEL=single); tests run both.
constexpr int kWarps = 8;                                // members per CTA; 4 CTAs = 32 members per SM

__global__ void __launch_bounds__(kWarps * 32, 4)
k0_inflate_bgzf(const uint8_t *__restrict__ gz, const inf::Member *__restrict__ members, uint32_t first,
                uint32_t count, uint8_t *text, uint32_t *__restrict__ status,
                unsigned long long *__restrict__ first_error)
{
   __shared__ inf::Tables s_tables[kWarps];
   __shared__ inf::MatchQueue s_queue[kWarps];
   const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
   const uint32_t idx = blockIdx.x * kWarps + warp;
   if (idx >= count) return;
   const inf::Member mb = members[first + idx];
   inf::Tables &t = s_tables[warp];
   inf::MatchQueue &q = s_queue[warp];
   uint8_t *out = text + mb.out_off;
   const uint32_t oend = mb.isize;

   inf::BitReader br;
   br.init(gz + mb.in_off, mb.in_len);                  // every lane holds a reader; only lane 0's advances
   uint32_t pos = 0, err = inf::OK, final_block = 0;

   while (!final_block && err == inf::OK) {
      uint32_t type = 0;
      if (lane == 0) err = inf::read_block_header(br, t, &type, &final_block);
      err = __shfl_sync(kFull, err, 0);
      type = __shfl_sync(kFull, type, 0);
      final_block = __shfl_sync(kFull, final_block, 0);
      if (err != inf::OK) break;

      if (type == 0) {                                   // stored: LEN, ~LEN, bytes
         uint32_t len = 0;
         unsigned long long src = 0;
         if (lane == 0) {
            br.align_byte();
            const bool over = br.refill();
            len = br.take(16);
            const uint32_t nlen = br.take(16);
            const uint8_t *sp = br.byte_ptr();
            if (over || len != (~nlen & 0xffffu)) err = inf::ERR_HEADER;
            else if (sp + len > br.src_end()) err = inf::ERR_INPUT;
            else if (len > oend - pos) err = inf::ERR_OUTPUT;
            src = (unsigned long long)(uintptr_t)sp;
         }
         err = __shfl_sync(kFull, err, 0);
         if (err != inf::OK) break;
         len = __shfl_sync(kFull, len, 0);
         src = __shfl_sync(kFull, src, 0);
         const uint8_t *sp = (const uint8_t *)(uintptr_t)src;
         for (uint32_t j = lane; j < len; j += 32) out[pos + j] = sp[j];
         pos += len;
         if (lane == 0) br.init(sp + len, (uint32_t)(br.src_end() - (sp + len)));
         __syncwarp();
         continue;
      }

      __syncwarp();                                      // counts and sorted symbols are lane 0's: publish
      // the scratch arrays alias t.lit: every lane has passed the barrier, nobody reads them any more
      for (uint32_t e = lane; e < inf::kLitN; e += 32) t.lit[e] = inf::make_lit_entry(t.lcnt, t.lsym, e);
      for (uint32_t e = lane; e < inf::kDistN; e += 32) t.dist[e] = inf::make_dist_entry(t.dcnt, t.dsym, e);
      __syncwarp();

      for (;;) {
         uint32_t nq = 0;
         int r = inf::R_EOB;
         if (lane == 0) r = inf::run_symbols(br, t, out, pos, oend, q, inf::kQueue, &nq);
         r = __shfl_sync(kFull, r, 0);
         nq = __shfl_sync(kFull, nq, 0);
         __syncwarp();                                   // lane 0's literal stores and queue entries: visible
         // resolve the queue (sqb_inflate.h: match_ready): lane i owns match i
         uint32_t mp = 0, ml = 0, md = 0;
         if (lane < nq) { const inf::MatchQueue::Entry qe = q.e[lane]; mp = qe.pos; ml = qe.ld & 0xffffu; md = qe.ld >> 16; }
         uint32_t pending = __ballot_sync(kFull, lane < nq);
         while (pending) {
            const int f = __ffs((int)pending) - 1;
            const uint32_t P = __shfl_sync(kFull, mp, f);
            const bool ready = ((pending >> lane) & 1u) && ((int)lane == f || inf::match_ready(mp, ml, md, P));
            const bool mine = ready && inf::match_by_lane(ml, md);
            const uint32_t rmask = __ballot_sync(kFull, ready);
            uint32_t wide = __ballot_sync(kFull, ready && !mine);
            if (mine) inf::copy_by_lane(out, mp, ml, md);
            while (wide) {                               // long or self-overlapping: the warp copies it
               const int i = __ffs((int)wide) - 1;
               wide &= wide - 1u;
               const uint32_t bp = __shfl_sync(kFull, mp, i), bl = __shfl_sync(kFull, ml, i), bd = __shfl_sync(kFull, md, i);
               if (bd >= bl) for (uint32_t j = lane; j < bl; j += 32) out[bp + j] = out[bp - bd + j];
               else for (uint32_t j = lane; j < bl; j += 32) out[bp + j] = out[inf::match_src(bp, bd, j)];
            }
            pending &= ~rmask;
            __syncwarp();                                // this round's text is final for the next round
         }
         if (r >= inf::R_ERR) err = (uint32_t)(r - inf::R_ERR);
         if (r != inf::R_FULL) break;
      }
      if (err != inf::OK) break;
      pos = __shfl_sync(kFull, pos, 0);
      __syncwarp();                                      // the tables are rewritten by the next header
   }

   if (lane == 0) {
      if (err == inf::OK && br.overrun() > 0) err = inf::ERR_INPUT;
      if (err == inf::OK && pos != oend) err = inf::ERR_SHORT;
      status[first + idx] = err;
      if (err != inf::OK) atomicMin(first_error, ((unsigned long long)(first + idx) << 8) | err);
   }
}

// ---- per-device state: the compressed bytes, the text and the member list in HBM ----------------------------------
constexpr int kLanes = 32;
struct BgzfState {
   std::mutex mu;
   uint8_t *d_gz = nullptr;        size_t gz_cap = 0;
   uint8_t *d_text = nullptr;      size_t text_cap = 0;
   inf::Member *d_members = nullptr; size_t members_cap = 0;
   inf::Member *h_members = nullptr; size_t h_members_cap = 0;     // pinned: read by the kernel where it lies
   uint32_t *d_status = nullptr;   size_t status_cap = 0;
   unsigned long long *d_first = nullptr, *h_first = nullptr;
   cudaStream_t copy = nullptr, work = nullptr;
   cudaStream_t lanes[kLanes] = {};                     // the slices of one buffer are inflated side by side
   cudaEvent_t lane_done[kLanes] = {};
   std::vector<cudaEvent_t> ev;
   cudaEvent_t t0 = nullptr, t1 = nullptr, begin = nullptr;
   uint64_t text_bytes = 0;        // of the last sqbScanHostBgzf
   bool ready = false;
};
BgzfState g_state[64];

void fail(const char *fmt, ...)
{
   char buf[480];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof buf, fmt, ap);
   va_end(ap);
   sqb_set_error(buf);
}

#define CUB(call)                                                                                             \
   do {                                                                                                       \
      cudaError_t e_ = (call);                                                                                \
      if (e_ != cudaSuccess) {                                                                                \
         fail("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call);              \
         return -1;                                                                                           \
      }                                                                                                       \
   } while (0)

template <class T> int grow(T **p, size_t *cap, size_t need)
{
   if (*cap >= need) return 0;
   if (*p) cudaFree(*p);
   *p = nullptr;
   *cap = 0;
   const size_t want = need + need / 8 + 256;
   CUB(cudaMalloc((void **)p, want * sizeof(T)));
   *cap = want;
   return 0;
}

int state_init(BgzfState &s)
{
   if (s.ready) return 0;
   CUB(cudaStreamCreateWithFlags(&s.copy, cudaStreamNonBlocking));
   CUB(cudaStreamCreateWithFlags(&s.work, cudaStreamNonBlocking));
   CUB(cudaMalloc((void **)&s.d_first, sizeof(unsigned long long)));
   CUB(cudaMallocHost((void **)&s.h_first, sizeof(unsigned long long)));
   CUB(cudaEventCreate(&s.t0));
   CUB(cudaEventCreate(&s.t1));
   CUB(cudaEventCreateWithFlags(&s.begin, cudaEventDisableTiming));
   for (int i = 0; i < kLanes; i++) {
      CUB(cudaStreamCreateWithFlags(&s.lanes[i], cudaStreamNonBlocking));
      CUB(cudaEventCreateWithFlags(&s.lane_done[i], cudaEventDisableTiming));
   }
   s.ready = true;
   return 0;
}

const char *err_text(uint32_t code)
{
   switch (code) {
   case inf::ERR_INPUT: return "the deflate stream runs past the member's data";
   case inf::ERR_OUTPUT: return "more text than ISIZE announces";
   case inf::ERR_CODE: return "invalid code";
   case inf::ERR_DIST: return "distance reaches in front of the member";
   case inf::ERR_HEADER: return "invalid block header";
   case inf::ERR_SHORT: return "less text than ISIZE announces";
   }
   return "?";
}

// members [first, first + count) on stream st
void launch_inflate(BgzfState &s, const uint8_t *d_gz, const inf::Member *members, uint32_t first, uint32_t count,
                    uint8_t *d_text, cudaStream_t st)
{
   if (count == 0) return;
   // 7 CTAs of 32.5 KB need the whole shared-memory carve-out of the SM (per device and kernel; a hint, cheap)
   cudaFuncSetAttribute(k0_inflate_bgzf_pair, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
   cudaFuncSetAttribute(k0_inflate_bgzf, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
   const char *env = getenv("SEEQ_B200_BGZF_KERNEL");
   if (env && strcmp(env, "single") == 0) {
      const uint32_t grid = (count + kWarps - 1) / kWarps;
      k0_inflate_bgzf<<<grid, kWarps * 32, 0, st>>>(d_gz, members, first, count, d_text, s.d_status, s.d_first);
   } else {
      const uint32_t per = 2 * kPairWarps;
      const uint32_t grid = (count + per - 1) / per;
      k0_inflate_bgzf_pair<<<grid, kPairWarps * 32, 0, st>>>(d_gz, members, first, count, d_text, s.d_status, s.d_first);
   }
}

// The members of a buffer, one after the other: each header says where the next one starts, so the walk is a chain
// of dependent cache misses (110 ns each: 1.9 ms for the 16 k members of a 330 MB buffer, with the device idle).
int index_serial(const uint8_t *gz, size_t nbytes, std::vector<inf::Member> &out, uint64_t *text_bytes)
{
   uint64_t off = 0, o = 0;
   out.clear();
   while (off < nbytes) {
      inf::Member m;
      const uint64_t next = inf::parse_member(gz, nbytes, off, &m);
      if (next == 0) {
         fail("not a BGZF member at byte %llu of %llu (plain gzip has no block index: recompress with bgzip)",
              (unsigned long long)off, (unsigned long long)nbytes);
         return -1;
      }
      if (m.isize > 65536u) { fail("BGZF member at byte %llu announces %u bytes of text (> 64 KiB)", (unsigned long long)off, m.isize); return -1; }
      m.out_off = o;
      if (m.isize) out.push_back(m);                   // the empty end-of-file member carries nothing
      o += m.isize;
      off = next;
   }
   *text_bytes = o;
   return 0;
}

// The same walk by several threads: thread t starts at the first byte sequence at or behind t/T of the buffer that
// parses as a member followed by another member (or by the end of the buffer) and walks to the share of thread t+1.
// A share is accepted when the walk of the share in front of it ends exactly where it started: by induction from
// byte 0 its start is a member boundary then.  Anything else -- a look-alike header inside deflate data, damage --
// goes back to the serial walk, which also words the error.
struct IndexShare {
   uint64_t first = 0, end = 0;                         // offset of the first member, offset the walk stopped at
   bool ok = false;
   std::vector<inf::Member> members;                    // out_off not yet set
};

void index_share(const uint8_t *gz, size_t nbytes, uint64_t from, uint64_t limit, bool search, IndexShare *sh)
{
   uint64_t off = from;
   inf::Member m;
   if (search) {
      for (;; off++) {
         if (off + 18 > nbytes) return;                 // no member starts in this share: not ok
         if (gz[off] != 0x1f || gz[off + 1] != 0x8b || gz[off + 2] != 8 || !(gz[off + 3] & 4)) continue;
         const uint64_t next = inf::parse_member(gz, nbytes, off, &m);
         if (next == 0) continue;
         inf::Member m2;
         if (next == nbytes || inf::parse_member(gz, nbytes, next, &m2) != 0) break;
      }
   }
   sh->first = off;
   while (off < limit && off < nbytes) {
      const uint64_t next = inf::parse_member(gz, nbytes, off, &m);
      if (next == 0 || m.isize > 65536u) return;
      if (m.isize) sh->members.push_back(m);
      off = next;
   }
   sh->end = off;
   sh->ok = true;
}

int index_members(const uint8_t *gz, size_t nbytes, std::vector<inf::Member> &out, uint64_t *text_bytes)
{
   static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
   const char *env = getenv("SEEQ_B200_BGZF_INDEX_THREADS");
   const unsigned want = env ? (unsigned)std::max(1, atoi(env)) : std::min(8u, hw);
   const unsigned T = (unsigned)std::min<uint64_t>(want, nbytes >> 22);          // 4 MiB or more per thread
   if (T < 2) return index_serial(gz, nbytes, out, text_bytes);
   std::vector<IndexShare> sh(T);
   std::vector<std::thread> th;
   for (unsigned t = 1; t < T; t++)
      th.emplace_back(index_share, gz, nbytes, (uint64_t)nbytes * t / T, t + 1 < T ? (uint64_t)nbytes * (t + 1) / T : (uint64_t)nbytes,
                      true, &sh[t]);
   index_share(gz, nbytes, 0, (uint64_t)nbytes / T, false, &sh[0]);
   for (std::thread &x : th) x.join();
   bool ok = sh[0].ok;
   for (unsigned t = 1; t < T && ok; t++) ok = sh[t].ok && sh[t].first == sh[t - 1].end;
   ok = ok && sh[T - 1].end == nbytes;
   if (!ok) return index_serial(gz, nbytes, out, text_bytes);
   size_t n = 0;
   for (const IndexShare &x : sh) n += x.members.size();
   out.clear();
   out.reserve(n);
   uint64_t o = 0;
   for (const IndexShare &x : sh)
      for (inf::Member m : x.members) {
         m.out_off = o;
         o += m.isize;
         out.push_back(m);
      }
   *text_bytes = o;
   return 0;
}

}  // namespace

extern "C" {

int sqbBgzfIndex(const void *gz, size_t nbytes, sqb_bgzf_member_t *members, uint64_t cap, uint64_t *count,
                 uint64_t *text_bytes)
{
   static_assert(sizeof(sqb_bgzf_member_t) == sizeof(inf::Member), "member layout");
   std::vector<inf::Member> v;
   uint64_t tb = 0;
   if (gz == NULL && nbytes) { fail("sqbBgzfIndex: no buffer"); return -1; }
   if (index_members((const uint8_t *)gz, nbytes, v, &tb)) return -1;
   if (count) *count = v.size();
   if (text_bytes) *text_bytes = tb;
   if (members) {
      if (cap < v.size()) { fail("sqbBgzfIndex: room for %llu members, %llu found", (unsigned long long)cap, (unsigned long long)v.size()); return -1; }
      if (!v.empty()) memcpy(members, v.data(), v.size() * sizeof(inf::Member));
   }
   return 0;
}

int sqbBgzfInflateDevice(int device, const void *d_gz, const sqb_bgzf_member_t *members, uint64_t count,
                         void *d_text, void *stream, double *kernel_ms)
{
   if (device < 0 || device >= 64) { fail("sqbBgzfInflateDevice: device %d", device); return -1; }
   if (count > 0xffffffffull) { fail("sqbBgzfInflateDevice: too many members"); return -1; }
   if (count == 0) { if (kernel_ms) *kernel_ms = 0; return 0; }
   if (d_gz == NULL || members == NULL || d_text == NULL) { fail("sqbBgzfInflateDevice: invalid arguments"); return -1; }
   CUB(cudaSetDevice(device));
   BgzfState &s = g_state[device];
   std::lock_guard<std::mutex> lock(s.mu);
   if (state_init(s)) return -1;
   cudaStream_t st = stream ? (cudaStream_t)stream : s.work;
   if (grow(&s.d_members, &s.members_cap, (size_t)count)) return -1;
   if (grow(&s.d_status, &s.status_cap, (size_t)count)) return -1;
   *s.h_first = ~0ull;
   CUB(cudaMemcpyAsync(s.d_first, s.h_first, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
   CUB(cudaMemcpyAsync(s.d_members, members, (size_t)count * sizeof(inf::Member), cudaMemcpyHostToDevice, st));
   CUB(cudaEventRecord(s.t0, st));
   launch_inflate(s, (const uint8_t *)d_gz, s.d_members, 0, (uint32_t)count, (uint8_t *)d_text, st);
   CUB(cudaGetLastError());
   CUB(cudaEventRecord(s.t1, st));
   CUB(cudaMemcpyAsync(s.h_first, s.d_first, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
   CUB(cudaStreamSynchronize(st));
   if (kernel_ms) {
      float ms = 0;
      CUB(cudaEventElapsedTime(&ms, s.t0, s.t1));
      *kernel_ms = ms;
   }
   if (*s.h_first != ~0ull) {
      const unsigned long long w = *s.h_first;
      fail("BGZF member %llu (text offset %llu): %s", w >> 8, (unsigned long long)members[w >> 8].out_off, err_text((uint32_t)(w & 0xff)));
      return -1;
   }
   return 0;
}

// sqbScanHost for a BGZF buffer: the compressed bytes go to the device in slices, the members of slice k are
// inflated while slice k+1 is on the link, the scan runs over the inflated text in HBM.
int sqbScanHostBgzf(sqb_engine_t *e, const void *gz, size_t nbytes, int options, sqb_stats_t *stats)
{
   if (e == NULL || (gz == NULL && nbytes)) { fail("sqbScanHostBgzf: invalid arguments"); return -1; }
   const int device = sqbEngineDevice(e);
   if (device < 0 || device >= 64) { fail("sqbScanHostBgzf: device %d", device); return -1; }
   // SEEQ_B200_BGZF_TRACE=1: host wall time of the phases of one call on stderr
   const bool trace = getenv("SEEQ_B200_BGZF_TRACE") != NULL;
   auto now = [] { return std::chrono::steady_clock::now(); };
   auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
   const auto t_start = now();
   CUB(cudaSetDevice(device));
   BgzfState &s = g_state[device];
   std::lock_guard<std::mutex> lock(s.mu);
   if (state_init(s)) return -1;
   if (grow(&s.d_gz, &s.gz_cap, nbytes + 64)) return -1;

   // 1. The compressed bytes leave for the device at once, in slices of a fixed size (a copy needs no index).
   const char *env = getenv("SEEQ_B200_BGZF_SLICE_MB");
   const long mb = env ? atol(env) : 32;                // measured on 330 MB: 67 / 70 / 73 GB/s of text with 8 / 16 / 32 MiB
   const size_t slice_bytes = (size_t)(mb > 0 ? mb : 32) << 20;
   const size_t nslices = (nbytes + slice_bytes - 1) / slice_bytes;
   while (s.ev.size() < nslices) {
      cudaEvent_t ev;
      CUB(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      s.ev.push_back(ev);
   }
   for (size_t k = 0; k < nslices; k++) {
      const size_t b0 = k * slice_bytes, b1 = std::min(nbytes, b0 + slice_bytes);
      CUB(cudaMemcpyAsync(s.d_gz + b0, (const uint8_t *)gz + b0, b1 - b0, cudaMemcpyHostToDevice, s.copy));
      CUB(cudaEventRecord(s.ev[k], s.copy));
   }
   // 2. The members are found while the bytes travel (1.9 ms of dependent cache misses for 16 k members).  From here
   // on nobody returns with copies out of the caller's buffer in flight.
   std::vector<inf::Member> mem;
   uint64_t text_bytes = 0;
   int bad = index_members((const uint8_t *)gz, nbytes, mem, &text_bytes);
   if (!bad && mem.size() > 0xffffffffull) { fail("sqbScanHostBgzf: too many members"); bad = 1; }
   const double ms_index = ms_since(t_start);
   if (!bad) bad = grow(&s.d_text, &s.text_cap, (size_t)text_bytes + 64) || grow(&s.d_status, &s.status_cap, mem.size() + 1);
   if (!bad && s.h_members_cap < mem.size() + 1) {
      if (s.h_members) cudaFreeHost(s.h_members);
      s.h_members = nullptr;
      s.h_members_cap = 0;
      const size_t want = mem.size() + mem.size() / 8 + 256;
      if (cudaMallocHost((void **)&s.h_members, want * sizeof(inf::Member)) != cudaSuccess) { fail("sqbScanHostBgzf: no pinned memory for the member list"); bad = 1; }
      else s.h_members_cap = want;
   }
   if (bad) {
      cudaStreamSynchronize(s.copy);
      return -1;
   }
   // The member list stays in pinned host memory and the kernel reads it there (24 bytes per member, once): a copy
   // of it would queue up behind the slices on the one host-to-device copy engine and hold every kernel back until
   // the last slice has arrived (measured: 18.7 ms instead of 15.9).  The error word is set by a memset for the same
   // reason.
   if (!mem.empty()) memcpy(s.h_members, mem.data(), mem.size() * sizeof(inf::Member));
   CUB(cudaMemsetAsync(s.d_first, 0xff, sizeof(unsigned long long), s.work));
   CUB(cudaEventRecord(s.begin, s.work));
   for (int i = 0; i < kLanes; i++) CUB(cudaStreamWaitEvent(s.lanes[i], s.begin, 0));
   // 3. The members that END in slice k (trailer included; the reader looks up to three bytes further and uses none of
   // them) are inflated on stream k mod 32 once the slice has arrived.  A member is one serial chain of 2-5 ms
   // whatever the size of the grid: a slice must not wait for the slice in front of it, and whatever has arrived by
   // the time the index is there fills the device at once.
   size_t m0 = 0, k = 0;
   for (k = 0; k < nslices; k++) {
      const size_t b1 = std::min(nbytes, (k + 1) * slice_bytes);
      size_t m1 = m0;
      while (m1 < mem.size() && mem[m1].in_off + mem[m1].in_len + 8 <= b1) m1++;
      if (m1 == m0) continue;
      CUB(cudaStreamWaitEvent(s.lanes[k % kLanes], s.ev[k], 0));
      launch_inflate(s, s.d_gz, s.h_members, (uint32_t)m0, (uint32_t)(m1 - m0), s.d_text, s.lanes[k % kLanes]);
      CUB(cudaGetLastError());
      m0 = m1;
   }
   for (int i = 0; i < kLanes; i++) {
      CUB(cudaEventRecord(s.lane_done[i], s.lanes[i]));
      CUB(cudaStreamWaitEvent(s.work, s.lane_done[i], 0));
   }
   CUB(cudaMemcpyAsync(s.h_first, s.d_first, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.work));
   const double ms_queued = ms_since(t_start);
   if (trace) CUB(cudaStreamSynchronize(s.copy));
   const double ms_copied = ms_since(t_start);
   CUB(cudaStreamSynchronize(s.work));
   const double ms_inflated = ms_since(t_start);
   if (*s.h_first != ~0ull) {
      const unsigned long long w = *s.h_first;
      fail("BGZF member %llu (text offset %llu): %s", w >> 8, (unsigned long long)mem[w >> 8].out_off, err_text((uint32_t)(w & 0xff)));
      return -1;
   }
   s.text_bytes = text_bytes;
   const int rc = sqbScanDeviceLarge(e, s.d_text, (size_t)text_bytes, options, NULL, stats);
   if (rc == 0 && stats) stats->launches += (uint32_t)k;
   if (trace)
      fprintf(stderr, "[sqbScanHostBgzf] %zu slices: index %.2f ms, queued %.2f, copied %.2f, inflated %.2f, scanned %.2f\n",
              k, ms_index, ms_queued, ms_copied, ms_inflated, ms_since(t_start));
   return rc;
}

// the inflated text of the last sqbScanHostBgzf on the engine's device (valid until the next BGZF call there)
const void *sqbBgzfDeviceText(sqb_engine_t *e, uint64_t *nbytes)
{
   const int device = sqbEngineDevice(e);
   if (device < 0 || device >= 64) return NULL;
   if (nbytes) *nbytes = g_state[device].text_bytes;
   return g_state[device].d_text;
}

}  // extern "C"
