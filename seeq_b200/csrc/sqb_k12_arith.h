// sqb_k12_arith.h -- the pure arithmetic of the fused tokenise + pack kernel (sqb_k12_fused.cuh), in a header that
// compiles for the host as well: tests/host_bitslice.cpp drives it on the CPU, exhaustively where the domain is small
// (tests/test_k12_host.py), so that the bit tricks of K12 are pinned without a GPU.
#ifndef SQB_K12_ARITH_H_
#define SQB_K12_ARITH_H_

#include <stdint.h>

#include "sqb_tables.h"

#ifdef __CUDACC__
#define SQB_K12_HD __host__ __device__ __forceinline__
#else
#define SQB_K12_HD static inline
#endif

namespace sqb {

// byte -> {p0, p1, p2, newline} in the four bytes of a word (bit 0 of each)
struct ClassTable32 {
   uint32_t w[256];
};
static inline void build_class_table32(const ClassTable &ct, ClassTable32 *out)
{
   for (int b = 0; b < 256; b++) {
      const uint32_t c = ct.code[b];
      out->w[b] = (c & 1u) | (((c >> 1) & 1u) << 8) | (((c >> 2) & 1u) << 16) | (((c >> 3) & 1u) << 24);
   }
}

// Newline flags of four text bytes: bit 7 of every byte that is '\n' (exact: no carry crosses a byte).
SQB_K12_HD uint32_t nl_flags(uint32_t w)
{
   const uint32_t t = ((w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;     // bit 7: the low seven bits differ from '\n'
   return ~(t | w) & 0x80808080u;                                         // ... and bit 7 of the byte itself is clear
}

// The flags of a 32-byte chunk live in ONE word ("u-space"): the flag of byte b of text word k (0..7) sits at
// bit u = 8 b + k -- the flags of word k shifted right by 7 - k (on the device one multiply-add, high half, each).
// Mask of the bytes in front of byte index v = 4 k + b (v = 0..32) in TEXT order:
SQB_K12_HD uint32_t chunk_before(uint32_t v)
{
   if (v >= 32u) return ~0u;
   const uint32_t k = v >> 2, b = v & 3u;
   return (((1u << k) - 1u) * 0x01010101u) | ((0x01010101u << k) & ((1u << (8u * b)) - 1u));
}
// byte index (0..31) inside its chunk of the flag at bit u
SQB_K12_HD uint32_t chunk_byte_of(uint32_t u) { return 4u * (u & 7u) + (u >> 3); }

// the flag word of the eight text words of a chunk (reference form of chunk_flags in the kernel)
SQB_K12_HD uint32_t chunk_flags_words(const uint32_t *w)
{
   uint32_t acc = 0;
   for (int k = 0; k < 8; k++) acc += nl_flags(w[k]) >> (7 - k);
   return acc;
}

// PRMT as the kernel uses it: byte i of the result = byte (sel >> 4 i) & 7 of {hi:lo}
SQB_K12_HD uint32_t k12_prmt(uint32_t lo, uint32_t hi, uint32_t sel)
{
#ifdef __CUDA_ARCH__
   return __byte_perm(lo, hi, sel);
#else
   const uint64_t v = ((uint64_t)hi << 32) | lo;
   uint32_t r = 0;
   for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
   return r;
#endif
}

// Four accumulators of one COLUMN (one per line octet q: bytes [- | p2 | p1 | p0], bit i = line 8 q + i) -> the three
// plane words of the column (bit r = line r)
SQB_K12_HD void k12_planes_of_column(const uint32_t *a, uint32_t &p0, uint32_t &p1, uint32_t &p2)
{
   const uint32_t u0 = k12_prmt(a[0], a[1], 0x5140u), u1 = k12_prmt(a[2], a[3], 0x5140u);
   const uint32_t u2 = k12_prmt(a[0], a[1], 0x7362u), u3 = k12_prmt(a[2], a[3], 0x7362u);
   p0 = k12_prmt(u0, u1, 0x5410u);
   p1 = k12_prmt(u0, u1, 0x7632u);
   p2 = k12_prmt(u2, u3, 0x5410u);
}

// plane units (uint4) of a group of `ncols` columns: [block of 32 columns][plane 0..2][32 columns] words
SQB_K12_HD uint32_t k12_group_units(uint32_t ncols) { return ((ncols + 31u) >> 5) * 24u; }

}  // namespace sqb
#endif
