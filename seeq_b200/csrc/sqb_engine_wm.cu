// sqb_engine_wm.cu -- instantiations of the bit-sliced matcher with the NFA-level
// automaton (tau <= 2, patterns of up to 32 positions; sqb_bitslice.h: bs_wm_step).
// A translation unit of its own so that it compiles next to sqb_engine.cu.
// Compiled twice (seeq_b200/build.py): -DSQB_WM_FUSED=0 (planes of k15_pack) and -DSQB_WM_FUSED=1 (group
// planes of the fused tokenise + pack kernel, sqb_k12_fused.cuh).
#include "sqb_k2_bitslice.cuh"

using namespace sqb;

#ifndef SQB_WM_FUSED
#define SQB_WM_FUSED 0
#endif
#if SQB_WM_FUSED
#define SQB_WM_ENTRY sqb_launch_bitslice_wm_fused
#else
#define SQB_WM_ENTRY sqb_launch_bitslice_wm
#endif

template <int R, int T, int MODE> static cudaError_t launch3(bool skip, int grid, cudaStream_t st, const K2BsArgs &a, const BsPattern &p)
{
   const size_t smem = (MODE == BS_ALL ? sizeof(BsWarpSmemAll) : sizeof(BsWarpSmem)) * kBsWarps;      // < 48 KiB
   if (p.ncustom > 0) {
      if (skip) k2_bitslice<R, 1, MODE, true, T, SQB_WM_FUSED != 0, true><<<grid, kBsThreads, smem, st>>>(a, p);
      else k2_bitslice<R, 1, MODE, false, T, SQB_WM_FUSED != 0, true><<<grid, kBsThreads, smem, st>>>(a, p);
   } else {
      if (skip) k2_bitslice<R, 1, MODE, true, T, SQB_WM_FUSED != 0, false><<<grid, kBsThreads, smem, st>>>(a, p);
      else k2_bitslice<R, 1, MODE, false, T, SQB_WM_FUSED != 0, false><<<grid, kBsThreads, smem, st>>>(a, p);
   }
   return cudaGetLastError();
}

template <int R, int T> static cudaError_t launch2(int bsmode, bool skip, int grid, cudaStream_t st, const K2BsArgs &a, const BsPattern &p)
{
   switch (bsmode) {
   case BS_FIRST: return launch3<R, T, BS_FIRST>(skip, grid, st, a, p);
   case BS_BEST: return launch3<R, T, BS_BEST>(skip, grid, st, a, p);
   default: return launch3<R, T, BS_ALL>(skip, grid, st, a, p);
   }
}

template <int R> static cudaError_t launch1(int levels, int bsmode, bool skip, int grid, cudaStream_t st, const K2BsArgs &a, const BsPattern &p)
{
   switch (levels) {
   case 1: return launch2<R, 1>(bsmode, skip, grid, st, a, p);
   case 2: return launch2<R, 2>(bsmode, skip, grid, st, a, p);
   default: return launch2<R, 3>(bsmode, skip, grid, st, a, p);
   }
}

// rows: R of the pattern's kernel shape (parts == 1); levels = tau + 1 (1..3)
cudaError_t SQB_WM_ENTRY(int rows, int levels, int bsmode, bool skip, int grid, cudaStream_t st,
                         const K2BsArgs &a, const BsPattern &p)
{
   switch (rows) {
   case 8: return launch1<8>(levels, bsmode, skip, grid, st, a, p);
   case 10: return launch1<10>(levels, bsmode, skip, grid, st, a, p);
   case 12: return launch1<12>(levels, bsmode, skip, grid, st, a, p);
   case 16: return launch1<16>(levels, bsmode, skip, grid, st, a, p);
   case 20: return launch1<20>(levels, bsmode, skip, grid, st, a, p);
   case 24: return launch1<24>(levels, bsmode, skip, grid, st, a, p);
   case 32: return launch1<32>(levels, bsmode, skip, grid, st, a, p);
   default: return cudaErrorInvalidValue;
   }
}
