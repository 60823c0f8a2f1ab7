// sqb_engine_bsf.cu -- instantiations of the single-part Myers bit-sliced matcher that reads the GROUP planes
// of the fused tokenise + pack kernel (sqb_k12_fused.cuh); a translation unit of its own so that it compiles
// next to sqb_engine.cu (which holds the instances that read the planes of k15_pack).
#include "sqb_k2_bitslice.cuh"

using namespace sqb;

// (first call per device: the opt-in above 48 KiB of dynamic shared memory, sqb_engine.cu: first_use)
bool sqb_first_use(const void *fn);

template <int R, int MODE> static cudaError_t launch2(bool skip, int grid, cudaStream_t st, const K2BsArgs &a, const BsPattern &p)
{
   const size_t smem = (MODE == BS_ALL ? sizeof(BsWarpSmemAll) : sizeof(BsWarpSmem)) * kBsWarps;
   if (sqb_first_use((const void *)k2_bitslice<R, 1, MODE, true, 0, true>)) {
      cudaError_t e = cudaFuncSetAttribute(k2_bitslice<R, 1, MODE, true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess)
         e = cudaFuncSetAttribute(k2_bitslice<R, 1, MODE, false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
   }
   if (skip) k2_bitslice<R, 1, MODE, true, 0, true><<<grid, kBsThreads, smem, st>>>(a, p);
   else k2_bitslice<R, 1, MODE, false, 0, true><<<grid, kBsThreads, smem, st>>>(a, p);
   return cudaGetLastError();
}

template <int R> static cudaError_t launch1(int bsmode, bool skip, int grid, cudaStream_t st, const K2BsArgs &a, const BsPattern &p)
{
   switch (bsmode) {
   case BS_FIRST: return launch2<R, BS_FIRST>(skip, grid, st, a, p);
   case BS_BEST: return launch2<R, BS_BEST>(skip, grid, st, a, p);
   default: return launch2<R, BS_ALL>(skip, grid, st, a, p);
   }
}

cudaError_t sqb_launch_bitslice_myers_fused(int rows, int bsmode, bool skip, int grid, cudaStream_t st,
                                            const K2BsArgs &a, const BsPattern &p)
{
   switch (rows) {
   case 8: return launch1<8>(bsmode, skip, grid, st, a, p);
   case 10: return launch1<10>(bsmode, skip, grid, st, a, p);
   case 12: return launch1<12>(bsmode, skip, grid, st, a, p);
   case 16: return launch1<16>(bsmode, skip, grid, st, a, p);
   case 20: return launch1<20>(bsmode, skip, grid, st, a, p);
   case 24: return launch1<24>(bsmode, skip, grid, st, a, p);
   case 32: return launch1<32>(bsmode, skip, grid, st, a, p);
   default: return cudaErrorInvalidValue;
   }
}
