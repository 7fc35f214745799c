/*
 * cs_tma.cuh -- the few lines of sm_90+/sm_100 machinery the tile kernels share: a 2-D byte tensor map made through the runtime's driver
 * entry point (no link-time dependency on libcuda), and one-shot "copy this box into shared memory and tell me when it is there"
 * (cp.async.bulk.tensor.2d + mbarrier) for kernels that stage one tile per CTA.  Out-of-bounds bytes of a box are zero-filled by the
 * copy engine, which is never what the image kernels want at a border (they reflect or replicate): callers use it for interior tiles only.
 */
#ifndef CS_TMA_CUH
#define CS_TMA_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

/* {width_bytes, rows} tensor of bytes with the given pitch, box box_w x box_h bytes.  False when the driver entry point is missing or
 * the layout does not qualify (base and pitch must be multiples of 16 bytes, box_w a multiple of 16 and at most 256). */
static inline bool cs_make_tmap_bytes(CUtensorMap *tm, const void *base, int64_t width_bytes, int64_t rows, int64_t pitch_bytes, int box_w, int box_h)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeFn)p;
        cudaGetLastError();
    }
    memset(tm, 0, sizeof(*tm));
    if (!fn || (pitch_bytes % 16) != 0 || (((uintptr_t)base) & 15) != 0 || (box_w % 16) != 0 || box_w > 256 || box_h > 256 || width_bytes < box_w || rows < box_h)
        return false;
    const cuuint64_t dims[2] = {(cuuint64_t)width_bytes, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t cs_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cs_mbar_init(unsigned long long *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cs_smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
/* one thread: expect `bytes` on the barrier and start the copy of the box at (x, y) into dst (128-byte aligned shared memory).
 * x, the innermost coordinate, must be a multiple of 16 BYTES: measured on B200 (tools/probe/tma_probe.cu), x = 30 on a byte tensor is an
 * illegal-instruction fault, x = 32 is fine -- callers round the start down and read the tile at an offset. */
__device__ __forceinline__ void cs_tma_load_2d(const CUtensorMap *tm, void *dst, unsigned long long *bar, int x, int y, uint32_t bytes)
{
    const uint32_t b = cs_smem_u32(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(cs_smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(b)
                 : "memory");
}
/* every thread: wait for the barrier's phase `parity`; bounded, returns false when the copy never arrived (the caller reports it) */
__device__ __forceinline__ bool cs_mbar_wait(unsigned long long *bar, uint32_t parity)
{
    const uint32_t b = cs_smem_u32(bar);
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; spin++)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b), "r"(parity) : "memory");
    return done != 0;
}
#endif

#endif /* CS_TMA_CUH */
