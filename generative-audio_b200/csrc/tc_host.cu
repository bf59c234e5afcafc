// Host side of the TMA plumbing: cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency).
#include "tc_common.cuh"
#include <cudaTypedefs.h>

namespace nppc {
namespace tc {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                      uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
    auto enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return NPPC_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu stride=%llu box=%ux%u", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_bytes, box_rows, box_cols);
        return NPPC_ERR_CUDA;
    }
    return NPPC_OK;
}

int make_tmap_f16_kslabs(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t n_slabs) {
    auto enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return NPPC_ERR_CUDA;
    }
    cuuint64_t gdim[3] = {64, rows, cols / 64};
    cuuint64_t gstride[2] = {cols * 2, 128};
    cuuint32_t box[3] = {64, box_rows, n_slabs};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed (%d) rows=%llu cols=%llu box=%ux%u", (int)r, (unsigned long long)rows,
                  (unsigned long long)cols, box_rows, n_slabs);
        return NPPC_ERR_CUDA;
    }
    return NPPC_OK;
}

}  // namespace tc
}  // namespace nppc

namespace nppc {
namespace tc {
// NHWC fp16 tensor [B][H][W][C] viewed as 4-D (C, W, H, B): box = [box_c channels][box_w][box_h][1], 128-byte swizzle.
// Out-of-range coordinates (the 3x3 halo: w = -1, h = H, ...) are zero-filled on load and clipped on store.
int make_tmap_f16_nhwc(CUtensorMap* map, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint32_t box_c,
                       uint32_t box_w, uint32_t box_h) {
    auto enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return NPPC_ERR_CUDA;
    }
    cuuint64_t gdim[4] = {C, W, H, B};
    cuuint64_t gstride[3] = {C * 2, C * W * 2, C * W * H * 2};
    cuuint32_t box[4] = {box_c, box_w, box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(nhwc) failed (%d) C=%llu W=%llu H=%llu B=%llu", (int)r, (unsigned long long)C,
                  (unsigned long long)W, (unsigned long long)H, (unsigned long long)B);
        return NPPC_ERR_CUDA;
    }
    return NPPC_OK;
}
}  // namespace tc
}  // namespace nppc
