// Host helper shared by the batched path and the tcgen05 vocoder: 2-D TMA tensor maps (cuTensorMapEncodeTiled through the
// runtime's driver entry point; libcuda is not linked). Included inside engine.cu's anonymous namespace.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}
// bf16 [rows][K] row-major, box = 64 columns (one 128-byte swizzle atom) x box_rows
int make_map(lqt_engine* h, CUtensorMap* m, const void* base, long long rows, int K, int box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) { h->err = "cuTensorMapEncodeTiled unavailable"; return 1; }
    if (K % 64 || box_rows > 256 || box_rows < 1) { h->err = "tensor map: K must be a multiple of 64 and the box at most 256 rows"; return 1; }
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return 1; }
    return 0;
}

