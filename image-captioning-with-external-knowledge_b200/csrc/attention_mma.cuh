// Device / host helpers shared by the bf16 tensor-core attention kernels (attention_mma.cu: forward + chunked fall-backs,
// attention_bwd_fused.cu: the fused backward): PTX wrappers, the 64B-swizzled tile addressing, mma.sync fragment helpers,
// the TMA tensor map of a head-layout tensor.  Everything is internal-linkage (static / inline in a header-local namespace).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ickattn {


constexpr int HD = 32;                   // padded head width: 64-byte rows
constexpr int TK = 64;                   // streamed rows per tile
constexpr int TILE_BYTES = TK * HD * 2;  // 4096
constexpr int CH = 10;                   // tiles resident at once
constexpr int NWMAX = 8;                 // warps per CTA
constexpr int SUB = 32;                  // streamed rows per register sub-step of the backward kernels

struct Dims {
    int B, H, Sq, Sk, dh;
    int causal;
    float scale, scale_log2;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---- shared-memory carve-up: [mbarriers (1 KiB)] [tensor 0: n tiles] [tensor 1: n tiles] [per-row scalars (dKV only)] ---------
struct Smem {
    uint32_t bars, t0, t1;
    float* scal;  // generic pointer to the scalar area
};
__device__ __forceinline__ Smem carve(uint8_t* raw, int ntc) {
    uint8_t* p = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    Smem s;
    s.bars = smem_u32(p);
    s.t0 = s.bars + 1024;
    s.t1 = s.t0 + ntc * TILE_BYTES;
    s.scal = (float*)(p + 1024 + 2 * ntc * TILE_BYTES);
    return s;
}
__device__ __forceinline__ void init_bars(const Smem& s, int ntc, const CUtensorMap* a, const CUtensorMap* b) {
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(b) : "memory");
        for (int t = 0; t < ntc; ++t) mbar_init(s.bars + 8 * t, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}
// one elected thread: tiles [tile0, tile0 + n) of both streamed tensors of image b, head h
__device__ __forceinline__ void issue_tiles(const Smem& s, const CUtensorMap* a, const CUtensorMap* b2, int tile0, int n, int h, int b) {
    for (int t = 0; t < n; ++t) {
        const uint32_t bar = s.bars + 8 * t;
        mbar_expect_tx(bar, 2 * TILE_BYTES);
        tma_load_3d(s.t0 + t * TILE_BYTES, a, bar, h * HD, (tile0 + t) * TK, b);
        tma_load_3d(s.t1 + t * TILE_BYTES, b2, bar, h * HD, (tile0 + t) * TK, b);
    }
}

// Byte offset of (row r, 16-byte chunk c) inside a 64B-swizzled tile whose base is 1024-aligned: r*64 + ((c ^ ((r>>1)&3)) << 4).
// Per-lane ldmatrix offsets; rows advance in multiples of 8 (non-trans) / 16 (trans), which leaves the swizzle term unchanged.
struct LaneOff {
    uint32_t nt;     // "tile^T" B operand: matrix m = lane>>3 is chunk m of row (lane&7)
    uint32_t tr[2];  // "P x tile" B operand (.trans): row (lane&7) + 8*((lane>>3)&1), chunk 2*nd + (lane>>4)
};
__device__ __forceinline__ LaneOff lane_offsets(int lane) {
    LaneOff o;
    const int r = lane & 7, c = lane >> 3;
    o.nt = r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
    const int r2 = (lane & 7) + 8 * ((lane >> 3) & 1), hi = lane >> 4;
#pragma unroll
    for (int nd = 0; nd < 2; ++nd) o.tr[nd] = r2 * 64 + (((2 * nd + hi) ^ ((r2 >> 1) & 3)) << 4);
    return o;
}

// A-operand fragments (16 rows x 32 k) of the warp's own rows, read from global memory; rows >= rmax and columns >= dh are zero
__device__ __forceinline__ uint32_t ld_frag(const bf16* row, bool ok, int c, int dh) {
    uint32_t v = ok ? *reinterpret_cast<const uint32_t*>(row + c) : 0u;
    if (c + 1 >= dh) v = c >= dh ? 0u : (v & 0xFFFFu);
    return v;
}
__device__ __forceinline__ void load_own(const bf16* base, size_t ld, int r0, int r1, int rmax, int dh, int tq, uint32_t (*a)[4]) {
    const bf16* p0 = base + (size_t)r0 * ld;
    const bf16* p1 = base + (size_t)r1 * ld;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = ld_frag(p0, r0 < rmax, 16 * ks + 2 * tq, dh);
        a[ks][1] = ld_frag(p1, r1 < rmax, 16 * ks + 2 * tq, dh);
        a[ks][2] = ld_frag(p0, r0 < rmax, 16 * ks + 2 * tq + 8, dh);
        a[ks][3] = ld_frag(p1, r1 < rmax, 16 * ks + 2 * tq + 8, dh);
    }
}

// acc[j] (16 x 8, j = 0..NJ-1) = A(16 x 32) * T^T for rows [8*j0, 8*(j0+NJ)) of the tile at shared address `tile`
template <int NJ>
__device__ __forceinline__ void mma_a_tT(float (*acc)[4], const uint32_t (*a)[4], uint32_t tile, int j0, const LaneOff& lo) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(b0, b1, b2, b3, tile + (j0 + j) * 512 + lo.nt);
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        mma16816(acc[j], a[0], b0, b1);
        mma16816(acc[j], a[1], b2, b3);
    }
}
// out[n] (16 x 8 over d, n = 0..3) += P(16 x 8*NJ) * T[rows 8*j0 ..], P given as packed A fragments pa[NJ/2][4]
template <int NJ>
__device__ __forceinline__ void mma_p_t(float (*out)[4], const uint32_t (*pa)[4], uint32_t tile, int j0, const LaneOff& lo) {
#pragma unroll
    for (int kk = 0; kk < NJ / 2; ++kk) {
#pragma unroll
        for (int nd = 0; nd < 2; ++nd) {
            uint32_t r0, r1, r2, r3;
            ldsm_x4_trans(r0, r1, r2, r3, tile + (j0 / 2 + kk) * 1024 + lo.tr[nd]);
            mma16816(out[2 * nd], pa[kk], r0, r1);
            mma16816(out[2 * nd + 1], pa[kk], r2, r3);
        }
    }
}
// accumulator layout p[NJ][4] -> packed A fragments
template <int NJ>
__device__ __forceinline__ void pack_p(const float (*p)[4], uint32_t (*pa)[4]) {
#pragma unroll
    for (int kk = 0; kk < NJ / 2; ++kk) {
        pa[kk][0] = pack2(p[2 * kk][0], p[2 * kk][1]);
        pa[kk][1] = pack2(p[2 * kk][2], p[2 * kk][3]);
        pa[kk][2] = pack2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        pa[kk][3] = pack2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    }
}

// write a 16 x 32 accumulator slab (4 n-tiles) as bf16 rows; columns >= dh are written as zero
__device__ __forceinline__ void store_slab(bf16* base, size_t ld, int row_g, int row_g8, int rmax, const float (*o)[4], float s0, float s1,
                                           int dh, int tq) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const int c = 8 * n + 2 * tq;
        const float m0 = c < dh ? 1.f : 0.f, m1 = c + 1 < dh ? 1.f : 0.f;
        if (row_g < rmax) *reinterpret_cast<uint32_t*>(base + (size_t)row_g * ld + c) = pack2(o[n][0] * s0 * m0, o[n][1] * s0 * m1);
        if (row_g8 < rmax) *reinterpret_cast<uint32_t*>(base + (size_t)row_g8 * ld + c) = pack2(o[n][2] * s1 * m0, o[n][3] * s1 * m1);
    }
}

// ---- dropout of the probabilities: element (row = (b,h,query), col = key) -------------------------------------------------
__device__ __forceinline__ uint64_t prob_row(int b, int H, int h, int Sq, int qi) { return ((uint64_t)b * H + h) * (uint64_t)Sq + qi; }
// 0xFFFF / 0x0000 per half-word from the two keep bits (byte permute with sign replication)
__device__ __forceinline__ uint32_t keep_mask_bf16x2(uint32_t kb) {
    uint32_t m;
    asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(kb));
    return m;
}


// ---- host side -------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    }
    return fn;
}
// 3-D bf16 map over a (B, S, ld) head-layout tensor: (column < H*32, position < S, image < B); box = 32 columns x 64 positions
static inline int make_tmap3(CUtensorMap* tm, const void* ptr, int H, int S, int B, int ld) {
    EncodeFn enc = get_encode();
    if (!enc) {
        ick_set_error("cuTensorMapEncodeTiled entry point not available");
        return ICK_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)H * HD, (cuuint64_t)S, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)S * ld * 2};
    cuuint32_t box[3] = {HD, TK, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ick_set_error("attention: cuTensorMapEncodeTiled failed (%d): ptr=%p H=%d S=%d B=%d ld=%d", (int)r, ptr, H, S, B, ld);
        return ICK_ERR_CUDA;
    }
    return ICK_OK;
}

static inline Dims make_dims(int B, int H, int Sq, int Sk, int dh, int causal) {
    Dims d;
    d.B = B; d.H = H; d.Sq = Sq; d.Sk = Sk; d.dh = dh;
    d.causal = causal;
    d.scale = 1.0f / sqrtf((float)dh);
    d.scale_log2 = d.scale * 1.4426950408889634f;
    return d;
}

// ---- per-warp environment shared by the tile bodies --------------------------------------------------------------------------
struct OwnRows {
    int r0, r1;         // own rows of the accumulator halves (lane group g and g + 8)
    int wrow;           // first own row of the warp
    uint32_t rm0, rm1;  // dropout row mixes of r0 / r1 (fwd, dQ)
};
struct TileEnv {
    Dims d;
    float c;        // scale * log2(e)
    float ik;       // 1 / keep
    uint32_t thr;   // dropout threshold (0 = off)
    uint32_t t16;   // keep threshold of the bit-parallel attention masks (common.cuh: ick_keepword)
    LaneOff lo;
    int tq;
};

__device__ __forceinline__ void zero16(float (*a)[4]) {
#pragma unroll
    for (int n = 0; n < 4; ++n) a[n][0] = a[n][1] = a[n][2] = a[n][3] = 0.f;
}
__device__ __forceinline__ TileEnv make_env(const Dims& d, const DropCfg& drop, int lane) {
    TileEnv e;
    e.d = d;
    e.c = d.scale_log2;
    e.ik = drop.inv_keep;
    e.thr = drop.thr;
    e.t16 = ick_attn_t16(drop.thr);
    e.lo = lane_offsets(lane);
    e.tq = lane & 3;
    return e;
}
__device__ __forceinline__ OwnRows own_rows(int wrow, int g, const DropCfg& drop, int b, int H, int h, int Sq, bool mix) {
    OwnRows r;
    r.wrow = wrow;
    r.r0 = wrow + g;
    r.r1 = r.r0 + 8;
    r.rm0 = r.rm1 = 0u;
    if (mix) {
        r.rm0 = ick_rowmix(drop.seed, drop.site, prob_row(b, H, h, Sq, r.r0));
        r.rm1 = ick_rowmix(drop.seed, drop.site, prob_row(b, H, h, Sq, r.r1));
    }
    return r;
}

// first own slab of compute warp `warp` in local item `li` (slabs are numbered globally: li * nslabs + slab)
template <int PNW>
__device__ __forceinline__ int first_slab(int li, int nslabs, int warp) { return (warp + PNW - (int)(((long long)li * nslabs) % PNW)) % PNW; }


}  // namespace ickattn
