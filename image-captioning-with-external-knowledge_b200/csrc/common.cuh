// Shared device helpers for the ickb200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define ICK_OK 0
#define ICK_ERR_ARG 1
#define ICK_ERR_CUDA 2
#define ICK_ERR_UNSUPPORTED 3

// dtype codes of the C-ABI
#define ICK_F32 0
#define ICK_BF16 1

void ick_set_error(const char* fmt, ...);
int ick_check_launch(const char* what);

#define ICK_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            ick_set_error(__VA_ARGS__); \
            return ICK_ERR_ARG;         \
        }                               \
    } while (0)

typedef __nv_bfloat16 bf16;

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
// ick_pdl_entry(): `launch_dependents` lets the NEXT kernel of the stream be scheduled onto SMs as they drain (its launch
// latency, CTA ramp-up and barrier/TMEM/tensor-map prologue overlap this kernel's tail), `wait` blocks until the PREVIOUS
// kernel has completed and its writes are visible.  Nothing before the wait may touch global memory.  Because every
// kernel waits, completion is transitive along the stream.  ICK_PDL=0 in the environment falls back to plain launches.
__device__ __forceinline__ void ick_pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void ick_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ick_pdl_entry() {
    ick_pdl_launch();
    ick_pdl_wait();
}
bool ick_pdl_enabled();  // loss_optim.cu
extern long long ick_launch_counter;  // loss_optim.cu: kernels launched by this library (ick_launch_count, bench.py's gpu_launches)

template <typename... KArgs>
struct IckLauncher {
    void (*kernel)(KArgs...);
    dim3 grid, block;
    size_t smem;
    cudaStream_t stream;
    template <typename... Args>
    void operator()(Args&&... args) const {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = ick_pdl_enabled() ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ++ick_launch_counter;
        cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through ick_check_launch()
    }
};
template <typename... KArgs>
static inline IckLauncher<KArgs...> ick_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
    return IckLauncher<KArgs...>{kernel, grid, block, smem, stream};
}

// ---- dtype conversion -------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

// load / store 2 consecutive elements (address must be 2-element aligned)
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld2(const bf16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st2(bf16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// load 8 consecutive elements as floats (address must be 8-element aligned: 32 B for f32, 16 B for bf16)
__device__ __forceinline__ void ld8(const float* p, float* v) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float* v) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void st8(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float* v) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}

// ---- warp reductions ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- counter-based dropout hash -------------------------------------------------------------------------------------
// keep(seed, site, row, col) is a pure function, so the backward pass regenerates the mask instead of storing it.
// An element is addressed as (row, col) = (flat index of the leading dimensions, index in the last dimension).  The
// row is mixed once (ick_rowmix); ONE 32-bit hash then serves the column pair (2k, 2k+1) as two 15-bit uniform fields
// (bits 0-14 for the even column, bits 16-30 for the odd one), compared against thr = floor(p * 32768): keep iff
// field >= thr.  15 bits leave the top bit of each half-word free, so both comparisons can be done with one add.  tests/dropout_ref.py is the numpy port used to inject the same masks into
// the oracle.
__host__ __device__ __forceinline__ uint32_t ick_rowmix(uint32_t seed, uint32_t site, uint64_t row) {
    uint32_t h = seed ^ (site * 0x7F4A7C15u) ^ ((uint32_t)row * 0x9E3779B1u) ^ ((uint32_t)(row >> 32) * 0x85EBCA77u);
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
// hash of column pair `pair` (= col >> 1) of a row: the row mix is already a full avalanche of (seed, site, row), so one
// add-multiply, one xor-shift and one multiply are enough for dropout-quality bits in both 15-bit fields (checked in
// tests/test_host_wiring.py: keep rate, neighbour correlations along rows / columns / pairs)
__host__ __device__ __forceinline__ uint32_t ick_pairhash_idx(uint32_t rowmix, uint32_t pair) {
    uint32_t h = rowmix + pair * 0x9E3779B1u;
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    return h;
}
__host__ __device__ __forceinline__ uint32_t ick_pairhash(uint32_t rowmix, uint32_t col) { return ick_pairhash_idx(rowmix, col >> 1); }
__host__ __device__ __forceinline__ bool ick_keep_lo(uint32_t pairhash, uint32_t thr) { return (pairhash & 0x7FFFu) >= thr; }
__host__ __device__ __forceinline__ bool ick_keep_hi(uint32_t pairhash, uint32_t thr) { return ((pairhash >> 16) & 0x7FFFu) >= thr; }
// multiplier (0 or 1/(1-p)) of column `col` given the hash of its pair
__device__ __forceinline__ float ick_keep(uint32_t thr, float inv_keep, uint32_t pairhash, uint32_t col) {
    return ((col & 1u) ? ick_keep_hi(pairhash, thr) : ick_keep_lo(pairhash, thr)) ? inv_keep : 0.0f;
}

// ---- attention-probability dropout: bit-parallel keep words ---------------------------------------------------------
// The (query, key) probability matrices are by far the largest dropout sites (116 M elements per entity layer at B = 128), and a
// hash per element PAIR made the attention kernels instruction-bound.  For these sites the mask is defined per WORD instead:
// the 32 keys [32*kg, 32*kg + 32) of a probability row (row mix as above) share up to 16 hashed words w_0.. (bit planes,
// most significant first) of 32 independent 16-bit uniform numbers U_b, one per key; key k keeps its probability iff
// U_b < t16, t16 = 65536 * (1 - p) = 65536 - 2 * thr, b = ick_keybit(k).  The comparison runs bit-serially on whole words
// (two logic ops per plane for 32 keys) and stops at the lowest set bit of t16: the reference's p = 0.5 needs ONE hash per 32
// probabilities.  Key k of a group sits at bit (k >> 1) + 16 * (k & 1) - the two keys of a bf16 pair at bits i and 16 + i,
// so a shift puts them on the two sign bits that one byte-permute expands into the AND mask of the packed pair.
__host__ __device__ __forceinline__ uint32_t ick_fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ uint32_t ick_keybit(uint32_t key) { return ((key & 31u) >> 1) + ((key & 1u) << 4); }
// keep word of key group kg of a row; t16 in [2, 65534].  t16 = 0x8000 (p = 0.5, the reference's default) is ONE plane: the
// complement of a single hash; the general case loops over the planes (never unrolled: it sits inside the attention tile loops).
__host__ __device__ __forceinline__ uint32_t ick_keepword(uint32_t rowmix, uint32_t kg, uint32_t t16) {
    const uint32_t w0 = ick_fmix32(rowmix + (kg * 16u) * 0x9E3779B1u);
    if (t16 == 0x8000u) return ~w0;
    uint32_t lt = 0u, eq = 0xFFFFFFFFu, w = w0;
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
        if ((t16 & (0xFFFFu >> i)) == 0u) break;  // U < t16 is decided: equal prefixes are not below
        if (i > 0) w = ick_fmix32(rowmix + (kg * 16u + (uint32_t)i) * 0x9E3779B1u);
        if ((t16 >> (15 - i)) & 1u) {
            lt |= eq & ~w;
            eq &= w;
        } else {
            eq &= ~w;
        }
    }
    return lt;
}
__host__ __device__ __forceinline__ uint32_t ick_attn_t16(uint32_t thr) { return 65536u - 2u * thr; }

struct DropCfg {
    uint32_t thr;    // floor(p * 32768), 0 = off
    float inv_keep;  // 1 / (1 - p)
    uint32_t seed;
    uint32_t site;
    const uint32_t* seed_dev;  // optional device-resident seed offset (lets a captured CUDA graph draw new masks per replay)
};
// host-side registry of the device seed offset (ick_set_seed_source); defined in loss_optim.cu
const uint32_t* ick_seed_source();
// kernels call this once on their by-value DropCfg copy
__device__ __forceinline__ void ick_resolve_seed(DropCfg& d) {
    if (d.seed_dev != nullptr) d.seed += *d.seed_dev;
}
// convenience for kernels that touch one element at a time
__device__ __forceinline__ float ick_drop_mul(const DropCfg& d, uint32_t rowmix, uint32_t col) {
    if (d.thr == 0u) return 1.0f;
    return ick_keep(d.thr, d.inv_keep, ick_pairhash(rowmix, col), col);
}
static inline DropCfg make_drop(float p, unsigned seed, unsigned site) {
    DropCfg d;
    d.seed_dev = ick_seed_source();
    if (p <= 0.f) {
        d.thr = 0u;
        d.inv_keep = 1.f;
    } else {
        double t = (double)p * 32768.0;
        d.thr = t >= 32767.0 ? 32767u : (uint32_t)t;
        d.inv_keep = 1.0f / (1.0f - p);
    }
    d.seed = seed;
    d.site = site;
    return d;
}
