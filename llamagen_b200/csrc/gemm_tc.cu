// Blackwell-native weight-streaming GEMM: tcgen05.mma (UMMA) with the accumulator in TMEM, operands staged
// in shared memory by TMA (cp.async.bulk.tensor) through an mbarrier full/empty ring.
//
//   y[r, n] = sum_k x[r, k] * W[n, k]        (nn.Linear without bias, gpt.py:161-163,199-200,287)
//
// "Swap-AB" orientation for decode: the WEIGHT tile is the UMMA A operand (M = 128 output features = the
// 128 TMEM lanes), the activations are the B operand (N = rows padded to 16, <= 256 = TMEM columns), so a
// skinny batch never wastes the 128-row MMA shape and each CTA streams a [128 x Kslice] weight slab from
// HBM exactly once. One CTA = one (n-tile, k-slice); split-K slabs (fp32) are reduced by the row-wise
// epilogue kernels in xf_kernels.cu, exactly like the mma.sync path it replaces (gemm.cu).
//
// Warp roles (256 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected lane)
// + TMEM allocator; when the accumulator is complete all 8 warps drain it (TMEM lane quarter = warp % 4,
// column half = warp / 4). The drain is written as a pointer walk (one IADD + one STG per value): a lone warp per
// scheduler is latency-bound, so every instruction in this loop costs ~4 cycles (measured: the first version
// spent 4.8 us of a 7 us kernel here).
#include "kernels.cuh"
#include "tma_utils.cuh"
#include "umma_utils.cuh"
#include <mutex>
#include <unordered_map>

namespace {

constexpr int kBlockN = 128;    // weight rows per CTA  (UMMA M)
constexpr int kBlockK = 64;     // bf16 elements per k-block = 128 B = one swizzle-128B row
constexpr int kMaxStages = 8;
constexpr int kThreads = 256;   // 8 warps: TMA producer, MMA issuer, then ALL of them drain the accumulator
constexpr int kATileBytes = kBlockN * kBlockK * 2;   // 16 KB
constexpr int kMaxRowsPerCta = 256;                  // UMMA N <= 256 = TMEM columns of one accumulator

using namespace tma;

using namespace umma;

struct TcArgs {
    int M, N, K;          // activations rows, weight rows, reduction
    int n_split;          // weight rows [0, n_split) come from map_wa, the rest from map_wb
    int rpad;             // rows of one row block rounded up to 16 (UMMA N); M > 256 is cut into gridDim.z blocks of rblk rows
    int rblk;             // rows per row block (== M when gridDim.z == 1)
    int kblocks_per_split;
    int tmem_cols;        // power of two >= max(32, rpad)
    int stages;           // smem ring depth (<= kMaxStages), sized to fit 227 KB
    int swap;             // 1: weights are the UMMA A operand (TMEM lane = feature); 0: activations are A (TMEM lane = row)
    float* partial;       // [ksplit][M][N]
    const char* pf0; unsigned long long pfb0;   // next GEMM's weights to pull into L2 (see GemmNext)
    const char* pf1; unsigned long long pfb1;
    unsigned long long* trace;   // debug: [cta][8] %globaltimer stamps of the pipeline phases (nullable)
    unsigned long long whint;    // L2 eviction hint of the weight tiles (0: default policy)
    int ld32;                    // drain with 32-column TMEM loads where possible
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TC_TRACE(slot)                                                                            \
    do {                                                                                          \
        if (a.trace) a.trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot)] = gtime(); \
    } while (0)

__device__ __forceinline__ void gemm_tc_body(const CUtensorMap& map_wa, const CUtensorMap& map_wb, const CUtensorMap& map_x,
                                             const TcArgs& a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // stage s: A tile at s*stage_bytes (1024-aligned), B tile right after
    const int b_tile_bytes = a.rpad * kBlockK * 2;
    // activations tile: rpad rows are loaded; when it is the UMMA A operand (M = 128) the full 128-row slot is reserved
    const int stage_bytes = kATileBytes + (a.swap ? ((b_tile_bytes + 1023) / 1024) * 1024 : kATileBytes);
    uint8_t* tiles = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int kStages = a.stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + kStages * stage_bytes);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full_bar = empty_bar + kMaxStages;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * kBlockN;
    const int ks = blockIdx.y;
    const int row0 = (int)blockIdx.z * a.rblk;               // first activation row of this CTA's row block (t2i prefill: M = R*120)
    const int Mb = min(a.rblk, a.M - row0);                  // valid rows in the block
    if (threadIdx.x == 0) TC_TRACE(0);
    const int total_kb = (a.K + kBlockK - 1) / kBlockK;
    const int kb0 = ks * a.kblocks_per_split;
    const int nkb = max(0, min(a.kblocks_per_split, total_kb - kb0));

    if (warp == 0 && lane == 0) {
        prefetch_map(&map_wa);
        prefetch_map(&map_wb);
        prefetch_map(&map_x);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_slot, (uint32_t)a.tmem_cols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;
    if (threadIdx.x == 0) TC_TRACE(1);
    lg_pdl_launch_dependents();
    if (warp != 0) lg_pdl_wait();   // warp 0 waits after it has requested the first weight tiles

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            const bool second = n0 >= a.n_split;
            const CUtensorMap* wmap = second ? &map_wb : &map_wa;
            const int wrow = second ? n0 - a.n_split : n0;
            const uint32_t tx = (uint32_t)(kATileBytes + b_tile_bytes);
            // Weights do not depend on the previous kernel: the first ring of weight tiles is requested BEFORE
            // the programmatic-dependency wait, so the HBM stream of this GEMM overlaps the tail of its producer.
            const int npre = min(nkb, kStages);
            for (int i = 0; i < npre; ++i) {
                mbar_expect_tx(&full_bar[i], tx);
                load_2d_hint(tiles + i * stage_bytes, wmap, &full_bar[i], (kb0 + i) * kBlockK, wrow, a.whint);
            }
            lg_pdl_wait();
            for (int i = 0; i < npre; ++i)
                load_2d(tiles + i * stage_bytes + kATileBytes, &map_x, &full_bar[i], (kb0 + i) * kBlockK, row0);
            for (int i = npre; i < nkb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (uint32_t)((i / kStages) & 1);
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], tx);
                uint8_t* sa = tiles + s * stage_bytes;
                load_2d_hint(sa, wmap, &full_bar[s], (kb0 + i) * kBlockK, wrow, a.whint);
                load_2d(sa + kATileBytes, &map_x, &full_bar[s], (kb0 + i) * kBlockK, row0);
            }
            // ask L2 for this CTA's share of the next GEMM's weights (static chain: qkv -> wo -> w1|w3 -> w2 -> next qkv)
            {
                const unsigned long long cta = (unsigned long long)blockIdx.y * gridDim.x + blockIdx.x;
                const unsigned long long ncta = (unsigned long long)gridDim.x * gridDim.y;
                const char* pp[2] = {a.pf0, a.pf1};
                const unsigned long long bb[2] = {a.pfb0, a.pfb1};
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (!bb[t]) continue;
                    const unsigned long long per = ((bb[t] + ncta - 1) / ncta + 4095ull) & ~4095ull;
                    unsigned long long off = cta * per;
                    const unsigned long long end = off + per < bb[t] ? off + per : bb[t];
                    for (; off < end; off += 16384ull) {
                        const unsigned long long len = end - off < 16384ull ? end - off : 16384ull;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pp[t] + off), "r"((uint32_t)(len & ~15ull)) : "memory");
                    }
                }
            }
            TC_TRACE(2);
        }
        __syncwarp();      // reconverge: the drain below uses warp-collective (.sync.aligned) TMEM loads
        lg_pdl_wait();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = make_idesc(a.swap ? a.rpad : kBlockN);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % kStages;
            const uint32_t ph = (uint32_t)((i / kStages) & 1);
            mbar_wait(&full_bar[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (i == 0 && lane == 0) TC_TRACE(3);
            if (i == nkb - 1 && lane == 0) TC_TRACE(4);
            if (elect_one()) {
                const uint32_t sa = smem_u32(tiles + s * stage_bytes);
                const uint64_t wdesc = make_desc_sw128(sa);
                const uint64_t xdesc = make_desc_sw128(sa + kATileBytes);
                const uint64_t adesc = a.swap ? wdesc : xdesc, bdesc = a.swap ? xdesc : wdesc;
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                    // advance 16 bf16 = 32 bytes inside the 128-byte swizzle span: +2 in the (addr >> 4) field
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((i | k) != 0));
                }
                umma_commit(&empty_bar[s]);                       // frees the smem slot when these MMAs retire
                if (i == nkb - 1) umma_commit(tmem_full_bar);     // accumulator complete
            }
            __syncwarp();
        }
    }
    // ---------------------------------------------------------------------- drain: TMEM -> registers -> fp32 slab
    if (a.swap) {
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = warp >> 2;                   // which half of the columns (activation rows) it drains
        const int n = n0 + q * 32 + lane;
        const int cols_half = ((a.rpad / 16 + 1) / 2) * 16;
        const int c_begin = half * cols_half, c_end = min(a.rpad, c_begin + cols_half);
        {
        float* out = a.partial + (size_t)ks * a.M * a.N + (size_t)row0 * a.N;
        if (nkb > 0) {
            mbar_wait(tmem_full_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (threadIdx.x == 64) TC_TRACE(5);
            float* p = out + (size_t)c_begin * a.N + n;
            const size_t stride = (size_t)a.N;
            int c0 = c_begin;
            // 32 columns per TMEM load while a full 32-row group of valid rows remains (64-row decode chains: ONE load per warp)
            if (a.ld32) {
                for (; c0 + 32 <= c_end && c0 + 32 <= Mb; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    if (n < a.N) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { *p = __uint_as_float(v[j]); p += stride; }
                    } else {
                        p += 32 * stride;
                    }
                }
            }
            for (; c0 < c_end; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (n < a.N) {
                    if (c0 + 16 <= Mb) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) { *p = __uint_as_float(v[j]); p += stride; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (c0 + j < Mb) *p = __uint_as_float(v[j]);
                            p += stride;
                        }
                    }
                }
            }
        } else if (n < a.N) {
            for (int r = c_begin; r < min(c_end, Mb); ++r) out[(size_t)r * a.N + n] = 0.f;
        }
        }
    } else {
        // activations are the A operand: TMEM lane = row r, columns = 128 features -> 64-byte vector stores per thread
        const int q = warp & 3, half = warp >> 2;
        const int r = q * 32 + lane;
        const int c_begin = half * (kBlockN / 2), c_end = c_begin + kBlockN / 2;
        float* out = a.partial + (size_t)ks * a.M * a.N + (size_t)r * a.N + n0;
        if (nkb > 0) {
            mbar_wait(tmem_full_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (threadIdx.x == 64) TC_TRACE(5);
            for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (r < a.M && n0 + c0 + 16 <= a.N) {
                    float4* p4 = reinterpret_cast<float4*>(out + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        p4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                            __uint_as_float(v[4 * j + 3]));
                } else if (r < a.M) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + c0 + j < a.N) out[c0 + j] = __uint_as_float(v[j]);
                }
            }
        } else if (r < a.M) {
            for (int c = c_begin; c < c_end; ++c)
                if (n0 + c < a.N) out[c] = 0.f;
        }
    }
    if (threadIdx.x == 64) TC_TRACE(6);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    if (threadIdx.x == 0) TC_TRACE(7);
}

__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap map_wa,
                                                              const __grid_constant__ CUtensorMap map_wb,
                                                              const __grid_constant__ CUtensorMap map_x, TcArgs a) {
    gemm_tc_body(map_wa, map_wb, map_x, a);
}

// ---------------------------------------------------------------------------------------------- host side
}  // namespace

namespace tma {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}
int make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                uint32_t box_cols) {
    EncodeTiledFn enc = get_encode();
    LG_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
               (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, box_cols);
    return 0;
}
int make_map_nhwc(CUtensorMap* m, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t box_h,
                  uint32_t box_w, uint32_t box_c, uint32_t pixel_stride) {
    EncodeTiledFn enc = get_encode();
    LG_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    // pixel_stride s > 1: the box traverses s*box_w x s*box_h input pixels and keeps every s-th one (a strided conv's taps)
    cuuint32_t box[4] = {box_c, box_w * pixel_stride, box_h * pixel_stride, 1};
    cuuint32_t estr[4] = {1, pixel_stride, pixel_stride, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed (%d) B=%llu H=%llu W=%llu C=%llu box=%ux%ux%u", (int)r,
               (unsigned long long)B, (unsigned long long)H, (unsigned long long)W, (unsigned long long)C, box_h, box_w, box_c);
    return 0;
}
}  // namespace tma

static unsigned long long* g_tc_trace = nullptr;
extern "C" void lg_debug_set_tc_trace(unsigned long long* dev_buf) { g_tc_trace = dev_buf; }

// Plan: number of k-slices so that (N/128) * ksplit is ~104 CTAs (measured best on B200 for the decode shapes:
// 148 -> 350 ms/step, 112 -> 341, 96 -> 340, 72 -> 352; fewer, fatter slices halve the fp32 slab traffic), >= 2 k-blocks per slice.
int gemm_tc_ksplit(int M, int N, int K) {
    // row blocks (M > 256: t2i prefill) already multiply the CTA count
    const int tiles = cdiv(N, kBlockN) * cdiv(M, kMaxRowsPerCta), kb = cdiv(K, kBlockK);
    int ks = std::max(1, lg_env_flag("LG_TC_CTAS", 104) / std::max(tiles, 1));
    ks = std::min(ks, std::max(1, kb / 2));
    return std::min(ks, 16);
}

bool gemm_tc_supported(int M, int N, int K, int dtype) {
    return dtype == LG_DTYPE_BF16 && M >= 1 && K % 8 == 0 && N % 2 == 0;
}

static int gemm_tc_launch(const void* X, int ldx, const void* Wa, const void* Wb, int n_split, int M, int N, int K,
                          float* partial, int* ksplit_out, cudaStream_t st, const GemmNext* next) {
    LG_REQUIRE(gemm_tc_supported(M, N, K, LG_DTYPE_BF16), "gemm_tc: unsupported shape %d %d %d", M, N, K);
    if (Wb == nullptr) { Wb = Wa; n_split = N; }
    LG_REQUIRE(n_split % kBlockN == 0 || n_split == N, "gemm_tc: weight segment boundary %d must be a multiple of %d", n_split, kBlockN);
    LG_REQUIRE(((uintptr_t)X & 15) == 0 && ((uintptr_t)Wa & 15) == 0 && ((uintptr_t)Wb & 15) == 0 && ldx % 8 == 0,
               "gemm_tc: operands must be 16-byte aligned");
    TcArgs a;
    a.M = M; a.N = N; a.K = K; a.n_split = n_split;
    a.rblk = M <= kMaxRowsPerCta ? M : kMaxRowsPerCta;
    const int zblocks = cdiv(M, a.rblk);
    LG_REQUIRE(zblocks <= 65535, "gemm_tc: %d row blocks not launchable", zblocks);
    a.rpad = ((a.rblk + 15) / 16) * 16;
    const int ks = gemm_tc_ksplit(M, N, K);
    const int kb = cdiv(K, kBlockK);
    a.kblocks_per_split = cdiv(kb, ks);
    // rows-as-lanes (swap = 0, 64-byte vector stores per thread) was measured SLOWER than features-as-lanes on B200:
    // each warp store then touches 32 different 128-byte lines (drain 2.8-3.2 us vs 1.1 us), so it stays opt-in.
    a.swap = (lg_env_flag("LG_TC_NOSWAP", 0) && M > 64 && M <= kBlockN && zblocks == 1) ? 0 : 1;
    a.tmem_cols = 32;
    while (a.tmem_cols < (a.swap ? a.rpad : kBlockN)) a.tmem_cols *= 2;
    a.partial = partial;
    if (ksplit_out) *ksplit_out = ks;

    CUtensorMap mwa, mwb, mx;
    LG_TRY(tma::make_map_2d(&mwa, Wa, (uint64_t)std::min(n_split, N), (uint64_t)K, (uint64_t)K, kBlockN, kBlockK));
    LG_TRY(tma::make_map_2d(&mwb, Wb, (uint64_t)std::max(N - n_split, Wb == Wa ? N : 1), (uint64_t)K, (uint64_t)K, kBlockN, kBlockK));
    LG_TRY(tma::make_map_2d(&mx, X, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, (uint32_t)a.rpad, kBlockK));   // rows >= M read as zero

    const int b_tile_bytes = a.rpad * kBlockK * 2;
    const int stage_bytes = kATileBytes + (a.swap ? ((b_tile_bytes + 1023) / 1024) * 1024 : kATileBytes);
    a.stages = std::min(std::min(kMaxStages, lg_env_flag("LG_TC_STAGES", 4)), (int)((225 * 1024 - 1024) / stage_bytes));
    // 4 stages (96 KB at the R = 64 rows of a decode chain) instead of filling shared memory: a CTA never owns more than ~8
    // k-blocks, and the smaller footprint lets the PDL-launched next kernel become resident next to this one (filling
    // shared memory: 335 ms/step, 3 stages at R = 128: 320). With two 64-row chains, 4 stages let the QKV GEMM request its
    // whole K range before the dependency wait: 3 -> 4 stages = 297.2 -> 294.9 ms/step (2 runs each), 5 no better.
    a.stages = std::min(a.stages, std::max(2, a.kblocks_per_split));   // never more stages than k-blocks
    a.trace = g_tc_trace;
    a.ld32 = lg_env_flag("LG_TC_LD32", 1);
    a.whint = (lg_env_flag("LG_L2_HINT", 1) & 2) ? tma::kL2EvictLast : 0ull;
    const bool pf = next && lg_env_flag("LG_L2_PREFETCH", 1);
    a.pf0 = pf ? (const char*)next->p0 : nullptr; a.pfb0 = pf ? next->b0 : 0;
    a.pf1 = pf ? (const char*)next->p1 : nullptr; a.pfb1 = pf ? next->b1 : 0;
    LG_REQUIRE(a.stages >= 2, "gemm_tc: tile too large for a 2-stage ring");
    const size_t smem = 1024 + (size_t)a.stages * stage_bytes + (2 * kMaxStages + 1) * sizeof(uint64_t) + 16;
    static DevOnce attr;
    if (lg_first_on_device(attr)) {
        LG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    LG_REQUIRE(smem <= 227 * 1024, "gemm_tc: shared memory %zu too large", smem);
    dim3 grid(cdiv(N, kBlockN), ks, zblocks);
    (void)lg_launch(gemm_tc_kernel, dim3(grid), dim3(kThreads), smem, st, mwa, mwb, mx, a);
    LG_LAUNCH_CHECK();
    return 0;
}

int gemm_tc_partial(const void* X, int ldx, const void* Wa, const void* Wb, int n_split, int M, int N, int K,
                    float* partial, int* ksplit_out, cudaStream_t st, const GemmNext* next) {
    return gemm_tc_launch(X, ldx, Wa, Wb, n_split, M, N, K, partial, ksplit_out, st, next);
}
