// Decode attention (one query per sequence row) for bf16 KV caches: the K and V prefixes of one (row, head)
// are contiguous [c, hd] streams in HBM, so they are pulled by TMA (cp.async.bulk.tensor, 128-byte swizzle)
// through a 3-stage mbarrier ring — no registers are spent on loads in flight — and the two tiny matrix
// products (q.K^T and P.V, M = 1 query padded to the m16 MMA shape) run on the tensor cores with ldmatrix
// operands; softmax stays fp32 in registers (FlashAttention-2 style register reuse of P).
//
// Replaces, for Tq == 1, gpt.py:229-236 (repeat_interleave copies + masked SDPA over all max_seq slots):
// only the valid prefix [0, pos] is read. Mask rule: generate.py:154-163 (emb_masks on the condition keys).
#include "kernels.cuh"
#include "tma_utils.cuh"
#include <algorithm>

namespace {

using namespace tma;

#ifndef LG_ATTN_KC
#define LG_ATTN_KC 32   // keys per stage. Round 1 (one chain): 48 keys / 3 warps / 8 CTAs per SM beat 64 / 4 / 6 (314.9 vs 319.7 ms/step).
                        // Round 2 (two chains, final tree): 32 keys / 2 warps / 12 CTAs per SM -> 274.7 ms/step vs 279.7 for 48 / 3 / 8
                        // (profiles/r2_s22_sweep_attn_kc.txt): the smaller CTAs leave shared memory for the other chain's GEMM CTAs.
#endif
constexpr int kKC = LG_ATTN_KC;   // keys per stage (64 -> 4 warps, 6 CTAs/SM; 48 -> 3 warps, 8 CTAs/SM)
constexpr int kStagesA = 2;       // 2 stages of K+V per CTA; contexts here are <= 1144 keys
constexpr int kWarps = kKC / 16;  // each warp owns 16 keys of a stage
constexpr int kDeepStages = kKC == 32 ? 8 : 6;   // few-item (batch-1) variant: the whole <= 256/288-key context is requested before the dependency wait
#ifdef LG_ATTN_CTAS
constexpr int kCtasPerSm64 = LG_ATTN_CTAS;
#else
constexpr int kCtasPerSm64 = kKC == 48 ? 8 : (kKC == 32 ? 12 : 6);
#endif

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&p);
}

struct AttnTmaArgs {
    const bf16* q;     // [R, D] (unfused path)
    bf16* out;         // [R, D]
    int R, H, maxS;
    const int* pos_dev; int pos_value;
    const int* pos_rows;         // per-row positions [R] or null (continuous batching)
    long long row_base;          // first cache row of this layer inside the tensor maps
    const float* emb_mask; int B, Tc;
    float scale;
    // fused QKV epilogue (FUSED kernels): split-K slabs of the QKV GEMM, RoPE table, this layer's cache bases
    const float* partial; int ksplit;
    const float* freqs;
    bf16* kcache; bf16* vcache;
    unsigned long long kvhint;   // L2 eviction hint of the K/V stream (0: default policy)
    int hd;                      // real head dim (<= HD): GPT-3B's hd = 100 runs the HD = 128 kernel over zero-padded tiles
    int hdp;                     // elements between consecutive cache rows (112 for hd = 100: the tensor map zero-fills 112..127)
};

// FUSED = true: the kernel also IS the QKV epilogue of gpt.py:214-230 for its (row, head): it reduces the split-K
// slabs of the QKV GEMM for its 3*hd columns, applies RoPE to q and k, writes the new K/V row into the cache
// (for future steps) and attends to it straight from shared memory. Every TMA load then only touches rows written
// in EARLIER steps, so the whole KV stream is requested before the programmatic-dependency wait and overlaps the
// QKV GEMM; one dependent kernel per layer disappears.
template <int HD, bool FUSED, int NST, bool PAR_ = (NST > 2)>
__global__ void __launch_bounds__(kWarps * 32 * (PAR_ ? NST : 1), PAR_ ? 1 : (HD == 64 ? (NST > 2 ? 4 : kCtasPerSm64) : (kKC == 32 ? 5 : 4))) attn_tma_kernel(const __grid_constant__ CUtensorMap kmap,
                                                               const __grid_constant__ CUtensorMap vmap,
                                                               const __grid_constant__ CUtensorMap kmap16,
                                                               const __grid_constant__ CUtensorMap vmap16, AttnTmaArgs a) {
    constexpr int NSUB = HD / 64;                 // 128-byte-wide sub-tiles per row
    constexpr int SUB_BYTES = kKC * 128;          // one [kKC keys][64 dims] bf16 sub-tile
    constexpr int TILE_BYTES = NSUB * SUB_BYTES;  // K (or V) of one stage
    // NST > 2 (few work items, batch-1 latency path): one warp group per ring stage, so the chunks of a context are processed
    // concurrently instead of one after the other; a stage is private to its group (named barrier, no CTA-wide sync per chunk)
    constexpr bool PAR = PAR_;
    constexpr int NW = PAR ? NST * kWarps : kWarps;
    constexpr int NSLOT = NW + (FUSED ? 1 : 0);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + NST * 2 * TILE_BYTES);
    float* merge = reinterpret_cast<float*>(full_bar + NST);            // [NSLOT][HD + 2]
    bf16* qbuf = reinterpret_cast<bf16*>(merge + NSLOT * (HD + 2));            // [3][HD]: q, k_new, v_new (FUSED)

    const int h = blockIdx.x, r = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    const int grp = PAR ? warp / kWarps : 0, wk = PAR ? warp % kWarps : warp;   // ring stage owned / 16-key slice inside a chunk
    const long long row0 = a.row_base + ((long long)r * a.H + h) * a.maxS;
    const int hdr = a.hd;                      // real head dim; q / slabs / out are [.., H * hdr]
    const int D = a.H * hdr;

    if (threadIdx.x == 0) {
        prefetch_map(&kmap);
        prefetch_map(&vmap);
        prefetch_map(&kmap16);
        prefetch_map(&vmap16);
        for (int s = 0; s < NST; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    // Programmatic dependent launch: the position counter and every key row of EARLIER steps were produced at least
    // one kernel before the predecessor, so their TMA loads are issued before the dependency wait.
    const int qpos = a.pos_rows ? a.pos_rows[r] : (a.pos_dev ? *a.pos_dev : 0) + a.pos_value;
    const int nkeys = FUSED ? qpos : qpos + 1;      // keys streamed from the cache (FUSED: the new key stays on chip)
    const int nchunks = (nkeys + kKC - 1) / kKC;

    auto issue = [&](int ci) {
        const int s = ci % NST;
        uint8_t* kt = tiles + s * 2 * TILE_BYTES;
        uint8_t* vt = kt + TILE_BYTES;
        const int row = (int)(row0 + (long long)ci * kKC);
        const int valid = nkeys - ci * kKC;            // keys of this chunk that exist
        if (valid >= kKC) {
            mbar_expect_tx(&full_bar[s], 2 * TILE_BYTES);
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                load_2d_hint(kt + sub * SUB_BYTES, &kmap, &full_bar[s], sub * 64, row, a.kvhint);
                load_2d_hint(vt + sub * SUB_BYTES, &vmap, &full_bar[s], sub * 64, row, a.kvhint);
            }
        } else {
            // tail chunk: 16-row boxes (one per warp's key group) so at most 15 rows beyond the context are read;
            // groups that are not loaded are never touched by the MMA loop (it skips j0 >= nkeys).
            const int n16 = (valid + 15) / 16;
            mbar_expect_tx(&full_bar[s], (uint32_t)(n16 * 2 * NSUB * 16 * 128));
            for (int i = 0; i < n16; ++i) {
#pragma unroll
                for (int sub = 0; sub < NSUB; ++sub) {
                    load_2d_hint(kt + sub * SUB_BYTES + i * 2048, &kmap16, &full_bar[s], sub * 64, row + 16 * i, a.kvhint);
                    load_2d_hint(vt + sub * SUB_BYTES + i * 2048, &vmap16, &full_bar[s], sub * 64, row + 16 * i, a.kvhint);
                }
            }
        }
    };
    const int npro = min(NST, nchunks);
    if (threadIdx.x == 0)
        for (int ci = 0; ci < npro; ++ci)
            if (FUSED || ci != nchunks - 1) issue(ci);
    // RoPE angles of this step's position: a table lookup that does not depend on the producer kernel -> issued before the wait
    float4 cs_pre = make_float4(1.f, 0.f, 1.f, 0.f);
    if (FUSED && threadIdx.x < 2 * HD / 4) {
        const int e = (threadIdx.x % (HD / 4)) * 4;
        if (e < hdr) cs_pre = __ldg(reinterpret_cast<const float4*>(a.freqs + ((size_t)qpos * (hdr / 2) + (e >> 1)) * 2));
    }
    lg_pdl_sync();
    if (!FUSED && threadIdx.x == 0 && nchunks - 1 < npro && nchunks > 0) issue(nchunks - 1);

    uint32_t qa[HD / 16][2];
    if (FUSED) {
        // ---- QKV epilogue for this (row, head): slab reduce -> dtype rounding -> RoPE -> cache write / smem
        for (int it = threadIdx.x; it < 3 * HD / 4; it += blockDim.x) {     // 48 / 96 items; a CTA may have only 64 threads (32-key stages)
            const int sec = it / (HD / 4), e = (it % (HD / 4)) * 4;
            const bool live = e < hdr;                 // dims [hdr, HD) are zero padding (hdr % 4 == 0)
            const size_t N3 = (size_t)3 * D, slab = (size_t)a.R * N3;
            const float* p = a.partial + (size_t)r * N3 + (size_t)sec * D + (size_t)h * hdr + e;
            float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                sv = *reinterpret_cast<const float4*>(p);
                for (int k = 1; k < a.ksplit; ++k) {
                    const float4 t = *reinterpret_cast<const float4*>(p + (size_t)k * slab);
                    sv.x += t.x; sv.y += t.y; sv.z += t.z; sv.w += t.w;
                }
            }
            float x0 = round_bf16(sv.x), x1 = round_bf16(sv.y), x2 = round_bf16(sv.z), x3 = round_bf16(sv.w);
            if (sec < 2 && live) {   // apply_rotary_emb (gpt.py:420-430): adjacent pairs, fp32, separate roundings
                // first pass: the angles were fetched before the dependency wait; a second pass (HD = 128 on 64 threads) loads its own
                const float4 cs = it == (int)threadIdx.x ? cs_pre
                                                         : __ldg(reinterpret_cast<const float4*>(a.freqs + ((size_t)qpos * (hdr / 2) + (e >> 1)) * 2));
                const float y0 = __fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y));
                const float y1 = __fadd_rn(__fmul_rn(x1, cs.x), __fmul_rn(x0, cs.y));
                const float y2 = __fsub_rn(__fmul_rn(x2, cs.z), __fmul_rn(x3, cs.w));
                const float y3 = __fadd_rn(__fmul_rn(x3, cs.z), __fmul_rn(x2, cs.w));
                x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            }
            uint2 pk;
            pk.x = pack_bf16(x0, x1);
            pk.y = pack_bf16(x2, x3);
            *reinterpret_cast<uint2*>(qbuf + sec * HD + e) = pk;
            if (sec > 0 && live) {
                bf16* cache = sec == 1 ? a.kcache : a.vcache;
                *reinterpret_cast<uint2*>(cache + (((size_t)r * a.H + h) * a.maxS + qpos) * a.hdp + e) = pk;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
            qa[kk][0] = 0; qa[kk][1] = 0;
            if (g == 0) {
                qa[kk][0] = *reinterpret_cast<const uint32_t*>(qbuf + kk * 16 + tg * 2);
                qa[kk][1] = *reinterpret_cast<const uint32_t*>(qbuf + kk * 16 + 8 + tg * 2);
            }
        }
    } else {
        // q as the A operand of m16n8k16: only MMA row 0 (lanes with g == 0) is real, the other 15 rows are zero
        const bf16* qp = a.q + (size_t)r * D + (size_t)h * hdr;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
            qa[kk][0] = 0; qa[kk][1] = 0;
            if (g == 0) {
                if (kk * 16 + tg * 2 < hdr) qa[kk][0] = *reinterpret_cast<const uint32_t*>(qp + kk * 16 + tg * 2);
                if (kk * 16 + 8 + tg * 2 < hdr) qa[kk][1] = *reinterpret_cast<const uint32_t*>(qp + kk * 16 + 8 + tg * 2);
            }
        }
    }
    const float* mrow = a.emb_mask ? a.emb_mask + (size_t)(r % a.B) * a.Tc : nullptr;

    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = 0.f; o[i][1] = 0.f; o[i][2] = 0.f; o[i][3] = 0.f; }
    float mx = -INFINITY, l = 0.f;

    for (int ci = grp; ci < nchunks; ci += PAR ? NST : 1) {
        const int s = ci % NST;
        mbar_wait(&full_bar[s], (uint32_t)((ci / NST) & 1));
        const uint32_t kt = smem_u32(tiles + s * 2 * TILE_BYTES);
        const uint32_t vt = kt + TILE_BYTES;
        const int j0 = ci * kKC + wk * 16;            // this warp's 16 keys
        if (j0 < nkeys) {                              // warp-uniform
            // ---- S = q K^T for 16 keys (two n8 tiles)
            float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) {
                uint32_t b0, b1, b2, b3;
                const int row = wk * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int chunk = (kk * 2 + ((lane >> 3) & 1)) & 7;
                ldsm_x4(kt + (kk / 4) * SUB_BYTES + swz(row, chunk), b0, b1, b2, b3);
                mma16816(sc[0], qa[kk][0], 0u, qa[kk][1], 0u, b0, b1);
                mma16816(sc[1], qa[kk][0], 0u, qa[kk][1], 0u, b2, b3);
            }
            // ---- online softmax on MMA row 0 (held by the quad g == 0; other quads carry zero rows)
            float pv[4];
            float lmax = -INFINITY;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = j0 + (i >> 1) * 8 + tg * 2 + (i & 1);
                bool vis = j < nkeys;
                if (vis && mrow && j < a.Tc && j != qpos) vis = mrow[j] != 0.f;
                pv[i] = vis ? sc[i >> 1][i & 1] * a.scale : -INFINITY;
                lmax = fmaxf(lmax, pv[i]);
            }
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, 1));
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, 2));
            const float mn = fmaxf(mx, lmax);
            float corr = 1.f, lsum = 0.f;
            if (mn != -INFINITY) {
                corr = __expf(mx - mn);
#pragma unroll
                for (int i = 0; i < 4; ++i) { pv[i] = __expf(pv[i] - mn); lsum += pv[i]; }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) pv[i] = 0.f;
            }
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
            l = l * corr + lsum;
            mx = mn;
#pragma unroll
            for (int i = 0; i < HD / 8; ++i) { o[i][0] *= corr; o[i][1] *= corr; }
            // ---- O += P V : P (bf16) is already in A-fragment layout (C layout of S == A layout of P)
            const uint32_t pa0 = pack_bf16(pv[0], pv[1]);   // keys tg*2, +1
            const uint32_t pa2 = pack_bf16(pv[2], pv[3]);   // keys 8 + tg*2, +1
#pragma unroll
            for (int np = 0; np < HD / 16; ++np) {
                uint32_t b0, b1, b2, b3;
                const int row = wk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int chunk = (np * 2 + (lane >> 4)) & 7;
                ldsm_x4_t(vt + (np / 4) * SUB_BYTES + swz(row, chunk), b0, b1, b2, b3);
                mma16816(o[2 * np], pa0, 0u, pa2, 0u, b0, b1);
                mma16816(o[2 * np + 1], pa0, 0u, pa2, 0u, b2, b3);
            }
        }
        if (PAR) {
            if (ci + NST < nchunks) {                   // uniform within the group: refill this group's stage
                asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kWarps * 32) : "memory");
                if (wk == 0 && lane == 0) issue(ci + NST);
            }
        } else {
            __syncthreads();                            // every warp is done with stage s
            if (threadIdx.x == 0 && ci + NST < nchunks) issue(ci + NST);
        }
    }

    // ---- merge the warps (each saw a disjoint key subset) and, when FUSED, the new key held in shared memory
    float* wrow = merge + warp * (HD + 2);
    if (g == 0) {
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            wrow[i * 8 + tg * 2] = o[i][0];
            wrow[i * 8 + tg * 2 + 1] = o[i][1];
        }
        if (tg == 0) { wrow[HD] = mx; wrow[HD + 1] = l; }
    }
    if (FUSED && warp == 0) {
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < HD / 32; ++i) {
            const int e = lane * (HD / 32) + i;
            dot = fmaf(__bfloat162float(qbuf[e]), __bfloat162float(qbuf[HD + e]), dot);
        }
        dot = warp_sum(dot);
        float* crow = merge + NW * (HD + 2);
#pragma unroll
        for (int i = 0; i < HD / 32; ++i) {
            const int e = lane * (HD / 32) + i;
            crow[e] = __bfloat162float(qbuf[2 * HD + e]);
        }
        if (lane == 0) { crow[HD] = dot * a.scale; crow[HD + 1] = 1.f; }
    }
    __syncthreads();
    bf16* op = a.out + (size_t)r * D + (size_t)h * hdr;
    for (int e = threadIdx.x; e < hdr; e += blockDim.x) {
        float M_ = -INFINITY;
#pragma unroll
        for (int w = 0; w < NSLOT; ++w) M_ = fmaxf(M_, merge[w * (HD + 2) + HD]);
        float L = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < NSLOT; ++w) {
            const float mw = merge[w * (HD + 2) + HD];
            const float c = mw == -INFINITY ? 0.f : __expf(mw - M_);
            L += merge[w * (HD + 2) + HD + 1] * c;
            O += merge[w * (HD + 2) + e] * c;
        }
        op[e] = __float2bfloat16_rn(O / L);
    }
}

// ---------------------------------------------------------------------------------------------------
// v2: persistent, one WARP per (row, head) work item. Every warp owns a private 3-stage TMA ring (its own
// mbarriers), so there is no CTA-wide synchronisation at all, the ring keeps streaming across item
// boundaries (the next item's first chunks are requested while the current item is still being reduced),
// and q for the next item is prefetched into registers. 4 warps x 3 stages x 16 KB = 192 KB of loads in
// flight per SM.
// ---------------------------------------------------------------------------------------------------
constexpr int kWarpsV2 = 4;
constexpr int kStagesV2 = 3;

template <int HD>
__global__ void __launch_bounds__(kWarpsV2 * 32, 1) attn_tma_v2_kernel(const __grid_constant__ CUtensorMap kmap,
                                                                       const __grid_constant__ CUtensorMap vmap, AttnTmaArgs a) {
    constexpr int NSUB = HD / 64;
    constexpr int SUB_BYTES = kKC * 128;
    constexpr int TILE_BYTES = NSUB * SUB_BYTES;
    constexpr int WARP_BYTES = kStagesV2 * 2 * TILE_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* tiles_all = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles_all + kWarpsV2 * WARP_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    uint8_t* tiles = tiles_all + warp * WARP_BYTES;
    uint64_t* full_bar = bars + warp * kStagesV2;
    if (lane == 0) {
        if (warp == 0) { prefetch_map(&kmap); prefetch_map(&vmap); }
        for (int s = 0; s < kStagesV2; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    lg_pdl_sync();     // K/V rows of this step and the position counter come from earlier kernels
    const int qpos = (a.pos_dev ? *a.pos_dev : 0) + a.pos_value;
    const int nkeys = qpos + 1;
    const int nchunks = (nkeys + kKC - 1) / kKC;
    const int D = a.H * HD;
    const int nitems = a.R * a.H;
    const int wid = blockIdx.x * kWarpsV2 + warp, nw = gridDim.x * kWarpsV2;
    const int my_items = wid < nitems ? (nitems - wid + nw - 1) / nw : 0;
    const long long total = (long long)my_items * nchunks;       // flattened (item, chunk) stream of this warp

    auto issue = [&](long long f) {                               // lane 0 only
        const int it = (int)(f / nchunks), ci = (int)(f - (long long)it * nchunks);
        const int item = wid + it * nw;
        const long long row0 = a.row_base + (long long)item * a.maxS;     // item = r*H + h
        const int s = (int)(f % kStagesV2);
        uint8_t* kt = tiles + s * 2 * TILE_BYTES;
        uint8_t* vt = kt + TILE_BYTES;
        mbar_expect_tx(&full_bar[s], 2 * TILE_BYTES);
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
            load_2d(kt + sub * SUB_BYTES, &kmap, &full_bar[s], sub * 64, (int)(row0 + (long long)ci * kKC));
            load_2d(vt + sub * SUB_BYTES, &vmap, &full_bar[s], sub * 64, (int)(row0 + (long long)ci * kKC));
        }
    };
    if (lane == 0)
        for (long long f = 0; f < total && f < kStagesV2; ++f) issue(f);

    auto load_q = [&](int item, uint32_t (*qa)[2]) {
        const bf16* qp = a.q + (size_t)(item / a.H) * D + (size_t)(item % a.H) * HD;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
            qa[kk][0] = 0; qa[kk][1] = 0;
            if (g == 0) {
                qa[kk][0] = *reinterpret_cast<const uint32_t*>(qp + kk * 16 + tg * 2);
                qa[kk][1] = *reinterpret_cast<const uint32_t*>(qp + kk * 16 + 8 + tg * 2);
            }
        }
    };
    uint32_t qa[HD / 16][2], qn[HD / 16][2];
    if (my_items > 0) load_q(wid, qa);

    long long f = 0;
    for (int it = 0; it < my_items; ++it) {
        const int item = wid + it * nw;
        const int r = item / a.H;
        if (it + 1 < my_items) load_q(item + nw, qn);             // prefetch the next item's query
        const float* mrow = a.emb_mask ? a.emb_mask + (size_t)(r % a.B) * a.Tc : nullptr;
        float o[HD / 8][2];
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) { o[i][0] = 0.f; o[i][1] = 0.f; }
        float mx = -INFINITY, l = 0.f;

        for (int ci = 0; ci < nchunks; ++ci, ++f) {
            const int s = (int)(f % kStagesV2);
            mbar_wait(&full_bar[s], (uint32_t)((f / kStagesV2) & 1));
            const uint32_t kt = smem_u32(tiles + s * 2 * TILE_BYTES);
            const uint32_t vt = kt + TILE_BYTES;
#pragma unroll 1
            for (int sub16 = 0; sub16 < kKC / 16; ++sub16) {
                const int j0 = ci * kKC + sub16 * 16;
                if (j0 >= nkeys) break;
                float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk) {
                    uint32_t b0, b1, b2, b3;
                    const int row = sub16 * 16 + (lane & 7) + ((lane >> 4) << 3);
                    const int chunk = (kk * 2 + ((lane >> 3) & 1)) & 7;
                    ldsm_x4(kt + (kk / 4) * SUB_BYTES + swz(row, chunk), b0, b1, b2, b3);
                    mma16816(sc[0], qa[kk][0], 0u, qa[kk][1], 0u, b0, b1);
                    mma16816(sc[1], qa[kk][0], 0u, qa[kk][1], 0u, b2, b3);
                }
                float pv[4];
                float lmax = -INFINITY;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = j0 + (i >> 1) * 8 + tg * 2 + (i & 1);
                    bool vis = j < nkeys;
                    if (vis && mrow && j < a.Tc && j != qpos) vis = mrow[j] != 0.f;
                    pv[i] = vis ? sc[i >> 1][i & 1] * a.scale : -INFINITY;
                    lmax = fmaxf(lmax, pv[i]);
                }
                lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, 1));
                lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, 2));
                const float mn = fmaxf(mx, lmax);
                float corr = 1.f, lsum = 0.f;
                if (mn != -INFINITY) {
                    corr = __expf(mx - mn);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { pv[i] = __expf(pv[i] - mn); lsum += pv[i]; }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) pv[i] = 0.f;
                }
                lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
                lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
                l = l * corr + lsum;
                mx = mn;
#pragma unroll
                for (int i = 0; i < HD / 8; ++i) { o[i][0] *= corr; o[i][1] *= corr; }
                const uint32_t pa0 = pack_bf16(pv[0], pv[1]);
                const uint32_t pa2 = pack_bf16(pv[2], pv[3]);
#pragma unroll
                for (int np = 0; np < HD / 16; ++np) {
                    uint32_t b0, b1, b2, b3;
                    const int row = sub16 * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                    const int chunk = (np * 2 + (lane >> 4)) & 7;
                    ldsm_x4_t(vt + (np / 4) * SUB_BYTES + swz(row, chunk), b0, b1, b2, b3);
                    float c0[4] = {o[2 * np][0], o[2 * np][1], 0.f, 0.f};
                    float c1[4] = {o[2 * np + 1][0], o[2 * np + 1][1], 0.f, 0.f};
                    mma16816(c0, pa0, 0u, pa2, 0u, b0, b1);
                    mma16816(c1, pa0, 0u, pa2, 0u, b2, b3);
                    o[2 * np][0] = c0[0]; o[2 * np][1] = c0[1];
                    o[2 * np + 1][0] = c1[0]; o[2 * np + 1][1] = c1[1];
                }
            }
            __syncwarp();                                          // all lanes are done reading stage s
            if (lane == 0 && f + kStagesV2 < total) issue(f + kStagesV2);
        }
        if (g == 0) {
            bf16* op = a.out + (size_t)r * D + (size_t)(item % a.H) * HD;
            const float inv = 1.0f / l;
#pragma unroll
            for (int i = 0; i < HD / 8; ++i)
                *reinterpret_cast<uint32_t*>(op + i * 8 + tg * 2) = pack_bf16(o[i][0] * inv, o[i][1] * inv);
        }
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) { qa[kk][0] = qn[kk][0]; qa[kk][1] = qn[kk][1]; }
    }
}

template <int HD>
int launch_v2(const CUtensorMap& kmap, const CUtensorMap& vmap, const AttnTmaArgs& a, cudaStream_t st) {
    constexpr int TILE_BYTES = (HD / 64) * kKC * 128;
    constexpr int WARP_BYTES = kStagesV2 * 2 * TILE_BYTES;
    const size_t smem = 1024 + (size_t)kWarpsV2 * WARP_BYTES + kWarpsV2 * kStagesV2 * sizeof(uint64_t);
    static DevOnce attr;
    static int sms = 148;
    if (lg_first_on_device(attr)) {
        LG_CUDA_OK(cudaFuncSetAttribute(attn_tma_v2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int nitems = a.R * a.H;
    const int ctas = std::min(sms, (nitems + kWarpsV2 - 1) / kWarpsV2);
    (void)lg_launch(attn_tma_v2_kernel<HD>, dim3(ctas), dim3(kWarpsV2 * 32), smem, st, kmap, vmap, a);
    LG_LAUNCH_CHECK();
    return 0;
}

template <int HD, bool FUSED, int NST = kStagesA, bool PAR = (NST > 2)>
int launch_t(const CUtensorMap& kmap, const CUtensorMap& vmap, const CUtensorMap& kmap16, const CUtensorMap& vmap16,
             const AttnTmaArgs& a, cudaStream_t st) {
    constexpr int TILE_BYTES = (HD / 64) * kKC * 128;
    constexpr int NW = PAR ? NST * kWarps : kWarps;
    const size_t smem = 1024 + (size_t)NST * 2 * TILE_BYTES + NST * sizeof(uint64_t) +
                        (NW + 1) * (HD + 2) * sizeof(float) + 3 * HD * sizeof(bf16) + 16;
    static DevOnce attr;
    if (lg_first_on_device(attr)) {
        LG_CUDA_OK(cudaFuncSetAttribute(attn_tma_kernel<HD, FUSED, NST, PAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(a.H, a.R);
    (void)lg_launch(attn_tma_kernel<HD, FUSED, NST, PAR>, dim3(grid), dim3(NW * 32), smem, st, kmap, vmap, kmap16, vmap16, a);
    LG_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// Prefill attention for the t2i condition (Tq = T <= 128 query positions per row, gpt.py:232-236 with the causal mask of
// gpt.py:354 and the emb_masks of generate.py:154-163): one CTA per (row, head). Q (post-RoPE, [R*T, D]) and the T freshly
// written K/V rows arrive by TMA (one box for Q, kKC-row boxes of the cache maps for K and V), S = Q K^T and O = P V run on
// mma.sync m16n8k16 with ldmatrix operands (each warp owns 16 query rows and skips the key blocks above its causal diagonal),
// softmax in fp32 registers, P reused from the score fragments. Replaces one CUDA-core CTA per (query, head).
// ---------------------------------------------------------------------------------------------------
constexpr int kPfRows = 128;                                   // query / key rows staged per (row, head)
constexpr int kPfBoxes = (kPfRows + kKC - 1) / kKC;            // K (or V) boxes of kKC rows

__global__ void __launch_bounds__(256, 2) attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap qmap,
                                                                 const __grid_constant__ CUtensorMap kmap,
                                                                 const __grid_constant__ CUtensorMap vmap, AttnTmaArgs a, int T) {
    constexpr int HD = 64;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* qs = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // [128][64] bf16, swizzle-128B
    uint8_t* ks = qs + kPfRows * 128;                                                              // [kPfBoxes * kKC][64]
    uint8_t* vs = ks + kPfBoxes * kKC * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(vs + kPfBoxes * kKC * 128);
    const int h = blockIdx.x, r = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    const long long row0 = a.row_base + ((long long)r * a.H + h) * a.maxS;
    const int D = a.H * HD;
    if (threadIdx.x == 0) {
        prefetch_map(&qmap);
        prefetch_map(&kmap);
        prefetch_map(&vmap);
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    lg_pdl_sync();                                   // q and the K/V rows were written by the QKV epilogue just before
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)(kPfRows * 128 + 2 * kPfBoxes * kKC * 128));
        load_2d(qs, &qmap, bar, h * HD, r * T);
        for (int i = 0; i < kPfBoxes; ++i) {
            load_2d(ks + i * kKC * 128, &kmap, bar, 0, (int)(row0 + (long long)i * kKC));
            load_2d(vs + i * kKC * 128, &vmap, bar, 0, (int)(row0 + (long long)i * kKC));
        }
    }
    mbar_wait(bar, 0);
    const int q0 = warp * 16;                        // this warp's query rows [q0, q0 + 16)
    if (q0 >= T) return;
    const uint32_t qb = smem_u32(qs), kb = smem_u32(ks), vb = smem_u32(vs);
    const int nkb = warp + 1;                        // 16-key blocks at or below the causal diagonal of these rows

    float sc[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { sc[i][0] = 0.f; sc[i][1] = 0.f; sc[i][2] = 0.f; sc[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(qb + swz(q0 + (lane & 15), (kk * 2 + (lane >> 4)) & 7), a0, a1, a2, a3);
#pragma unroll
        for (int kb16 = 0; kb16 < 8; ++kb16) {
            if (kb16 < nkb) {
                uint32_t b0, b1, b2, b3;
                const int row = kb16 * 16 + (lane & 7) + ((lane >> 4) << 3);
                ldsm_x4(kb + swz(row, (kk * 2 + ((lane >> 3) & 1)) & 7), b0, b1, b2, b3);
                mma16816(sc[2 * kb16], a0, a1, a2, a3, b0, b1);
                mma16816(sc[2 * kb16 + 1], a0, a1, a2, a3, b2, b3);
            }
        }
    }
    // ---- mask + softmax; this thread holds rows t0 = q0 + g (values [..][0..1]) and t1 = t0 + 8 (values [..][2..3])
    const float* mrow = a.emb_mask ? a.emb_mask + (size_t)(r % a.B) * a.Tc : nullptr;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int t = q0 + g + ((e >> 1) << 3), j = nt * 8 + tg * 2 + (e & 1);
            bool vis = j <= t && nt < 2 * nkb;
            if (vis && mrow && j < a.Tc && j != t) vis = mrow[j] != 0.f;
            const float x = vis ? sc[nt][e] * a.scale : -INFINITY;
            sc[nt][e] = x;
            mx[e >> 1] = fmaxf(mx[e >> 1], x);
        }
    }
    float l[2] = {0.f, 0.f};
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
    }
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float pv = __expf(sc[nt][e] - mx[e >> 1]);      // the diagonal key is always visible: mx is finite
            sc[nt][e] = pv;
            l[e >> 1] += pv;
        }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        l[hh] += __shfl_xor_sync(0xffffffffu, l[hh], 1);
        l[hh] += __shfl_xor_sync(0xffffffffu, l[hh], 2);
    }
    // ---- O = P V
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = 0.f; o[i][1] = 0.f; o[i][2] = 0.f; o[i][3] = 0.f; }
#pragma unroll
    for (int kb16 = 0; kb16 < 8; ++kb16) {
        if (kb16 < nkb) {
            const uint32_t pa0 = pack_bf16(sc[2 * kb16][0], sc[2 * kb16][1]), pa1 = pack_bf16(sc[2 * kb16][2], sc[2 * kb16][3]);
            const uint32_t pa2 = pack_bf16(sc[2 * kb16 + 1][0], sc[2 * kb16 + 1][1]), pa3 = pack_bf16(sc[2 * kb16 + 1][2], sc[2 * kb16 + 1][3]);
#pragma unroll
            for (int np = 0; np < HD / 16; ++np) {
                uint32_t b0, b1, b2, b3;
                const int row = kb16 * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                ldsm_x4_t(vb + swz(row, (np * 2 + (lane >> 4)) & 7), b0, b1, b2, b3);
                mma16816(o[2 * np], pa0, pa1, pa2, pa3, b0, b1);
                mma16816(o[2 * np + 1], pa0, pa1, pa2, pa3, b2, b3);
            }
        }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int t = q0 + g + hh * 8;
        if (t < T) {
            const float inv = 1.0f / l[hh];
            bf16* op = a.out + ((size_t)r * T + t) * D + (size_t)h * HD;
#pragma unroll
            for (int i = 0; i < HD / 8; ++i)
                *reinterpret_cast<uint32_t*>(op + i * 8 + tg * 2) = pack_bf16(o[i][2 * hh] * inv, o[i][2 * hh + 1] * inv);
        }
    }
}

// KV-cache tensor maps: the whole K (or V) region of the workspace as one [rows, hd] bf16 matrix
int attn_tma_make_map(void* map_out, const void* cache_base, long long total_rows, int hdp, int tail16) {
    // hdp = cache row width in elements (64, 128, or 112 for head_dim 100: the second 64-wide box then reads 112..127 as zeros)
    return tma::make_map_2d(reinterpret_cast<CUtensorMap*>(map_out), cache_base, (uint64_t)total_rows, (uint64_t)hdp, (uint64_t)hdp,
                            tail16 ? 16 : kKC, 64);
}

bool attn_tma_enabled() { return lg_env_flag("LG_ATTN_TMA", 1) != 0; }

bool attn_prefill_tc_supported(const AttnArgs& a) {
    return a.dtype == LG_DTYPE_BF16 && a.hd == 64 && (a.hdp == 0 || a.hdp == 64) && a.Tq > 1 && a.Tq <= kPfRows && a.kmap && a.vmap &&
           a.pos.dev == nullptr && a.pos.rows == nullptr && a.pos.value == 0 && a.R <= 65535 && a.maxS >= kPfBoxes * kKC &&
           lg_env_flag("LG_ATTN_PREFILL_TC", 1) != 0;
}

int launch_attention_prefill_tc(const AttnArgs& a, cudaStream_t st) {
    AttnTmaArgs t{};
    t.q = (const bf16*)a.q; t.out = (bf16*)a.out; t.R = a.R; t.H = a.H; t.maxS = a.maxS;
    t.row_base = a.cache_row_base; t.emb_mask = a.emb_mask; t.B = a.B; t.Tc = a.Tc; t.scale = a.scale;
    t.hd = a.hd; t.hdp = a.hd;
    CUtensorMap qmap;
    LG_TRY(tma::make_map_2d(&qmap, a.q, (uint64_t)a.R * a.Tq, (uint64_t)a.H * a.hd, (uint64_t)a.H * a.hd, kPfRows, 64));
    const size_t smem = 1024 + (size_t)kPfRows * 128 + 2 * (size_t)kPfBoxes * kKC * 128 + 16;
    static DevOnce attr;
    if (lg_first_on_device(attr)) {
        LG_CUDA_OK(cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(a.H, a.R);
    (void)lg_launch(attn_prefill_tc_kernel, dim3(grid), dim3(256), smem, st, qmap, *reinterpret_cast<const CUtensorMap*>(a.kmap),
                    *reinterpret_cast<const CUtensorMap*>(a.vmap), t, a.Tq);
    LG_LAUNCH_CHECK();
    return 0;
}

bool attn_tma_supported(const AttnArgs& a) {
    const bool shape = a.hd == 64 || a.hd == 128 || (a.hd == 100 && a.hdp == 112);
    return a.dtype == LG_DTYPE_BF16 && a.Tq == 1 && shape && a.kmap && a.vmap && a.kmap16 && a.vmap16 && a.R <= 65535;
}

int launch_attention_tma(const AttnArgs& a, cudaStream_t st) {
    AttnTmaArgs t;
    t.q = (const bf16*)a.q; t.out = (bf16*)a.out; t.R = a.R; t.H = a.H; t.maxS = a.maxS;
    t.pos_dev = a.pos.dev; t.pos_value = a.pos.value; t.pos_rows = a.pos.rows; t.row_base = a.cache_row_base;
    t.emb_mask = a.emb_mask; t.B = a.B; t.Tc = a.Tc; t.scale = a.scale;
    t.partial = a.qkv_partial; t.ksplit = a.qkv_ksplit; t.freqs = a.freqs;
    t.kcache = (bf16*)const_cast<void*>(a.kcache); t.vcache = (bf16*)const_cast<void*>(a.vcache);
    t.kvhint = (lg_env_flag("LG_L2_HINT", 1) & 1) ? tma::kL2EvictFirst : 0ull;
    t.hd = a.hd; t.hdp = a.hdp ? a.hdp : a.hd;
    const CUtensorMap& km = *reinterpret_cast<const CUtensorMap*>(a.kmap);
    const CUtensorMap& vm = *reinterpret_cast<const CUtensorMap*>(a.vmap);
    const CUtensorMap& km16 = *reinterpret_cast<const CUtensorMap*>(a.kmap16);
    const CUtensorMap& vm16 = *reinterpret_cast<const CUtensorMap*>(a.vmap16);
    if (a.qkv_partial) {     // fused QKV epilogue
        // few (row, head) items (batch-1 latency path): a 6-stage ring holds a whole 288-key context, so every K/V byte is
        // requested before the dependency wait instead of two stages at a time
        if (a.hd == 64 && a.R * a.H <= 2 * 148 && lg_env_flag("LG_ATTN_DEEP", 1)) return launch_t<64, true, kDeepStages>(km, vm, km16, vm16, t, st);
        // deeper sequential ring (A/B switch): more keys requested before the dependency wait, fewer refill round trips
        const int nst = lg_env_flag("LG_ATTN_NST", 2);
        if (a.hd == 64 && nst == 3) return launch_t<64, true, 3, false>(km, vm, km16, vm16, t, st);
        if (a.hd == 64 && nst == 4) return launch_t<64, true, 4, false>(km, vm, km16, vm16, t, st);
        if (a.hd == 64) return launch_t<64, true>(km, vm, km16, vm16, t, st);
        return launch_t<128, true>(km, vm, km16, vm16, t, st);
    }
    // v2 (persistent warp-per-item, LG_ATTN_V2=1) measured SLOWER than the CTA-per-item kernel on B200 (25.7 vs
    // 19.1 us at R=128, c=128: with one warp per scheduler the ldmatrix->mma->softmax chain is latency-bound), so it
    // stays opt-in; profiles/ keeps both ncu captures.
    const bool v2 = lg_env_flag("LG_ATTN_V2", 0) && a.R * a.H >= 4 * 148 && a.hd == 64 && !a.pos.rows;   // (hd 64 only)
    if (v2) return launch_v2<64>(km, vm, t, st);
    if (a.hd == 64) return launch_t<64, false>(km, vm, km16, vm16, t, st);
    return launch_t<128, false>(km, vm, km16, vm16, t, st);
}
