// NN potential for (H2O)2: Coulomb-matrix descriptor + the shipped 15-120-120-120-1 Keras MLP
// (TensorflowPots/call_sample_model.py:4-9; descriptor: tensorflow_descriptors/distance_descriptors.py
// :102-113,154-168; model layout read from sample_h4o2_nn.h5: Dense(120,swish) x3, Dense(1,relu), float32).
//
// Three kernels, newest last in this file:
//   k_nn_h4o2      float32 FMA on the CUDA cores (cross-check path, PVD_NN_FP32=1): a CTA of 128 threads owns 64
//                  walkers; activations live in shared memory k-major so that one LDS.128 feeds four FMAs
//   k_nn_h4o2_tc   tcgen05, activations staged through shared memory, one tile in flight (PVD_NN_TC1=1)
//   k_nn_h4o2_tc2  tcgen05, activations in TMEM, two tiles in flight (default)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "pvd_step.cuh"

constexpr int NN_IN = 15, NN_H = 120, NN_TILE = 64, NN_THREADS = 128;
constexpr int NN_W0 = 0, NN_B0 = NN_W0 + NN_IN * NN_H, NN_W1 = NN_B0 + NN_H, NN_B1 = NN_W1 + NN_H * NN_H,
              NN_W2 = NN_B1 + NN_H, NN_B2 = NN_W2 + NN_H * NN_H, NN_W3 = NN_B2 + NN_H, NN_B3 = NN_W3 + NN_H,
              NN_NPARAM = NN_B3 + 1;
static_assert(NN_NPARAM == 31081, "packed weight vector size");

__device__ __forceinline__ float swish_f32(float z) { return z / (1.0f + expf(-z)); }

// one dense layer on the tile: out[j][w] = act( b[j] + sum_k in[k][w] * W[k][j] )
template <int K, bool SWISH>
__device__ __forceinline__ void nn_layer(const float *__restrict__ W, const float *__restrict__ b,
                                         const float (*in)[NN_TILE], float (*out)[NN_TILE])
{
    const int j = threadIdx.x;
    if (j < NN_H) {
        float acc[NN_TILE];
        const float bj = b[j];
#pragma unroll
        for (int w = 0; w < NN_TILE; ++w) acc[w] = bj;
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            const float wkj = __ldg(&W[k * NN_H + j]);
            const float4 *row = reinterpret_cast<const float4 *>(in[k]);
#pragma unroll
            for (int q = 0; q < NN_TILE / 4; ++q) {
                const float4 a = row[q];
                acc[4 * q + 0] = fmaf(a.x, wkj, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(a.y, wkj, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(a.z, wkj, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(a.w, wkj, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int w = 0; w < NN_TILE; ++w) out[j][w] = SWISH ? swish_f32(acc[w]) : acc[w];
    }
}

// coords: SOA (stride cap) when soa != 0, else AoS (n,6,3).  n from the state when st != nullptr.
__global__ void __launch_bounds__(NN_THREADS) k_nn_h4o2(const double *__restrict__ xyz, int soa, long long cap, const DevState *st, int parity,
                                                        long long n_fixed, const float *__restrict__ P, double *__restrict__ v,
                                                        double *__restrict__ desc_out)
{
    extern __shared__ __align__(16) float nn_smem[];
    float (*s_a)[NN_TILE] = reinterpret_cast<float (*)[NN_TILE]>(nn_smem);
    float (*s_b)[NN_TILE] = reinterpret_cast<float (*)[NN_TILE]>(nn_smem + NN_H * NN_TILE);
    const long long n = st ? st[parity].n : n_fixed;
    if (st && st[parity].err) return;
    const double zs[6] = {8.0, 1.0, 1.0, 8.0, 1.0, 1.0};
    for (long long base = (long long)blockIdx.x * NN_TILE; base < n; base += (long long)gridDim.x * NN_TILE) {
        // descriptor: Z_i Z_j / r_ij over itertools.combinations(range(6), 2), fp64 then cast to fp32
        for (int t = threadIdx.x; t < NN_TILE * NN_IN; t += NN_THREADS) {
            const int w = t % NN_TILE, p = t / NN_TILE;
            const long long i = base + w;
            float feat = 0.0f;
            if (i < n) {
                int a = 0, rem = p;
                while (rem >= 5 - a) { rem -= 5 - a; ++a; }
                const int bb = a + 1 + rem;
                double d2 = 0.0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const double xa = soa ? xyz[(a * 3 + d) * cap + i] : xyz[i * 18 + a * 3 + d];
                    const double xb = soa ? xyz[(bb * 3 + d) * cap + i] : xyz[i * 18 + bb * 3 + d];
                    d2 += (xa - xb) * (xa - xb);
                }
                const double c = zs[a] * zs[bb] / sqrt(d2);
                if (desc_out) desc_out[i * NN_IN + p] = c;
                feat = (float)c;
            }
            s_a[p][w] = feat;
        }
        __syncthreads();
        nn_layer<NN_IN, true>(P + NN_W0, P + NN_B0, s_a, s_b);
        __syncthreads();
        nn_layer<NN_H, true>(P + NN_W1, P + NN_B1, s_b, s_a);
        __syncthreads();
        nn_layer<NN_H, true>(P + NN_W2, P + NN_B2, s_a, s_b);
        __syncthreads();
        if (threadIdx.x < NN_TILE) {
            const int w = threadIdx.x;
            float acc = P[NN_B3];
            for (int k = 0; k < NN_H; ++k) acc = fmaf(s_b[k][w], __ldg(&P[NN_W3 + k]), acc);
            acc = fmaxf(acc, 0.0f);                                         // relu
            const long long i = base + w;
            // Constants.convert(float32 array, 'wavenumbers'): NumPy keeps float32 for array * Python float
            if (i < n) v[i] = (double)(acc * 4.556335281212229e-6f);
        }
        __syncthreads();
    }
}

// =====================================================================================================
// tcgen05 version of the MLP (the only tensor-core use in the library).
//
// Tile: 128 walkers per CTA (UMMA M = 128, one walker per TMEM lane / per thread), N = 128 output
// neurons (120 padded), fp32 accumulators in TMEM (128 columns).  fp32 accuracy is kept by splitting
// every activation and every weight into two fp16 pieces and issuing the four
// cross terms of a two-piece split (x = x1 + x2: a1w1, a1w2, a2w1, a2w2; residual ~2^-22, tighter than the TF32
// path TensorFlow itself takes for float32 matmuls on tensor-core GPUs) as kind::f16 MMAs with
// K = 16.  Operands are K-major, un-swizzled "core matrix" layout in shared memory:
//     byte(row, k) = (k/8) * (128*16) + (row/8) * 128 + (row%8) * 16 + (k%8) * 2      (LBO = 2048, SBO = 128)
// so that thread `row` writes its eight consecutive k values as one conflict-free 16-byte store.
// All three weight images (pre-formatted on the host, 136 KB) are loaded once per persistent CTA by bulk TMA copies;
// per layer 4 x K/16 MMAs are issued by one thread,
// tcgen05.commit -> mbarrier, then every thread drains its own TMEM lane (tcgen05.ld 32x32b.x32), adds
// the bias, applies swish, splits into fp16 pieces and stores the next layer's A operand.
// =====================================================================================================
constexpr int TC_M = 128, TC_N = 128, TC_THREADS = 512;
constexpr int TC_PIECE_BYTES_K128 = 128 * 128 * 2;          // one fp16 operand piece, K = 128
constexpr int TC_PIECE_BYTES_K16 = 128 * 16 * 2;
constexpr int TC_NPIECE = 2;                                 // fp16 pieces per operand (x = x1 + x2, residual ~2^-22)
constexpr int TC_IMG_L0 = TC_NPIECE * TC_PIECE_BYTES_K16;    // weights image, layer 0 (K padded 15 -> 16)
constexpr int TC_IMG_L12 = TC_NPIECE * TC_PIECE_BYTES_K128;  // weights image, layers 1 and 2
constexpr int TC_IMG_TOTAL = TC_IMG_L0 + 2 * TC_IMG_L12;     // bytes of pre-formatted fp16 weight images
constexpr int TC_VEC_FLOATS = 3 * 128 + 128 + 4;             // b0,b1,b2 (padded), W3 (padded), b3
constexpr size_t TC_SMEM_BYTES = 3 * TC_IMG_L12 + TC_IMG_L0 + 64 + TC_VEC_FLOATS * 4 + 128 * 4 * 4;   // A, W1, W2, W0, barriers, vectors

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr)
{
    // start address, LBO = 2048 B, SBO = 128 B (all >> 4), version 1 (Blackwell), no swizzle
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                 "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}
// bulk global -> shared copy completing on an mbarrier (TMA engine, 1-D)
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// split two floats into two packed fp16x2 pieces each (round to nearest even at both levels): x = p1 + p2 up to ~2^-22.
// fp16 rather than bf16 pieces: 11 + 11 mantissa bits instead of 8 + 8 at the same tensor-core cost; the values of this
// network (Coulomb features <= 64 / r, swish activations, trained weights) sit well inside the fp16 range, and the
// Coulomb features are clamped to +-65504 so that two colliding atoms cannot turn into inf (activations are not: an
// activation that large means the walker is far outside anything the network was trained on).
template <bool CLAMP = true>
__device__ __forceinline__ void split2x2(float x0, float x1, uint32_t &p1, uint32_t &p2)
{
    if (CLAMP) {
        x0 = fminf(fmaxf(x0, -65504.0f), 65504.0f);
        x1 = fminf(fmaxf(x1, -65504.0f), 65504.0f);
    }
    __half2 b = __floats2half2_rn(x0, x1);
    p1 = *reinterpret_cast<uint32_t *>(&b);
    const float2 f = __half22float2(b);
    b = __floats2half2_rn(x0 - f.x, x1 - f.y);
    p2 = *reinterpret_cast<uint32_t *>(&b);
}
// z * sigmoid(z) with the two SFU ops issued as single instructions (flush-to-zero forms: no denormal fix-up code around
// them): 6 instructions per activation instead of the ~17 __expf / __fdividef expand to without -use_fast_math.
__device__ __forceinline__ float swish_fast(float z)
{
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return z * r;
}
// Two activations with ONE reciprocal: z0 / (1 + e0) = z0 (1 + e1) / ((1 + e0)(1 + e1)): three SFU instructions per pair instead
// of four (the XU pipe is this kernel's busiest: 58 % in profiles/r01_profiles.md) for two caps and four more FMA-pipe
// multiplies (the exponent is capped at 60 so that the product of the denominators stays finite).  TRIED AND REJECTED
// (profiles/r02_s3.txt): the stand-alone kernel went from 0.704 to 0.775 ms per 2e6 walkers -- the epilogue is bound by issue
// slots and the dependency chain of a lane's 16 activations, not by the SFU's throughput.  Kept behind the switch as the record.
#ifndef PVD_NN_SWISH_PAIR
#define PVD_NN_SWISH_PAIR 0
#endif
__device__ __forceinline__ void swish_fast2(float z0, float z1, float &h0, float &h1)
{
#if PVD_NN_SWISH_PAIR
    float e0, e1, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(z0 * -1.4426950408889634f, 60.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(z1 * -1.4426950408889634f, 60.0f)));
    const float a0 = 1.0f + e0, a1 = 1.0f + e1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a0 * a1));
    h0 = (z0 * a1) * r;
    h1 = (z1 * a0) * r;
#else
    h0 = swish_fast(z0);
    h1 = swish_fast(z1);
#endif
}
// eight consecutive k values of one row -> one 16-byte store per piece
__device__ __forceinline__ void store_chunk(unsigned char *sA, uint32_t off, const float (&h)[8])
{
    uint32_t a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split2x2(h[2 * q], h[2 * q + 1], a[q], b[q]);
    *reinterpret_cast<uint4 *>(sA + 0 * TC_PIECE_BYTES_K128 + off) = make_uint4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<uint4 *>(sA + 1 * TC_PIECE_BYTES_K128 + off) = make_uint4(b[0], b[1], b[2], b[3]);
}

// images: [L0 | L1 | L2] fp16 weight images (see host builder); vecs: b0[128] b1[128] b2[128] W3[128] b3
// 512 threads: warp w owns TMEM lanes (walkers) 32*(w%4).. and the 32-column quarter w/4 of the 128 outputs,
// i.e. four threads share a walker during the epilogues (4 warps per scheduler hide the TMEM / SFU latencies).
__global__ void __launch_bounds__(TC_THREADS, 1)
k_nn_h4o2_tc(const double *__restrict__ xyz, int soa, long long cap, const DevState *st, int parity, long long n_fixed,
             const unsigned char *__restrict__ images, const float *__restrict__ vecs, double *__restrict__ v)
{
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char *sA = tc_smem;                                   // 2 pieces x 32 KB: activations (A operand)
    unsigned char *sW = tc_smem + TC_IMG_L12;                      // 2 x (2 pieces x 32 KB): weights of layers 1 and 2, resident
    unsigned char *sW0 = tc_smem + 3 * TC_IMG_L12;                 // 8 KB: weights of layer 0, resident
    uint64_t *bar = reinterpret_cast<uint64_t *>(sW0 + TC_IMG_L0);  // [0] MMA done, [2] weights landed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sW0 + TC_IMG_L0 + 32);
    float *s_vec = reinterpret_cast<float *>(sW0 + TC_IMG_L0 + 64);                 // biases, W3, b3
    float *s_out = s_vec + TC_VEC_FLOATS;                                          // [4 quarters][128 rows] partial outputs
    const long long n = st ? st[parity].n : n_fixed;
    if (st && st[parity].err) return;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int row = (warp & 3) * 32 + lane;                        // walker of this thread inside the tile
    const int quarter = warp >> 2;                                 // which 32 output columns this thread drains
    const uint32_t mma_bar = smem_u32(&bar[0]), w0_bar = smem_u32(&bar[2]);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mma_bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(w0_bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int k = t; k < TC_VEC_FLOATS; k += TC_THREADS) s_vec[k] = vecs[k];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (t == 0) {                                                  // all weights: loaded once per (persistent) CTA by the TMA engine
        mbar_expect_tx(w0_bar, TC_IMG_TOTAL);
        bulk_load(smem_u32(sW0), images, TC_IMG_L0, w0_bar);
        for (int q = 0; q < 2 * TC_NPIECE; ++q)
            bulk_load(smem_u32(sW + q * TC_PIECE_BYTES_K128), images + TC_IMG_L0 + q * TC_PIECE_BYTES_K128, TC_PIECE_BYTES_K128, w0_bar);
    }
    // instruction descriptor: D = F32, A = B = F16 (format 0), K-major both, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const double zs[6] = {8.0, 1.0, 1.0, 8.0, 1.0, 1.0};
    uint32_t mma_phase = 0;
    const uint32_t row_off = (uint32_t)((row >> 3) * 128 + (row & 7) * 16);        // this walker's row inside a k-chunk
    const int pa[4] = {0, 0, 1, 1}, pb[4] = {0, 1, 0, 1};         // cross terms a1w1 + a1w2 + a2w1 + a2w2

    for (long long base = (long long)blockIdx.x * TC_M; base < n; base += (long long)gridDim.x * TC_M) {
        const long long i = base + row;
        // ---- layer-0 A operand: Coulomb descriptor, 15 features + 1 zero pad; quarters 0 and 1 take 8 features each
        if (quarter < 2) {
            float feat[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) feat[p] = 0.0f;
            if (i < n) {
                double c[18];
#pragma unroll
                for (int k = 0; k < 18; ++k) c[k] = soa ? xyz[k * cap + i] : xyz[i * 18 + k];
                int p = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = a + 1; b < 6; ++b) {
                        if ((p >> 3) == quarter) {
                            const double dx = c[3 * a] - c[3 * b], dy = c[3 * a + 1] - c[3 * b + 1], dz = c[3 * a + 2] - c[3 * b + 2];
                            feat[p & 7] = (float)(zs[a] * zs[b] * rsqrt(dx * dx + dy * dy + dz * dz));
                        }
                        ++p;
                    }
            }
            store_chunk(sA, (uint32_t)quarter * 2048u + row_off, feat);
        }
        float out = 0.0f;
#pragma unroll 1
        for (int layer = 0; layer < 3; ++layer) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // A stores (generic proxy) -> visible to the MMA (async proxy)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (t == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int kblocks = layer == 0 ? 1 : 8;                        // K / 16
                const unsigned char *wb = layer == 0 ? sW0 : sW + (layer - 1) * TC_IMG_L12;
                const int wpiece = layer == 0 ? TC_PIECE_BYTES_K16 : TC_PIECE_BYTES_K128;
                mbar_wait(w0_bar, 0);                                          // completes once; later waits return immediately
                uint32_t accumulate = 0;
                for (int term = 0; term < 4; ++term)
                    for (int kb = 0; kb < kblocks; ++kb) {
                        const uint64_t da = umma_desc(smem_u32(sA + pa[term] * TC_PIECE_BYTES_K128 + kb * 4096));
                        const uint64_t db = umma_desc(smem_u32(wb + pb[term] * wpiece + kb * 4096));
                        umma_f16(tmem, da, db, idesc, accumulate);
                        accumulate = 1;
                    }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mma_bar) : "memory");
            }
            mbar_wait(mma_bar, mma_phase);
            mma_phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // ---- epilogue: this thread's TMEM lane, its 32 columns
            const float *bias = s_vec + layer * 128;
            uint32_t r[32];
            const int col0 = quarter * 32;
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                           "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                           "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int j = col0 + ch * 8 + e;
                    h[e] = swish_fast(__uint_as_float(r[ch * 8 + e]) + bias[j]);
                    if (layer == 2) out = fmaf(h[e], s_vec[3 * 128 + j], out);
                }
                if (layer < 2) store_chunk(sA, (uint32_t)((col0 >> 3) + ch) * 2048u + row_off, h);
            }
        }
        // the four column quarters of a walker live in warps w, w+4, w+8, w+12: combine through shared memory
        s_out[quarter * 128 + row] = out;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (quarter == 0 && i < n) {
            const float e = fmaxf(s_out[row] + s_out[128 + row] + s_out[256 + row] + s_out[384 + row] + s_vec[3 * 128 + 128], 0.0f);   // + b3, relu
            v[i] = (double)(e * 4.556335281212229e-6f);
        }
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}

// =====================================================================================================
// Pipelined tcgen05 version: activations never touch shared memory.  The A operand of every layer lives in
// TMEM (tcgen05.mma accepts A from tensor memory), written by the epilogue threads with tcgen05.st, so the
// 64 KB activation buffer is gone and TWO tiles of 128 walkers are in flight per SM:
//     TMEM columns [256 g, 256 g + 128)        fp32 accumulator of tile slot g        (g = 0, 1)
//                  [256 g + 128 + 64 p, + 64)  fp16 piece p of the slot's activations (two fp16 per column)
// Warps 0-7 own slot 0, warps 8-15 slot 1; each group has its own named barrier, its own MMA-completion
// mbarrier and its own issuing thread, so while one group applies bias + swish on the CUDA cores the tensor
// core works on the other group's layer.  Inside a group warp w owns TMEM lanes 32 (w % 4).. (one walker per
// lane) and the 64-column half (w / 4) % 2 of the 128 outputs.
// =====================================================================================================
constexpr int TC2_TERMS = 3;                                  // cross terms issued: a1w1 + a1w2 + a2w1 (a2w2 ~ 2^-22, below the split residual;
                                                              // measured: same error against float64 as with all four)
constexpr size_t TC2_SMEM_BYTES = 2 * TC_IMG_L12 + TC_IMG_L0 + 64 + TC_VEC_FLOATS * 4 + 2 * 4 * 128 * 4;
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st2(uint32_t taddr, const uint32_t (&r)[2])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

// THREADS = 512: a group is 8 warps, a thread drains 64 columns of its walker; THREADS = 1024: 16 warps, 32 columns.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
k_nn_h4o2_tc2(const double *__restrict__ xyz, int soa, long long cap, const DevState *st, int parity, long long n_fixed,
              const unsigned char *__restrict__ images, const float *__restrict__ vecs, double *__restrict__ v, int nterms)
{
    constexpr int GT = THREADS / 2;                                // threads per group (tile slot)
    constexpr int NQ = GT / 128;                                   // column parts per walker: 2 or 4
    constexpr int COLS = 128 / NQ;                                 // output columns per thread
    constexpr int NFEAT = 16 / NQ;                                 // descriptor features per thread (padded 15 -> 16)
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char *sW = tc_smem;                                   // 2 x (2 pieces x 32 KB): weights of layers 1 and 2, resident
    unsigned char *sW0 = tc_smem + 2 * TC_IMG_L12;                 // 8 KB: weights of layer 0, resident
    uint64_t *bar = reinterpret_cast<uint64_t *>(sW0 + TC_IMG_L0);  // [0], [1]: MMA done (slot 0, 1); [2] weights landed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sW0 + TC_IMG_L0 + 32);
    float *s_vec = reinterpret_cast<float *>(sW0 + TC_IMG_L0 + 64);                 // biases, W3, b3
    float *s_out = s_vec + TC_VEC_FLOATS;                                          // [slot][part][128 rows] partial outputs
    const long long n = st ? st[parity].n : n_fixed;
    if (st && st[parity].err) return;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int g = t / GT;                                          // tile slot / thread group
    const int part = ((t % GT) >> 7);                              // which COLS output columns this thread drains
    const int row = (warp & 3) * 32 + lane;                        // walker of this thread inside the tile (== its TMEM lane)
    const bool leader = (t % GT) == 0;
    const uint32_t mma_bar = smem_u32(&bar[g]), w0_bar = smem_u32(&bar[2]);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(w0_bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int k = t; k < TC_VEC_FLOATS; k += THREADS) s_vec[k] = vecs[k];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (t == 0) {                                                  // all weights: loaded once per (persistent) CTA by the TMA engine
        mbar_expect_tx(w0_bar, TC_IMG_TOTAL);
        bulk_load(smem_u32(sW0), images, TC_IMG_L0, w0_bar);
        for (int q = 0; q < 2 * TC_NPIECE; ++q)
            bulk_load(smem_u32(sW + q * TC_PIECE_BYTES_K128), images + TC_IMG_L0 + q * TC_PIECE_BYTES_K128, TC_PIECE_BYTES_K128, w0_bar);
    }
    // instruction descriptor: D = F32, A = B = F16 (format 0), K-major both, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const double zs[6] = {8.0, 1.0, 1.0, 8.0, 1.0, 1.0};
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;   // TMEM lane field of this warp
    const uint32_t tD = tmem + (uint32_t)(256 * g);                // accumulator of this slot
    const uint32_t tA = tD + 128u;                                 // activation pieces of this slot (64 columns each)
    uint32_t mma_phase = 0;

    for (long long base = ((long long)blockIdx.x * 2 + g) * TC_M; base < n; base += (long long)gridDim.x * 2 * TC_M) {
        const long long i = base + row;
        // ---- layer-0 A operand: Coulomb descriptor, 15 features + 1 zero pad; the parts of a walker share the features
        {
            float feat[NFEAT];
#pragma unroll
            for (int p = 0; p < NFEAT; ++p) feat[p] = 0.0f;
            if (i < n) {
                double c[18];
#pragma unroll
                for (int k = 0; k < 18; ++k) c[k] = soa ? xyz[k * cap + i] : xyz[i * 18 + k];
                int p = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = a + 1; b < 6; ++b) {
                        if (p / NFEAT == part) {
                            const double dx = c[3 * a] - c[3 * b], dy = c[3 * a + 1] - c[3 * b + 1], dz = c[3 * a + 2] - c[3 * b + 2];
                            feat[p % NFEAT] = (float)(zs[a] * zs[b] * rsqrt(dx * dx + dy * dy + dz * dz));
                        }
                        ++p;
                    }
            }
            uint32_t w1[NFEAT / 2], w2[NFEAT / 2];
#pragma unroll
            for (int q = 0; q < NFEAT / 2; ++q) split2x2(feat[2 * q], feat[2 * q + 1], w1[q], w2[q]);
            const uint32_t fa = tA + lane_sel + (uint32_t)(part * (NFEAT / 2));
            if constexpr (NFEAT == 8) { tmem_st4(fa, w1); tmem_st4(fa + 64u, w2); }
            else { tmem_st2(fa, w1); tmem_st2(fa + 64u, w2); }
        }
        float out = 0.0f;
#pragma unroll 1
        for (int layer = 0; layer < 3; ++layer) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");       // this thread's A stores have landed in TMEM
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, %1;" :: "r"(1 + g), "n"(GT) : "memory");
            if (leader) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int kblocks = layer == 0 ? 1 : 8;                        // K / 16
                const unsigned char *wb = layer == 0 ? sW0 : sW + (layer - 1) * TC_IMG_L12;
                const int wpiece = layer == 0 ? TC_PIECE_BYTES_K16 : TC_PIECE_BYTES_K128;
                mbar_wait(w0_bar, 0);                                          // completes once; later waits return immediately
                uint32_t accumulate = 0;
                for (int term = 0; term < nterms; ++term) {                    // cross terms a1w1 + a1w2 + a2w1 (+ a2w2)
                    const int pa = term >> 1, pb = term & 1;
                    const uint64_t db0 = umma_desc(smem_u32(wb + pb * wpiece));
                    for (int kb = 0; kb < kblocks; ++kb) {
                        umma_f16_ts(tD, tA + (uint32_t)(pa * 64 + kb * 8), db0 + (uint64_t)(kb * (4096 >> 4)), idesc, accumulate);
                        accumulate = 1;
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mma_bar) : "memory");
            }
            mbar_wait(mma_bar, mma_phase);
            mma_phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // ---- epilogue: this thread's TMEM lane, its COLS columns, 16 at a time, the next chunk's load in flight
            const float *bias = s_vec + layer * 128;
            uint32_t r[2][16];
            tmem_ld16_nowait(tD + lane_sel + (uint32_t)(part * COLS), r[0]);
#pragma unroll
            for (int ch = 0; ch < COLS / 16; ++ch) {
                const int col0 = part * COLS + ch * 16;
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ch + 1 < COLS / 16) tmem_ld16_nowait(tD + lane_sel + (uint32_t)(col0 + 16), r[(ch + 1) & 1]);
                float h[16];
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    swish_fast2(__uint_as_float(r[ch & 1][e]) + bias[col0 + e], __uint_as_float(r[ch & 1][e + 1]) + bias[col0 + e + 1],
                                h[e], h[e + 1]);
                    if (layer == 2) out = fmaf(h[e + 1], s_vec[3 * 128 + col0 + e + 1], fmaf(h[e], s_vec[3 * 128 + col0 + e], out));
                }
                if (layer < 2) {
                    uint32_t w1[8], w2[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) split2x2<false>(h[2 * q], h[2 * q + 1], w1[q], w2[q]);
                    tmem_st8(tA + lane_sel + (uint32_t)(col0 >> 1), w1);
                    tmem_st8(tA + 64u + lane_sel + (uint32_t)(col0 >> 1), w2);
                }
            }
        }
        // the column parts of a walker live in different warps of the group: combine through shared memory
        s_out[(g * NQ + part) * 128 + row] = out;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, %1;" :: "r"(1 + g), "n"(GT) : "memory");
        if (part == 0 && i < n) {
            float e = s_vec[3 * 128 + 128];                                     // b3
#pragma unroll
            for (int q = 0; q < NQ; ++q) e += s_out[(g * NQ + q) * 128 + row];
            v[i] = (double)(fmaxf(e, 0.0f) * 4.556335281212229e-6f);            // relu, cm-1 -> Hartree in float32 like the reference
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

// ---------------------------------------------------------------- host side: weight images + launch
struct NNDeviceWeights {
    float *packed = nullptr;            // raw packed float32 weights (CUDA-core kernel)
    unsigned char *images = nullptr;    // fp16-split, core-matrix formatted weight images (tcgen05 kernel)
    float *vecs = nullptr;              // biases, W3, b3
};

static inline unsigned short host_f16_rn(float x)
{
    const __half h = __float2half_rn(x);
    unsigned short u;
    memcpy(&u, &h, 2);
    return u;
}
static inline float host_f16_to_f(unsigned short u)
{
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
}

// images of the B operands: B[n][k] = W[k][n] (K-major), N padded to 128, K padded to 16 / 128
static void nn_build_images(const float *P, std::vector<unsigned char> &img, std::vector<float> &vecs)
{
    img.assign(TC_IMG_TOTAL, 0);
    vecs.assign(TC_VEC_FLOATS, 0.0f);
    const int woff[3] = {NN_W0, NN_W1, NN_W2}, boff[3] = {NN_B0, NN_B1, NN_B2}, kreal[3] = {NN_IN, NN_H, NN_H};
    size_t base = 0;
    for (int l = 0; l < 3; ++l) {
        const int kpad = l == 0 ? 16 : 128;
        const size_t piece = (size_t)128 * kpad * 2;
        for (int n = 0; n < NN_H; ++n)
            for (int k = 0; k < kreal[l]; ++k) {
                const float w = P[woff[l] + k * NN_H + n];
                const unsigned short h1 = host_f16_rn(w);
                const float r1 = w - host_f16_to_f(h1);
                const unsigned short h2 = host_f16_rn(r1);
                const size_t off = (size_t)(k / 8) * 2048 + (size_t)(n / 8) * 128 + (size_t)(n % 8) * 16 + (size_t)(k % 8) * 2;
                const unsigned short hs[TC_NPIECE] = {h1, h2};
                for (int p = 0; p < TC_NPIECE; ++p) memcpy(&img[base + p * piece + off], &hs[p], 2);
            }
        for (int n = 0; n < NN_H; ++n) vecs[l * 128 + n] = P[boff[l] + n];
        base += TC_NPIECE * piece;
    }
    for (int n = 0; n < NN_H; ++n) vecs[3 * 128 + n] = P[NN_W3 + n];
    vecs[3 * 128 + 128] = P[NN_B3];
}

static int nn_upload_weights(NNDeviceWeights &w, const float *packed)
{
    std::vector<unsigned char> img;
    std::vector<float> vecs;
    nn_build_images(packed, img, vecs);
    if (!w.packed) PVD_CUDA(cudaMalloc((void **)&w.packed, (size_t)NN_NPARAM * 4));
    if (!w.images) PVD_CUDA(cudaMalloc((void **)&w.images, TC_IMG_TOTAL));
    if (!w.vecs) PVD_CUDA(cudaMalloc((void **)&w.vecs, TC_VEC_FLOATS * 4));
    PVD_CUDA(cudaMemcpy(w.packed, packed, (size_t)NN_NPARAM * 4, cudaMemcpyHostToDevice));
    PVD_CUDA(cudaMemcpy(w.images, img.data(), TC_IMG_TOTAL, cudaMemcpyHostToDevice));
    PVD_CUDA(cudaMemcpy(w.vecs, vecs.data(), TC_VEC_FLOATS * 4, cudaMemcpyHostToDevice));
    return PVD_OK;
}

constexpr size_t NN_SMEM_BYTES = 2 * NN_H * NN_TILE * sizeof(float);
// Which kernel evaluates the network: chosen once (environment at first use, or pvd_nn_config), never looked up per launch.
struct NNPathCfg {
    int path = 0;        // 0 tcgen05, two tiles in flight (default); 1 tcgen05, one tile (activations through shared memory); 2 float32 FMA on CUDA cores
    int terms = TC2_TERMS;
    int threads = 512;
    bool init = false;
};
static NNPathCfg g_nn_cfg;
static const NNPathCfg &nn_cfg()
{
    if (!g_nn_cfg.init) {
        const char *e = getenv("PVD_NN_FP32"), *e1 = getenv("PVD_NN_TC1"), *et = getenv("PVD_NN_TERMS"), *eth = getenv("PVD_NN_THREADS");
        if (e && e[0] == '1') g_nn_cfg.path = 2;
        else if (e1 && e1[0] == '1') g_nn_cfg.path = 1;
        if (et && (et[0] == '3' || et[0] == '4')) g_nn_cfg.terms = et[0] - '0';
        if (eth && atoi(eth) == 1024) g_nn_cfg.threads = 1024;
        g_nn_cfg.init = true;
    }
    return g_nn_cfg;
}
static bool nn_use_cuda_cores() { return nn_cfg().path == 2; }
static int nn_prepare_launch()
{
    static bool done = false;
    if (!done) {
        PVD_CUDA(cudaFuncSetAttribute(k_nn_h4o2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NN_SMEM_BYTES));
        PVD_CUDA(cudaFuncSetAttribute(k_nn_h4o2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
        PVD_CUDA(cudaFuncSetAttribute(k_nn_h4o2_tc2<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC2_SMEM_BYTES));
        PVD_CUDA(cudaFuncSetAttribute(k_nn_h4o2_tc2<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC2_SMEM_BYTES));
        done = true;
    }
    return PVD_OK;
}
static NNDeviceWeights g_nn;     // process-wide weights for the stand-alone entry point

// coords: soa != 0 -> SoA with stride cap and n from the device state; else AoS with n_fixed walkers
static int nn_launch(cudaStream_t stream, const double *x, int soa, long long cap, const DevState *st, int parity, long long n_fixed,
                     long long n_upper, double *v, const NNDeviceWeights &w)
{
    if (!w.packed) return pvd_fail(PVD_E_STATE, "NN potential: weights not set");
    if (int rc = nn_prepare_launch()) return rc;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (nn_use_cuda_cores()) {
        long long g = (n_upper + NN_TILE - 1) / NN_TILE;
        if (g > 3ll * sms) g = 3ll * sms;
        k_nn_h4o2<<<(int)(g < 1 ? 1 : g), NN_THREADS, NN_SMEM_BYTES, stream>>>(x, soa, cap, st, parity, n_fixed, w.packed, v, nullptr);
    } else {
        if (nn_cfg().path == 1) {                                // previous generation (activations through shared memory), for A/B runs
            long long g = (n_upper + TC_M - 1) / TC_M;
            if (g > sms) g = sms;                                // persistent: one CTA per SM (192 KB of shared memory each)
            k_nn_h4o2_tc<<<(int)(g < 1 ? 1 : g), TC_THREADS, TC_SMEM_BYTES, stream>>>(x, soa, cap, st, parity, n_fixed, w.images, w.vecs, v);
        } else {
            long long g = (n_upper + 2 * TC_M - 1) / (2 * TC_M);
            if (g > sms) g = sms;                                // persistent: one CTA per SM (all 512 TMEM columns), two tiles in flight
            const int nterms = nn_cfg().terms;
            if (nn_cfg().threads == 1024)                        // measured: 2.77e9 walkers/s vs 2.88e9 with 512 threads
                k_nn_h4o2_tc2<1024><<<(int)(g < 1 ? 1 : g), 1024, TC2_SMEM_BYTES, stream>>>(x, soa, cap, st, parity, n_fixed, w.images, w.vecs, v, nterms);
            else
                k_nn_h4o2_tc2<512><<<(int)(g < 1 ? 1 : g), 512, TC2_SMEM_BYTES, stream>>>(x, soa, cap, st, parity, n_fixed, w.images, w.vecs, v, nterms);
        }
    }
    PVD_CHECK_LAUNCH();
    return PVD_OK;
}
