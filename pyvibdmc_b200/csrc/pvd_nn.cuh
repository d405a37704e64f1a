// NN potential for (H2O)2: Coulomb-matrix descriptor + the shipped 15-120-120-120-1 Keras MLP
// (TensorflowPots/call_sample_model.py:4-9; descriptor: tensorflow_descriptors/distance_descriptors.py
// :102-113,154-168; model layout read from sample_h4o2_nn.h5: Dense(120,swish) x3, Dense(1,relu), float32).
//
// Round-1 implementation: float32 FMA on the CUDA cores.  A CTA of 128 threads owns a tile of 64
// walkers; activations live in shared memory k-major ([feature][walker]) so that one LDS.128
// feeds four FMAs, thread j accumulates output neuron j for all 64 walkers in registers, weights
// stream through L1 (row k of a layer is one coalesced 480-byte read per CTA).
// TODO(round 2): move the three 120x120 layers onto tcgen05 (M = 128 walkers, K/N padded to 128,
// accumulators in TMEM, swish in the epilogue) -- see DESIGN.md.
#pragma once
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

constexpr size_t NN_SMEM_BYTES = 2 * NN_H * NN_TILE * sizeof(float);
static int nn_prepare_launch()
{
    static bool done = false;
    if (!done) {
        PVD_CUDA(cudaFuncSetAttribute(k_nn_h4o2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NN_SMEM_BYTES));
        done = true;
    }
    return PVD_OK;
}
static float *g_nn_weights = nullptr;     // process-wide copy for the stand-alone entry point

static int nn_launch_soa(cudaStream_t stream, const double *x, const DevState *st, int parity, long long cap, double *v, int grid,
                         const float *weights)
{
    if (!weights) return pvd_fail(PVD_E_STATE, "NN potential: weights not set (pvd_sim_set_nn_weights)");
    if (int rc = nn_prepare_launch()) return rc;
    k_nn_h4o2<<<grid, NN_THREADS, NN_SMEM_BYTES, stream>>>(x, 1, cap, st, parity, 0, weights, v, nullptr);
    PVD_CHECK_LAUNCH();
    return PVD_OK;
}
